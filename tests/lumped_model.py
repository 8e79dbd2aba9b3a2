"""numpy model of the LUMPED-state engine's arithmetic -- test infrastructure.

The lumped engine (laml_b200/csrc/pmm_lumped.cuh) does not store a node's whole state
vector.  Because edits are irreversible, the inside vector of any node is, at every site,

    x_alpha = xb + [alpha == h] (xh - xb)         (one "hot" edited state at most)

and hot implies xb = 0 and x_m = 0 exactly (a subtree that shows an edited state cannot sit
under a silenced or differently edited ancestor); the outside vector is

    out_alpha = oB q_alpha + [alpha == k] oD .

So a (node, site) needs the record (x0, xm, xz, qh, h) -- xz is xb without a hot state and the
hot value with one -- and (o0, om, oB, oD, k): every sum over the alphabet has a closed form
in qsum = sum_alpha q_alpha and q_h.  This is SURVEY.md section 8's "exact lumping" taken per
(node, site) instead of per site.  The model below is vectorised over sites and loops over
nodes exactly as one GPU thread does for its site; formulas: SURVEY.md 3.5,
EM_solver.py:91-300.
"""
import math

import numpy as np

from fullstate_model import branch_consts


def _expo(mx):
    """exponent e with mx * 2^-e in [1,2); 0 for zero/denormal maxima"""
    m, e = np.frexp(mx)
    return np.where(mx >= 2.2250738585072014e-308, e - 1, 0).astype(np.int64)


def _leaf_view(code, qoff, qval, phi):
    """(x0, xm, xz, qh, h, cum) of a leaf from its state codes (EM_solver.py:103-116)"""
    K = code.shape[0]
    x0 = np.where(code == 0, 1.0 - phi, np.where(code == -1, phi, 0.0))
    xm = np.where(code < 0, 1.0, 0.0)
    xz = np.where(code == -1, phi, np.where(code > 0, 1.0 - phi, 0.0))
    h = np.where(code > 0, code - 1, -1)
    qh = np.where(code > 0, qval[np.minimum(qoff + np.maximum(h, 0), len(qval) - 1)], 0.0)
    return x0, xm, xz, qh, h, np.zeros(K, np.int64)


def _message(view, bc, qsum):
    """child message: E0, Em, Eb (non-hot slots), Eh (hot slot), plus sq = sum q x and xb"""
    x0, xm, xz, qh, h, _ = view
    a, b, c, g = bc
    hot = h >= 0
    xb = np.where(hot, 0.0, xz)
    sq = np.where(hot, xz * qh, xz * qsum)
    E0 = a * x0 + b * sq + g * xm
    Eb = c * xb + g * xm
    Eh = c * xz + g * xm
    return E0, xm, Eb, Eh, sq, xb


def estep(forest, leaf_row, codes, qptr, qval, d, nu, phi, dump=False):
    n = forest.parent.shape[0]
    N = forest.n_leaves
    K = codes.shape[1]
    qptr = np.asarray(qptr, np.int64)
    qval = np.asarray(qval, np.float64) if len(qval) else np.zeros(1)
    qoff = qptr[:-1]
    qsum = np.array([qval[qptr[s]:qptr[s + 1]].sum() for s in range(K)])
    A, B, C, G = branch_consts(np.asarray(d, np.float64), nu)
    left, right, parent = forest.left, forest.right, forest.parent
    S = np.zeros((n, 5))
    site_llh = np.zeros(K)
    Rt = 0.0
    rec = [None] * n
    for v in range(N):
        rec[v] = _leaf_view(codes[leaf_row[v]].astype(np.int64), qoff, qval, phi)

    # ---------------- inside
    for u in forest.order:
        if left[u] < 0:
            continue
        v, w = left[u], right[u]
        vv, vw = rec[v], rec[w]
        E0v, Emv, Ebv, Ehv, _, _ = _message(vv, (A[v], B[v], C[v], G[v]), qsum)
        E0w, Emw, Ebw, Ehw, _, _ = _message(vw, (A[w], B[w], C[w], G[w]), qsum)
        hv, hw = vv[4], vw[4]
        x0 = E0v * E0w
        xm = Emv * Emw
        xb = Ebv * Ebw
        both = (hv >= 0) & (hw >= 0)
        same = both & (hv == hw)
        onlyv = (hv >= 0) & (hw < 0)
        onlyw = (hw >= 0) & (hv < 0)
        h = np.where(same | onlyv, hv, np.where(onlyw, hw, -1))
        xz = np.where(same, Ehv * Ehw, np.where(onlyv, Ehv * Ebw, np.where(onlyw, Ebv * Ehw, xb)))
        qh = np.where(same | onlyv, vv[3], np.where(onlyw, vw[3], 0.0))
        # conflicting hot states: every edited slot is exactly zero
        assert np.all(xb[both] == 0.0) and np.all((Ehv * Ebw)[both & ~same] == 0.0)
        e = _expo(np.maximum(np.maximum(x0, xm), xz))
        sc = np.ldexp(1.0, -e)
        rec[u] = (x0 * sc, xm * sc, xz * sc, qh, h, vv[5] + vw[5] + e)

    # ---------------- roots: site likelihood
    invL = {}
    root_llh = {}
    for r in np.nonzero(parent < 0)[0]:
        E0, _, _, _, sq, _ = _message(rec[r], (A[r], B[r], C[r], G[r]), qsum)
        with np.errstate(divide="ignore"):
            root_llh[r] = np.where(E0 > 0, np.log(E0), -np.inf) + rec[r][5] * math.log(2.0)
            site_llh += root_llh[r]
        inv = np.where(E0 > 0, 1.0 / np.where(E0 > 0, E0, 1.0), 0.0)
        invL[r] = inv
        S[r, 0] += (A[r] * rec[r][0] * inv).sum()
        S[r, 1] += (B[r] * sq * inv).sum()
        S[r, 2] += (G[r] * rec[r][1] * inv).sum()

    # ---------------- outside + counts
    root_of = np.zeros(n, np.int64)
    for v in forest.order[::-1]:
        root_of[v] = v if parent[v] < 0 else root_of[parent[v]]
    out = [None] * n
    for u in forest.order[::-1]:
        if left[u] < 0:
            continue
        if parent[u] < 0:
            z = np.zeros(K)
            out[u] = (A[u] + z, G[u] + z, B[u] + z, z, np.full(K, -1, np.int64), -rec[u][5])
        o0, om, oB, oD, k, F = out[u]
        kids = (left[u], right[u])
        views = [rec[c] for c in kids]
        bcs = [(A[c], B[c], C[c], G[c]) for c in kids]
        msgs = [_message(vw_, bc, qsum) for vw_, bc in zip(views, bcs)]
        delta = np.clip(F + views[0][5] + views[1][5], -1000, 1000)
        fac = np.ldexp(invL[root_of[u]], delta)
        for i in (0, 1):
            v, w = kids[i], kids[1 - i]
            (x0v, xmv, xzv, qhv, hv, _), (_, _, _, qhw, hw, cumw) = views[i], views[1 - i]
            av, bv, cv, gv = bcs[i]
            _, _, _, _, sqv, xbv = msgs[i]
            E0w, Emw, Ebw, Ehw, _, _ = msgs[1 - i]
            hotw, hotv, hotk = hw >= 0, hv >= 0, k >= 0
            T0 = o0 * E0w
            Tm = om * Emw
            sqE = Ebw * qsum + np.where(hotw, (Ehw - Ebw) * qhw, 0.0)            # sum_a q_a E^w_a
            Ewk = np.where(hotk & (k == hw), Ehw, Ebw)                              # E^w at slot k
            sT = oB * sqE + np.where(hotk, oD * Ewk, 0.0)                          # sum_a T_a
            # out[v]
            n0 = av * T0
            nm = gv * (T0 + sT) + Tm
            nB = bv * T0 + cv * oB * Ebw
            nD = np.where(hotw, cv * ((oB * qhw + np.where(k == hw, oD, 0.0)) * Ehw - oB * qhw * Ebw),
                          np.where(hotk, cv * oD * Ebw, 0.0))
            nk = np.where(hotw, hw, np.where(hotk, k, -1))
            # counts of the branch above v
            Ewhv = np.where(hotv & (hv == hw), Ehw, Ebw)                            # E^w at slot h_v
            sqEx = xbv * sqE + np.where(hotv, (xzv - xbv) * qhv * Ewhv, 0.0)       # sum_a q_a E^w_a x^v_a
            xvk = np.where(hotk & (k == hv), xzv, xbv)                              # x^v at slot k
            sTx = oB * sqEx + np.where(hotk, oD * Ewk * xvk, 0.0)                   # sum_a T_a x^v_a
            S[v, 0] += (T0 * av * x0v * fac).sum()
            S[v, 1] += (T0 * bv * sqv * fac).sum()
            S[v, 2] += (T0 * gv * xmv * fac).sum()
            S[v, 3] += (cv * sTx * fac).sum()
            S[v, 4] += (gv * xmv * sT * fac).sum()
            if left[v] < 0:
                q = codes[leaf_row[v]] == -1
                Rt += (1.0 - (nm * xmv * fac)[q]).sum()
            else:
                e = _expo(np.maximum(np.maximum(n0, nm), np.maximum(nB, nD)))
                sc = np.ldexp(1.0, -e)
                out[v] = (n0 * sc, nm * sc, nB * sc, nD * sc, nk, F + cumw + e)
    res = {"llh": float(site_llh.sum()), "site_llh": site_llh, "S": S, "Rtilde": float(Rt)}
    if dump:        # the records themselves, as the engine stores them (pmm_lumped.cuh): inside (x0, xm, xz, qh, h, cum), outside (o0, om, oB, oD, k, F)
        res.update(rec=rec, out=out, invL=invL, root_llh=root_llh, bc=np.stack([A, B, C, G], axis=1), qsum=qsum)
    return res
