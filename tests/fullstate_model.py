"""numpy model of the ENGINE's arithmetic (full-state vectors, linear space,
power-of-two lazy scaling) -- test infrastructure.

It mirrors, site by site and in plain loops, exactly what the CUDA kernels in
laml_b200/csrc/pmm_kernels.cu compute (same formulas, same scaling rule), so
that the engine's *math* can be checked against the oracle on a machine without
a GPU, and so that a failing GPU parity test can be bisected per node.
Spec: SURVEY.md section 3.5; reference formulas cited there.
"""
import math

import numpy as np


def branch_consts(d, nu):
    """a = p^(1+nu), b = p^nu (1-p), c = p^nu, g = 1-p^nu   (EM_solver.py:124-133)"""
    a = np.exp(-(1.0 + nu) * d)
    c = np.exp(-nu * d)
    b = c * (-np.expm1(-d))
    g = -np.expm1(-nu * d)
    return a, b, c, g


def expo(mx):
    """exponent e with mx * 2^-e in [1,2); 0 for zero/denormal"""
    if mx <= 0 or mx < 2.2250738585072014e-308:
        return 0
    return math.frexp(mx)[1] - 1


def leaf_vector(code, A, phi):
    """[x0, xm, x_1..x_A]   (EM_solver.py:103-116 in full-state form)"""
    x = np.zeros(A + 2)
    if code == 0:
        x[0] = 1 - phi
    elif code == -1:
        x[:] = phi
        x[1] = 1.0
    elif code == -2:
        x[1] = 1.0
    else:
        x[1 + code] = 1 - phi
    return x


def estep(forest, leaf_row, codes, qptr, qval, d, nu, phi):
    n = forest.parent.shape[0]
    N = forest.n_leaves
    K = codes.shape[1]
    a, b, c, g = branch_consts(np.asarray(d, np.float64), nu)
    left, right, parent = forest.left, forest.right, forest.parent
    S = np.zeros((n, 5))
    site_llh = np.zeros(K)
    Rt = 0.0
    root_of = np.zeros(n, np.int64)
    for v in forest.order[::-1]:
        root_of[v] = v if parent[v] < 0 else root_of[parent[v]]
    for s in range(K):
        q = qval[qptr[s]:qptr[s + 1]]
        A = q.shape[0]
        x = [None] * n
        cum = np.zeros(n, np.int64)
        for v in range(N):
            x[v] = leaf_vector(int(codes[leaf_row[v], s]), A, phi)

        def message(v):
            """normalised child message E (as [E0, Em, E_alpha...]) and its total exponent"""
            xv = x[v]
            e = expo(xv.max())
            xh = np.ldexp(xv, -e)
            sq = float(np.dot(q, xh[2:]))
            E = np.empty(A + 2)
            E[0] = a[v] * xh[0] + b[v] * sq + g[v] * xh[1]
            E[1] = xh[1]
            E[2:] = c[v] * xh[2:] + g[v] * xh[1]
            return E, cum[v] + e, xh, sq

        for u in forest.order:
            if left[u] < 0:
                continue
            Ev, ev, _, _ = message(left[u])
            Ew, ew, _, _ = message(right[u])
            x[u] = Ev * Ew
            cum[u] = ev + ew
        Lm, Le = {}, {}
        for r in np.nonzero(parent < 0)[0]:
            E, e, xh, sq = message(r)
            L = E[0]
            Lm[r], Le[r] = L, e
            site_llh[s] += (math.log(L) if L > 0 else -math.inf) + e * math.log(2.0)
            if L > 0:
                S[r, 0] += a[r] * xh[0] / L
                S[r, 1] += b[r] * sq / L
                S[r, 2] += g[r] * xh[1] / L
        out = [None] * n
        oexp = np.zeros(n, np.int64)
        for u in forest.order[::-1]:
            if left[u] < 0:
                continue
            if parent[u] < 0:
                o = np.empty(A + 2)
                o[0] = a[u]; o[1] = g[u]; o[2:] = b[u] * q
                oe = 0
            else:
                o, oe = out[u], oexp[u]
            eo = expo(o.max())
            oh = np.ldexp(o, -eo)
            kids = (left[u], right[u])
            msgs = [message(k) for k in kids]
            r = root_of[u]
            for (v, (Ev, ev, xh, sq)), (w, (Ew, ew, _, _)) in zip(zip(kids, msgs), zip(kids[::-1], msgs[::-1])):
                T = oh * Ew
                sT = float(T[2:].sum())
                ov = np.empty(A + 2)
                ov[0] = a[v] * T[0]
                ov[2:] = b[v] * q * T[0] + c[v] * T[2:]
                ov[1] = g[v] * (T[0] + sT) + T[1]
                out[v] = ov
                oexp[v] = oe + eo + ew
                delta = int(oe + eo + ew + ev - Le[r])
                fac = math.ldexp(1.0 / Lm[r], delta) if Lm[r] > 0 else 0.0
                S[v, 0] += T[0] * a[v] * xh[0] * fac
                S[v, 1] += T[0] * b[v] * sq * fac
                S[v, 2] += T[0] * g[v] * xh[1] * fac
                S[v, 3] += c[v] * float(np.dot(T[2:], xh[2:])) * fac
                S[v, 4] += g[v] * xh[1] * sT * fac
                if left[v] < 0 and codes[leaf_row[v], s] == -1:
                    Rt += 1.0 - ov[1] * xh[1] * fac
    return {"llh": float(site_llh.sum()), "site_llh": site_llh, "S": S, "Rtilde": Rt}
