"""TEST INFRASTRUCTURE ONLY: a stand-in for ``laml_b200.capi.Engine`` backed by the CPU oracle.

``GPU_EM_solver`` (laml_b200/solver.py) is host bookkeeping around the C ABI: newick handling,
initial points, fixed-branch marking, multi-start selection, status strings, the ultrametric
M-step on engine statistics.  On a machine without a GPU that logic would otherwise go untested,
so the ``-m "not gpu"`` suite swaps the engine class for this one (``monkeypatch`` of
``laml_b200.solver.Engine``) and runs the reference's own solver tests through the host code.
Nothing under ``laml_b200/`` imports this module; the product path has no CPU fallback and fails
loudly without the CUDA library (tests/test_capi_symbols.py::test_no_cpu_fallback).

Same method names, argument order and return shapes as ``capi.Engine``; the arithmetic is
``oracle/pmm_oracle.c`` (the reference's compressed z/alpha/? formulation), i.e. NOT the
engine's formulation -- agreement of the two is what the ``-m gpu`` parity tests establish.
"""
import numpy as np

from laml_b200.flatten import forest_from_parent
from laml_b200.nni import apply_nni
from oracle.pmm_oracle import OracleProblem


class OracleEngine:
    instances = 0                      # how many handles were created (tests check reuse across topologies)

    def __init__(self, n_leaves, qptr, qval, codes, device=0, mem_budget_bytes=0):
        codes = np.ascontiguousarray(codes, np.int32)
        if codes.ndim != 2 or codes.shape[0] != n_leaves:
            raise ValueError("codes must be (n_leaves, n_sites) in leaf-id order")
        self.n_leaves, self.n_sites = int(n_leaves), int(codes.shape[1])
        self.qptr, self.qval, self.codes = np.asarray(qptr, np.int32), np.asarray(qval, np.float64), codes
        self.n_nodes = 0
        self.device = device
        self._p = None
        self.closed = False
        self.set_tree_calls = 0
        OracleEngine.instances += 1

    # -- plumbing
    def close(self):
        self.closed = True

    def set_tree(self, parent, left, right):
        self._forest = forest_from_parent(np.asarray(parent, np.int32), np.asarray(left, np.int32),
                                          np.asarray(right, np.int32))
        self._p = OracleProblem(self._forest, np.arange(self.n_leaves, dtype=np.int32), self.codes, self.qptr, self.qval)
        self.n_nodes = int(len(parent))
        self.set_tree_calls += 1

    def info(self):
        return {"kernel_launches": 0, "n_nodes": self.n_nodes, "n_sites": self.n_sites}

    # -- host-buffer entry points
    def inside(self, d, nu, phi, want_site_llh=False):
        out = self._p.inside(d, nu, phi)
        return (out["llh"], out["site_llh"]) if want_site_llh else out["llh"]

    def estep(self, d, nu, phi, want_site_llh=False, out_S=None):
        r = self._p.estep(d, nu, phi)
        S = np.ascontiguousarray(r["S"].T)
        if out_S is not None:
            out_S[...] = S
            S = out_S
        out = {"S": S, "Rtilde": r["Rtilde"], "llh": r["llh"]}
        if want_site_llh:
            out["site_llh"] = r["site_llh"]
        return out

    def posteriors(self, d, nu, phi):
        r = self._p.estep(d, nu, phi, dump=True)
        return r["post0"], r["post1"]

    def mstep(self, S, Rtilde, R, n_cells, d, nu, phi, free_mask=None, optimize_nu=True, optimize_phi=True,
              dmin=0.005, dmax=10.0, eps_s=1e-6, eps_nu=1e-5):
        d_new, nu_new, phi_new, _ = self._p.mstep(np.ascontiguousarray(np.asarray(S).T), Rtilde, d, nu, free_mask,
                                                  optimize_phi, optimize_nu, phi, dmin, dmax, eps_s, eps_nu)
        return d_new, nu_new, phi_new

    def em(self, d, nu, phi, R, n_cells, free_mask=None, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0,
           conv_eps=1e-8, max_iter=1000):
        r = self._p.em(d, nu, phi, free_mask, optimize_phi, optimize_nu, dmin, dmax, conv_eps, max_iter)
        r["status"] = "optimal"
        return r

    def _check_ultrametric(self, d, tol=1e-9):
        f = self._forest
        hi, lo = np.zeros(f.n_nodes), np.zeros(f.n_nodes)
        order = [int(r) for r in np.nonzero(f.parent < 0)[0]]
        for v in order:
            if f.left[v] >= 0:
                order += [int(f.left[v]), int(f.right[v])]
        for v in reversed(order):
            if f.left[v] >= 0:
                a, b = int(f.left[v]), int(f.right[v])
                hi[v], lo[v] = max(hi[a] + d[a], hi[b] + d[b]), min(lo[a] + d[a], lo[b] + d[b])
                if hi[v] - lo[v] > tol:
                    raise ValueError("pmm_score_nni_batch_ultra: the current lengths are not ultrametric")
        tops = [hi[r] + d[r] for r in np.nonzero(f.parent < 0)[0]]
        if len(tops) > 1 and max(tops) - min(tops) > tol:
            raise ValueError("pmm_score_nni_batch_ultra: the trees of the forest are not equally high")

    def _ultra_em(self, prob, parent, five, d, nu, phi, dmin, dmax, conv_eps, max_iter):
        """EM with the five branches free under the ultrametric equalities: whole-tree host M-step
        (laml_b200.ultrametric.mstep, everything else FIXED) around the oracle's E-step -- the loop of
        ``GPU_EM_solver.EM_optimization(ultra_constr=True)``"""
        from laml_b200 import ultrametric as um
        n = self.n_nodes
        kind = np.full(n, um.FIXED, np.int8)
        kind[five] = um.FREE
        fm = (kind == um.FREE).astype(np.uint8)
        P = um.UltrametricProblem(parent, kind, d, dmin, dmax)
        d = np.array(d, np.float64, copy=True)
        pre = prob.inside(d, nu, phi)["llh"]
        H, it, cur = None, 1, None
        while True:
            r = prob.estep(d, nu, phi)
            d_new, _, _, H_new, status = um.mstep(P, np.ascontiguousarray(r["S"].T), r["Rtilde"], 1.0, 1, self.n_sites, fm, d,
                                                  nu, phi, H_prev=H, optimize_phi=False, optimize_nu=False)
            if status != "optimal":
                return None
            H = H_new
            d = np.where(fm.astype(bool), d_new, d)
            cur = prob.inside(d, nu, phi)["llh"]
            if abs((cur - pre) / pre) < conv_eps or it >= max_iter:
                break
            pre = cur
            it += 1
        return {"nllh": -cur, "d": d, "iters": it}

    def score_nni_batch(self, cand_u, cand_k, d, nu, phi, dmin=0.005, dmax=10.0, conv_eps=1e-8, max_iter=1000, ultra=False):
        """Every candidate by a full-tree EM with the five local branches free on the explicitly
        re-wired topology -- what ``Topology_search.apply_nni`` does with ``local_brlen_opt``
        (``Topology_search.py:253-300``).  Same output layout as ``pmm_score_nni_batch``: ``nodes`` =
        (v, u, moved child, other child, former sibling), include/pmm_b200.h."""
        f = self._forest
        m = len(cand_u)
        nllh = np.empty(m); d5 = np.empty((m, 5)); iters = np.empty(m, np.int32); nodes = np.empty((m, 5), np.int32)
        d = np.asarray(d, np.float64)
        if ultra:
            self._check_ultrametric(d)
        for i, (u, k) in enumerate(zip(cand_u, cand_k)):
            u, k = int(u), int(k)
            v = int(f.parent[u])
            w = int(f.right[v] if f.left[v] == u else f.left[v])
            moved = int(f.right[u] if k else f.left[u])
            stay = int(f.left[u] if k else f.right[u])
            p2, l2, r2 = apply_nni(f.parent, f.left, f.right, u, k)
            prob = OracleProblem(forest_from_parent(p2, l2, r2), np.arange(self.n_leaves, dtype=np.int32), self.codes,
                                 self.qptr, self.qval)
            five = [v, u, moved, stay, w]
            nodes[i] = five
            if ultra:
                r = self._ultra_em(prob, p2, five, d, nu, phi, dmin, dmax, conv_eps, max_iter)
                if r is None:                       # constrained M-step infeasible / failed: no score
                    nllh[i], iters[i], d5[i] = np.nan, -1, d[five]
                    continue
            else:
                fm = np.zeros(self.n_nodes, np.uint8)
                fm[five] = 1
                r = prob.em(d, nu, phi, fm, False, False, dmin, dmax, conv_eps, max_iter)
            nllh[i], iters[i] = r["nllh"], r["iters"]
            d5[i] = r["d"][five]
            nodes[i] = five
        return {"nllh": nllh, "d5": d5, "iters": iters, "nodes": nodes}
