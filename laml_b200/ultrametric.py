"""Ultrametric-constrained branch-length M-step without a dense constraint matrix.

The reference maximises (``laml_libs/EM_solver.py:351-357``)

    sum_b  -(nu+1) S0_b d_b - nu (S1_b + S3_b) d_b + S1_b log(1 - e^{-d_b}) + (S2_b + S4_b) log(1 - e^{-nu d_b})

over ``dmin <= d <= dmax`` subject to the ultrametric equalities ``M d = b`` it builds at
``EM_solver.py:53-89`` (one length-B Python list per leaf: O(N B) memory) and hands the program to
cvxpy -> MOSEK (commercial; not installable here).  The same program is solved here exactly, in O(n)
memory, by re-parametrising with node heights: every node gets a height ``H`` above the leaves
(leaves at 0), a free branch is ``d_v = H_parent - H_v``, a fixed branch (``mark_fixed``) is the
equality ``H_parent - H_v = length`` and a polytomy-marked branch counts as 0 in the paths exactly as
in the reference (``EM_solver.py:60-79``).  A virtual super-root above the roots turns the root
branches (and the reference's "all trees have the same height" rows, ``:80-83``) into ordinary edges.
The equalities vanish; what is left is a concave objective in ``H`` with box constraints on
differences along tree edges.  It is solved by a log-barrier Newton method whose linear systems have
the tree's own sparsity: one children-to-parents sweep eliminates every node (the Schur complement of
an edge is the "series" combination ``w b / (w + b)`` of its curvature with the child's), one
parents-to-children sweep back-substitutes.  The objective is strictly concave in ``d`` so the
optimum is unique: it is the one MOSEK returns.

The product path is the C++ routine behind ``pmm_ultrametric_solve`` (:func:`solve`); this module keeps the
topology-dependent set-up (:class:`UltrametricProblem`) and the numpy statement of the algorithm
(:func:`solve_numpy`) it is checked against.
"""
from __future__ import annotations

import numpy as np

FREE, FIXED, ZERO = 0, 1, 2        # edge kinds: variable in [lo,hi]; given length; contributes 0 (polytomy mark)


class UltrametricProblem:
    """Topology-dependent part (levels, edge kinds); reusable across EM iterations."""

    def __init__(self, parent, kind, fixed_len, lo, hi):
        parent = np.asarray(parent, np.int64)
        n = parent.shape[0]
        # virtual super-root = node n; every root hangs below it by its own (free/fixed) branch
        self.n = n
        self.par = np.where(parent < 0, n, parent)
        self.kind = np.asarray(kind, np.int8)
        self.fixed_len = np.asarray(fixed_len, np.float64)
        self.lo, self.hi = float(lo), float(hi)
        is_leaf = np.ones(n + 1, bool)
        is_leaf[self.par] = False
        is_leaf[n] = False
        self.is_leaf = is_leaf[:n]
        depth = np.zeros(n + 1, np.int64)
        # depth by walking up (trees are shallow compared with n; vectorised pointer jumping)
        cur = self.par.copy()
        d = np.ones(n, np.int64)
        alive = cur != n
        while alive.any():
            d[alive] += 1
            cur[alive] = self.par[cur[alive]]
            alive = cur != n
        depth[:n] = d
        self.depth = depth
        self.levels = [np.nonzero(depth[:n] == l)[0] for l in range(1, int(d.max()) + 1)]   # level 1 = roots
        self.free = self.kind == FREE

    # -- feasibility helpers ------------------------------------------------------------------
    def interior_heights(self, margin=None):
        """A strictly feasible height assignment: every free edge strictly inside its box."""
        n = self.n
        H = np.zeros(n + 1)
        lo = self.lo
        margin = margin if margin is not None else max(1e-3, 0.5 * lo)
        for nodes in self.levels[::-1]:                     # deepest first: children before parents
            need = H[nodes] + np.where(self.kind[nodes] == FREE, lo + margin,
                                       np.where(self.kind[nodes] == FIXED, self.fixed_len[nodes], 0.0))
            np.maximum.at(H, self.par[nodes], need)
        # rigid edges need equality: a parent lifted by another child lifts the rigid child (and with it the
        # child's whole subtree, whose own free edges only get longer); a lifted leaf means infeasible and is
        # caught by feasible()
        for nodes in self.levels:
            rg = nodes[(self.kind[nodes] != FREE) & ~self.is_leaf[nodes]]
            ln = np.where(self.kind[rg] == FIXED, self.fixed_len[rg], 0.0)
            H[rg] = H[self.par[rg]] - ln
        return H

    def edge_lengths(self, H):
        return H[self.par] - H[: self.n]

    def feasible(self, H, strict=True, tol=1e-12):
        x = self.edge_lengths(H)
        f = self.free
        ok_free = ((x[f] > self.lo) & (x[f] < self.hi)).all() if strict else ((x[f] >= self.lo - tol) & (x[f] <= self.hi + tol)).all()
        fx = self.kind == FIXED
        z = self.kind == ZERO
        ok_fix = np.allclose(x[fx], self.fixed_len[fx], atol=1e-9) and np.allclose(x[z], 0.0, atol=1e-9)
        return bool(ok_free and ok_fix and (H[: self.n][self.is_leaf] == 0).all())


def _f_derivs(x, k, S1, S24, nu, use_nu):
    """value, first and second derivative of the per-branch objective at lengths x"""
    em = np.expm1(x)
    val = -k * x + S1 * np.log(-np.expm1(-x))
    g = -k + S1 / em
    h = -S1 * (em + 1.0) / (em * em)
    if use_nu:
        emn = np.expm1(nu * x)
        val = val + S24 * np.log(-np.expm1(-nu * x))
        g = g + nu * S24 / emn
        h = h - nu * nu * S24 * (emn + 1.0) / (emn * emn)
    return val, g, h


def solve(problem: UltrametricProblem, S, nu, eps_nu=1e-5, H0=None, mu_final=1e-11, verbose=False, warm_theta=1e-6,
          warm_mu=1e-9):
    """argmax of EM_solver.py:351-357 under the ultrametric equalities and the box: the library's host routine
    ``pmm_ultrametric_solve`` (``laml_b200/csrc/pmm_ultrametric.cpp``, declared in ``include/pmm_b200.h``).
    ``S`` : (n,5) clamped / rescaled statistics of the branches (rows of non-free branches are ignored).
    Returns ``(d, H, info)`` with ``d`` the lengths of all branches (fixed / zero ones as given), or
    ``(None, None, {"status": "infeasible"})``.  :func:`solve_numpy` is the same algorithm in numpy -- the model
    the C++ is tested against (tests/test_ultrametric_host.py)."""
    from .capi import ultrametric_solve
    P = problem
    parent = np.where(P.par[: P.n] == P.n, -1, P.par[: P.n]).astype(np.int32)
    d, H, rc, iters = ultrametric_solve(parent, P.kind, P.fixed_len, P.lo, P.hi, S, nu, eps_nu=eps_nu, H0=H0,
                                        mu_final=mu_final, warm_theta=warm_theta, warm_mu=warm_mu)
    if rc == 1:
        return None, None, {"status": "infeasible"}
    return d, H, {"status": "optimal" if rc == 0 else "failed_d", "newton_iters": iters}


def solve_numpy(problem: UltrametricProblem, S, nu, eps_nu=1e-5, H0=None, mu_final=1e-11, verbose=False, warm_theta=1e-6,
                warm_mu=1e-9):
    """argmax of EM_solver.py:351-357 under the ultrametric equalities and the box.
    ``S`` : (n,5) clamped / rescaled statistics of the branches (rows of non-free branches are ignored).
    Returns ``(d, H, info)`` with ``d`` the lengths of all branches (fixed / zero ones as given).

    ``H0`` (the heights returned for the previous EM iteration) warm-starts the barrier path: the start is
    ``H0`` pulled ``warm_theta`` of the way to an interior point and the barrier weight starts at ``warm_mu``
    instead of 1 -- consecutive M-steps of an EM run differ little, so the path from mu = 1 is mostly wasted.
    Same optimum to 10 digits on the test problems with 24-32 instead of 118-128 Newton steps per M-step."""
    P = problem
    n = P.n
    free = P.free
    S = np.asarray(S, np.float64)
    k = (nu + 1.0) * S[:, 0] + nu * (S[:, 1] + S[:, 3])
    S1 = S[:, 1]
    S24 = S[:, 2] + S[:, 4]
    use_nu = bool(nu > eps_nu)
    lo, hi = P.lo, P.hi
    Hint = P.interior_heights()
    warm = H0 is not None and P.feasible(H0, strict=False)
    if warm:
        H = (1.0 - warm_theta) * np.asarray(H0, np.float64) + warm_theta * Hint      # convex combination: strictly interior
    else:
        H = Hint
    if not P.feasible(H, strict=True):
        return None, None, {"status": "infeasible"}
    par = P.par
    fidx = np.nonzero(free)[0]
    rigid = ~free

    def objective(Hc, mu):
        x = (Hc[par] - Hc[:n])[fidx]
        if (x <= lo).any() or (x >= hi).any():
            return -np.inf
        val, _, _ = _f_derivs(x, k[fidx], S1[fidx], S24[fidx], nu, use_nu)
        return float(val.sum() + mu * (np.log(x - lo).sum() + np.log(hi - x).sum()))

    def newton_direction(Hc, mu):
        x = Hc[par] - Hc[:n]
        g = np.zeros(n); w = np.zeros(n)
        xf = x[fidx]
        _, gf, hf = _f_derivs(xf, k[fidx], S1[fidx], S24[fidx], nu, use_nu)
        g[fidx] = gf + mu * (1.0 / (xf - lo) - 1.0 / (hi - xf))
        w[fidx] = hf - mu * (1.0 / (xf - lo) ** 2 + 1.0 / (hi - xf) ** 2)
        # children-to-parents elimination: reduced model a*delta + b*delta^2/2 per node (b <= 0; -inf = grounded)
        a = np.zeros(n + 1); b = np.zeros(n + 1)
        grounded = np.zeros(n + 1, bool)
        grounded[:n] = P.is_leaf
        alpha = np.zeros(n); beta = np.zeros(n)
        for nodes in P.levels[::-1]:
            p = par[nodes]
            gr = grounded[nodes]
            rg = rigid[nodes]
            A = np.zeros(len(nodes)); B = np.zeros(len(nodes))
            # free edge, grounded child: delta_c = 0
            m = ~rg & gr
            A[m] = g[nodes[m]]; B[m] = w[nodes[m]]
            # free edge, free child: series combination
            m = ~rg & ~gr
            if m.any():
                c = nodes[m]
                s = w[c] + b[c]
                alpha[c] = (g[c] - a[c]) / s
                beta[c] = w[c] / s
                A[m] = (g[c] * b[c] + a[c] * w[c]) / s
                B[m] = w[c] * b[c] / s
            # rigid edge: delta_c = delta_p
            m = rg & ~gr
            if m.any():
                c = nodes[m]
                A[m] = a[c]; B[m] = b[c]
                beta[c] = 1.0
            np.add.at(a, p, A)
            np.add.at(b, p, B)
            # a rigid edge to a grounded child grounds the parent
            m = rg & gr
            if m.any():
                grounded[p[m]] = True
        delta = np.zeros(n + 1)
        if not grounded[n] and b[n] < 0:
            delta[n] = -a[n] / b[n]
        for nodes in P.levels:
            gr = grounded[nodes]
            c = nodes[~gr]
            delta[c] = alpha[c] + beta[c] * delta[par[c]]
        slope = float((g * (delta[par] - delta[:n])).sum())     # directional derivative of the barrier objective
        return delta, slope

    n_free = max(len(fidx), 1)
    mu = warm_mu if warm else 1.0
    iters = 0
    status = "optimal"
    while True:
        for _ in range(60):
            delta, slope = newton_direction(H, mu)
            iters += 1
            if not np.isfinite(slope):
                status = "failed_d"; break
            if slope <= 1e-13 * n_free * max(mu, 1e-12) + 1e-15:
                break
            # largest step that keeps every free edge strictly inside its box
            dx = (delta[par] - delta[:n])[fidx]
            x = (H[par] - H[:n])[fidx]
            with np.errstate(divide="ignore", invalid="ignore"):
                tmax = np.min(np.where(dx < 0, (x - lo) / (-dx), np.where(dx > 0, (hi - x) / dx, np.inf)))
            t = min(1.0, 0.99 * tmax)
            f0 = objective(H, mu)
            while t > 1e-14:
                Hn = H + t * delta
                if objective(Hn, mu) >= f0 + 1e-4 * t * slope:
                    break
                t *= 0.5
            else:
                break
            H = Hn
            if t * slope < 1e-14 * n_free:
                break
        if status != "optimal" or mu <= mu_final:
            break
        mu = max(mu * 0.1, mu_final)
    d = P.edge_lengths(H)
    d[P.kind == FIXED] = P.fixed_len[P.kind == FIXED]
    d[free] = np.clip(d[free], lo, hi)
    return d, H, {"status": status, "newton_iters": iters, "mu": mu}


def nu_given_d(S, fm, d, numsites, eps_s=1e-6):
    """argmax over nu in [0,1] of EM_solver.py:394-399 with d fixed: a scalar concave program on O(n) sums.
    ``S`` (n,5) raw statistics; rows are clamped and rescaled as EM_solver.py:343-346."""
    idx = np.nonzero(fm)[0]
    s = np.maximum(S[idx], eps_s)
    s = s / s.sum(axis=1, keepdims=True) * numsites
    dd = d[idx]
    lin = float(((s[:, 0] + s[:, 1] + s[:, 3]) * dd).sum())
    w = s[:, 2] + s[:, 4]

    def grad(x):
        return float((w * dd / np.expm1(x * dd)).sum()) - lin

    def grad2(x):
        em = np.expm1(x * dd)
        return -float((w * dd * dd * (em + 1.0) / (em * em)).sum())

    if grad(1.0) >= 0:
        return 1.0
    lo, hi = 0.5, 1.0
    while grad(lo) <= 0:
        hi, lo = lo, lo * 0.5
        if lo < 1e-300:
            return 0.0
    # grad is decreasing and convex on (0, 1]: Newton from the left end of the bracket rises monotonically to
    # the root; the bracket is kept only as a safeguard
    x = lo
    for _ in range(100):
        g = grad(x)
        if g > 0:
            lo = x
        else:
            hi = x
        nx = x - g / grad2(x)
        if not (lo < nx < hi):
            nx = 0.5 * (lo + hi)
        if abs(nx - x) <= 4e-16 * x:
            x = nx
            break
        x = nx
    return x


def mstep(problem: UltrametricProblem, S, Rtilde, R, n_cells, numsites, fm, d, nu, phi, H_prev=None, optimize_phi=True,
          optimize_nu=True, eps_nu=1e-5, eps_s=1e-6):
    """One M-step under the ultrametric constraint from all-reduced statistics (``EM_solver.Mstep`` with
    ``ultra_constr_cache``, EM_solver.py:307-459): branch lengths by :func:`solve`, then nu with the new lengths
    fixed (:427-433) and phi in closed form (:318-326).  Pure function of its arguments -- every rank of a
    site-sharded run computes the same result from the same reduced statistics.
    Returns ``(d_new, nu_new, phi_new, H, status)`` with status "optimal" / "failed_d" / "infeasible"."""
    S = np.asarray(S, np.float64)
    s = np.maximum(S, eps_s)                                         # EM_solver.py:343-346
    s = s / s.sum(axis=1, keepdims=True) * numsites
    d_sol, H, info = solve(problem, s, nu, eps_nu=eps_nu, H0=H_prev)
    if d_sol is None:
        return d, nu, phi, H_prev, "infeasible"
    free = np.asarray(fm).astype(bool)
    d_new = np.array(d, np.float64, copy=True)
    d_new[free] = d_sol[free]
    nu_new = nu_given_d(S, fm, d_new, numsites, eps_s) if optimize_nu else nu
    phi_new = phi
    if optimize_phi:
        phi_new = Rtilde / (R + Rtilde)
        if abs(phi_new) < 1 / (numsites * n_cells):
            phi_new = 0
    return d_new, nu_new, phi_new, H, info["status"]
