"""The constrained M-step of the batch NNI kernel (laml_b200/csrc/pmm_nni_ultra.cuh: two unknown node heights per
candidate) against the whole-tree solver it specialises (pmm_ultrametric_solve with every branch but the five
around the move fixed) -- both are host code of libpmm_b200.so, so this runs without a GPU.  What the batch kernel
adds around it (statistics, EM loop) is covered on the GPU by tests/test_zzz_gpu_nni_ultra.py."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from laml_b200 import ultrametric as um                      # noqa: E402
from laml_b200.capi import nni_ultra_mstep, ultrametric_solve  # noqa: E402
from laml_b200.nni import apply_nni, nni_candidates            # noqa: E402
from laml_b200.synth import random_tree                        # noqa: E402

LO, HI = 0.005, 10.0


def ultrametric_lengths(forest, rng, step=(0.02, 0.4)):
    """random node heights (leaves at 0) -> branch lengths; the root branch is random"""
    n = forest.n_nodes
    H = np.zeros(n)
    order = []
    stack = [int(r) for r in np.nonzero(forest.parent < 0)[0]]
    while stack:
        v = stack.pop()
        order.append(v)
        if forest.left[v] >= 0:
            stack += [int(forest.left[v]), int(forest.right[v])]
    for v in reversed(order):
        if forest.left[v] >= 0:
            H[v] = max(H[forest.left[v]], H[forest.right[v]]) + rng.uniform(*step)
    d = np.where(forest.parent >= 0, H[np.maximum(forest.parent, 0)] - H, rng.uniform(*step))
    return d, H


def local_problem(forest, d, H, u, k):
    """arrays of the whole-tree problem with five free branches, and the constants of the two-height problem"""
    parent, left, right = apply_nni(forest.parent, forest.left, forest.right, u, k)
    v = int(forest.parent[u])
    w = int(forest.right[v] if forest.left[v] == u else forest.left[v])
    ck = int(forest.right[u] if k else forest.left[u])
    co = int(forest.left[u] if k else forest.right[u])
    five = [v, u, ck, co, w]
    kind = np.full(forest.n_nodes, um.FIXED, np.int8)
    kind[five] = um.FREE
    g = int(forest.parent[v])
    consts = dict(hA=H[co], hB=H[w], hC=H[ck], hTop=H[g] if g >= 0 else 0.0, top_free=g < 0)
    return parent, kind, five, consts


def random_stats(rng, n, K=40.0):
    S = rng.gamma(0.7, 1.0, size=(n, 5)) + 1e-6
    return S / S.sum(axis=1, keepdims=True) * K


@pytest.mark.parametrize("seed,nu", [(1, 0.15), (2, 0.0), (3, 0.6)])
def test_two_height_mstep_matches_whole_tree_solver(seed, nu):
    rng = np.random.default_rng(seed)
    forest = random_tree(40, seed=seed)
    d, H = ultrametric_lengths(forest, rng)
    cu, ck = nni_candidates(forest.parent, forest.left, forest.n_leaves)
    n_ok = n_bad = 0
    for u, k in zip(cu.tolist(), ck.tolist()):
        parent, kind, five, c = local_problem(forest, d, H, u, k)
        S = random_stats(rng, forest.n_nodes)
        d_full, H_full, rc_full, it_full = ultrametric_solve(parent, kind, d, LO, HI, S, nu)
        d5, H3, rc, it = nni_ultra_mstep(S[five], nu, LO, HI, **c)
        assert rc == rc_full, (u, k, rc, rc_full)
        if rc != 0:
            n_bad += 1
            continue
        n_ok += 1
        # both stop on the same Newton-decrement test (1e-13 n_free mu): lengths agree to the solver's own precision
        np.testing.assert_allclose(d5, d_full[five], rtol=2e-7, atol=2e-8)
        assert abs(it - it_full) <= 2                      # the same barrier path, step for step
        # warm start from the optimum with new statistics (the next EM iteration)
        S2 = S * rng.uniform(0.8, 1.25, size=S.shape)
        d_full2, _, rc_full2, it_full2 = ultrametric_solve(parent, kind, d, LO, HI, S2, nu, H0=H_full)
        d5b, _, rc2, it2 = nni_ultra_mstep(S2[five], nu, LO, HI, H=H3, **c)
        assert rc2 == rc_full2 == 0
        np.testing.assert_allclose(d5b, d_full2[five], rtol=2e-7, atol=2e-8)
        assert it2 < it and abs(it2 - it_full2) <= 2       # warm starts are shorter
        # the five lengths keep the tree ultrametric: heights re-derived from them close the equalities
        Hu, Hv = c["hA"] + d5b[3], c["hC"] + d5b[2]
        assert abs(c["hB"] + d5b[4] - Hu) < 1e-12 and abs(Hu + d5b[1] - Hv) < 1e-12
        if not c["top_free"]:
            assert abs(Hv + d5b[0] - c["hTop"]) < 1e-12
    assert n_ok > 20


def test_infeasible_candidates_are_reported():
    """short branches above the move leave no room for u' and v' between the former sibling and the top"""
    rng = np.random.default_rng(7)
    forest = random_tree(30, seed=5)
    d, H = ultrametric_lengths(forest, rng, step=(0.006, 0.012))
    cu, ck = nni_candidates(forest.parent, forest.left, forest.n_leaves)
    seen = set()
    for u, k in zip(cu.tolist(), ck.tolist()):
        parent, kind, five, c = local_problem(forest, d, H, u, k)
        S = random_stats(rng, forest.n_nodes)
        _, _, rc_full, _ = ultrametric_solve(parent, kind, d, LO, HI, S, 0.1)
        _, _, rc, _ = nni_ultra_mstep(S[five], 0.1, LO, HI, **c)
        assert rc == rc_full
        seen.add(rc)
    assert 1 in seen


def test_forest_root_hangs_from_the_common_top():
    """v a root of a two-tree forest: its branch ends at the common height of the trees (EM_solver.py:80-83)"""
    from laml_b200.flatten import forest_from_parent
    rng = np.random.default_rng(11)
    t1, t2 = random_tree(12, seed=3), random_tree(9, seed=4)
    # merge: leaves of both first (ids 0..N-1), then the internal nodes
    n1l, n2l = t1.n_leaves, t2.n_leaves
    N = n1l + n2l

    def remap(t, leaf0, int0):
        return lambda x: -1 if x < 0 else (leaf0 + x if x < t.n_leaves else int0 + x - t.n_leaves)
    m1, m2 = remap(t1, 0, N), remap(t2, n1l, N + t1.n_nodes - n1l)
    n = t1.n_nodes + t2.n_nodes
    parent = np.full(n, -1, np.int32); left = np.full(n, -1, np.int32); right = np.full(n, -1, np.int32)
    for t, m in ((t1, m1), (t2, m2)):
        for x in range(t.n_nodes):
            parent[m(x)], left[m(x)], right[m(x)] = m(int(t.parent[x])), m(int(t.left[x])), m(int(t.right[x]))
    forest = forest_from_parent(parent, left, right)
    d, H = ultrametric_lengths(forest, rng)
    roots = np.nonzero(forest.parent < 0)[0]
    top = max(H[r] for r in roots) + 0.3
    d[roots] = top - H[roots]
    checked = 0
    for u, k in zip(*[a.tolist() for a in nni_candidates(forest.parent, forest.left, forest.n_leaves)]):
        if forest.parent[forest.parent[u]] >= 0:
            continue
        p2, kind, five, c = local_problem(forest, d, H, u, k)
        c["hTop"], c["top_free"] = top, False
        S = random_stats(rng, n)
        d_full, _, rc_full, _ = ultrametric_solve(p2, kind, d, LO, HI, S, 0.2)
        d5, _, rc, _ = nni_ultra_mstep(S[five], 0.2, LO, HI, **c)
        assert rc == rc_full
        if rc == 0:
            np.testing.assert_allclose(d5, d_full[five], rtol=2e-7, atol=2e-8)
            checked += 1
    assert checked >= 2
