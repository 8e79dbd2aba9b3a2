"""The clade schedule of the lumped-state engine (host logic of pmm_set_tree, pure C++: runs without a GPU).

What the kernels in laml_b200/csrc/pmm_lumped.cuh rely on, checked on random, caterpillar and forest
topologies for several clade capacities:
  * every internal node is exactly one inside step and one outside step; a clade has at most `cap` steps and
    belongs to one tree; launches (phases) partition the clades;
  * inside: a chained child (kind 2) is the second child and was produced by the step just before; a loaded child
    (kind 1) was produced at least two steps earlier in the same clade, or in a clade of an earlier launch
    (then, and only then, its "earlier launch" bit is set);
  * outside: a step whose out-record is chained (how = 2) follows the step of its parent, which lists it first and
    has the keep bit; a loaded out-record (how = 1) was stored (store bit) by the parent's step at least two steps
    earlier in the same clade or in an earlier launch; a root loads too (the inside pass stored its out-record);
  * leaves are flagged as leaves;
  * dataflow launches (several tiers in one launch): the wait list of a clade holds exactly the clades that produce
    what it loads from other clades, every one of them at an earlier position of the pass; and the item order of a
    launch (a re-statement of ``df_item``) visits every (clade, site block) once, producers before consumers.
"""
import numpy as np
import pytest

from laml_b200.capi import lumped_schedule
from laml_b200.synth import random_tree

KIND_LEAF, KIND_LOAD, KIND_PREV = 0, 1, 2


def forest_of(trees):
    """concatenate binary trees (leaves first) into one forest"""
    N = sum(t.n_leaves for t in trees)
    n = sum(t.n_nodes for t in trees)
    parent = np.full(n, -1, np.int32); left = np.full(n, -1, np.int32); right = np.full(n, -1, np.int32)
    lo, io = 0, N
    for t in trees:
        Nt, nt = t.n_leaves, t.n_nodes

        def m(a):
            return np.where(a < 0, -1, np.where(a < Nt, a + lo, a - Nt + io))
        ids = m(np.arange(nt))
        parent[ids] = m(t.parent); left[ids] = m(t.left); right[ids] = m(t.right)
        lo += Nt; io += nt - Nt
    return N, parent, left, right


def check(N, parent, left, right, cap, top):
    n = len(parent)
    n_int = n - N
    S = lumped_schedule(N, parent, left, right, cap, top)
    ncl, nph = S["n_clades"], S["n_phases"]
    assert ncl >= 1 and nph >= 1
    root_of = np.arange(n)
    order = [v for v in range(n) if parent[v] < 0]
    for v in order:
        if left[v] >= 0:
            for c in (left[v], right[v]):
                root_of[c] = root_of[v]; order.append(c)
    roots = sorted(v for v in range(n) if parent[v] < 0)
    tree_index = {r: i for i, r in enumerate(roots)}

    for which in ("in", "out"):
        steps, clades, pp = S[which + "_steps"], S[which + "_clades"], S[which + "_phase_ptr"]
        assert sorted(steps[:, 0].tolist()) == list(range(N, n))                       # every internal node once
        assert pp[0] == 0 and pp[-1] == ncl and (np.diff(pp) >= 0).all()
        assert sorted(clades[:, 3].tolist()) == list(range(ncl))                       # clade ids
        assert clades[0, 0] == 0 and clades[-1, 1] == n_int and (clades[1:, 0] == clades[:-1, 1]).all()
        sizes = clades[:, 1] - clades[:, 0]
        assert (sizes >= 1).all()
        phase_of_clade = np.repeat(np.arange(nph), np.diff(pp))
        step_clade = np.repeat(np.arange(ncl), sizes)
        step_pos = {int(u): i for i, u in enumerate(steps[:, 0])}
        for ci in range(ncl):
            us = steps[clades[ci, 0]:clades[ci, 1], 0]
            assert (np.array([tree_index[root_of[u]] for u in us]) == clades[ci, 2]).all()   # one tree per clade
            cap_here = cap if phase_of_clade[ci] == (0 if which == "in" else nph - 1) else top
            assert sizes[ci] <= min(cap_here, 127)
        for i, (u, v, w, f) in enumerate(steps.tolist()):
            assert {v, w} == {int(left[u]), int(right[u])}
            ph, cl = phase_of_clade[step_clade[i]], step_clade[i]
            if which == "in":
                kv, kw = f & 3, (f >> 2) & 3
                assert bool(f & 16) == (parent[u] < 0)
                for c, k, far in ((v, kv, f & 32), (w, kw, f & 64)):
                    assert (k == KIND_LEAF) == (c < N)
                    if k == KIND_LOAD:
                        j = step_pos[c]
                        same = step_clade[j] == cl
                        assert bool(far) == (not same)
                        assert (j <= i - 2) if same else (phase_of_clade[step_clade[j]] < ph)
                    else:
                        assert not far
                assert kv != KIND_PREV
                if kw == KIND_PREV:
                    assert i > clades[cl, 0] and steps[i - 1, 0] == w                 # produced by the step just before
            else:
                how = f & 3
                assert bool(f & 4) == (v < N) and bool(f & 8) == (w < N)
                if parent[u] < 0:
                    assert how == KIND_LOAD                    # a root's out-record is stored by the inside pass' root step
                if parent[u] >= 0:
                    j = step_pos[int(parent[u])]
                    pf = steps[j, 3]
                    first = steps[j, 1] == u
                    if how == KIND_PREV:
                        assert j == i - 1 and first and (pf & 64) and step_clade[j] == cl
                    else:
                        assert pf & (16 if first else 32)                              # the parent's step stored it
                        same = step_clade[j] == cl
                        assert (j <= i - 2) if same else (phase_of_clade[step_clade[j]] < ph)
                if f & 64:
                    assert v >= N and i + 1 < clades[cl, 1] and steps[i + 1, 0] == v and (steps[i + 1, 3] & 3) == KIND_PREV
                for c, bit in ((v, 16), (w, 32)):
                    if c >= N and not ((f & 64) and c == v):
                        assert f & bit                                                 # everything not chained is stored
    # ---- dataflow waits: exactly the other clades a clade loads from, all at earlier positions
    for which in ("in", "out"):
        steps, clades = S[which + "_steps"], S[which + "_clades"]
        ptr, deps = S[which + "_dep_ptr"], S[which + "_deps"]
        sizes = clades[:, 1] - clades[:, 0]
        step_clade = np.repeat(np.arange(ncl), sizes)
        pos_of_node = {int(u): int(step_clade[i]) for i, u in enumerate(steps[:, 0])}
        assert len(ptr) == ncl + 1 and ptr[0] == 0 and ptr[-1] == len(deps)
        for ci in range(ncl):
            want = set()
            for u, v, w, f in steps[clades[ci, 0]:clades[ci, 1]].tolist():
                if which == "in":
                    want |= {pos_of_node[c] for c, far in ((v, f & 32), (w, f & 64)) if far}
                elif parent[u] >= 0 and pos_of_node[int(parent[u])] != ci:
                    want.add(pos_of_node[int(parent[u])])
            got = deps[ptr[ci]:ptr[ci + 1]].tolist()
            assert sorted(got) == sorted(want) and len(set(got)) == len(got)
            assert all(p < ci for p in got)
    return S


def df_item(i, n1, n2, nblk):
    """the item order of a dataflow launch, as laml_b200/csrc/pmm_lumped.cuh: df_item"""
    if i < n1:
        return i, 0
    j = i - n1
    g, r = divmod(j, n1 + n2)
    if g < nblk - 1:
        return r, (g + 1 if r < n1 else g)
    return n1 + r, nblk - 1


@pytest.mark.parametrize("cap,top", [(3, 2), (16, 6), (40, 8)])
def test_dataflow_item_order(cap, top):
    """every (clade, site block) of a dataflow launch is visited once, and after all of its producers -- for the launch
    over all tiers (mode 2) and for the launch over the top tiers alone (mode 1), both passes, 1 to 5 site blocks"""
    t = random_tree(300, seed=8)
    S = lumped_schedule(t.n_leaves, t.parent, t.left, t.right, cap, top)
    ncl, nt0 = S["n_clades"], int(S["in_phase_ptr"][1])
    assert S["n_phases"] >= 3
    for which, groups in (("in", ((0, ncl, nt0, ncl - nt0), (nt0, ncl, 0, ncl - nt0))),
                          ("out", ((0, ncl, ncl - nt0, nt0), (0, ncl - nt0, ncl - nt0, 0)))):
        ptr, deps = S[which + "_dep_ptr"], S[which + "_deps"]
        for c0, c1, n1, n2 in groups:
            assert n1 + n2 == c1 - c0
            for nblk in (1, 2, 5):
                seen = {}
                for i in range((c1 - c0) * nblk):
                    cl, blk = df_item(i, n1, n2, nblk)
                    assert 0 <= cl < c1 - c0 and 0 <= blk < nblk and (cl, blk) not in seen
                    for p in deps[ptr[c0 + cl]:ptr[c0 + cl + 1]].tolist():
                        assert p < c0 or (p - c0, blk) in seen           # produced by an earlier launch, or earlier in this one
                    seen[(cl, blk)] = i
                assert len(seen) == (c1 - c0) * nblk


@pytest.mark.parametrize("cap,top", [(1, 1), (2, 3), (7, 2), (16, 16), (40, 8), (127, 127)])
def test_schedule_invariants_random_and_caterpillar(cap, top):
    for seed, n_leaves, bias in ((1, 2, 0.0), (2, 3, 0.0), (3, 57, 0.0), (4, 400, 0.0), (5, 150, 1.0), (6, 300, 0.5)):
        t = random_tree(n_leaves, seed=seed, caterpillar_bias=bias)
        check(t.n_leaves, t.parent, t.left, t.right, cap, top)


def test_schedule_invariants_forest():
    trees = [random_tree(n, seed=10 + i) for i, n in enumerate((40, 2, 75, 3))]
    N, parent, left, right = forest_of(trees)
    for cap, top in ((1, 1), (5, 5), (16, 4), (100, 16)):
        S = check(N, parent, left, right, cap, top)
        assert len(set(S["in_clades"][:, 2].tolist())) == 4


def test_one_launch_covers_most_of_a_big_tree():
    """tier 0 (every maximal subtree with at most `cap` internal nodes) holds the bulk of a random tree"""
    t = random_tree(20000, seed=21)
    S = lumped_schedule(t.n_leaves, t.parent, t.left, t.right, 64, 16)
    c = S["in_clades"]
    tier0 = c[:S["in_phase_ptr"][1]]
    steps0 = int((tier0[:, 1] - tier0[:, 0]).sum())
    assert steps0 >= 0.9 * (t.n_nodes - t.n_leaves)
    assert S["n_phases"] <= 8
    # best-fit packing: clades of tier 0 are nearly full
    assert (tier0[:, 1] - tier0[:, 0]).mean() >= 0.85 * 64
