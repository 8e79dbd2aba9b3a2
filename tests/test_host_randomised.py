"""Seeded randomised checks of the host-side pieces added around the hot path (CPU only): array-level NNI moves
against the pointer swaps on tree objects, the z / alpha / ? tags against the oracle's partition, the native
ultrametric solver against its numpy model on random trees with random fixed / zero branches, and the converter
to the reference's input form."""
import numpy as np
import pytest

from laml_b200 import ultrametric as um
from laml_b200.annotate import TAG_MISSING, TAG_Z, az_tags
from laml_b200.flatten import flatten_trees, forest_from_parent
from laml_b200.nni import apply_nni, nni_candidates
from laml_b200.synth import Workload, make_workload, random_tree, reference_inputs
from laml_b200.tree import read_tree_newick
from oracle.pmm_oracle import OracleProblem


def _clades(parent, left, right, n_leaves):
    """the set of leaf sets of a forest given as arrays"""
    n = len(parent)
    sets = [None] * n
    order = forest_from_parent(parent, left, right).order
    for v in order:
        sets[v] = frozenset([int(v)]) if left[v] < 0 else sets[left[v]] | sets[right[v]]
    return set(sets)


@pytest.mark.parametrize("seed", range(6))
def test_array_nni_equals_pointer_swap(seed):
    rng = np.random.default_rng(seed)
    w = make_workload(int(rng.integers(6, 40)), 3, 1, 4, 0.1, 0.1, seed=seed)
    newicks, msa, Q = reference_inputs(w)
    tree = read_tree_newick(newicks[0])
    f = flatten_trees([tree], leaf_key={"c%d" % i: i for i in range(w.forest.n_leaves)})
    cu, ck = nni_candidates(f.parent, f.left, f.n_leaves)
    assert len(cu) == 2 * (f.n_leaves - 2)
    for i in rng.choice(len(cu), size=min(8, len(cu)), replace=False):
        u_id, k = int(cu[i]), int(ck[i])
        p2, l2, r2 = apply_nni(f.parent, f.left, f.right, u_id, k)
        # the reference's pointer swap (Topology_search.py:281-287) on a fresh copy of the objects
        t2 = read_tree_newick(newicks[0])
        g = flatten_trees([t2], leaf_key={"c%d" % j: j for j in range(f.n_leaves)})
        u = g.nodes[u_id]; v = u.parent
        wn = next(c for c in v.children if c is not u)
        child = u.children[k]
        u.remove_child(child); v.add_child(child); v.remove_child(wn); u.add_child(wn)
        h = flatten_trees([t2], leaf_key={"c%d" % j: j for j in range(f.n_leaves)})
        assert _clades(p2, l2, r2, f.n_leaves) == _clades(h.parent, h.left, h.right, f.n_leaves)
        # ids, and with them branch lengths, stay with their nodes
        assert (np.sort(p2[p2 >= 0]) == np.sort(f.parent[f.parent >= 0])).all()
        assert p2[child_id := int(f.right[u_id] if k else f.left[u_id])] == f.parent[u_id]


@pytest.mark.parametrize("seed", range(5))
def test_az_tags_equal_oracle_partition(seed):
    w = make_workload(30 + 7 * seed, 25, 1, 6 + seed, 0.3, 0.4, seed=40 + seed, caterpillar_bias=0.3 * (seed % 2))
    ref = OracleProblem(w.forest, np.arange(w.forest.n_leaves, dtype=np.int32), w.codes, w.qptr, w.qval).az_partition()
    got = az_tags(w.forest, w.codes)
    assert ((got == TAG_Z) == (ref == -1)).all() and ((got == TAG_MISSING) == (ref == -2)).all()
    assert (got[got > 0] == ref[ref > 0]).all()


@pytest.mark.parametrize("seed", range(6))
def test_native_ultrametric_equals_numpy_on_random_problems(seed):
    rng = np.random.default_rng(100 + seed)
    f = random_tree(int(rng.integers(3, 300)), seed=seed, caterpillar_bias=0.5 * (seed % 2))
    n = f.n_nodes
    kind = np.zeros(n, np.int8)
    fixed = np.zeros(n)
    internal = np.nonzero((f.left >= 0) & (f.parent >= 0))[0]
    if len(internal):
        pick = rng.choice(internal, size=max(1, len(internal) // 6), replace=False)
        kind[pick[: len(pick) // 2 + 1]] = um.FIXED
        fixed[kind == um.FIXED] = rng.uniform(0.02, 0.3, int((kind == um.FIXED).sum()))
        kind[pick[len(pick) // 2 + 1:]] = um.ZERO
    S = rng.random((n, 5)) * np.array([5, 1, 0.3, 1, 0.3]) + 1e-6
    S = S / S.sum(axis=1, keepdims=True) * 50
    P = um.UltrametricProblem(f.parent, kind, fixed, 0.005, 10.0)
    nu = float(rng.choice([0.0, 0.05, 0.7]))
    a, Ha, ia = um.solve_numpy(P, S, nu)
    b, Hb, ib = um.solve(P, S, nu)
    if a is None:
        assert b is None
        return
    assert ia["status"] == ib["status"]
    np.testing.assert_allclose(b, a, rtol=0, atol=1e-7)
    assert P.feasible(Hb, strict=False, tol=1e-9)
    # tips level: every root-to-leaf path has the same length (zero-kind edges count as 0)
    d = b.copy(); d[kind == um.ZERO] = 0.0
    depth = np.zeros(n)
    for v in f.order[::-1]:
        depth[v] = d[v] + (depth[f.parent[v]] if f.parent[v] >= 0 else 0.0)
    tips = depth[: f.n_leaves]
    assert np.ptp(tips) < 1e-9 * max(1.0, tips.mean())


def test_reference_inputs_round_trip():
    w = make_workload(50, 20, 1, 9, 0.2, 0.3, seed=8)
    newicks, msa, Q = reference_inputs(w)
    t = read_tree_newick(newicks[0])
    f = flatten_trees([t], leaf_key={"c%d" % i: i for i in range(50)})
    assert _clades(f.parent, f.left, f.right, 50) == _clades(w.forest.parent, w.forest.left, w.forest.right, 50)
    by_label = {nd.label: nd.edge_length for nd in t.traverse_leaves()}
    assert all(by_label["c%d" % i] == float(w.d[i]) for i in range(50))
    from laml_b200.flatten import encode_sites
    enc = encode_sites(msa, Q)
    assert (enc.codes == w.codes).all() and np.allclose(enc.qval, w.qval) and (enc.qptr == w.qptr).all()
