"""Host-side logic: tree container, flattening, site encoding, site partitioning,
synthetic generator -- CPU only."""
import numpy as np
import pytest

from laml_b200.dist import partition_sites
from laml_b200.flatten import encode_sites, flatten_trees
from laml_b200.synth import make_workload, random_priors, random_tree
from laml_b200.tree import read_tree_newick


def test_newick_roundtrip_and_orders():
    t = read_tree_newick("[&R] ((a:1,b:0.5)e:1,(c:1,d:2.25)f:1)r:0.2;")
    assert t.newick() == "[&R] ((a:1,b:0.5)e:1,(c:1,d:2.25)f:1)r:0.2;"
    assert [n.label for n in t.traverse_postorder()] == ["a", "b", "e", "c", "d", "f", "r"]
    assert [n.label for n in t.traverse_preorder()] == ["r", "f", "d", "c", "e", "b", "a"]   # last child first
    assert t.root.edge_length == 0.2 and abs(t.height() - 3.45) < 1e-12


def test_unifurcations_and_polytomies():
    t = read_tree_newick("(((a:1):2,b:1):1,(c:1,d:1,e:1):1);")
    t.suppress_unifurcations()
    a = t.label_to_node()["a"]
    assert a.edge_length == 3
    t.resolve_polytomies()
    assert all(len(n.children) in (0, 2) for n in t.traverse_preorder())
    new = [n for n in t.traverse_preorder() if n.edge_length == 0 and n.label is None and n.children]
    assert len(new) == 1


def test_flatten_leaves_first_and_order():
    t = read_tree_newick("((a:1,b:1)e:1,(c:1,d:1)f:1)r:1;")
    f = flatten_trees([t])
    assert f.n_leaves == 4 and f.n_nodes == 7
    assert (f.left[:4] == -1).all() and (f.left[4:] >= 0).all()
    pos = {v: i for i, v in enumerate(f.order)}
    for v in range(7):
        if f.left[v] >= 0:
            assert pos[f.left[v]] < pos[v] and pos[f.right[v]] < pos[v]
    assert f.depth[f.roots[0]] == 0 and f.height[f.roots[0]] == 2
    f2 = flatten_trees([t], leaf_key={"d": 0, "c": 1, "b": 2, "a": 3})
    assert f2.leaf_labels == ["d", "c", "b", "a"]


def test_encode_sites():
    msa = {"x": [0, 5, "?"], "y": [3, -1, 7]}
    Q = [{3: 2.0, 9: 2.0}, {0: 0, 5: 1.0}, {7: 0.25, 2: 0.75}]
    enc = encode_sites(msa, Q)
    assert enc.qptr.tolist() == [0, 2, 3, 5]
    np.testing.assert_allclose(enc.qval, [0.5, 0.5, 1.0, 0.75, 0.25])
    assert enc.codes.tolist() == [[0, 1, -1], [1, -2, 2]]
    with pytest.raises(KeyError):
        encode_sites({"x": [4]}, [{3: 1.0}])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_sites_balanced(world):
    qptr, _ = random_priors(400, 2, 100, seed=3)
    for balance in ("sites", "slots"):
        parts = partition_sites(qptr, world, balance=balance)
        assert parts[0][0] == 0 and parts[-1][1] == 400
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and all(hi > lo for lo, hi in parts)
        if balance == "slots":          # equal bytes for the full-state row engine
            slots = [int(qptr[hi] - qptr[lo] + 2 * (hi - lo)) for lo, hi in parts]
            assert max(slots) - min(slots) <= 2 * 102
        else:                           # equal (node, site) records for the lumped engine
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_shard_consistency():
    full = make_workload(200, 40, 2, 20, 0.2, 0.2, seed=9)
    part = make_workload(200, 40, 2, 20, 0.2, 0.2, seed=9, sites=(13, 29))
    assert (full.codes[:, 13:29] == part.codes).all()
    assert ((full.codes >= -1) & (full.codes <= np.diff(full.qptr)[None, :])).all()
    t = random_tree(50, seed=4, caterpillar_bias=0.9)
    assert t.depth.max() > 15
