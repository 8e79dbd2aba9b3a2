"""Shared helpers for the test-suite: load golden fixtures into the array form
both the oracle and the engine take."""
import json
import os

import numpy as np

from laml_b200.flatten import encode_sites, flatten_trees
from laml_b200.tree import read_tree_newick

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

NPZ_CASES = ["ref_rand_small", "ref_rand_ragged", "ref_rand_nu0", "ref_rand_phi0", "ref_rand_nomissing",
             "ref_rand_zeroprior", "ref_rand_forest", "ref_rand_deep", "ref_test4_interior", "ref_test4_zero",
             "ref_test45_test1", "ref_example1_optimum", "ref_example1_start"]


def matrix_to_msa(cells, matrix):
    """Inverse of make_golden.msa_to_array: -1 -> '?', -2 -> literal -1."""
    msa = {}
    for c, row in zip(cells, matrix):
        msa[str(c)] = ["?" if x == -1 else (-1 if x == -2 else int(x)) for x in row]
    return msa


def csr_to_Q(qptr, qstate, qraw):
    return [{int(qstate[j]): float(qraw[j]) for j in range(qptr[s], qptr[s + 1])} for s in range(len(qptr) - 1)]


class Case:
    """A problem in both reference form (newick, msa dict, Q dicts) and flat form."""

    def __init__(self, trees, msa, Q, phi, nu):
        self.newicks = list(trees)
        self.msa, self.Q, self.phi, self.nu = msa, Q, float(phi), float(nu)
        self.trees = [read_tree_newick(t) for t in trees]
        self.forest = flatten_trees(self.trees)
        self.enc = encode_sites(msa, Q)
        row = {c: i for i, c in enumerate(self.enc.cells)}
        self.leaf_row = np.array([row[l] for l in self.forest.leaf_labels], np.int32)
        self.d = self.forest.edge_lengths()
        self.label_to_id = {str(nd.label): i for i, nd in enumerate(self.forest.nodes)}

    def ids(self, labels, tree_of=None):
        """node ids for the golden's (label, tree index) rows"""
        if tree_of is None or len(self.trees) == 1:
            return np.array([self.label_to_id[str(l)] for l in labels])
        per_tree = []
        for t in self.trees:
            per_tree.append({str(nd.label): id(nd) for nd in t.traverse_postorder()})
        pos = {id(nd): i for i, nd in enumerate(self.forest.nodes)}
        return np.array([pos[per_tree[int(ti)][str(l)]] for l, ti in zip(labels, tree_of)])

    def oracle(self):
        from oracle.pmm_oracle import OracleProblem
        return OracleProblem(self.forest, self.leaf_row, self.enc.codes, self.enc.qptr, self.enc.qval)


def load_npz_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    msa = matrix_to_msa(z["cells"], z["matrix"])
    Q = csr_to_Q(z["qptr"], z["qstate"], z["qraw"])
    case = Case([str(t) for t in z["trees"]], msa, Q, float(z["phi"]), float(z["nu"]))
    return case, z


def load_json(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def json_case(rec):
    msa = {c: [x for x in row] for c, row in rec["msa"].items()}
    Q = [{int(k): v for k, v in q.items()} for q in rec["Q"]]
    return Case([rec["T"]], msa, Q, rec["phi"], rec["nu"])


def unjson(rows):
    return np.array([[float("-inf") if x == "-inf" else x for x in row] for row in rows], np.float64)


def assert_close_loginf(a, b, rtol, atol, what=""):
    """compare arrays that may hold -inf at matching places"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    ia, ib = np.isneginf(a), np.isneginf(b)
    assert (ia == ib).all(), "%s: -inf pattern differs at %s" % (what, np.argwhere(ia != ib)[:5])
    np.testing.assert_allclose(a[~ia], b[~ib], rtol=rtol, atol=atol, err_msg=what)


def make_engine(case, device=0, **kw):
    """Engine for a :class:`Case` (codes re-ordered to leaf-id order)."""
    from laml_b200.capi import Engine
    eng = Engine(case.forest.n_leaves, case.enc.qptr, case.enc.qval, case.enc.codes[case.leaf_row], device=device, **kw)
    eng.set_tree(case.forest.parent, case.forest.left, case.forest.right)
    return eng


class env_override:
    """temporarily set environment variables read by pmm_create / pmm_set_tree"""

    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
