"""``laml_b200.fast_em.EMOptimizer`` -- the reference's array-level EM interface (``fast_em/laml.py:196-276``)
-- on CPU with the oracle-backed engine stand-in: array conventions (dense zero-padded priors, recoded states,
networkx tree with leaves first, branch mask) against the dict / newick route through ``GPU_EM_solver``."""
import networkx as nx
import numpy as np
import pytest

import laml_b200.fast_em as fe
import laml_b200.solver as solver_mod
from helpers import load_npz_case
from oracle_engine import OracleEngine


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    monkeypatch.setattr(fe, "Engine", OracleEngine)
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)


def arrays_of(case):
    """what fastEM_solver hands to EMOptimizer (fastEM_solver.py:46-146, fast_em/phylogeny.py:48-114), built
    from a Case: leaves 0..N-1 = forest leaf ids, states recoded to 1..A per site, -1 unknown"""
    f, enc = case.forest, case.enc
    K = enc.n_sites
    A = int(np.diff(enc.qptr).max())
    pri = np.zeros((K, A))
    for s in range(K):
        pri[s, : enc.qptr[s + 1] - enc.qptr[s]] = enc.qval[enc.qptr[s]:enc.qptr[s + 1]]
    cm = enc.codes[case.leaf_row].copy()
    assert (cm != -2).all()                          # fast_em has no "known silenced" code
    g = nx.DiGraph()
    g.add_nodes_from(range(f.n_nodes))
    for v in range(f.n_nodes):
        if f.left[v] >= 0:
            g.add_edge(v, int(f.left[v])); g.add_edge(v, int(f.right[v]))
    return pri, cm, g


@pytest.mark.parametrize("name", ["ref_rand_nu0", "ref_test4_interior"])
def test_optimize_matches_the_solver_route(name):
    case, z = load_npz_case(name)
    if (case.enc.codes == -2).any():
        pytest.skip("case holds literal -1 entries")
    pri, cm, g = arrays_of(case)
    opt = fe.EMOptimizer(pri, cm)
    d0 = np.maximum(case.d, 0.01)
    out = opt.optimize(g, d0, 0.1, 0.1, np.ones(case.forest.n_nodes, bool))
    assert set(out) == {"nllh", "branch_lengths", "nu", "phi", "em_iterations"}
    ref = case.oracle().em(d0, 0.1, 0.1)
    assert out["em_iterations"] == ref["iters"] and abs(out["nllh"] - ref["nllh"]) <= 1e-12 * ref["nllh"]
    np.testing.assert_allclose(out["branch_lengths"], ref["d"], rtol=1e-12)
    # the same through the reference-protocol class: evaluate the optimum as a newick + dicts problem
    s = solver_mod.GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': out["phi"], 'nu': out["nu"]})
    f = s._ensure()
    by_label = {str(nd.label): float(length) for nd, length in zip(case.forest.nodes, out["branch_lengths"])}
    for nd in f.nodes:                               # the solver numbers its leaves by matrix row: match by label
        nd.edge_length = by_label[str(nd.label)]
    assert abs(s.negative_llh() - out["nllh"]) <= 1e-9 * out["nllh"]


def test_branch_mask_and_fixed_parameters():
    case, z = load_npz_case("ref_rand_nu0")
    pri, cm, g = arrays_of(case)
    opt = fe.EMOptimizer(pri, cm)
    n = case.forest.n_nodes
    d0 = np.maximum(case.d, 0.01)
    mask = np.zeros(n, bool); mask[::3] = True
    out = opt.optimize(g, d0, 0.2, 0.15, mask, fixed_nu=True, fixed_phi=True)
    assert out["nu"] == 0.2 and out["phi"] == 0.15
    assert (out["branch_lengths"][~mask] == d0[~mask]).all() and (out["branch_lengths"][mask] != d0[mask]).any()
    start = -case.oracle().inside(d0, 0.2, 0.15)["llh"]
    assert out["nllh"] <= start + 1e-9


def test_usage_errors():
    case, z = load_npz_case("ref_rand_nu0")
    pri, cm, g = arrays_of(case)
    with pytest.raises(ValueError):
        fe.EMOptimizer(pri[:-1], cm)
    bad = cm.copy(); bad[0, 0] = pri.shape[1] + 1
    with pytest.raises(ValueError):
        fe.EMOptimizer(pri, bad)
    opt = fe.EMOptimizer(pri, cm)
    h = g.copy(); h.remove_node(case.forest.n_nodes - 1)
    with pytest.raises(ValueError):
        opt.optimize(h, case.d, 0.1, 0.1, np.ones(case.forest.n_nodes, bool))


def test_reference_fastEM_test_1():
    """laml_unit_tests/unit_tests_fastEMSolver.py:15-44 at the EMOptimizer level: test4's data from the start
    fastEM_solver uses (nu = phi = 0.5, fastEM_solver.py:232-233) reaches nllh 1007.009158767873 with
    phi = nu = 0 -- the reference accepts |diff| < 10 for its float32 gradient M-step; here 1e-4 relative as
    for EM_solver (unit_tests_EMSolver.py:556-566)"""
    case, z = load_npz_case("ref_test4_zero")
    pri, cm, g = arrays_of(case)
    opt = fe.EMOptimizer(pri, cm)
    out = opt.optimize(g, np.maximum(case.d, 0.01), 0.5, 0.5, np.ones(case.forest.n_nodes, bool))
    assert abs(out["nllh"] - 1007.009158767873) < 1e-4 * 1007.009158767873
    assert abs(out["phi"]) < 1e-4 and abs(out["nu"]) < 1e-4
