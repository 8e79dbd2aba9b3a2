"""``GPU_EM_solver``'s HOST logic on CPU: the reference's solver tests
(laml_unit_tests/unit_tests_EMSolver.py) through the class with the engine swapped for the
oracle-backed stand-in (tests/oracle_engine.py -- test infrastructure; the product has no CPU
path).  The same tests run against the real engine in tests/test_gpu_solver.py."""
import numpy as np
import pytest

import laml_b200.solver as solver_mod
from helpers import json_case, load_json, load_npz_case
from laml_b200.solver import DEFAULT_STRATEGY, GPU_EM_solver
from oracle_engine import OracleEngine

KAT = load_json("kat_likelihood.json")
OPT = load_json("stored_optima.json")


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    OracleEngine.instances = 0


def _solver(case, phi=None, nu=None):
    return GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q},
                         {'phi': case.phi if phi is None else phi, 'nu': case.nu if nu is None else nu})


def _test4():
    case, z = load_npz_case("ref_test4_zero")
    return case


def _tip_depths(tree):
    dist = {}
    for nd in tree.traverse_preorder():
        dist[nd] = (dist[nd.parent] if nd.parent is not None else 0.0) + (nd.edge_length or 0.0)
    return np.array([dist[nd] for nd in tree.traverse_leaves()])


@pytest.mark.parametrize("rec", KAT, ids=lambda r: "test_%d" % r["no"])
def test_negative_llh_protocol(rec):
    """unit_tests_EMSolver.py tests 1-20, 46: constructor -> az_partition -> negative_llh"""
    s = _solver(json_case(rec))
    s.az_partition()
    assert abs(s.negative_llh() - rec["true_nllh"]) < 5e-6        # places=5 in the reference
    assert s.get_params() == {'phi': rec["phi"], 'nu': rec["nu"]}
    s.close()


def test_47_optimize():
    """unit_tests_EMSolver.py:535-566"""
    s = _solver(_test4(), 0, 0)
    nllh, status = s.optimize(initials=1, random_seeds=1221, verbose=-1, ultra_constr=False)
    ref = OPT["test_47"]
    assert status == "optimal"
    assert abs(ref["nllh"] - nllh) / ref["nllh"] < 1e-4
    assert abs(s.params.phi - ref["phi"]) < 1e-4 and abs(s.params.nu - ref["nu"]) < 1e-4
    # the optimised lengths are on the trees and come back through the newick (ML_solver.py:316-318)
    assert abs(GPU_EM_solver(s.get_tree_newick(), {'charMtrx': s.charMtrx}, {'Q': s.Q},
                             s.get_params()).negative_llh() - nllh) < 1e-6 * nllh


def test_48_score_tree():
    """unit_tests_EMSolver.py:568-602"""
    s = _solver(_test4(), 0, 0)
    strategy = dict(DEFAULT_STRATEGY)
    strategy['fixed_brlen'] = None
    score, status = s.score_tree(strategy=strategy)
    assert status == "optimal"
    assert abs(OPT["test_47"]["nllh"] + score) / OPT["test_47"]["nllh"] < 1e-4


def test_multi_start_keeps_the_best_and_is_seeded():
    """ML_solver.py:256-336: several starts, the best one is kept; the same seeds give the same result"""
    case, z = load_npz_case("ref_rand_small")
    a = _solver(case)
    n1, st1 = a.optimize(initials=3, random_seeds=[5, 6, 7], verbose=-1)
    # later starts begin from the lengths the previous start left on the trees (ini_brlens keeps existing
    # lengths, ML_solver.py:195-203), so only the first start equals a single-start run; the best is kept
    b = _solver(case)
    first = b.optimize(initials=1, random_seeds=5, verbose=-1)[0]
    assert st1 == "optimal" and n1 <= first
    c = _solver(case)
    n2, _ = c.optimize(initials=3, random_seeds=[5, 6, 7], verbose=-1)
    assert n1 == n2 and a.get_tree_newick() == c.get_tree_newick()


def test_fixed_branches_and_fixed_params():
    """score_tree with fixed_brlen / fixed_phi / fixed_nu (Topology_search.py:256-300, ML_solver.py:296-305):
    marked branches and fixed parameters do not move, the others do"""
    case = _test4()
    s = _solver(case, 0.05, 0.1)
    leaves = [nd.label for nd in s.trees[0].traverse_leaves()]
    anchor = (leaves[0], leaves[0])                    # a leaf's own branch
    strategy = dict(DEFAULT_STRATEGY)
    strategy.update(fixed_brlen=[{anchor: 0.123}], fixed_phi=0.05, fixed_nu=0.1)
    score, status = s.score_tree(strategy=strategy)
    assert status == "optimal" and score is not None
    assert s.params.phi == 0.05 and s.params.nu == 0.1
    lab = s.trees[0].label_to_node()
    assert lab[leaves[0]].edge_length == 0.123
    assert len({round(nd.edge_length, 9) for nd in s.trees[0].traverse_postorder()}) > 3


def test_polytomies_are_binarised_and_marked():
    """EM_solver.py:19-43: inserted nodes get length dmin*0.01, are never optimised; test6 of case_tests/old:
    EM_solver's likelihood on a polytomous tree differs from ML_solver's by that binarisation"""
    msa = {"a": [1, 0], "b": [1, 2], "c": [0, 2], "d": [1, '?']}
    Q = [{1: 1.0}, {2: 0.5, 3: 0.5}]
    s = GPU_EM_solver(["(a:0.3,b:0.2,c:0.4,d:0.1)r:0.1;"], {'charMtrx': msa}, {'Q': Q}, {'phi': 0.1, 'nu': 0.1})
    assert s.has_polytomy and s.num_polytomy_mark == 2
    marked = [nd for nd in s.trees[0].traverse_postorder() if nd.polytomy_mark]
    assert all(abs(nd.edge_length - 5e-5) < 1e-15 for nd in marked)
    nllh, status = s.optimize(initials=1, random_seeds=3, verbose=-1)
    assert status == "optimal" and np.isfinite(nllh)
    assert all(abs(nd.edge_length - 5e-5) < 1e-15 for nd in marked)


def test_topology_change_reuses_the_engine_handle():
    """a new topology over the same cells is a pmm_set_tree on the live handle"""
    case, z = load_npz_case("ref_rand_small")
    s = _solver(case)
    base = s.negative_llh()
    eng = s._engine
    t = s.trees[0]
    u = next(nd for nd in t.traverse_preorder() if not nd.is_leaf() and not nd.is_root())
    v = u.parent
    w = next(c for c in v.children if c is not u)
    c0 = u.children[0]
    u.remove_child(c0); v.remove_child(w); u.add_child(w); v.add_child(c0)
    s._forest = None                                   # what update_from_solver / a new newick would trigger
    moved = s.negative_llh()
    assert s._engine is eng and OracleEngine.instances == 1 and eng.set_tree_calls == 2
    assert np.isfinite(moved) and moved != base
    fresh = GPU_EM_solver(s.get_tree_newick(), {'charMtrx': case.msa}, {'Q': case.Q}, s.get_params())
    assert abs(fresh.negative_llh() - moved) < 1e-9 * abs(moved)


def test_ultrametric_em_host_path():
    """ultra_constr=True (runner.py:91,287): Estep statistics -> laml_b200/ultrametric.py -> nu, phi"""
    case, z = load_npz_case("ref_rand_ragged")
    s = _solver(case, 0.1, 0.1)
    nllh, status = s.optimize(initials=1, random_seeds=7, verbose=-1, ultra_constr=True)
    tips = _tip_depths(s.trees[0])
    assert status == "optimal" and np.isfinite(nllh)
    assert np.ptp(tips) < 1e-6 * tips.mean() + 1e-3
    assert 0 <= s.params.phi < 1 and s.params.nu >= 0
    # EM does not decrease the likelihood: the optimum is at least as good as its own start
    t = _solver(case, 0.1, 0.1)
    seed_start = t.optimize(initials=1, random_seeds=7, verbose=-1, ultra_constr=True)[0]
    assert seed_start == nllh


def test_example1_ultrametric_em_reaches_published_optimum():
    """BASELINE config 1 through the host path (same test as tests/test_gpu_solver.py, oracle statistics)"""
    case, z = load_npz_case("ref_example1_start")
    ref = OPT["example1"]
    s = _solver(case, 0.1, 0.1)
    for tree in s.trees:
        for nd in tree.traverse_postorder():
            nd.edge_length = None
    nllh, status = s.optimize(initials=1, random_seeds=1984, verbose=-1, ultra_constr=True)
    assert status == "optimal"
    assert abs(nllh - ref["nllh"]) < 1e-3
    assert abs(s.params.phi - ref["phi"]) < 3e-4 and abs(s.params.nu - ref["nu"]) < 1e-3
    assert np.ptp(_tip_depths(s.trees[0])) < 1e-6


def test_optimize_argument_errors_return_none():
    """ML_solver.py:262-281: too few seeds / a wrong seed type are reported by returning None, not by raising"""
    case, z = load_npz_case("ref_rand_small")
    s = _solver(case)
    assert s.optimize(initials=3, random_seeds=[1, 2], verbose=-1) is None
    assert s.optimize(initials=1, random_seeds="7", verbose=-1) is None
    out = s.optimize(initials=2, random_seeds=[1, 2, 3], verbose=-1)           # surplus seeds are ignored
    assert out is not None and out[1] == "optimal"


def test_uninformative_sites_are_dropped_to_match_the_prior(capsys):
    """ML_solver.py:47-67: when Q has fewer entries than the matrix has columns and exactly that many columns
    are uninformative (only 0 / -1 / '?'), those columns are dropped"""
    msa = {"a": [1, 0, 2], "b": [1, '?', 0], "c": [0, 0, 2]}
    Q = [{1: 1.0}, {2: 0.5, 3: 0.5}]
    s = GPU_EM_solver(["((a:0.2,b:0.3):0.1,c:0.4):0.1;"], {'charMtrx': msa}, {'Q': Q}, {'phi': 0.1, 'nu': 0.1})
    assert s.numsites == 2 and s.charMtrx["a"] == [1, 2]
    assert "uninformative" in capsys.readouterr().out
    full = GPU_EM_solver(["((a:0.2,b:0.3):0.1,c:0.4):0.1;"], {'charMtrx': {k: [v[0], v[2]] for k, v in msa.items()}},
                         {'Q': Q}, {'phi': 0.1, 'nu': 0.1})
    assert s.negative_llh() == full.negative_llh()


def test_missing_cell_and_unknown_state_raise():
    with pytest.raises(KeyError):
        GPU_EM_solver(["(a:1,b:1):1;"], {'charMtrx': {"a": [1]}}, {'Q': [{1: 1.0}]}, {'phi': 0, 'nu': 0}).negative_llh()
    with pytest.raises(KeyError):
        GPU_EM_solver(["(a:1,b:1):1;"], {'charMtrx': {"a": [1], "b": [5]}}, {'Q': [{1: 1.0}]}, {'phi': 0, 'nu': 0})


def test_estep_leaves_statistics_and_posteriors_on_the_nodes():
    """runner.py:391-435 reads node.S1.. and node.post0/post1 after Estep()"""
    case, z = load_npz_case("ref_rand_small")
    s = _solver(case)
    out = s.Estep(posteriors=True)
    ids = {str(nd.label): nd for nd in s._forest.nodes}
    lab = [str(x) for x in z["labels"]]
    for k in range(5):
        got = np.array([getattr(ids[l], "S%d" % k) for l in lab])
        np.testing.assert_allclose(got, z["Ssum"][k], rtol=1e-10, atol=1e-12)
    p0 = np.array([ids[l].post0 for l in lab])
    from helpers import assert_close_loginf
    assert_close_loginf(p0, z["post0"], 1e-10, 1e-12, "post0")
    assert abs(out["llh"] + float(z["nllh"])) <= 1e-12 * float(z["nllh"])


def test_mstep_accepts_what_ultrametric_constr_returns():
    """EM_solver.py:470,489: ``cache = self.ultrametric_constr()`` ... ``self.Mstep(ultra_constr_cache=cache)`` -- the
    dense ``(M, b)`` the protocol method returns must be accepted by ``Mstep`` (same M-step as with the height
    problem the EM loop uses internally)"""
    case, z = load_npz_case("ref_rand_small")
    out = []
    for dense in (True, False):
        s = _solver(case)
        s._ensure()
        s.Estep()
        cache = s.ultrametric_constr() if dense else s._ultrametric_problem()
        ok, status = s.Mstep(ultra_constr_cache=cache)
        assert ok and status == "optimal"
        out.append((s._d(), s.params.nu, s.params.phi))
        if dense:                                    # a second M-step with the same object reuses the built problem
            s.Estep()
            assert s.Mstep(ultra_constr_cache=cache)[0]
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=0, atol=1e-12)
    assert out[0][1:] == out[1][1:]
