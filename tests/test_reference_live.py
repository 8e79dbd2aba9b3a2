"""Live cross-checks against the reference itself, imported verbatim from /root/reference (tests/golden/ref_loader.py).
They run wherever the reference tree is present -- this container -- and skip themselves elsewhere (the GPU box has
no /root/reference).  CPU only; the engine is the oracle-backed stand-in (tests/oracle_engine.py).

* fresh random problems (not among the committed fixtures): reference EM_solver vs the oracle and vs the numpy
  models of both engines' arithmetic;
* the reference's OWN ``Topology_search`` class driving ``GPU_EM_solver`` as its solver plug-in -- the drop-in
  INTEGRATION.md describes -- on the reference's own search tests;
* ``GPU_Topology_search`` against that run: same optimum."""
import os
import sys
from copy import deepcopy
from random import seed

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

import laml_b200.solver as solver_mod  # noqa: E402
from helpers import Case  # noqa: E402
from laml_b200.solver import DEFAULT_STRATEGY, GPU_EM_solver  # noqa: E402
from laml_b200.topology import GPU_Topology_search  # noqa: E402
from oracle_engine import OracleEngine  # noqa: E402
from test_topology_host import MSA_A, MSA_B, Q5, _brute_force  # noqa: E402


@pytest.fixture(autouse=True)
def oracle_engine(monkeypatch):
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


@pytest.mark.parametrize("sd", [101, 102, 103, 104])
def test_fresh_random_problems(R, sd):
    import make_golden as G
    kw = dict(seed=sd, n_leaves=10 + 7 * (sd % 5), K=6 + sd % 7, max_A=4 + sd % 9, phi=0.05 * (sd % 4), nu=0.1 * (sd % 3),
              n_trees=1 + (sd % 2))
    trees, msa, Q = G.random_case(**kw)
    ref = G.run_reference(trees, msa, Q, kw["phi"], kw["nu"])
    case = Case(ref["newick_out"], msa, Q, kw["phi"], kw["nu"])
    ids = case.ids(ref["labels"], ref["tree_of"])
    out = case.oracle().estep(case.d, case.nu, case.phi, dump=True)
    assert abs(-out["llh"] - ref["nllh"]) <= 1e-12 * abs(ref["nllh"])
    for k in range(5):
        np.testing.assert_allclose(out["S%d" % k][ids], ref["S%d" % k], rtol=1e-10, atol=1e-12)
    from fullstate_model import estep as full_estep
    from lumped_model import estep as lumped_estep
    Ssum = np.stack([ref["S%d" % k].sum(axis=1) for k in range(5)])
    for model in (lumped_estep, full_estep):
        m = model(case.forest, case.leaf_row, case.enc.codes, case.enc.qptr, case.enc.qval, case.d, case.nu, case.phi)
        assert abs(-m["llh"] - ref["nllh"]) <= 1e-10 * abs(ref["nllh"])
        np.testing.assert_allclose(m["S"][ids].T, Ssum, rtol=1e-8, atol=1e-10 * case.enc.n_sites)
    # and through the solver class
    s = GPU_EM_solver(trees, {'charMtrx': msa}, {'Q': Q}, {'phi': kw["phi"], 'nu': kw["nu"]})
    assert abs(s.negative_llh() - ref["nllh"]) <= 1e-12 * abs(ref["nllh"])


@pytest.mark.parametrize("no,msa,T0,resolve_only,stored", [
    (1, MSA_A, '((a,c),(b,d));', False, None), (3, MSA_A, '(a,b,c,d);', False, None), (4, MSA_A, '((a,b),c,d);', True, None),
    (6, MSA_B, '((a,b),(c,d));', False, None),
    (8, MSA_B, '((c,d),a,b);', True, 10.08306053772235)])       # only the polytomy may move: the reference's stored value
def test_reference_topology_search_drives_the_solver_class(R, no, msa, T0, resolve_only, stored):
    """laml_libs/Topology_search.py, unmodified, with solver=GPU_EM_solver (unit_tests_TopoSearch.py tests 1, 3, 4,
    6, 8): every protocol call it makes -- constructor from newick strings, score_tree(strategy) with fixed_brlen
    anchors / fixed_nu / fixed_phi, get_tree_newick, get_params, solver_name -- is served by the class"""
    from laml_libs.Topology_search import Topology_search
    Q = [dict(q) for q in Q5]
    nllh_bf = stored if stored is not None else _brute_force(msa, Q)
    seed(a=2000 + no)
    ts = Topology_search([T0], GPU_EM_solver, data={'charMtrx': msa}, prior={'Q': Q}, params={'phi': 0, 'nu': 0})
    st = deepcopy(DEFAULT_STRATEGY)
    st['resolve_search_only'] = resolve_only
    best_tree, max_score, best_params = ts.search(maxiter=200, verbose=False, strategy=st, nreps=1)
    assert abs(nllh_bf + max_score) < 5e-5, (no, -max_score, nllh_bf)                    # the reference's places=4
    # the batched search class ends on the same optimum
    seed(a=2000 + no)
    mine = GPU_Topology_search([T0], GPU_EM_solver, data={'charMtrx': msa}, prior={'Q': Q}, params={'phi': 0, 'nu': 0})
    _, my_score, my_params = mine.search(maxiter=200, verbose=False, strategy=st, nreps=1)
    assert abs(my_score - max_score) < 5e-5
    assert set(my_params) == set(best_params) == {'phi', 'nu'}


def test_reference_apply_nni_scores_equal_the_batch(R):
    """the reference's apply_nni (its own code: anchors, LCAs, new solver, score_tree) on every branch of a small
    tree, with an acceptance rule that rejects everything so that all candidates get scored; the scores it sees are
    the batch scorer's (clamp=True)"""
    from laml_libs.Topology_search import Topology_search
    from helpers import load_npz_case
    from laml_b200.topology import score_neighbourhood
    case, z = load_npz_case("ref_rand_small")
    params = {'phi': case.phi, 'nu': case.nu}
    seen = []

    class Recording(Topology_search):
        def _Topology_search__accept_proposal__(self, curr, new, t):      # not name-mangled in the reference, kept for safety
            seen.append(new); return False

        def __accept_proposal__(self, curr, new, t):
            seen.append(new); return False

    ts = Recording(case.newicks, GPU_EM_solver, data={'charMtrx': case.msa}, prior={'Q': case.Q}, params=params)
    seed(a=5)
    score, n_attempts, took = ts.single_nni(-1e18, 0, deepcopy(DEFAULT_STRATEGY))
    assert not took and n_attempts == len(seen) == 14
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, params)
    nb = score_neighbourhood(s, clamp=True)
    np.testing.assert_allclose(np.sort(seen), np.sort(nb.score), rtol=1e-9)


@pytest.mark.parametrize("sd", [201, 202, 203])
def test_annotation_writer_on_fresh_problems(R, sd, tmp_path):
    """the reference's own annotation block (runner.py:367-435, executed from the reference tree) against
    laml_b200.annotate on problems that are not among the committed files: byte-identical"""
    import make_golden as G
    from make_annotation_golden import reference_writer, strip_internal_labels
    from laml_b200.annotate import annotate
    kw = dict(seed=sd, n_leaves=8 + 5 * (sd % 4), K=5 + sd % 6, max_A=3 + sd % 8, phi=0.1 + 0.05 * (sd % 3), nu=0.15 * (sd % 3),
              n_trees=1 + (sd % 2))
    trees, msa, Q = G.random_case(**kw)
    if sd % 2:
        trees = [strip_internal_labels(t) for t in trees]
    params = {"phi": kw["phi"], "nu": kw["nu"]}
    ref_out, my_out = tmp_path / "ref.txt", tmp_path / "mine.txt"
    reference_writer(R)(list(trees), msa, Q, params, str(ref_out))
    annotate(list(trees), msa, Q, params, str(my_out))
    assert my_out.read_text() == ref_out.read_text()


# ---- the reference's own M-step / EM code, verbatim, with only the numerical solver behind cvxpy substituted -----------------
@pytest.fixture()
def R_with_solver(R):
    ref_loader.enable_cvxpy_shim()           # tests/golden/cvxpy_shim.py: the cvxpy calls of EM_solver.py:349-407 on scipy
    return R


def _ref_solver(R, case, phi, nu):
    s = R.EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': phi, 'nu': nu})
    for tree in s.trees:
        for nd in tree.traverse_postorder():
            nd.mark_fixed = False             # what optimize() does before every start (ML_solver.py:296-299)
    return s


@pytest.mark.parametrize("name,phi,nu", [("ref_rand_small", 0.17, 0.23), ("ref_rand_ragged", 0.17, 0.23), ("ref_rand_forest", 0.12, 0.4),
                                         ("ref_rand_nu0", 0.2, 0.0)])
def test_reference_mstep_unconstrained(R_with_solver, name, phi, nu):
    """EM_solver.Estep + EM_solver.Mstep of the reference against the M-step the engine is tested against
    (oracle: per-branch bisection; the GPU kernels: per-branch Newton) from the same statistics"""
    from helpers import load_npz_case
    case, z = load_npz_case(name)
    s = _ref_solver(R_with_solver, case, phi, nu)
    s.az_partition(); s.Estep()
    ok, status = s.Mstep(verbose=-1)
    assert ok and status == "optimal"
    orc = case.oracle()
    est = orc.estep(case.d, nu, phi)
    d1, nu1, phi1, _ = orc.mstep(est["S"], est["Rtilde"], case.d, nu)
    # statistics as Mstep prepares them (EM_solver.py:341-346), the objective (:351-357) and its curvature
    Sn = np.maximum(est["S"].T, 1e-6)
    Sn = Sn / Sn.sum(axis=1, keepdims=True) * case.enc.n_sites
    use_nu = nu > 1e-5

    def objective(d):
        f = -(nu + 1) * Sn[:, 0] * d - nu * (Sn[:, 1] + Sn[:, 3]) * d + Sn[:, 1] * np.log(-np.expm1(-d))
        return f + ((Sn[:, 2] + Sn[:, 4]) * np.log(-np.expm1(-nu * d)) if use_nu else 0.0)

    by_key = {(str(nd.label), ti): i for i, (nd, ti) in enumerate(_with_tree_index(case))}
    d_ref = np.empty_like(d1)
    for ti, tree in enumerate(s.trees):
        for nd in tree.traverse_postorder():
            d_ref[by_key[(str(nd.label), ti)]] = nd.edge_length
    f_mine, f_ref = objective(d1), objective(d_ref)
    assert f_mine.sum() >= f_ref.sum() - 1e-9 * abs(f_ref.sum())           # at least as optimal as the reference's solve
    assert (f_mine >= f_ref - 1e-9).all()                                  # ... branch by branch (the program is separable)
    em = np.expm1(d1)
    curv = Sn[:, 1] * (em + 1) / em ** 2 + (nu * nu * (Sn[:, 2] + Sn[:, 4]) * (np.expm1(nu * d1) + 1) / np.expm1(nu * d1) ** 2 if use_nu else 0.0)
    sharp = curv > 1e-2                                                    # where the optimum is well determined
    assert sharp.sum() >= len(d1) // 3
    np.testing.assert_allclose(d1[sharp], d_ref[sharp], rtol=2e-6, atol=2e-6)
    assert abs(s.params.nu - nu1) < 2e-6 and abs(s.params.phi - phi1) < 1e-12


def _with_tree_index(case):
    where = {}
    for ti, t in enumerate(case.trees):
        for nd in t.traverse_postorder():
            where[id(nd)] = ti
    return [(nd, where[id(nd)]) for nd in case.forest.nodes]


@pytest.mark.parametrize("name", ["ref_rand_small", "ref_rand_forest"])
def test_reference_mstep_ultrametric(R_with_solver, name):
    """the reference's ultrametric program -- its own dense rows ``M d = b`` (EM_solver.py:53-89) in its own Mstep --
    against pmm_ultrametric_solve on the same statistics"""
    from helpers import load_npz_case
    from laml_b200 import ultrametric as um
    case, z = load_npz_case(name)
    phi, nu = 0.1, 0.2
    s = _ref_solver(R_with_solver, case, phi, nu)
    s.az_partition(); s.Estep()
    cache = s.ultrametric_constr(local_brlen_opt=True)
    ok, status = s.Mstep(verbose=-1, ultra_constr_cache=cache)
    assert ok and status == "optimal"
    est = case.oracle().estep(case.d, nu, phi)
    S = np.maximum(est["S"].T, 1e-6)
    S = S / S.sum(axis=1, keepdims=True) * case.enc.n_sites
    n = case.forest.n_nodes
    prob = um.UltrametricProblem(case.forest.parent, np.zeros(n, np.int8), np.zeros(n), 0.005, 10.0)
    d, H, info = um.solve(prob, S, nu)
    assert info["status"] == "optimal"
    mine = {(str(nd.label), ti): d[i] for i, (nd, ti) in enumerate(_with_tree_index(case))}
    worst = max(abs(mine[(str(nd.label), ti)] - nd.edge_length) for ti, tree in enumerate(s.trees) for nd in tree.traverse_postorder())
    assert worst < 5e-6, worst


@pytest.mark.parametrize("name,ultra,sd", [("ref_rand_small", False, 31), ("ref_rand_small", True, 32), ("ref_rand_nu0", False, 33)])
def test_reference_optimize_end_to_end(R_with_solver, name, ultra, sd):
    """EM_solver.optimize of the reference (initial point from the seed, EM loop, convergence test, multi-start
    bookkeeping, newick write-back) against GPU_EM_solver.optimize with the same seed: same optimum"""
    from helpers import load_npz_case
    case, z = load_npz_case(name)
    ref = R_with_solver.EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0.1, 'nu': 0.1})
    r_nllh, r_status = ref.optimize(initials=1, random_seeds=sd, verbose=-1, ultra_constr=ultra)
    mine = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0.1, 'nu': 0.1})
    m_nllh, m_status = mine.optimize(initials=1, random_seeds=sd, verbose=-1, ultra_constr=ultra)
    assert r_status == m_status == "optimal"
    assert abs(r_nllh - m_nllh) <= 1e-6 * abs(r_nllh), (r_nllh, m_nllh)
    assert abs(ref.params.phi - mine.params.phi) < 1e-5 and abs(ref.params.nu - mine.params.nu) < 1e-4
    # the trees that come back evaluate to the same likelihood under the reference's own likelihood code
    back = R_with_solver.EM_solver(mine.get_tree_newick(), {'charMtrx': case.msa}, {'Q': case.Q}, mine.get_params())
    assert abs(back.negative_llh() - m_nllh) <= 1e-9 * abs(m_nllh)


def _run_reference_unittests(R, module_name, names=None, rebind="EM_solver"):
    """load a unittest file of the reference, rebind its ``EM_solver`` (or ``rebind``) to the drop-in class, run it"""
    import importlib
    import types
    import unittest
    if "laml_unit_tests" not in sys.modules:
        pkg = types.ModuleType("laml_unit_tests")
        pkg.__path__ = [os.path.join(R.root, "laml_unit_tests")]
        sys.modules["laml_unit_tests"] = pkg
    mod = importlib.import_module("laml_unit_tests." + module_name)
    original = getattr(mod, rebind)
    setattr(mod, rebind, GPU_EM_solver)
    cls = next(c for c in vars(mod).values() if isinstance(c, type) and issubclass(c, unittest.TestCase))
    suite = (unittest.defaultTestLoader.loadTestsFromTestCase(cls) if names is None
             else unittest.TestSuite([cls(n) for n in names]))
    cwd = os.getcwd()
    os.chdir(os.path.join(R.root, "laml_unit_tests"))                 # the reference's tests open 'test_data/...' relatively
    try:
        with open(os.devnull, "w") as sink:
            res = unittest.TextTestRunner(stream=sink, verbosity=0).run(suite)
    finally:
        os.chdir(cwd)
        setattr(mod, rebind, original)
    return res


def test_reference_unittest_file_toposearch(R_with_solver):
    """laml_unit_tests/unit_tests_TopoSearch.py, all 10 tests, unmodified, with ``EM_solver`` bound to GPU_EM_solver:
    the reference's own Topology_search drives the class and the reference's own assertions judge the result
    (brute force through the reference's ML_solver, ultrametric tests 9-10 included)"""
    seed(a=4)
    res = _run_reference_unittests(R_with_solver, "unit_tests_TopoSearch")
    assert res.testsRun == 10 and not res.failures and not res.errors, (res.failures + res.errors)[:1]


def test_reference_unittest_file_emsolver(R_with_solver):
    """laml_unit_tests/unit_tests_EMSolver.py: the tests that use the solver PROTOCOL (1-20, 46: likelihood known
    answers; 47: optimize; 48: score_tree) against GPU_EM_solver.  Tests 21-45 call EM_solver internals
    (Estep_in_llh / Estep_out_llh and per-node log arrays of the compressed representation); their inputs are
    covered through the dumps in tests/test_oracle_golden.py and tests/test_gpu_parity.py"""
    names = ["test_%d" % i for i in list(range(1, 21)) + [46, 47, 48]]
    res = _run_reference_unittests(R_with_solver, "unit_tests_EMSolver", names)
    assert res.testsRun == 23 and not res.failures and not res.errors, (res.failures + res.errors)[:1]


def test_reference_unittest_file_mlsolver(R_with_solver):
    """laml_unit_tests/unit_tests_MLSolver.py (21 likelihood known answers written for ML_solver) with ``ML_solver``
    bound to GPU_EM_solver: on binary trees the two reference solvers define the same likelihood"""
    res = _run_reference_unittests(R_with_solver, "unit_tests_MLSolver", rebind="ML_solver")
    assert res.testsRun == 21 and not res.failures and not res.errors, (res.failures + res.errors)[:1]


def test_find_lcas_matches_the_reference(R):
    """laml_libs/lca_lib.find_LCAs (Euler tour + segment tree) against solver.find_LCAs on random trees and anchors"""
    from laml_libs.lca_lib import find_LCAs as ref_find
    from laml_b200.solver import find_LCAs
    from laml_b200.synth import make_workload, reference_inputs
    from laml_b200.tree import read_tree_newick
    rng = np.random.default_rng(3)
    for sd in range(4):
        w = make_workload(int(rng.integers(5, 60)), 2, 1, 3, 0.1, 0.1, seed=60 + sd, caterpillar_bias=0.4 * (sd % 2))
        newicks, _, _ = reference_inputs(w)
        t1, t2 = read_tree_newick(newicks[0]), read_tree_newick(newicks[0])
        leaves = [nd.label for nd in t1.traverse_leaves()]
        queries = [tuple(rng.choice(leaves, size=2)) for _ in range(40)] + [(l, l) for l in leaves[:5]]
        a = ref_find(t1, queries)
        b = find_LCAs(t2, queries)
        def leafset(nd):
            out, stack = set(), [nd]
            while stack:
                x = stack.pop()
                if x.children:
                    stack.extend(x.children)
                else:
                    out.add(x.label)
            return frozenset(out)
        assert [leafset(x) for x in a] == [leafset(x) for x in b]
