"""The reference's PUBLISHED runs as known answers.  example1 (``examples/out_example1``, LAML 0.0.2): the published
likelihood is pinned in tests/test_oracle_golden.py / tests/test_gpu_solver.py; here its published annotation file.
The reference's PUBLISHED example2 run (``examples/out_example2``: LAML 1.0.0 with MOSEK and treeswift, topology
search, de-duplication) as known answers (tests/golden/make_published_goldens.py): the published negative
log-likelihood 1124.4446364799137 at the published trees / phi / nu, and the published posterior annotation file,
byte for byte.  CPU: oracle, numpy models of both engines, the solver class and the annotation writer on the
oracle-backed stand-in; tests/test_zz_gpu_annotate.py and tests/test_zz_gpu_case_tests.py run the engine."""
import json
import os

import numpy as np
import pytest

import laml_b200.solver as solver_mod
from helpers import GOLDEN, load_npz_case
from oracle_engine import OracleEngine

PUBLISHED_NLLH = 1124.4446364799137


def annotation_case():
    rec = json.load(open(os.path.join(GOLDEN, "example2_annotation_case.json")))
    Q = [{int(k): v for k, v in q.items()} for q in rec["Q"]]
    return rec["trees"], rec["msa"], Q, rec["published"]


def test_published_likelihood_oracle_and_models():
    case, z = load_npz_case("ref_example2_published")
    out = case.oracle().estep(case.d, case.nu, case.phi)
    assert abs(-out["llh"] - PUBLISHED_NLLH) <= 1e-12 * PUBLISHED_NLLH
    from fullstate_model import estep as full_estep
    from lumped_model import estep as lumped_estep
    for model in (lumped_estep, full_estep):
        m = model(case.forest, case.leaf_row, case.enc.codes, case.enc.qptr, case.enc.qval, case.d, case.nu, case.phi)
        assert abs(-m["llh"] - PUBLISHED_NLLH) <= 1e-11 * PUBLISHED_NLLH
        ids = case.ids(z["labels"], z["tree_of"])
        np.testing.assert_allclose(m["S"][ids].T, z["Ssum"], rtol=1e-8, atol=1e-10 * case.enc.n_sites)


def test_published_annotation_file_byte_for_byte(tmp_path, monkeypatch):
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    from laml_b200.annotate import annotate
    trees, msa, Q, pub = annotation_case()
    out = tmp_path / "annot.txt"
    annotate(trees, msa, Q, {"phi": pub["phi"], "nu": pub["nu"]}, str(out))
    assert out.read_text() == open(os.path.join(GOLDEN, "annot_example2_published.txt")).read()


def test_published_optimum_is_a_fixed_point_of_the_em(monkeypatch):
    """EM from the published point (the reference's MOSEK optimum; the published tree is not ultrametric, so the
    run's final optimisation was the unconstrained one) stops at once: the likelihood moves by < 1e-8 relative, phi
    and nu by < 1e-5 -- the north-star tolerance for converged parameters -- and no branch by more than 5e-4"""
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    case, z = load_npz_case("ref_example2_published")
    s = solver_mod.GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': case.phi, 'nu': case.nu})
    start = s.negative_llh()
    s._ensure()
    d0 = s._d().copy()
    nllh, iters, status = s.EM_optimization(verbose=-1, ultra_constr=False)
    assert status == "optimal" and iters <= 2
    assert nllh <= start and abs(nllh - start) <= 1e-8 * start
    assert abs(s.params.phi - case.phi) < 1e-5 and abs(s.params.nu - case.nu) < 1e-5
    assert np.abs(s._d() - d0).max() < 5e-4


def test_published_example1_annotation_file_byte_for_byte(tmp_path, monkeypatch):
    """examples/out_example1/example1_annotations.txt (501 lines; 250 cells x 30 sites) from the published trees
    (time-scaled lengths x mutation rate = ref_example1_optimum.npz) and the published phi / nu"""
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    from laml_b200.annotate import annotate
    case, z = load_npz_case("ref_example1_optimum")
    trees = [t.replace("[&R] ", "") for t in case.newicks]      # the published input carried no rooting comment
    out = tmp_path / "annot.txt"
    annotate(trees, case.msa, case.Q, {"phi": case.phi, "nu": case.nu}, str(out))
    assert out.read_text() == open(os.path.join(GOLDEN, "annot_example1_published.txt")).read()


# ---- the other published example2 runs: run_laml_api.py (ultrametric, MOSEK) and the two fastEM-cpu runs -------------------
RUNS = json.load(open(os.path.join(GOLDEN, "published_runs.json")))


@pytest.mark.parametrize("name", sorted(RUNS))
def test_other_published_likelihoods(name):
    """the ``Negative-llh`` each run printed (examples/<run>/*_params.txt) at its published trees / phi / nu, on the
    de-duplicated data it refers to: reproduced to 1e-12 by the oracle and by the numpy model of the lumped engine"""
    case, z = load_npz_case(name)
    out = case.oracle().inside(case.d, case.nu, case.phi)
    assert abs(-out["llh"] - RUNS[name]["nllh"]) <= 1e-12 * RUNS[name]["nllh"]
    from lumped_model import estep
    m = estep(case.forest, case.leaf_row, case.enc.codes, case.enc.qptr, case.enc.qval, case.d, case.nu, case.phi)
    assert abs(-m["llh"] - RUNS[name]["nllh"]) <= 1e-11 * RUNS[name]["nllh"]


def _tip_spread(tree):
    dist = {}
    for nd in tree.traverse_preorder():
        dist[nd] = (dist[nd.parent] if nd.parent is not None else 0.0) + (nd.edge_length or 0.0)
    tips = np.array([dist[nd] for nd in tree.traverse_leaves()])
    return float(np.ptp(tips))


@pytest.mark.parametrize("name", ["ref_example2_api", "ref_example2_fastem"])
def test_published_ultrametric_optima_are_fixed_points(name, monkeypatch):
    """``run_laml_api.py`` ran the default, ultrametric mode with MOSEK: its published tree is ultrametric (tips level
    to 3e-8) and is a fixed point of the ultrametric EM here -- pmm_ultrametric_solve against a real MOSEK optimum:
    one iteration, likelihood unchanged to 1e-9 relative, phi and nu to 1e-5 (the north-star tolerance), branches
    to 1e-4.  Likewise the fastEM-cpu run's tree (its M-step is a float32 gradient method: 2e-5 on nu)."""
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    case, z = load_npz_case(name)
    s = solver_mod.GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': case.phi, 'nu': case.nu})
    assert _tip_spread(s.trees[0]) < 1e-7
    start = s.negative_llh()
    d0 = s._d().copy()
    nllh, iters, status = s.EM_optimization(verbose=-1, ultra_constr=True)
    assert status == "optimal" and iters <= 2
    assert abs(nllh - start) <= 1e-8 * start
    tol_nu = 1e-5 if name == "ref_example2_api" else 2e-5
    assert abs(s.params.phi - case.phi) < 1e-5 and abs(s.params.nu - case.nu) < tol_nu
    assert np.abs(s._d() - d0).max() < 1e-4 and _tip_spread(s.trees[0]) < 1e-9


@pytest.mark.parametrize("name", sorted(RUNS))
def test_other_published_annotation_files(name, tmp_path, monkeypatch):
    """the annotation files those runs wrote, from their published trees (with the re-inserted duplicates) and
    parameters"""
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    from annot_compare import assert_annotations_match
    from laml_b200.annotate import annotate
    _, msa, Q, _ = annotation_case()
    run = RUNS[name]
    out = tmp_path / "annot.txt"
    annotate(run["annotation_trees"], msa, Q, {"phi": run["phi"], "nu": run["nu"]}, str(out))
    ref = open(os.path.join(GOLDEN, run["annotation"])).read()
    assert_annotations_match(out.read_text(), ref, exact=False)
    assert out.read_text() == ref


def test_published_example1_optimum_under_the_ultrametric_em(monkeypatch):
    """example1's published run (LAML 0.0.2, ultrametric, MOSEK; nllh 6987.352449704049): its tree is ultrametric to
    5e-7; the ultrametric EM here from that point finds a slightly better one within the reference's own convergence
    slack (2e-4 in nllh after two iterations) and moves phi by < 1e-5, nu by < 3e-5, no branch by more than 5e-4"""
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    case, z = load_npz_case("ref_example1_optimum")
    s = solver_mod.GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': case.phi, 'nu': case.nu})
    assert _tip_spread(s.trees[0]) < 1e-6
    start = s.negative_llh()
    assert abs(start - 6987.352449704049) <= 1e-12 * start
    d0 = s._d().copy()
    nllh, iters, status = s.EM_optimization(verbose=-1, ultra_constr=True)
    assert status == "optimal" and iters <= 3
    assert start - 5e-4 < nllh <= start
    assert abs(s.params.phi - case.phi) < 1e-5 and abs(s.params.nu - case.nu) < 3e-5
    assert np.abs(s._d() - d0).max() < 5e-4


def test_published_uninformative_sites_case(tmp_path, monkeypatch):
    """case_tests/test1 (LAML 1.0.0, "testcase with uninformative sites"): three columns, a prior for one of them; the
    two uninformative columns are dropped (ML_solver.py:47-67).  Published: nllh 3.3033497524754867, phi 0,
    nu 2.481710016212404e-06, mutation rate 1.1061115018134888, and the annotation file below."""
    monkeypatch.setattr(solver_mod, "Engine", OracleEngine)
    from laml_b200.annotate import annotate
    from laml_b200.tree import read_tree_newick
    msa = {'a': [0, 0, '?'], 'b': [0, 1, '?'], 'c': [0, 2, '?']}
    Q = [{1: 0.5, 2: 0.5}]
    t = read_tree_newick("(((a:0.9909593184497377,b:0.9909593184497377):0.004520341420182489,c:0.9954796598699202):0.004520340130079882);")
    for nd in t.traverse_preorder():
        if nd.edge_length is not None:
            nd.edge_length *= 1.1061115018134888
    params = {'phi': 0, 'nu': 2.481710016212404e-06}
    s = solver_mod.GPU_EM_solver([t.newick()], {'charMtrx': msa}, {'Q': Q}, params)
    assert s.numsites == 1
    assert abs(s.negative_llh() - 3.3033497524754867) <= 1e-14
    out = tmp_path / "annot.txt"
    annotate([t.newick()], msa, Q, params, str(out))
    assert out.read_text() == "(((a:0,b:1)I2:0,c:1)I1:0)I0;\nI0,0\nI1,0\nc,2\nI2,0\nb,1\na,0\n"
