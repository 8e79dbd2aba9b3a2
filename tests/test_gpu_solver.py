"""The reference-facing solver class on the GPU engine: the reference's own
solver tests (laml_unit_tests/unit_tests_EMSolver.py) re-stated against
``GPU_EM_solver``."""
import numpy as np
import pytest

from helpers import json_case, load_json, load_npz_case
from laml_b200.solver import DEFAULT_STRATEGY, GPU_EM_solver

pytestmark = pytest.mark.gpu

KAT = load_json("kat_likelihood.json")
OPT = load_json("stored_optima.json")


@pytest.mark.parametrize("rec", KAT[:6] + KAT[-3:], ids=lambda r: "test_%d" % r["no"])
def test_negative_llh(rec):
    case = json_case(rec)
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': case.phi, 'nu': case.nu})
    s.az_partition()
    assert abs(s.negative_llh() - rec["true_nllh"]) < 5e-6        # places=5 in the reference
    s.close()


def _test4():
    case, z = load_npz_case("ref_test4_zero")
    return case


def test_47_optimize():
    """unit_tests_EMSolver.py:535-566: EM from seed 1221 reaches nllh 1007.009158767873, phi = nu = 0"""
    case = _test4()
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0, 'nu': 0})
    nllh, status = s.optimize(initials=1, random_seeds=1221, verbose=-1, ultra_constr=False)
    ref = OPT["test_47"]
    assert status == "optimal"
    assert abs(ref["nllh"] - nllh) / ref["nllh"] < 1e-4
    assert abs(s.params.phi - ref["phi"]) < 1e-4 and abs(s.params.nu - ref["nu"]) < 1e-4
    s.close()


def test_48_score_tree():
    case = _test4()
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0, 'nu': 0})
    strategy = dict(DEFAULT_STRATEGY)
    strategy['fixed_brlen'] = None
    score, status = s.score_tree(strategy=strategy)
    assert abs(OPT["test_47"]["nllh"] + score) / OPT["test_47"]["nllh"] < 1e-4
    assert abs(s.params.phi) < 1e-4 and abs(s.params.nu) < 1e-4
    s.close()


def test_example1_published_optimum():
    """examples/out_example1/example1_params.txt: the published tree / phi / nu evaluate to the
    published nllh, and one more EM iteration from there does not move the parameters."""
    case, z = load_npz_case("ref_example1_optimum")
    ref = OPT["example1"]
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': ref["phi"], 'nu': ref["nu"]})
    nllh = s.negative_llh()
    assert abs(nllh - ref["nllh"]) / ref["nllh"] < 1e-9
    s.close()


def test_fixed_branches_and_fixed_params():
    """score_tree with fixed_brlen / fixed_phi / fixed_nu (Topology_search.py:256-300): marked branches
    and fixed parameters must not move"""
    case = _test4()
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0.05, 'nu': 0.1})
    leaves = [nd.label for nd in s.trees[0].traverse_leaves()]
    anchor = (leaves[0], leaves[1])
    strategy = dict(DEFAULT_STRATEGY)
    strategy.update(fixed_brlen=[{anchor: 0.123}], fixed_phi=0.05, fixed_nu=0.1)
    score, status = s.score_tree(strategy=strategy)
    assert status == "optimal" and score is not None
    assert s.params.phi == 0.05 and s.params.nu == 0.1
    s.close()


def test_example1_ultrametric_em_reaches_published_optimum():
    """BASELINE config 1 (examples/example1, the runner's default ultra_constr=True): EM from a random start on
    the shipped starting tree reaches the published optimum (examples/out_example1/example1_params.txt:
    nllh 6987.352449704049, phi 0.14378501771164973, nu 0.13359737967697907).  Tolerances are the spread of the
    reference's own five random starts (examples/example1/LAML_output.log:11-36: 4.5e-4 in nllh, ~1.5e-4 in
    phi, ~5e-4 in nu); the E-step runs on the GPU, the ultrametric M-step in laml_b200/ultrametric.py."""
    case, z = load_npz_case("ref_example1_start")
    ref = OPT["example1"]
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0.1, 'nu': 0.1})
    for tree in s.trees:                                   # the shipped starting tree carries no lengths
        for nd in tree.traverse_postorder():
            nd.edge_length = None
    nllh, status = s.optimize(initials=1, random_seeds=1984, verbose=-1, ultra_constr=True)
    assert status == "optimal"
    assert abs(nllh - ref["nllh"]) < 1e-3
    assert abs(s.params.phi - ref["phi"]) < 3e-4 and abs(s.params.nu - ref["nu"]) < 1e-3
    tree = s.trees[0]
    dist = {}
    for nd in tree.traverse_preorder():
        dist[nd] = (dist[nd.parent] if nd.parent is not None else 0.0) + (nd.edge_length or 0.0)
    tips = np.array([dist[nd] for nd in tree.traverse_leaves()])
    assert np.ptp(tips) < 1e-6
    s.close()


def test_ultrametric_em():
    """ultra_constr=True (the runner's default, runner.py:91,287): host program on GPU statistics;
    the result must be ultrametric and not worse than the start"""
    case, z = load_npz_case("ref_rand_ragged")
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0.1, 'nu': 0.1})
    start = s.negative_llh()
    nllh, status = s.optimize(initials=1, random_seeds=7, verbose=-1, ultra_constr=True)
    tree = s.trees[0]
    dist = {}
    for nd in tree.traverse_preorder():
        dist[nd] = (dist[nd.parent] if nd.parent is not None else 0.0) + (nd.edge_length or 0.0)
    tips = np.array([dist[nd] for nd in tree.traverse_leaves()])
    assert np.ptp(tips) < 1e-6 * tips.mean() + 1e-3      # collapsed near-zero branches may shift tips by dmin*0.01
    assert np.isfinite(nllh)
    s.close()
