"""The reference's shipped case tests (``case_tests/old/test{1,2,3,6}``) through the real engine, both
engines: statistics against the reference's EM_solver dump at the stored point, and the number the reference's
own run printed (tests/test_case_test_goldens.py has the same checks for the oracle and the numpy models)."""
import numpy as np
import pytest

from helpers import env_override, load_json, load_npz_case, make_engine
from test_case_test_goldens import _em_binarised_to_polytomy

pytestmark = pytest.mark.gpu
CASES = load_json("case_tests.json")


@pytest.mark.parametrize("engine", ["lumped", "fullstate"])
@pytest.mark.parametrize("name", sorted(CASES, key=lambda s: (int(s[4:].rstrip('u')), s)))
def test_case_test_known_answers(name, engine):
    rec = CASES[name]
    case, z = load_npz_case(rec["fixture"])
    with env_override(PMM_ENGINE=engine):
        eng = make_engine(case)
    out = eng.estep(case.d, case.nu, case.phi)
    ids = case.ids(z["labels"], z["tree_of"])
    K = case.enc.n_sites
    assert abs(-out["llh"] - rec["em_solver_nllh"]) <= 1e-9 * rec["em_solver_nllh"]
    np.testing.assert_allclose(out["S"][ids].T, z["Ssum"], rtol=1e-8, atol=1e-10 * K)
    assert abs(out["Rtilde"] - float(z["Rtilde"])) <= 1e-8 * max(1.0, float(z["Rtilde"]))
    ml = -eng.inside(_em_binarised_to_polytomy(case), case.nu, case.phi)
    if rec["ml_solver_nllh"] is not None:
        assert abs(ml - rec["ml_solver_nllh"]) <= 1e-9 * rec["ml_solver_nllh"]
    stored = rec["stored_nllh"]
    exact = name in ("test5", "test6", "test8u", "test9", "test10")
    assert min(abs(ml - stored), abs(-out["llh"] - stored)) <= (1e-9 if exact else 1e-8) * stored
    eng.close()


@pytest.mark.parametrize("engine", ["lumped", "fullstate"])
def test_published_example2_likelihood(engine):
    """examples/out_example2/example2_params.txt: 1124.4446364799137 at the published trees / phi / nu"""
    case, z = load_npz_case("ref_example2_published")
    with env_override(PMM_ENGINE=engine):
        eng = make_engine(case)
    out = eng.estep(case.d, case.nu, case.phi)
    assert abs(-out["llh"] - 1124.4446364799137) <= 1e-9 * 1124.4446364799137
    ids = case.ids(z["labels"], z["tree_of"])
    np.testing.assert_allclose(out["S"][ids].T, z["Ssum"], rtol=1e-8, atol=1e-10 * case.enc.n_sites)
    eng.close()


@pytest.mark.parametrize("engine", ["lumped", "fullstate"])
@pytest.mark.parametrize("name", ["ref_example2_api", "ref_example2_fastem", "ref_example2_fastem_nostart"])
def test_other_published_example2_likelihoods(name, engine):
    """the likelihoods the other published example2 runs printed (tests/test_published_runs.py)"""
    runs = load_json("published_runs.json")
    case, z = load_npz_case(name)
    with env_override(PMM_ENGINE=engine):
        eng = make_engine(case)
    llh = eng.inside(case.d, case.nu, case.phi)
    assert abs(-llh - runs[name]["nllh"]) <= 1e-9 * runs[name]["nllh"]
    eng.close()


def test_published_ultrametric_optimum_is_a_fixed_point():
    """MOSEK's ultrametric optimum of the published run_laml_api.py run is a fixed point of the ultrametric EM:
    E-step on the GPU, pmm_ultrametric_solve on the host (same assertions as tests/test_published_runs.py)"""
    from laml_b200.solver import GPU_EM_solver
    case, z = load_npz_case("ref_example2_api")
    s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': case.phi, 'nu': case.nu})
    start = s.negative_llh()
    d0 = s._d().copy()
    nllh, iters, status = s.EM_optimization(verbose=-1, ultra_constr=True)
    assert status == "optimal" and iters <= 2 and abs(nllh - start) <= 1e-8 * start
    assert abs(s.params.phi - case.phi) < 1e-5 and abs(s.params.nu - case.nu) < 1e-5
    assert np.abs(s._d() - d0).max() < 1e-4
    s.close()
