"""Pin oracle/pmm_oracle.c (the C restatement of the reference's algorithm)
against the reference's own known answers and against dumps produced by the
reference itself (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from helpers import (NPZ_CASES, assert_close_loginf, json_case, load_json, load_npz_case, unjson)

KAT = load_json("kat_likelihood.json")
PROP = load_json("outside_cases.json")


@pytest.mark.parametrize("rec", KAT, ids=["test_%d" % r["no"] for r in KAT])
def test_known_answer_likelihood(rec):
    """laml_unit_tests/unit_tests_EMSolver.py tests 1-20, 46 (places=5 there)."""
    case = json_case(rec)
    out = case.oracle().inside(case.d, case.nu, case.phi)
    assert abs(-out["llh"] - rec["true_nllh"]) < 5e-6
    assert abs(-out["llh"] - rec["ref_nllh"]) <= 1e-12 * max(1.0, abs(rec["ref_nllh"]))


@pytest.mark.parametrize("rec", PROP, ids=["test_%d" % r["no"] for r in PROP])
def test_outside_posterior_small(rec):
    """inputs of unit tests 21-44; outputs dumped from the reference"""
    case = json_case(rec)
    out = case.oracle().estep(case.d, case.nu, case.phi, dump=True)
    ids = case.ids(rec["labels"])
    assert abs(-out["llh"] - rec["nllh"]) <= 1e-12 * max(1.0, abs(rec["nllh"]))
    for k in ["L0", "L1", "out0", "out1", "post0", "post1"]:
        assert_close_loginf(out[k][ids], unjson(rec[k]), 1e-12, 1e-12, k)
    for k in ["S0", "S1", "S2", "S3", "S4"]:
        np.testing.assert_allclose(out[k][ids], unjson(rec[k]), rtol=1e-11, atol=1e-13, err_msg=k)


@pytest.mark.parametrize("name", NPZ_CASES)
def test_reference_dumps(name):
    case, z = load_npz_case(name)
    orc = case.oracle()
    out = orc.estep(case.d, case.nu, case.phi, dump=True)
    ids = case.ids(z["labels"], z["tree_of"])
    nllh = float(z["nllh"])
    assert abs(-out["llh"] - nllh) <= 1e-12 * abs(nllh)
    np.testing.assert_allclose(out["site_llh"], z["site_llh"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["S"][:, ids], z["Ssum"], rtol=1e-10, atol=1e-11)
    assert abs(out["Rtilde"] - float(z["Rtilde"])) <= 1e-10 * max(1.0, float(z["Rtilde"]))
    assert orc.R() == float(z["R"])
    if "L0" in z.files:
        tag = orc.az_partition()[ids]
        ztag = z["tag"]
        # golden tags hold original state ids; the oracle's hold alphabet ranks
        assert ((tag == -1) == (ztag == -1)).all() and ((tag == -2) == (ztag == -2)).all()
        for k in ["L0", "L1", "out0", "out1", "post0", "post1"]:
            assert_close_loginf(out[k][ids], z[k], 1e-11, 1e-11, k)
        for k in ["S0", "S1", "S2", "S3", "S4"]:
            np.testing.assert_allclose(out[k][ids], z[k], rtol=1e-10, atol=1e-12, err_msg=k)


def test_threads_agree():
    case, z = load_npz_case("ref_rand_deep")
    orc = case.oracle()
    a = orc.estep(case.d, case.nu, case.phi, nthreads=1)
    b = orc.estep(case.d, case.nu, case.phi, nthreads=4)
    np.testing.assert_allclose(a["S"], b["S"], rtol=1e-12, atol=1e-12)
    assert abs(a["llh"] - b["llh"]) < 1e-9
