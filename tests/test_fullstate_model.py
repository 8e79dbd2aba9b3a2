"""The engine's arithmetic (numpy model of the CUDA kernels) against the oracle
and the reference dumps.  CPU only; pins the *formulas* the kernels implement."""
import numpy as np
import pytest

from fullstate_model import estep
from helpers import NPZ_CASES, load_npz_case

SMALL = [c for c in NPZ_CASES if c not in ("ref_example1_optimum", "ref_example1_start", "ref_test45_test1")]


@pytest.mark.parametrize("name", SMALL)
def test_model_matches_reference(name):
    case, z = load_npz_case(name)
    out = estep(case.forest, case.leaf_row, case.enc.codes, case.enc.qptr, case.enc.qval, case.d, case.nu, case.phi)
    ids = case.ids(z["labels"], z["tree_of"])
    nllh = float(z["nllh"])
    assert abs(-out["llh"] - nllh) <= 1e-11 * abs(nllh)          # north_star: 1e-9 relative
    np.testing.assert_allclose(out["site_llh"], z["site_llh"], rtol=1e-11, atol=1e-11)
    K = case.enc.n_sites
    np.testing.assert_allclose(out["S"][ids].T, z["Ssum"], rtol=1e-8, atol=1e-10 * K)
    assert abs(out["Rtilde"] - float(z["Rtilde"])) <= 1e-9 * max(1.0, float(z["Rtilde"]))
