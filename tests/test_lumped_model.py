"""The lumped-state engine's arithmetic (numpy model of laml_b200/csrc/pmm_lumped.cuh) against
the reference dumps.  CPU only; pins the closed forms the lumped kernels implement."""
import numpy as np
import pytest

from helpers import NPZ_CASES, load_npz_case
from lumped_model import estep


@pytest.mark.parametrize("name", NPZ_CASES)
def test_lumped_model_matches_reference(name):
    case, z = load_npz_case(name)
    out = estep(case.forest, case.leaf_row, case.enc.codes, case.enc.qptr, case.enc.qval, case.d, case.nu, case.phi)
    ids = case.ids(z["labels"], z["tree_of"])
    nllh = float(z["nllh"])
    assert abs(-out["llh"] - nllh) <= 1e-11 * abs(nllh)          # north_star: 1e-9 relative
    np.testing.assert_allclose(out["site_llh"], z["site_llh"], rtol=1e-11, atol=1e-11)
    K = case.enc.n_sites
    np.testing.assert_allclose(out["S"][ids].T, z["Ssum"], rtol=1e-8, atol=1e-10 * K)
    assert abs(out["Rtilde"] - float(z["Rtilde"])) <= 1e-9 * max(1.0, float(z["Rtilde"]))
