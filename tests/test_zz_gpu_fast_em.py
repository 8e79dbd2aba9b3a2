"""``laml_b200.fast_em.EMOptimizer`` (the reference's array-level EM interface, ``fast_em/laml.py:196-276``)
through the real engine, against the oracle's EM from the same start -- the problem and tolerances of
tests/test_gpu_parity.py::test_mstep_and_em_vs_oracle, entered through the dense-array / networkx boundary."""
import numpy as np
import pytest

from helpers import load_npz_case
from laml_b200.fast_em import EMOptimizer
from test_fast_em_host import arrays_of

pytestmark = pytest.mark.gpu


def test_optimize_matches_oracle_em():
    case, z = load_npz_case("ref_test4_zero")
    pri, cm, g = arrays_of(case)
    opt = EMOptimizer(pri, cm)
    d0 = np.maximum(case.d, 0.01)
    n = case.forest.n_nodes
    got = opt.optimize(g, d0, 0.1, 0.1, np.ones(n, bool))
    ref = case.oracle().em(d0, 0.1, 0.1)
    assert got["em_iterations"] == ref["iters"]
    assert abs(got["nllh"] - ref["nllh"]) <= 1e-9 * ref["nllh"]
    np.testing.assert_allclose(got["branch_lengths"], ref["d"], rtol=1e-5, atol=1e-7)
    assert abs(got["nu"] - ref["nu"]) <= 1e-5 and abs(got["phi"] - ref["phi"]) <= 1e-5
    # masked branches and fixed parameters stay put
    mask = np.zeros(n, bool); mask[::3] = True
    out = opt.optimize(g, d0, 0.2, 0.15, mask, fixed_nu=True, fixed_phi=True)
    assert out["nu"] == 0.2 and out["phi"] == 0.15 and (out["branch_lengths"][~mask] == d0[~mask]).all()
    opt.close()


def test_reference_fastEM_test_1():
    """unit_tests_fastEMSolver.py:15-44 at the EMOptimizer level (see tests/test_fast_em_host.py)"""
    case, z = load_npz_case("ref_test4_zero")
    pri, cm, g = arrays_of(case)
    opt = EMOptimizer(pri, cm)
    out = opt.optimize(g, np.maximum(case.d, 0.01), 0.5, 0.5, np.ones(case.forest.n_nodes, bool))
    assert abs(out["nllh"] - 1007.009158767873) < 1e-4 * 1007.009158767873
    assert abs(out["phi"]) < 1e-4 and abs(out["nu"]) < 1e-4
    opt.close()
