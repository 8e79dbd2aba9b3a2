"""Parity AT THE SIZES THAT ARE BENCHMARKED, on the GPU, against the oracle (the reference's algorithm,
``EM_solver.py:91-300`` restated in C and pinned to the reference's dumps by tests/test_oracle_golden.py):

* BASELINE config 3 (10k cells x 500 sites): the whole E-step against the threaded oracle, and the first EM iterations;
* BASELINE config 4 (100k x 2k): see tests/test_gpu_parity.py::test_full_size_properties_config4 (oracle on a site
  sample of the 100k-cell tree, per-site likelihoods of the whole run) and the ``parity`` key of every bench line;
* BASELINE config 5 (5k x 300): batch NNI scores against the definition -- a full-tree EM per candidate on the
  explicitly re-wired tree, oracle E-step (``Topology_search.py:242-322``);
* the M-step and whole EM runs on fixtures with nu > 0, phi > 0, a forest, fixed branches;
* the NULL free-mask convention of the C ABI.
"""
import os

import numpy as np
import pytest

from helpers import load_npz_case, make_engine
from laml_b200.capi import Engine
from laml_b200.synth import CONFIGS, make_workload

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]


def host_threads():
    return max(1, len(os.sched_getaffinity(0)))


def oracle_of(w):
    from oracle.pmm_oracle import OracleProblem
    return OracleProblem(w.forest, np.arange(w.forest.n_leaves, dtype=np.int32), w.codes, w.qptr, w.qval)


def assert_estep_equal(got, ref, K):
    """the north-star tolerances: nllh 1e-9 relative, counts 1e-8 relative (floor 1e-10 K: S1 / S3 are differences in the
    reference, EM_solver.py:299-300), R~ 1e-8"""
    assert abs(got["llh"] - ref["llh"]) <= 1e-9 * abs(ref["llh"])
    np.testing.assert_allclose(got["S"], ref["S"].T, rtol=1e-8, atol=1e-10 * K)
    assert abs(got["Rtilde"] - ref["Rtilde"]) <= 1e-8 * max(1.0, ref["Rtilde"])
    if "site_llh" in got:
        np.testing.assert_allclose(got["site_llh"], ref["site_llh"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("engine", ["lumped", "fullstate"])
def test_config3_whole_estep_vs_oracle(engine, monkeypatch):
    monkeypatch.setenv("PMM_ENGINE", engine)
    w = make_workload(seed=3, **CONFIGS["config3_10k_x_500"])
    N, K = w.forest.n_leaves, w.codes.shape[1]
    ref = oracle_of(w).estep(w.d, w.nu, w.phi, nthreads=host_threads())
    eng = Engine(N, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    got = eng.estep(w.d, w.nu, w.phi, want_site_llh=True)
    eng.close()
    assert_estep_equal(got, ref, K)


def test_config3_first_em_iterations_vs_oracle():
    """EM_optimization (EM_solver.py:461-512) at config-3 size, stopped after 3 iterations: same iterates as the oracle's
    EM (its M-step is a bisection on the same concave programs), i.e. the M-step is right at this size too"""
    w = make_workload(seed=3, **CONFIGS["config3_10k_x_500"])
    N = w.forest.n_leaves
    orc = oracle_of(w)
    d0 = np.maximum(w.d, 0.01)
    ref = orc.em(d0, w.nu, w.phi, max_iter=3, nthreads=host_threads())
    eng = Engine(N, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    got = eng.em(d0, w.nu, w.phi, orc.R(), N, max_iter=3)
    eng.close()
    assert got["iters"] == ref["iters"]
    assert abs(got["nllh"] - ref["nllh"]) <= 1e-9 * abs(ref["nllh"])
    np.testing.assert_allclose(got["d"], ref["d"], rtol=1e-5, atol=1e-7)
    assert abs(got["nu"] - ref["nu"]) <= 1e-5 and abs(got["phi"] - ref["phi"]) <= 1e-5


def test_config5_batch_scores_vs_the_definition():
    """a sample of the 9 996 candidates of BASELINE config 5, each scored by the definition on the CPU: the re-wired
    tree, five free branches, oracle E-step, EM to convergence"""
    from laml_b200.flatten import forest_from_parent
    from laml_b200.nni import apply_nni, nni_candidates
    from oracle.pmm_oracle import OracleProblem
    w = make_workload(seed=1, **CONFIGS["config5_5k_x_300"])
    f, N = w.forest, w.forest.n_leaves
    d = np.maximum(w.d, 0.005)
    eng = Engine(N, w.qptr, w.qval, w.codes)
    eng.set_tree(f.parent, f.left, f.right)
    cu, ck = nni_candidates(f.parent, f.left, N)
    assert len(cu) == 2 * (N - 2)
    res = eng.score_nni_batch(cu, ck, d, w.nu, w.phi)
    eng.close()
    assert np.isfinite(res["nllh"]).all()
    rng = np.random.default_rng(5)
    depth_of_u = f.depth[cu]
    picks = list(rng.choice(len(cu), 4, replace=False)) + [int(np.argmin(depth_of_u)), int(np.argmax(depth_of_u))]
    for i in picks:
        p2, l2, r2 = apply_nni(f.parent, f.left, f.right, int(cu[i]), int(ck[i]))
        prob = OracleProblem(forest_from_parent(p2, l2, r2), np.arange(N, dtype=np.int32), w.codes, w.qptr, w.qval)
        fm = np.zeros(f.n_nodes, np.uint8)
        fm[res["nodes"][i]] = 1
        ref = prob.em(d, w.nu, w.phi, fm, False, False, nthreads=host_threads())
        assert abs(res["nllh"][i] - ref["nllh"]) <= 1e-9 * abs(ref["nllh"]), (i, res["nllh"][i], ref["nllh"])
        assert res["iters"][i] == ref["iters"]
        np.testing.assert_allclose(res["d5"][i], ref["d"][res["nodes"][i]], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name,nu,phi,fixed", [("ref_test4_zero", 0.1, 0.1, 0), ("ref_rand_ragged", 0.25, 0.2, 0),
                                               ("ref_rand_forest", 0.15, 0.3, 0), ("ref_rand_small", 0.3, 0.05, 5),
                                               ("ref_rand_deep", 0.05, 0.4, 9), ("ref_example1_start", 0.2, 0.2, 0)])
def test_mstep_and_em_vs_oracle_on_more_fixtures(name, nu, phi, fixed):
    """M-step (EM_solver.py:307-459 without the equalities) and whole EM runs against the oracle: nu > 0, phi > 0, a
    forest, fixed branches (mark_fixed, ML_solver.py:296-305)"""
    case, z = load_npz_case(name)
    orc = case.oracle()
    eng = make_engine(case)
    n = case.forest.n_nodes
    fm = None
    if fixed:
        fm = np.ones(n, np.uint8)
        fm[np.random.default_rng(fixed).choice(n, fixed, replace=False)] = 0
    d_start = np.maximum(case.d, 0.01)
    est = eng.estep(d_start, nu, phi)
    R = orc.R()
    d1, nu1, phi1 = eng.mstep(est["S"], est["Rtilde"], R, orc.n_cells, d_start, nu, phi, free_mask=fm)
    d0, nu0, phi0, _ = orc.mstep(est["S"].T.copy(), est["Rtilde"], d_start, nu, free_mask=fm)
    np.testing.assert_allclose(d1, d0, rtol=1e-9, atol=1e-12)
    assert abs(nu1 - nu0) <= 1e-9 and abs(phi1 - phi0) <= 1e-12
    if fm is not None:
        assert (d1[fm == 0] == d_start[fm == 0]).all()
    got = eng.em(d_start, nu, phi, R, orc.n_cells, free_mask=fm)
    ref = orc.em(d_start, nu, phi, free_mask=fm)
    assert got["iters"] == ref["iters"]
    assert abs(got["nllh"] - ref["nllh"]) <= 1e-9 * abs(ref["nllh"])
    np.testing.assert_allclose(got["d"], ref["d"], rtol=1e-5, atol=1e-7)
    assert abs(got["nu"] - ref["nu"]) <= 1e-5 and abs(got["phi"] - ref["phi"]) <= 1e-5
    eng.close()


def test_null_free_mask_means_every_branch_is_free():
    """include/pmm_b200.h: a NULL mask never means "the mask of an earlier call" -- not after a call with a mask, and
    not after pmm_set_tree on the live handle (node ids mean something else under the new topology)"""
    from laml_b200.synth import lognormal_branch_lengths, random_tree
    w = make_workload(120, 20, 2, 25, 0.2, 0.2, seed=61)
    n = w.forest.n_nodes
    R = float((w.codes != -1).sum())
    d0 = np.maximum(w.d, 0.01)
    eng = Engine(120, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    est = eng.estep(d0, w.nu, w.phi)
    all_free = eng.mstep(est["S"], est["Rtilde"], R, 120, d0, w.nu, w.phi)
    fm = np.ones(n, np.uint8); fm[::3] = 0
    masked = eng.mstep(est["S"], est["Rtilde"], R, 120, d0, w.nu, w.phi, free_mask=fm)
    assert (masked[0][fm == 0] == d0[fm == 0]).all() and not np.array_equal(masked[0], all_free[0])
    again = eng.mstep(est["S"], est["Rtilde"], R, 120, d0, w.nu, w.phi)                  # NULL after a mask
    assert np.array_equal(again[0], all_free[0]) and again[1:] == all_free[1:]
    eng.upload_params(d0, w.nu, w.phi)
    eng.estep_dev()
    eng.mstep_dev(R, 120, fm)
    eng.upload_params(d0, w.nu, w.phi)
    eng.estep_dev()
    eng.mstep_dev(R, 120)                                                                  # device-resident entry point
    assert np.array_equal(eng.download_params()[0], all_free[0])
    em_masked = eng.em(d0, w.nu, w.phi, R, 120, free_mask=fm, max_iter=2)
    em_free = eng.em(d0, w.nu, w.phi, R, 120, max_iter=2)
    assert (em_masked["d"][fm == 0] == d0[fm == 0]).all() and (em_free["d"][fm == 0] != d0[fm == 0]).any()
    # a new topology on the live handle: the old mask is gone
    eng.mstep(est["S"], est["Rtilde"], R, 120, d0, w.nu, w.phi, free_mask=fm)
    f2 = random_tree(120, seed=62)
    d2 = np.maximum(lognormal_branch_lengths(f2, seed=162), 0.01)
    eng.set_tree(f2.parent, f2.left, f2.right)
    eng.upload_params(d2, w.nu, w.phi)
    eng.estep_dev()
    eng.mstep_dev(R, 120)
    got = eng.download_params()[0]
    fresh = Engine(120, w.qptr, w.qval, w.codes)
    fresh.set_tree(f2.parent, f2.left, f2.right)
    e2 = fresh.estep(d2, w.nu, w.phi)
    want = fresh.mstep(e2["S"], e2["Rtilde"], R, 120, d2, w.nu, w.phi)[0]
    assert np.array_equal(got, want)
    fresh.close(); eng.close()
