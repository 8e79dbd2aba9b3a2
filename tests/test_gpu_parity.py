"""GPU parity: the CUDA engine, called through the C ABI, against
(1) the reference's own outputs (golden fixtures), (2) the reference's
known-answer tests, (3) the C oracle on seeded synthetic inputs.

Tolerances are BASELINE.json's: total nllh 1e-9 relative, per-branch expected
counts 1e-8 relative (absolute floor 1e-10*K: S1 and S3 are differences in the
reference, EM_solver.py:299-300)."""
import numpy as np
import pytest

from helpers import (NPZ_CASES, env_override, json_case, load_json, load_npz_case, make_engine)

pytestmark = pytest.mark.gpu



@pytest.fixture(autouse=True, params=["lumped", "fullstate"])
def engine_mode(request, monkeypatch):
    """every parity test runs on both E-step engines: the lumped-state records (default) and the
    full-state rows (PMM_ENGINE=fullstate; also what the NNI batch scorer caches)"""
    monkeypatch.setenv("PMM_ENGINE", request.param)
    return request.param


KAT = load_json("kat_likelihood.json")
PROP = load_json("outside_cases.json")


def check_against_golden(case, z, eng):
    out = eng.estep(case.d, case.nu, case.phi, want_site_llh=True)
    nllh = float(z["nllh"])
    K = case.enc.n_sites
    assert abs(-out["llh"] - nllh) <= 1e-9 * abs(nllh)
    np.testing.assert_allclose(out["site_llh"], z["site_llh"], rtol=1e-9, atol=1e-10)
    ids = case.ids(z["labels"], z["tree_of"])
    np.testing.assert_allclose(out["S"][ids].T, z["Ssum"], rtol=1e-8, atol=1e-10 * K)
    assert abs(out["Rtilde"] - float(z["Rtilde"])) <= 1e-8 * max(1.0, float(z["Rtilde"]))
    llh_only = eng.inside(case.d, case.nu, case.phi)
    assert llh_only == out["llh"]          # same kernels, same order: bit-identical
    return out


@pytest.mark.parametrize("name", NPZ_CASES)
def test_reference_dumps(name):
    case, z = load_npz_case(name)
    eng = make_engine(case)
    check_against_golden(case, z, eng)
    eng.close()


@pytest.mark.parametrize("rec", KAT, ids=["test_%d" % r["no"] for r in KAT])
def test_known_answer_likelihood(rec):
    case = json_case(rec)
    eng = make_engine(case)
    llh = eng.inside(case.d, case.nu, case.phi)
    assert abs(-llh - rec["true_nllh"]) < 5e-6                      # the reference's own assertion (places=5)
    assert abs(-llh - rec["ref_nllh"]) <= 1e-9 * max(1.0, abs(rec["ref_nllh"]))
    eng.close()


@pytest.mark.parametrize("rec", PROP, ids=["test_%d" % r["no"] for r in PROP])
def test_outside_cases(rec):
    """unit tests 21-44: per-branch counts and posteriors against the reference dump"""
    from helpers import unjson, assert_close_loginf
    case = json_case(rec)
    eng = make_engine(case)
    out = eng.estep(case.d, case.nu, case.phi)
    ids = case.ids(rec["labels"])
    assert abs(-out["llh"] - rec["nllh"]) <= 1e-9 * max(1.0, abs(rec["nllh"]))
    for k in range(5):
        np.testing.assert_allclose(out["S"][ids, k], unjson(rec["S%d" % k]).sum(axis=1), rtol=1e-8, atol=1e-10)
    p0, p1 = eng.posteriors(case.d, case.nu, case.phi)
    assert_close_loginf(p0[ids], unjson(rec["post0"]), 1e-8, 1e-9, "post0")
    assert_close_loginf(p1[ids], unjson(rec["post1"]), 1e-8, 1e-9, "post1")
    eng.close()


@pytest.mark.parametrize("ts,group,chunks", [(64, 4, 1), (128, 8, 1), (256, 16, 2), (512, 32, 3), (2048, 8, 1)])
def test_layout_variants(ts, group, chunks, engine_mode):
    """tile size, lane-group width and site chunking must not change results"""
    case, z = load_npz_case("ref_rand_ragged")
    with env_override(PMM_TILE_SLOTS=ts, PMM_GROUP=group, PMM_CHUNKS=chunks):
        eng = make_engine(case)
        info = eng.info()
        assert info["n_chunks"] == chunks
        check_against_golden(case, z, eng)
        eng.close()
    case, z = load_npz_case("ref_test45_test1")
    with env_override(PMM_TILE_SLOTS=ts, PMM_GROUP=group, PMM_CHUNKS=chunks):
        eng = make_engine(case)
        check_against_golden(case, z, eng)
        eng.close()


@pytest.mark.parametrize("clade,top,chunks,store_all", [(1, 1, 1, 0), (2, 2, 1, 1), (3, 16, 2, 0), (7, 3, 1, 0), (16, 16, 3, 1),
                                                      (100000, 100000, 1, 0)])
def test_lumped_schedule_variants(clade, top, chunks, store_all, engine_mode):
    """clade capacity of the bottom tier and of the top tiers (1 = one launch per tree level ... 127 steps
    per CTA), site chunking, storing the register-chained out-records and programmatic dependent launch
    on / off must not change the lumped engine's results"""
    if engine_mode != "lumped":
        pytest.skip("lumped engine only")
    for name in ("ref_rand_ragged", "ref_rand_forest", "ref_rand_deep", "ref_test45_test1"):
        case, z = load_npz_case(name)
        with env_override(PMM_CLADE=clade, PMM_CLADE_TOP=top, PMM_CHUNKS=chunks, PMM_STORE_ALL=store_all, PMM_PDL=store_all):
            eng = make_engine(case)
            info = eng.info()
            assert info["engine"] == 0 and info["n_chunks"] == chunks and info["clade_target"] == min(clade, 127)
            check_against_golden(case, z, eng)
            eng.close()


@pytest.mark.parametrize("cfg,seed", [("config2_1k_x_30", 1), ("config2_1k_x_30", 2)])
def test_synthetic_vs_oracle(cfg, seed):
    from laml_b200.capi import Engine
    from laml_b200.synth import CONFIGS, make_workload
    from oracle.pmm_oracle import OracleProblem
    w = make_workload(seed=seed, **CONFIGS[cfg])
    N = w.forest.n_leaves
    orc = OracleProblem(w.forest, np.arange(N, dtype=np.int32), w.codes, w.qptr, w.qval)
    ref = orc.estep(w.d, w.nu, w.phi, nthreads=4)
    eng = Engine(N, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    out = eng.estep(w.d, w.nu, w.phi, want_site_llh=True)
    K = w.codes.shape[1]
    assert abs(out["llh"] - ref["llh"]) <= 1e-9 * abs(ref["llh"])
    np.testing.assert_allclose(out["site_llh"], ref["site_llh"], rtol=1e-9)
    np.testing.assert_allclose(out["S"], ref["S"].T, rtol=1e-8, atol=1e-10 * K)
    assert abs(out["Rtilde"] - ref["Rtilde"]) <= 1e-8 * ref["Rtilde"]
    # size-independent property: the five counts and the silenced-stays-silenced mass
    # partition each site, so every branch's counts sum to at most K
    assert (out["S"].sum(axis=1) <= K * (1 + 1e-9)).all() and (out["S"] >= -1e-9).all()
    eng.close()


def test_site_slices_are_bit_identical(engine_mode):
    """site slices on auxiliary streams (pmm_set_option "slices") only reorder launches: same bits"""
    if engine_mode != "lumped":
        pytest.skip("lumped engine only")
    from laml_b200.capi import Engine, PMMError
    from laml_b200.synth import make_workload
    w = make_workload(400, 700, 2, 30, 0.2, 0.2, seed=71)          # 700 sites = 6 site blocks of 128
    eng = Engine(400, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    ref = None
    for slices in (1, 2, 3, 4):
        eng.set_option("slices", slices)
        out = eng.estep(w.d, w.nu, w.phi, want_site_llh=True)
        tm = eng.last_timing()
        assert (tm["outside_ms"] > 0) == (slices == 1)              # the split exists only without slices
        if ref is None:
            ref = out
        else:
            assert out["llh"] == ref["llh"] and (out["S"] == ref["S"]).all() and out["Rtilde"] == ref["Rtilde"]
            assert (out["site_llh"] == ref["site_llh"]).all()
    with pytest.raises(PMMError):
        eng.set_option("no_such_option", 1)
    eng.close()


def test_engines_agree_at_config3_size(monkeypatch):
    """BASELINE config 3 (10k cells x 500 sites, alphabets up to 100 states): the two independent GPU engines (lumped
    records vs full-state rows) must agree, and the size-independent properties of the counts must hold.  (Both are
    compared with the threaded oracle on all 500 sites in tests/test_zz_gpu_parity_sizes.py.)"""
    from laml_b200.capi import Engine
    from laml_b200.synth import CONFIGS, make_workload
    w = make_workload(seed=3, **CONFIGS["config3_10k_x_500"])
    N, K = w.forest.n_leaves, w.codes.shape[1]
    out = {}
    for mode in ("lumped", "fullstate"):
        monkeypatch.setenv("PMM_ENGINE", mode)
        eng = Engine(N, w.qptr, w.qval, w.codes)
        eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
        out[mode] = eng.estep(w.d, w.nu, w.phi, want_site_llh=True)
        eng.close()
    a, b = out["lumped"], out["fullstate"]
    assert abs(a["llh"] - b["llh"]) <= 1e-12 * abs(b["llh"])
    np.testing.assert_allclose(a["site_llh"], b["site_llh"], rtol=1e-12)
    np.testing.assert_allclose(a["S"], b["S"], rtol=1e-8, atol=1e-10 * K)
    assert abs(a["Rtilde"] - b["Rtilde"]) <= 1e-10 * b["Rtilde"]
    # the five counts and the silenced-stays-silenced mass partition each site: 0 <= sum_k S_k <= K per branch
    assert (a["S"] >= -1e-9).all() and (a["S"].sum(axis=1) <= K * (1 + 1e-9)).all()
    # a leaf observed in an unsilenced state cannot have a silenced parent edge: its S2 + S4 vanish on those sites
    obs = (w.codes >= 0).sum(axis=1)
    assert (a["S"][:N, 2] + a["S"][:N, 4] <= (K - obs) + 1e-6).all()


def test_full_size_properties_config4():
    """BASELINE config 4 (100k cells x 2k sites) on the default engine: properties that need no oracle.
    Additivity over sites (two site shards sum to the whole -- what the multi-GPU all-reduce relies on), counts
    partition every site, llh = sum of site log-likelihoods, and likelihood-only pass == E-step's likelihood."""
    import torch
    from laml_b200.capi import Engine
    from laml_b200.synth import CONFIGS, make_workload
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~35 GB of device memory")
    w = make_workload(seed=1, device="cuda", **CONFIGS["config4_100k_x_2k"])
    N, K = w.forest.n_leaves, w.codes.shape[1]

    def run(lo, hi):
        qp = (w.qptr[lo:hi + 1] - w.qptr[lo]).astype(np.int32)
        qv = w.qval[w.qptr[lo]:w.qptr[hi]].copy()
        eng = Engine(N, qp, qv, np.ascontiguousarray(w.codes[:, lo:hi]))
        eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
        out = eng.estep(w.d, w.nu, w.phi, want_site_llh=True)
        llh_only = eng.inside(w.d, w.nu, w.phi)
        eng.close()
        assert llh_only == out["llh"]
        return out

    whole = run(0, K)
    a, b = run(0, 1000), run(1000, K)
    # the oracle (the reference's algorithm) on a site sample of THIS 100k-cell tree: per-site likelihoods of the whole
    # run, and the expected counts of an engine holding exactly those sites, at the north-star tolerances
    import os
    from oracle.pmm_oracle import OracleProblem
    ks = 16
    qp, qv = (w.qptr[:ks + 1] - w.qptr[0]).astype(np.int32), w.qval[:w.qptr[ks]].copy()
    ref = OracleProblem(w.forest, np.arange(N, dtype=np.int32), np.ascontiguousarray(w.codes[:, :ks]), qp, qv).estep(
        w.d, w.nu, w.phi, nthreads=max(1, len(os.sched_getaffinity(0))))
    np.testing.assert_allclose(whole["site_llh"][:ks], ref["site_llh"], rtol=1e-9)
    sample = run(0, ks)
    assert abs(sample["llh"] - ref["llh"]) <= 1e-9 * abs(ref["llh"])
    np.testing.assert_allclose(sample["S"], ref["S"].T, rtol=1e-8, atol=1e-10 * ks)
    assert abs(sample["Rtilde"] - ref["Rtilde"]) <= 1e-8 * ref["Rtilde"]
    assert np.isfinite(whole["llh"]) and abs(whole["llh"] - whole["site_llh"].sum()) <= 1e-9 * abs(whole["llh"])
    assert abs((a["llh"] + b["llh"]) - whole["llh"]) <= 1e-12 * abs(whole["llh"])
    assert (np.concatenate([a["site_llh"], b["site_llh"]]) == whole["site_llh"]).all()       # per-site results are shard-independent
    np.testing.assert_allclose(a["S"] + b["S"], whole["S"], rtol=1e-10, atol=1e-9)
    assert abs((a["Rtilde"] + b["Rtilde"]) - whole["Rtilde"]) <= 1e-10 * whole["Rtilde"]
    S = whole["S"]
    assert (S >= -1e-9).all() and (S.sum(axis=1) <= K * (1 + 1e-9)).all()
    obs = (w.codes >= 0).sum(axis=1)
    assert (S[:N, 2] + S[:N, 4] <= (K - obs) + 1e-6).all()       # an observed leaf is not silenced on its own branch


def test_mstep_and_em_vs_oracle():
    case, z = load_npz_case("ref_test4_zero")
    orc = case.oracle()
    eng = make_engine(case)
    est = eng.estep(case.d, 0.1, 0.1)
    R = orc.R()
    d1, nu1, phi1 = eng.mstep(est["S"], est["Rtilde"], R, orc.n_cells, case.d, 0.1, 0.1)
    d0, nu0, phi0, _ = orc.mstep(est["S"].T.copy(), est["Rtilde"], case.d, 0.1)
    np.testing.assert_allclose(d1, d0, rtol=1e-12, atol=1e-14)
    assert abs(nu1 - nu0) <= 1e-12 and abs(phi1 - phi0) <= 1e-14
    # full EM from the reference's starting lengths
    d_init = np.maximum(case.d, 0.01)
    got = eng.em(d_init, 0.1, 0.1, R, orc.n_cells)
    ref = orc.em(d_init, 0.1, 0.1)
    assert got["iters"] == ref["iters"]
    assert abs(got["nllh"] - ref["nllh"]) <= 1e-9 * ref["nllh"]
    np.testing.assert_allclose(got["d"], ref["d"], rtol=1e-5, atol=1e-7)
    assert abs(got["nu"] - ref["nu"]) <= 1e-5 and abs(got["phi"] - ref["phi"]) <= 1e-5
    eng.close()


def test_set_tree_reuse_matches_fresh_engine():
    """pmm_set_tree on a live handle (topology search re-scores many trees over the same cells)
    must give the same numbers as a fresh handle"""
    from laml_b200.capi import Engine
    from laml_b200.synth import make_workload, random_tree, lognormal_branch_lengths
    w = make_workload(120, 20, 2, 25, 0.2, 0.2, seed=61)
    eng = Engine(120, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    a0 = eng.estep(w.d, w.nu, w.phi)
    launches0 = eng.info()["device_bytes"]
    for seed in (62, 63):
        f2 = random_tree(120, seed=seed)
        d2 = lognormal_branch_lengths(f2, seed=seed + 100)
        eng.set_tree(f2.parent, f2.left, f2.right)
        got = eng.estep(d2, w.nu, w.phi)
        fresh = Engine(120, w.qptr, w.qval, w.codes)
        fresh.set_tree(f2.parent, f2.left, f2.right)
        ref = fresh.estep(d2, w.nu, w.phi)
        assert got["llh"] == ref["llh"] and (got["S"] == ref["S"]).all() and got["Rtilde"] == ref["Rtilde"]
        fresh.close()
    assert eng.info()["device_bytes"] == launches0        # nothing was reallocated
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    a1 = eng.estep(w.d, w.nu, w.phi)
    assert a1["llh"] == a0["llh"] and (a1["S"] == a0["S"]).all()
    eng.close()
