"""The lumped-state kernels THEMSELVES on the CPU: the unchanged source of laml_b200/csrc/pmm_lumped.cuh and
pmm_nni_lumped.cuh compiled for the host under a small CUDA shim (tests/native/cuda_host_shim.h: one std::thread
per CUDA thread, ``__shared__`` as function-local statics, barriers for __syncthreads / warp primitives).

* ``pmm_lumped_inside`` + ``pmm_lumped_outside`` -- the E-step, the hot path -- on the clade schedule the library
  derives (``pmm_debug_lumped_schedule``), launch by launch as ``run_lumped`` issues them, against the CPU oracle at
  the GPU parity tolerances;
* ``pmm_nni_lumped<ULTRA>`` -- the batch NNI scorer, unconstrained and under the ultrametric constraint -- fed with
  the inside / outside records of the numpy model of the lumped engine (tests/lumped_model.py), against the
  definition (tests/oracle_engine.py: a full-tree EM per candidate on the re-wired topology).

Covers what a machine without a GPU otherwise cannot: the kernels' own index arithmetic, record packing, closed
forms, prefetch / register-set rotation, warp reductions, EM loop and M-steps.  Says nothing about memory ordering,
registers or speed -- the ``-m gpu`` tests and the profiles do."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from laml_b200.nni import nni_candidates            # noqa: E402
from laml_b200.synth import make_workload           # noqa: E402
from lumped_model import estep as lumped_estep      # noqa: E402
from oracle_engine import OracleEngine              # noqa: E402
from test_nni_ultra_host import ultrametric_lengths  # noqa: E402

CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers not found")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu_nni.so")
    src = os.path.join(HERE, "native", "emu_lumped.cpp")
    # PMM_EMU_CXXFLAGS="-fsanitize=address,undefined -fno-omit-frame-pointer -g" (with libasan preloaded) turns this file into
    # a memory / UB checker of the kernels' source on the numpy-owned record arrays: scripts/native/sanitize_emulated_kernels.sh
    extra = os.environ.get("PMM_EMU_CXXFLAGS", "").split()
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread", "-w", "-I" + CUDA_INC] + extra + [src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.emu_nni_lumped.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu_warp(tmp_path_factory):
    """the same kernels with the constrained M-step in its WARP form (pmm_nni_ultra.cuh: the transcendental terms of the
    five edges spread over twenty lanes, handed round by shuffles) -- on the GPU that is the only form"""
    so = str(tmp_path_factory.mktemp("emu_warp") / "libemu_nni_warp.so")
    src = os.path.join(HERE, "native", "emu_lumped.cpp")
    extra = os.environ.get("PMM_EMU_CXXFLAGS", "").split()
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread", "-w", "-DPMM_ULTRA2_EMU_WARP", "-I" + CUDA_INC] + extra + [src, "-o", so],
                   check=True)
    lib = C.CDLL(so)
    lib.emu_nni_lumped.restype = C.c_int
    return lib


def pack_records(forest, m, K):
    """the model's records in the engine's memory layout: X [n_int][Kp] x (x0, xz, t, h|cum), O (o0, om, oB, oD), OI (k, F)"""
    N, n = forest.n_leaves, forest.n_nodes
    Kp = (K + 31) // 32 * 32
    X = np.zeros((n - N, Kp, 4)); O = np.zeros((n - N, Kp, 4)); OI = np.zeros((n - N, Kp, 2), np.int32)
    for u in range(N, n):
        x0, xm, xz, qh, h, cum = m["rec"][u]
        hot = h >= 0
        assert np.all(xm[hot] == 0.0)                                  # x_m and q_h share a slot
        X[u - N, :K, 0], X[u - N, :K, 1] = x0, xz
        X[u - N, :K, 2] = np.where(hot, qh, xm)
        packed = (np.asarray(cum, np.int64).astype(np.uint64) << np.uint64(32)) | (np.asarray(h, np.int64).astype(np.uint64) & np.uint64(0xffffffff))
        X[u - N, :K, 3] = packed.view(np.float64)
        o0, om, oB, oD, k, F = m["out"][u]
        O[u - N, :K] = np.stack([o0, om, oB, oD], axis=1)
        OI[u - N, :K, 0], OI[u - N, :K, 1] = k, F
    return Kp, X, O, OI


def run_kernel(emu, w, d, nu, phi, ultra, max_iter=1000, records=None):
    """``records``: the output of :func:`run_estep_kernels` with ``store_all`` -- the NNI kernel then reads what the
    emulated E-step kernels wrote, as on the GPU; default: records laid out from the numpy model"""
    f = w.forest
    N, n, K = f.n_leaves, f.n_nodes, w.codes.shape[1]
    roots = [int(r) for r in np.nonzero(f.parent < 0)[0]]
    if records is None:
        m = lumped_estep(f, np.arange(N), w.codes, w.qptr, w.qval, d, nu, phi, dump=True)
        Kp, X, O, OI = pack_records(f, m, K)
        invL = np.stack([m["invL"][r] for r in roots])
        rsum = np.array([m["root_llh"][r].sum() for r in roots])
    else:
        Kp, X, O, OI, invL = (K + 31) // 32 * 32, records["X"], records["O"], records["OI"], records["invL"]
        rsum = records["root_llh"].sum(axis=1)
        m = {"llh": records["llh"], "qsum": records["qsum"], "bc": records["bc"]}
    root_index = np.zeros(n, np.int32)
    order = list(roots)
    for v in order:
        if f.left[v] >= 0:
            order += [int(f.left[v]), int(f.right[v])]
    for i, r in enumerate(roots):
        root_index[r] = i
    for v in order:
        if f.parent[v] >= 0:
            root_index[v] = root_index[f.parent[v]]
    cu, ck = nni_candidates(f.parent, f.left, N)
    nc = len(cu)
    cand = np.zeros((nc, 8), np.int32); d5 = np.zeros((nc, 5)); ul = np.zeros((nc, 5)); nodes = np.zeros((nc, 5), np.int32)
    height = np.zeros(n)
    for v in reversed(order):
        if f.left[v] >= 0:
            height[v] = max(height[f.left[v]] + d[f.left[v]], height[f.right[v]] + d[f.right[v]])
    top = max(height[r] + d[r] for r in roots)
    for i, (u, k) in enumerate(zip(cu.tolist(), ck.tolist())):
        v = int(f.parent[u]); wn = int(f.right[v] if f.left[v] == u else f.left[v])
        c_k = int(f.right[u] if k else f.left[u]); c_o = int(f.left[u] if k else f.right[u])
        g = int(f.parent[v]); s = -1 if g < 0 else int(f.right[g] if f.left[g] == v else f.left[g])
        cand[i] = (g, s, c_o, wn, c_k, v, u, root_index[u])
        nodes[i] = (v, u, c_k, c_o, wn)
        d5[i] = d[nodes[i]]
        ul[i] = (height[c_o], height[wn], height[c_k], height[g] if g >= 0 else top, float(g < 0 and len(roots) == 1))
    codes = np.ascontiguousarray(w.codes, np.int32)
    site_ptr = np.ascontiguousarray(w.qptr, np.int32); prior = np.ascontiguousarray(w.qval, np.float64)
    qsum = np.ascontiguousarray(m["qsum"]); bc = np.ascontiguousarray(m["bc"])
    out_llh = np.zeros(nc); out_d5 = np.zeros((nc, 5)); out_it = np.zeros(nc, np.int32); out_done = np.zeros(nc, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    rc = emu.emu_nni_lumped(int(ultra), nc, p(cand), p(d5), p(ul), K, Kp, N, n, p(codes), p(site_ptr), p(prior), p(qsum),
                            C.c_double(nu), C.c_double(phi), p(bc), p(X), p(O), p(OI), p(np.ascontiguousarray(invL)), p(rsum),
                            C.c_double(m["llh"]), K, C.c_double(0.005), C.c_double(10.0), C.c_double(1e-8), max_iter,
                            p(out_llh), p(out_d5), p(out_it), p(out_done))
    assert rc == 0
    return cu, ck, nodes, -out_llh, out_d5, out_it, out_done


@pytest.mark.parametrize("seed,n,K,amax,nu,phi", [(71, 12, 150, 12, 0.15, 0.1), (73, 10, 20, 6, 0.3, 0.0)])
def test_unconstrained_kernel_on_the_host(emu, seed, n, K, amax, nu, phi):
    """the kernel of pmm_score_nni_batch (already green on a B200) through the same emulation: validates the shim"""
    w = make_workload(n, K, 1, amax, max(nu, 0.1), max(phi, 0.1), seed=seed)
    d = np.maximum(w.d, 0.01)
    cu, ck, nodes, nllh, d5, iters, done = run_kernel(emu, w, d, nu, phi, ultra=False)
    model = OracleEngine(n, w.qptr, w.qval, w.codes)
    model.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    want = model.score_nni_batch(cu, ck, d, nu, phi)
    assert (done == 1).all() and (nodes == want["nodes"]).all()
    np.testing.assert_allclose(nllh, want["nllh"], rtol=1e-9)
    np.testing.assert_allclose(d5, want["d5"], rtol=1e-6, atol=1e-8)
    assert (iters == want["iters"]).all()


@pytest.mark.parametrize("seed,n,K,amax,nu,phi,step", [(91, 12, 150, 12, 0.15, 0.1, (0.02, 0.3)), (92, 14, 9, 40, 0.0, 0.2, (0.006, 0.012))])
def test_constrained_kernel_on_the_host(emu, seed, n, K, amax, nu, phi, step):
    """pmm_nni_lumped<true> -- the kernel behind pmm_score_nni_batch_ultra, not yet run on a GPU -- against the
    definition: scores to 1e-9, the five lengths, EM iteration counts, and which candidates have no room"""
    w = make_workload(n, K, 1, amax, max(nu, 0.1), max(phi, 0.1), seed=seed)
    d, _ = ultrametric_lengths(w.forest, np.random.default_rng(seed), step=step)
    cu, ck, nodes, nllh, d5, iters, done = run_kernel(emu, w, d, nu, phi, ultra=True)
    model = OracleEngine(n, w.qptr, w.qval, w.codes)
    model.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    want = model.score_nni_batch(cu, ck, d, nu, phi, ultra=True)
    bad = np.isnan(want["nllh"])
    assert ((done == 2) == bad).all() and (done[~bad] == 1).all()
    if seed == 92:
        assert bad.sum() >= 3
    ok = ~bad
    np.testing.assert_allclose(nllh[ok], want["nllh"][ok], rtol=1e-9)
    np.testing.assert_allclose(d5[ok], want["d5"][ok], rtol=1e-5, atol=1e-7)
    assert (np.abs(iters[ok] - want["iters"][ok]) <= 1).all()


def test_constrained_kernel_warp_form_on_the_host(emu, emu_warp):
    """the warp form of the constrained M-step takes the serial form's decisions: same scores, lengths, iteration
    counts and infeasible candidates, to the bit (the sums are formed in the serial order on every lane)"""
    seed, n, K, amax, nu, phi, step = 92, 14, 9, 40, 0.15, 0.2, (0.006, 0.012)
    w = make_workload(n, K, 1, amax, nu, phi, seed=seed)
    d, _ = ultrametric_lengths(w.forest, np.random.default_rng(seed), step=step)
    a = run_kernel(emu, w, d, nu, phi, ultra=True)
    b = run_kernel(emu_warp, w, d, nu, phi, ultra=True)
    for x, y in zip(a[3:], b[3:]):
        assert np.asarray(x).tobytes() == np.asarray(y).tobytes()
    assert (a[6] == 2).sum() >= 1 and (a[6] == 1).sum() >= 5


# ---------------------------------------------------------------------------------------------------------------------
# the E-step kernels themselves (pmm_lumped_inside / pmm_lumped_outside), on the clade schedule the library derives
def run_estep_kernels(emu, w, d, nu, phi, clade_cap, clade_cap_top, store_all=0, posteriors=False, df_mode=0):
    from laml_b200.capi import lumped_schedule
    f = w.forest
    N, n, K = f.n_leaves, f.n_nodes, w.codes.shape[1]
    sch = lumped_schedule(N, f.parent, f.left, f.right, clade_cap, clade_cap_top)
    Kp, G = (K + 31) // 32 * 32, (K + 31) // 32
    n_roots = int((f.parent < 0).sum())
    qptr = np.asarray(w.qptr, np.int64)
    qsum = np.array([w.qval[qptr[s]:qptr[s + 1]].sum() for s in range(K)])
    X = np.zeros((n - N, Kp, 4)); O = np.zeros((n - N, Kp, 4)); OI = np.zeros((n - N, Kp, 2), np.int32)
    invL = np.zeros((n_roots, K)); rllh = np.zeros((n_roots, K))
    Spart = np.zeros((G, n, 5)); Rpart = np.zeros((G, max(sch["n_clades"], 1))); bc = np.zeros((n, 4))
    codes = np.ascontiguousarray(w.codes, np.int32)
    site_ptr = np.ascontiguousarray(w.qptr, np.int32); prior = np.ascontiguousarray(w.qval, np.float64)
    dd = np.ascontiguousarray(d, np.float64)
    arrs = {k: np.ascontiguousarray(sch[k], np.int32) for k in ("in_steps", "in_clades", "in_phase_ptr", "out_steps", "out_clades", "out_phase_ptr",
                                                                "in_dep_ptr", "in_deps", "out_dep_ptr", "out_deps")}
    post0 = np.full((n, K), np.nan) if posteriors else None
    post1 = np.full((n, K), np.nan) if posteriors else None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)   # noqa: E731
    rc = emu.emu_lumped_estep(K, Kp, N, n, sch["n_clades"], p(codes), p(site_ptr), p(prior), p(qsum), C.c_double(nu), C.c_double(phi),
                              p(dd), p(arrs["in_steps"]), p(arrs["in_clades"]), p(arrs["in_phase_ptr"]), len(arrs["in_phase_ptr"]) - 1,
                              p(arrs["out_steps"]), p(arrs["out_clades"]), p(arrs["out_phase_ptr"]), len(arrs["out_phase_ptr"]) - 1,
                              store_all, p(X), p(O), p(OI), p(invL), p(rllh), p(Spart), p(Rpart), p(bc), p(post0), p(post1),
                              df_mode, p(arrs["in_dep_ptr"]), p(arrs["in_deps"]), p(arrs["out_dep_ptr"]), p(arrs["out_deps"]))
    assert rc == 0, "a dataflow CTA found a producer's mark missing" if rc == 2 else rc
    return {"S": Spart.sum(axis=0), "llh": float(rllh.sum()), "Rtilde": float(Rpart.sum()), "site_llh": rllh.sum(axis=0),
            "n_phases": sch["n_phases"], "n_clades": sch["n_clades"], "X": X, "O": O, "OI": OI, "invL": invL, "root_llh": rllh,
            "bc": bc, "qsum": qsum, "post0": post0, "post1": post1}


@pytest.mark.parametrize("seed,n,K,amax,nu,phi,caps", [(31, 40, 70, 12, 0.2, 0.15, (6, 4)), (32, 24, 33, 40, 0.0, 0.3, (127, 127)),
                                                       (33, 30, 140, 5, 0.4, 0.0, (3, 3))])
def test_estep_kernels_on_the_host(emu, seed, n, K, amax, nu, phi, caps):
    """pmm_lumped_inside + pmm_lumped_outside, unchanged source, on the host: expected counts, likelihood and R~
    against the CPU oracle at the GPU parity tolerances (several launches per pass with small clade capacities,
    one with 127; K not a multiple of 32 or 128: idle lanes, a second site block)"""
    from oracle.pmm_oracle import OracleProblem
    w = make_workload(n, K, 1, amax, max(nu, 0.1), max(phi, 0.1), seed=seed)
    got = run_estep_kernels(emu, w, w.d, nu, phi, *caps)
    ref = OracleProblem(w.forest, np.arange(n, dtype=np.int32), w.codes, w.qptr, w.qval).estep(w.d, nu, phi)
    assert abs(got["llh"] - ref["llh"]) <= 1e-9 * abs(ref["llh"])
    np.testing.assert_allclose(got["S"], ref["S"].T, rtol=1e-8, atol=1e-10 * K)
    assert abs(got["Rtilde"] - ref["Rtilde"]) <= 1e-8 * max(1.0, abs(ref["Rtilde"]))
    np.testing.assert_allclose(got["site_llh"], ref["site_llh"], rtol=1e-9, atol=1e-12)
    if caps[0] < 10:
        assert got["n_phases"] > 1 and got["n_clades"] > 3


@pytest.mark.parametrize("seed,n,K,caps,forest", [(34, 60, 300, (5, 3), False), (35, 33, 40, (2, 2), False), (36, 90, 150, (16, 4), False),
                                                   (37, 24, 130, (3, 2), True)])
def test_dataflow_launches_on_the_host(emu, seed, n, K, caps, forest):
    """the dataflow variants of both kernels (several tiers in ONE launch: atomic tickets, waits on the producers'
    marks) against the launch-per-tier run of the same schedule: every record, count partial and likelihood to the
    bit, for the launch over all tiers (mode 2) and over the top tiers only (mode 1), one to three site blocks"""
    w = make_workload(n, K, 1, 14, 0.2, 0.15, seed=seed)
    if forest:
        from types import SimpleNamespace
        from laml_b200.flatten import forest_from_parent
        from test_lumped_schedule import forest_of
        from laml_b200.synth import random_tree
        trees = [random_tree(m, seed=seed + i) for i, m in enumerate((n - 10, 3, 7))]
        N, parent, left, right = forest_of(trees)
        f = forest_from_parent(parent, left, right)
        w = SimpleNamespace(forest=f, codes=w.codes[:N], qptr=w.qptr, qval=w.qval, d=np.resize(w.d, f.n_nodes))
    base = run_estep_kernels(emu, w, w.d, 0.2, 0.15, *caps)
    assert base["n_phases"] >= 2
    for mode in (2, 1):
        got = run_estep_kernels(emu, w, w.d, 0.2, 0.15, *caps, df_mode=mode)
        for key in ("X", "O", "OI", "invL", "root_llh", "S"):
            assert got[key].tobytes() == base[key].tobytes(), (mode, key)      # (packed h | cum words may look like NaN)
        assert got["llh"] == base["llh"] and got["Rtilde"] == base["Rtilde"]
    post = run_estep_kernels(emu, w, w.d, 0.2, 0.15, *caps, posteriors=True)
    got = run_estep_kernels(emu, w, w.d, 0.2, 0.15, *caps, posteriors=True, df_mode=2)
    assert np.array_equal(got["post0"], post["post0"]) and np.array_equal(got["post1"], post["post1"])


def test_posterior_variant_on_the_host(emu):
    """pmm_lumped_outside<true> (the annotation path, pmm_posteriors): log posteriors of state 0 / silenced per
    (node, site) against the oracle's dump of the reference's ``post0`` / ``post1`` (EM_solver.py:276-277)"""
    from oracle.pmm_oracle import OracleProblem
    w = make_workload(26, 45, 1, 9, 0.25, 0.2, seed=41)
    got = run_estep_kernels(emu, w, w.d, 0.25, 0.2, 5, 5, posteriors=True)
    ref = OracleProblem(w.forest, np.arange(26, dtype=np.int32), w.codes, w.qptr, w.qval).estep(w.d, 0.25, 0.2, dump=True)
    assert not np.isnan(got["post0"]).any() and not np.isnan(got["post1"]).any()     # every (node, site) was written
    for name in ("post0", "post1"):
        a, b = np.exp(got[name]), np.exp(np.asarray(ref[name]))
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-12)


def test_whole_batch_call_on_the_host(emu):
    """pmm_score_nni_batch_ultra end to end as the library runs it: E-step kernels with every out-record stored
    (``store_all``), then the NNI kernel on THOSE records -- every kernel of the call emulated -- against the definition"""
    seed, n, K, nu, phi = 95, 14, 40, 0.2, 0.1
    w = make_workload(n, K, 1, 10, nu, phi, seed=seed)
    d, _ = ultrametric_lengths(w.forest, np.random.default_rng(seed), step=(0.01, 0.2))
    rec = run_estep_kernels(emu, w, d, nu, phi, 4, 3, store_all=1)
    cu, ck, nodes, nllh, d5, iters, done = run_kernel(emu, w, d, nu, phi, ultra=True, records=rec)
    model = OracleEngine(n, w.qptr, w.qval, w.codes)
    model.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    want = model.score_nni_batch(cu, ck, d, nu, phi, ultra=True)
    bad = np.isnan(want["nllh"])
    assert ((done == 2) == bad).all() and (~bad).sum() > len(cu) // 2
    np.testing.assert_allclose(nllh[~bad], want["nllh"][~bad], rtol=1e-9)
    np.testing.assert_allclose(d5[~bad], want["d5"][~bad], rtol=1e-5, atol=1e-7)


def _check_estep(emu, forest, codes, qptr, qval, d, nu, phi, caps):
    from types import SimpleNamespace
    from oracle.pmm_oracle import OracleProblem
    w = SimpleNamespace(forest=forest, codes=codes, qptr=qptr, qval=qval)
    got = run_estep_kernels(emu, w, d, nu, phi, *caps)
    N, K = forest.n_leaves, codes.shape[1]
    ref = OracleProblem(forest, np.arange(N, dtype=np.int32), codes, qptr, qval).estep(d, nu, phi)
    assert abs(got["llh"] - ref["llh"]) <= 1e-9 * abs(ref["llh"])
    np.testing.assert_allclose(got["S"], ref["S"].T, rtol=1e-8, atol=1e-10 * K)
    assert abs(got["Rtilde"] - ref["Rtilde"]) <= 1e-8 * max(1.0, abs(ref["Rtilde"]))
    return got


def test_estep_kernels_larger_and_edge_shapes(emu):
    """the shapes the GPU edge-case tests use, through the emulated kernels: a larger tree at the engine's default
    clade capacities, a caterpillar-biased (deep) tree, a forest of two trees over the same cells, the smallest
    trees, a site with a single state"""
    from laml_b200.flatten import forest_from_parent
    from laml_b200.synth import lognormal_branch_lengths, random_priors, random_tree, simulate_codes
    w = make_workload(400, 200, 2, 60, 0.3, 0.3, seed=61)
    got = _check_estep(emu, w.forest, w.codes, w.qptr, w.qval, w.d, 0.3, 0.3, (64, 16))
    assert got["n_clades"] > 5
    w = make_workload(120, 40, 1, 8, 0.1, 0.2, seed=62, caterpillar_bias=0.85)
    _check_estep(emu, w.forest, w.codes, w.qptr, w.qval, w.d, 0.1, 0.2, (16, 16))
    # forest (layout as tests/test_gpu_edge_cases.py::test_forest_of_two_trees)
    N = 30
    f1, f2 = random_tree(N, seed=51), random_tree(N, seed=52)
    n1 = f1.n_nodes

    def shift(f, leaf_off, int_off):
        def m(a):
            return np.where(a < 0, -1, np.where(a < N, a + leaf_off, a - N + int_off))
        return m(f.parent), m(f.left), m(f.right)
    p1, l1, r1 = shift(f1, 0, 2 * N)
    p2, l2, r2 = shift(f2, N, 2 * N + (n1 - N))
    forest = forest_from_parent(*(np.concatenate([a[:N], b[:N], a[N:], b[N:]]).astype(np.int32)
                                  for a, b in ((p1, p2), (l1, l2), (r1, r2))))
    d = lognormal_branch_lengths(forest, seed=53)
    qptr, qval = random_priors(10, 2, 15, seed=54)
    c1 = simulate_codes(f1, lognormal_branch_lengths(f1, seed=55), qptr, qval, 0.2, 0.2, seed=56)
    _check_estep(emu, forest, np.concatenate([c1, c1]), qptr, qval, d, 0.2, 0.2, (8, 4))
    for n_leaves in (2, 3):
        f = random_tree(n_leaves, seed=91)
        d = lognormal_branch_lengths(f, seed=92)
        for codes in ([1, 0, -1], [-1, -1, -1], [1, 1, 1], [0, -2, 1]):
            c = np.array(codes[:n_leaves], np.int32).reshape(n_leaves, 1)
            _check_estep(emu, f, c, np.array([0, 1], np.int32), np.array([1.0]), d, 0.2, 0.1, (16, 16))
