"""The device ultrametric M-step kernel (``pmm_um_solve``, laml_b200/csrc/pmm_ultrametric_dev.cuh) run on the CPU under the
CUDA host shim -- its unchanged source as one CTA -- against the library's host solver ``pmm_ultrametric_solve``
(itself pinned to the reference's program on SLSQP and to MOSEK's published optima): same status, lengths equal to the
solvers' stopping precision, Newton step counts within 2, cold and warm, free / fixed / polytomy-marked branches, a
forest, an infeasible problem, nu = 0, and a tree wide enough for the grid-wide level phases."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from laml_b200.capi import ultrametric_solve                    # noqa: E402
from laml_b200.flatten import forest_from_parent                # noqa: E402
from laml_b200.synth import random_tree                         # noqa: E402

CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers not found")
NARROW = 512


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu_um") / "libemu_um.so")
    extra = os.environ.get("PMM_EMU_CXXFLAGS", "").split()
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread", "-w", "-I" + CUDA_INC] + extra +
                   [os.path.join(HERE, "native", "emu_um.cpp"), "-o", so], check=True)
    lib = C.CDLL(so)
    lib.emu_um_solve.restype = C.c_int
    return lib


def level_lists(f):
    """what pmm_set_tree keeps: internal nodes by height (children first) and by depth (roots first), as (u, a, b, _)"""
    n, N = f.n_nodes, f.n_leaves
    height = np.zeros(n, np.int64)
    for v in f.order:                                       # children before parents
        v = int(v)
        if f.left[v] >= 0:
            height[v] = 1 + max(height[f.left[v]], height[f.right[v]])
    def build(key, lo):
        items, ptr = [], [0]
        for l in range(lo, int(key[N:].max()) + 1):
            for u in np.nonzero(key[N:] == l)[0] + N:
                items.append((u, f.left[u], f.right[u], 0))
            ptr.append(len(items))
        return np.array(items, np.int32).reshape(-1, 4), np.array(ptr, np.int32)
    up, up_ptr = build(height, 1)
    dn, dn_ptr = build(np.asarray(f.depth, np.int64), 0)
    def phases(ptr):
        ph, l, nl = [], 0, len(ptr) - 1
        while l < nl:
            if ptr[l + 1] - ptr[l] > NARROW:
                ph.append((l, l + 1, 1, 0)); l += 1; continue
            e = l
            while e < nl and ptr[e + 1] - ptr[e] <= NARROW:
                e += 1
            ph.append((l, e, 0, 0)); l = e
        return np.array(ph, np.int32).reshape(-1, 4)
    return up, up_ptr, dn, dn_ptr, phases(up_ptr), phases(dn_ptr)


def device_solve(lib, f, kind, fixed_len, S_raw, nu, K_total, lo=0.005, hi=10.0, H_prev=None):
    n = f.n_nodes
    up, up_ptr, dn, dn_ptr, up_ph, dn_ph = level_lists(f)
    roots = np.nonzero(f.parent < 0)[0].astype(np.int32)
    parent, left = np.ascontiguousarray(f.parent, np.int32), np.ascontiguousarray(f.left, np.int32)
    kind = np.ascontiguousarray(kind, np.int8); fixed_len = np.ascontiguousarray(fixed_len, np.float64)
    stats = np.ascontiguousarray(S_raw, np.float64)
    Snorm = np.zeros((n, 5)); params = np.array([nu, 0.0]); d_out = np.zeros(n); result = np.zeros(2, np.int32)
    Hp = np.zeros(n + 1) if H_prev is None else np.array(H_prev, np.float64, copy=True)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emu_um_solve(n, len(roots), p(parent), p(left), p(roots), p(up), p(dn), p(up_ptr), p(dn_ptr), p(up_ph), len(up_ph), p(dn_ph),
                          len(dn_ph), p(kind), p(fixed_len), p(stats), p(Snorm), p(params), C.c_double(lo), C.c_double(hi),
                          C.c_double(1e-6), C.c_double(1e-5), int(K_total), int(H_prev is not None), p(Hp), p(d_out), p(result), 256)
    assert rc == 0
    return d_out, Hp, int(result[0]), int(result[1]), Snorm


def normalised(S_raw, K_total, eps_s=1e-6):
    s = np.maximum(S_raw, eps_s)
    return s / s.sum(axis=1, keepdims=True) * K_total


def random_problem(n_leaves, seed, forest=False, n_fixed=0, n_zero=0):
    rng = np.random.default_rng(seed)
    if forest:
        from test_lumped_schedule import forest_of
        trees = [random_tree(k, seed=seed + i) for i, k in enumerate((n_leaves // 2, 3, n_leaves - n_leaves // 2 - 3))]
        N, parent, left, right = forest_of(trees)
        f = forest_from_parent(parent, left, right)
    else:
        f = random_tree(n_leaves, seed=seed)
    n = f.n_nodes
    S = rng.gamma(0.6, 1.0, (n, 5)) * rng.choice([1e-8, 0.1, 3.0], (n, 5))
    kind = np.zeros(n, np.int8); fixed_len = np.zeros(n)
    internal = np.nonzero(f.left >= 0)[0]
    if n_zero:
        z = rng.choice(internal[np.asarray(f.parent)[internal] >= 0], n_zero, replace=False)
        kind[z] = 2
    if n_fixed:             # internal branches: a fixed leaf branch pins its parent's height and is rarely feasible
        cand = np.nonzero((kind == 0) & (np.asarray(f.left) >= 0) & (np.asarray(f.parent) >= 0))[0]
        fx = rng.choice(cand, n_fixed, replace=False)
        kind[fx] = 1; fixed_len[fx] = rng.uniform(0.01, 0.05, n_fixed)
    return f, kind, fixed_len, S


@pytest.mark.parametrize("n_leaves,seed,nu,forest,n_fixed,n_zero", [(30, 1, 0.2, False, 0, 0), (80, 2, 0.0, False, 0, 0),
                                                                     (60, 3, 0.4, True, 0, 0), (50, 4, 0.15, False, 0, 3), (70, 7, 0.2, False, 6, 2), (64, 8, 0.3, True, 4, 0),
                                                                     (2, 5, 0.3, False, 0, 0), (3, 6, 0.3, False, 0, 0)])
def test_device_kernel_equals_the_host_solver(emu, n_leaves, seed, nu, forest, n_fixed, n_zero):
    f, kind, fixed_len, S = random_problem(n_leaves, seed, forest, n_fixed, n_zero)
    K = 37
    d_ref, H_ref, rc_ref, it_ref = ultrametric_solve(f.parent, kind, fixed_len, 0.005, 10.0, normalised(S, K), nu)
    d, H, rc, it, Snorm = device_solve(emu, f, kind, fixed_len, S, nu, K)
    assert rc == rc_ref == 0
    np.testing.assert_allclose(Snorm, normalised(S, K), rtol=1e-14)
    np.testing.assert_allclose(d, d_ref, rtol=0, atol=2e-7)              # the solvers' own stopping precision
    assert abs(it - it_ref) <= 2
    # warm: the second M-step of an EM run starts from the first one's heights
    S2 = S * np.random.default_rng(seed + 100).uniform(0.9, 1.1, S.shape)
    d_ref2, _, rc_ref2, it_ref2 = ultrametric_solve(f.parent, kind, fixed_len, 0.005, 10.0, normalised(S2, K), nu, H0=H_ref)
    d2, _, rc2, it2, _ = device_solve(emu, f, kind, fixed_len, S2, nu, K, H_prev=H)
    assert rc2 == rc_ref2 == 0
    np.testing.assert_allclose(d2, d_ref2, rtol=0, atol=2e-7)
    assert abs(it2 - it_ref2) <= 3 and it2 < it                           # the warm start pays


def test_fixed_branches_and_infeasible_problem(emu):
    f = random_tree(40, seed=9)
    n = f.n_nodes
    rng = np.random.default_rng(9)
    S = rng.gamma(0.6, 1.0, (n, 5))
    # fix the branches of one cherry to different lengths: no ultrametric tree has them
    cherry = next(u for u in range(f.n_leaves, n) if f.left[u] < f.n_leaves and f.right[u] < f.n_leaves)
    kind = np.zeros(n, np.int8); fixed_len = np.zeros(n)
    kind[[f.left[cherry], f.right[cherry]]] = 1
    fixed_len[f.left[cherry]], fixed_len[f.right[cherry]] = 0.02, 0.05
    _, _, rc_ref, _ = ultrametric_solve(f.parent, kind, fixed_len, 0.005, 10.0, normalised(S, 20), 0.1)
    _, _, rc, it, _ = device_solve(emu, f, kind, fixed_len, S, 0.1, 20)
    assert rc == rc_ref == 1 and it == 0
    # the same lengths on both: feasible, and the fixed branches come back as given
    fixed_len[f.right[cherry]] = 0.02
    extra = [v for v in range(n) if kind[v] == 0 and f.left[v] >= 0 and f.parent[v] >= 0][:2]
    d_ref, _, rc_ref, it_ref = ultrametric_solve(f.parent, kind, fixed_len, 0.005, 10.0, normalised(S, 20), 0.1)
    d, _, rc, it, _ = device_solve(emu, f, kind, fixed_len, S, 0.1, 20)
    assert rc == rc_ref == 0 and abs(it - it_ref) <= 2
    np.testing.assert_allclose(d, d_ref, rtol=0, atol=2e-7)
    assert d[f.left[cherry]] == 0.02 and d[f.right[cherry]] == 0.02


def test_wide_levels_take_the_grid_phases(emu):
    """a tree whose lower levels exceed the narrow-level bound: the grid-wide phases (here a one-CTA grid) and the
    narrow runs above and below them give the host's optimum"""
    f, kind, fixed_len, S = random_problem(7000, 21)
    up, up_ptr, dn, dn_ptr, up_ph, dn_ph = level_lists(f)
    assert up_ph[:, 2].any() and dn_ph[:, 2].any() and (up_ph[:, 2] == 0).any()
    d_ref, H_ref, rc_ref, it_ref = ultrametric_solve(f.parent, kind, fixed_len, 0.005, 10.0, normalised(S, 50), 0.25)
    d, H, rc, it, _ = device_solve(emu, f, kind, fixed_len, S, 0.25, 50)
    assert rc == rc_ref == 0 and abs(it - it_ref) <= 3
    np.testing.assert_allclose(d, d_ref, rtol=0, atol=5e-7)
    tips = H[f.parent[np.arange(f.n_leaves)]] - d[: f.n_leaves]          # leaves sit at height 0
    assert np.abs(tips).max() <= 1e-9
