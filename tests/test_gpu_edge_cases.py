"""Edge cases of the hot path on the GPU, against the oracle: deep (caterpillar) trees,
degenerate sites, extreme alphabets and parameters, impossible data, forests."""
import numpy as np
import pytest

from laml_b200.capi import Engine, PMMError
from laml_b200.flatten import forest_from_parent
from laml_b200.synth import make_workload, random_priors, random_tree, simulate_codes, lognormal_branch_lengths

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["lumped", "fullstate"])
def engine_mode(request, monkeypatch):
    """both E-step engines (lumped-state records = default; full-state rows) face every edge case"""
    monkeypatch.setenv("PMM_ENGINE", request.param)
    return request.param


def oracle_of(forest, codes, qptr, qval):
    from oracle.pmm_oracle import OracleProblem
    return OracleProblem(forest, np.arange(forest.n_leaves, dtype=np.int32), codes, qptr, qval)


def compare(forest, codes, qptr, qval, d, nu, phi, rtol_S=1e-8, rtol_llh=1e-9):
    K = codes.shape[1]
    ref = oracle_of(forest, codes, qptr, qval).estep(d, nu, phi, nthreads=4)
    eng = Engine(forest.n_leaves, qptr, qval, codes)
    eng.set_tree(forest.parent, forest.left, forest.right)
    got = eng.estep(d, nu, phi, want_site_llh=True)
    eng.close()
    dead = np.isneginf(ref["site_llh"])
    assert (np.isneginf(got["site_llh"]) == dead).all()            # impossible sites agree
    np.testing.assert_allclose(got["site_llh"][~dead], ref["site_llh"][~dead], rtol=rtol_llh, atol=1e-10)
    assert np.isfinite(got["S"]).all()
    if dead.any():
        return got                                                 # the oracle's counts are NaN there, like the reference's
    assert abs(got["llh"] - ref["llh"]) <= rtol_llh * abs(ref["llh"])
    np.testing.assert_allclose(got["S"], ref["S"].T, rtol=rtol_S, atol=1e-10 * K)
    assert abs(got["Rtilde"] - ref["Rtilde"]) <= 1e-8 * max(1.0, ref["Rtilde"])
    return got


def test_caterpillar_tree():
    """one internal node per level: 299 levels, every launch is a single node"""
    f = random_tree(300, seed=5, caterpillar_bias=1.0)
    assert f.depth.max() >= 298
    d = lognormal_branch_lengths(f, seed=6)
    qptr, qval = random_priors(20, 1, 25, seed=7)
    codes = simulate_codes(f, d * 40, qptr, qval, 0.2, 0.2, seed=8)
    compare(f, codes, qptr, qval, d * 40, 0.2, 0.2)


def test_degenerate_sites_and_alphabets():
    f = random_tree(64, seed=11)
    d = lognormal_branch_lengths(f, seed=12)
    A = np.array([1, 1, 2, 3, 600, 7, 40, 1])                      # tiny and huge alphabets side by side
    qptr = np.concatenate([[0], np.cumsum(A)]).astype(np.int32)
    rng = np.random.default_rng(13)
    qval = np.concatenate([rng.dirichlet(np.ones(a)) for a in A])
    codes = simulate_codes(f, d, qptr, qval, 0.3, 0.3, seed=14)
    codes[:, 0] = -1                                               # a column that is entirely missing
    codes[:, 1] = 0                                                # a column that is entirely unedited
    codes[:, 2] = 2                                                # every cell shows the same edited state
    codes[:5, 3] = -2                                              # known-silenced entries
    compare(f, codes, qptr, qval, d, 0.3, 0.3)


def test_alphabet_larger_than_default_tile():
    f = random_tree(16, seed=21)
    d = lognormal_branch_lengths(f, seed=22)
    A = np.array([2500, 3])
    qptr = np.concatenate([[0], np.cumsum(A)]).astype(np.int32)
    rng = np.random.default_rng(23)
    qval = np.concatenate([rng.dirichlet(np.ones(a)) for a in A])
    codes = simulate_codes(f, d, qptr, qval, 0.1, 0.1, seed=24)
    compare(f, codes, qptr, qval, d, 0.1, 0.1)


@pytest.mark.parametrize("nu,phi", [(0.0, 0.0), (1.0, 0.99), (1e-9, 1e-9), (0.5, 0.0), (0.0, 0.5)])
def test_extreme_parameters(nu, phi):
    w = make_workload(80, 12, 1, 12, 0.25, 0.25, seed=31)
    codes = w.codes.copy()
    if phi == 0.0 and nu == 0.0:
        codes[codes == -1] = 0                                     # no way to explain a '?' otherwise
    elif phi == 0.0 or nu == 0.0:
        pass
    d = w.d.copy()
    d[3::11] = 10.0                                                # dmax
    # nu = phi = 1e-9: the reference (and the oracle, which restates it) forms 1 - p**nu and loses ~7
    # digits of the silencing probability to cancellation (EM_solver.py:107,126); the engine uses
    # -expm1(-nu d) and keeps them, so the two differ by a few 1e-9 relative there
    tol = dict(rtol_llh=1e-8, rtol_S=1e-6) if 0 < nu < 1e-6 else {}
    got = compare(w.forest, codes, w.qptr, w.qval, d, nu, phi, **tol)
    assert np.isfinite(got["llh"])
    d[w.forest.n_leaves::5] = 0.0                                  # zero-length internal branches (test_46 of the reference)
    compare(w.forest, codes, w.qptr, w.qval, d, nu, phi, **tol)


def test_impossible_data_is_minus_infinity_not_nan():
    """nu = phi = 0 cannot produce a '?': the reference's log_sum_exp yields NaN there
    (EM_solver.py:8-13); the engine and the oracle define -inf likelihood and zero counts."""
    w = make_workload(40, 6, 2, 6, 0.2, 0.2, seed=41)
    assert (w.codes == -1).any()
    eng = Engine(40, w.qptr, w.qval, w.codes)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    got = eng.estep(w.d, 0.0, 0.0, want_site_llh=True)
    dead = (w.codes == -1).any(axis=0)
    assert np.isneginf(got["site_llh"][dead]).all() and np.isfinite(got["site_llh"][~dead]).all()
    assert got["llh"] == -np.inf and np.isfinite(got["S"]).all()
    ref = oracle_of(w.forest, w.codes, w.qptr, w.qval).inside(w.d, 0.0, 0.0)
    assert np.isneginf(ref["site_llh"][dead]).all()
    np.testing.assert_allclose(got["site_llh"][~dead], ref["site_llh"][~dead], rtol=1e-9)
    eng.close()


def test_forest_of_two_trees():
    """several trees over the same cells share nu, phi (ML_solver.py:26-32)"""
    f1, f2 = random_tree(50, seed=51), random_tree(50, seed=52)
    n1 = f1.n_nodes
    N = 50

    def shift(f, leaf_off, int_off):
        def m(a):
            return np.where(a < 0, -1, np.where(a < N, a + leaf_off, a - N + int_off))
        return m(f.parent), m(f.left), m(f.right)

    p1, l1, r1 = shift(f1, 0, 2 * N)
    p2, l2, r2 = shift(f2, N, 2 * N + (n1 - N))
    parent = np.concatenate([p1[:N], p2[:N], p1[N:], p2[N:]]).astype(np.int32)
    left = np.concatenate([l1[:N], l2[:N], l1[N:], l2[N:]]).astype(np.int32)
    right = np.concatenate([r1[:N], r2[:N], r1[N:], r2[N:]]).astype(np.int32)
    forest = forest_from_parent(parent, left, right)
    d = lognormal_branch_lengths(forest, seed=53)
    qptr, qval = random_priors(10, 2, 15, seed=54)
    c1 = simulate_codes(f1, lognormal_branch_lengths(f1, seed=55), qptr, qval, 0.2, 0.2, seed=56)
    codes = np.concatenate([c1, c1])                               # both trees see the same cells
    compare(forest, codes, qptr, qval, d, 0.2, 0.2)


def test_smallest_problems():
    """two leaves under one root, a single site; and a three-leaf tree with one site of one state"""
    for n_leaves in (2, 3):
        f = random_tree(n_leaves, seed=91)
        d = lognormal_branch_lengths(f, seed=92)
        qptr = np.array([0, 1], np.int32)
        qval = np.array([1.0])
        for codes in ([1, 0, -1], [-1, -1, -1], [1, 1, 1], [0, -2, 1]):
            c = np.array(codes[:n_leaves], np.int32).reshape(n_leaves, 1)
            compare(f, c, qptr, qval, d, 0.2, 0.1)


def test_usage_errors_are_loud():
    with pytest.raises(PMMError):
        Engine(4, np.array([0, 2], np.int32), np.array([0.5, 0.5]), np.array([[3], [0], [0], [0]], np.int32))   # state 3 of 2
    eng = Engine(4, np.array([0, 2], np.int32), np.array([0.5, 0.5]), np.zeros((4, 1), np.int32))
    with pytest.raises(PMMError):                                  # not binary / leaves not first
        eng.set_tree(np.array([4, 4, 4, 4, -1]), np.array([-1, -1, -1, -1, 0]), np.array([-1, -1, -1, -1, 1]))
    with pytest.raises(PMMError):
        eng.inside(np.ones(7), 0.1, 0.1)                           # no tree yet
    eng.close()
