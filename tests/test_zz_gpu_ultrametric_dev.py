"""SURVEY 8 row f3 ON THE GPU: the ultrametric-constrained M-step as one cooperative launch (``pmm_mstep_ultra_dev``,
laml_b200/csrc/pmm_ultrametric_dev.cuh) and the device-resident EM around it (``pmm_em_ultra``), against the library's
host solver (``pmm_ultrametric_solve``: pinned to the reference's own program on SLSQP and to MOSEK's published optima),
on real grids (many CTAs, grid barriers, wide and narrow level phases).  The kernel's logic also runs on the CPU under
the CUDA shim (tests/test_ultrametric_dev_emulated.py)."""
import numpy as np
import pytest

from helpers import load_npz_case, make_engine
from laml_b200 import ultrametric as um
from laml_b200.capi import Engine
from laml_b200.synth import make_workload

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def objective(est, K, fm, d, nu, eps_s=1e-6):
    """the M-step's objective EM_solver.py:351-357 at lengths d (statistics clamped and rescaled as :343-346)"""
    s = np.maximum(est["S"], eps_s)
    s = s / s.sum(axis=1, keepdims=True) * K
    free = fm.astype(bool)
    x = d[free]; s = s[free]
    val = -((nu + 1.0) * s[:, 0] + nu * (s[:, 1] + s[:, 3])) * x + s[:, 1] * np.log(-np.expm1(-x))
    if nu > 1e-5:
        val = val + (s[:, 2] + s[:, 4]) * np.log(-np.expm1(-nu * x))
    return float(val.sum())


def host_mstep(problem, est, R, n_cells, K, fm, d, nu, phi, H_prev=None):
    return um.mstep(problem, est["S"], est["Rtilde"], R, n_cells, K, fm, d, nu, phi, H_prev=H_prev)


@pytest.mark.parametrize("n,K,seed,nu,n_fixed", [(60, 14, 1, 0.2, 0), (400, 30, 2, 0.0, 0), (3000, 24, 3, 0.3, 0),
                                                 (20000, 16, 4, 0.15, 0), (1500, 20, 5, 0.25, 40)])
def test_device_mstep_equals_the_host_solver(n, K, seed, nu, n_fixed):
    w = make_workload(n, K, 2, 20, max(nu, 0.1), 0.2, seed=seed)
    f = w.forest
    nn = f.n_nodes
    d0 = np.maximum(w.d, 0.01)
    kind = np.zeros(nn, np.int8); fixed_len = np.zeros(nn)
    if n_fixed:
        # internal branches (a fixed leaf branch pins its parent's height: infeasible unless the sibling agrees)
        inner = np.nonzero((f.left >= 0) & (f.parent >= 0))[0]
        fx = np.random.default_rng(seed).choice(inner, n_fixed, replace=False)
        kind[fx] = um.FIXED; fixed_len[fx] = 0.03
        d0[fx] = 0.03
    fm = (kind == um.FREE).astype(np.uint8)
    problem = um.UltrametricProblem(f.parent, kind, fixed_len, 0.005, 10.0)
    R = float((w.codes != -1).sum())
    eng = Engine(n, w.qptr, w.qval, w.codes)
    eng.set_tree(f.parent, f.left, f.right)
    est = eng.estep(d0, nu, 0.2)
    d_ref, nu_ref, phi_ref, H_ref, st_ref = host_mstep(problem, est, R, n, K, fm, d0, nu, 0.2)
    eng.upload_params(d0, nu, 0.2)
    eng.estep_dev()
    st, iters = eng.mstep_ultra_dev(R, n, kind, fixed_len, warm=False)
    d_dev, nu_dev, phi_dev = eng.download_params()
    assert st == st_ref == "optimal" and iters > 0
    # lengths to the solvers' own stopping precision (both stop on the Newton decrement at barrier weight 1e-11; the optimum
    # is flat to ~1e-7 in single lengths), the objective -- what the program maximises -- to 1e-9 relative (the duality gap of
    # the barrier at weight 1e-11 and the line search's resolution are ~1e-10 of it)
    np.testing.assert_allclose(d_dev, d_ref, rtol=0, atol=2e-7)
    assert abs(objective(est, K, fm, d_dev, nu) - objective(est, K, fm, d_ref, nu)) <= 1e-9 * abs(objective(est, K, fm, d_ref, nu))
    assert abs(nu_dev - nu_ref) <= 1e-8 and abs(phi_dev - phi_ref) <= 1e-12
    assert (d_dev[kind == um.FIXED] == d0[kind == um.FIXED]).all()
    # second iteration, warm-started on both sides from the first one's heights
    est2 = eng.estep(d_ref, nu_ref, phi_ref)
    d_ref2, nu_ref2, phi_ref2, _, st_ref2 = host_mstep(problem, est2, R, n, K, fm, d_ref, nu_ref, phi_ref, H_prev=H_ref)
    eng.upload_params(d_ref, nu_ref, phi_ref)
    eng.estep_dev()
    st2, iters2 = eng.mstep_ultra_dev(R, n, kind, fixed_len, warm=True)
    d_dev2, nu_dev2, _ = eng.download_params()
    assert st2 == st_ref2 == "optimal" and iters2 > 0
    np.testing.assert_allclose(d_dev2, d_ref2, rtol=0, atol=2e-7)
    assert abs(objective(est2, K, fm, d_dev2, nu_ref) - objective(est2, K, fm, d_ref2, nu_ref)) <= 1e-9 * abs(objective(est2, K, fm, d_ref2, nu_ref))
    assert abs(nu_dev2 - nu_ref2) <= 1e-8
    eng.close()


def test_infeasible_problem_updates_nothing():
    w = make_workload(50, 10, 2, 10, 0.2, 0.2, seed=7)
    f = w.forest
    nn = f.n_nodes
    cherry = next(u for u in range(50, nn) if f.left[u] < 50 and f.right[u] < 50)
    kind = np.zeros(nn, np.int8); fixed_len = np.zeros(nn)
    kind[[f.left[cherry], f.right[cherry]]] = um.FIXED
    fixed_len[f.left[cherry]], fixed_len[f.right[cherry]] = 0.02, 0.05
    d0 = np.maximum(w.d, 0.01)
    eng = Engine(50, w.qptr, w.qval, w.codes)
    eng.set_tree(f.parent, f.left, f.right)
    eng.upload_params(d0, 0.2, 0.2)
    eng.estep_dev()
    st, iters = eng.mstep_ultra_dev(float((w.codes != -1).sum()), 50, kind, fixed_len)
    d1, nu1, phi1 = eng.download_params()
    assert st == "infeasible" and iters == 0 and np.array_equal(d1, d0) and nu1 == 0.2 and phi1 == 0.2
    res = eng.em_ultra(d0, 0.2, 0.2, float((w.codes != -1).sum()), 50, kind, fixed_len)
    assert res["status"] == "d_infeasible" and res["iters"] == 1 and np.array_equal(res["d"], d0)
    eng.close()


@pytest.mark.parametrize("name", ["ref_example1_start", "ref_rand_forest", "ref_rand_ragged"])
def test_device_em_equals_the_host_loop(name, monkeypatch):
    """GPU_EM_solver.EM_optimization(ultra_constr=True): device-resident (pmm_em_ultra) vs the host-driven loop on the host
    solver -- same iteration count, likelihood, parameters; the trees that come back are ultrametric"""
    from laml_b200.solver import GPU_EM_solver
    from laml_b200.topology import is_ultrametric
    case, z = load_npz_case(name)
    out = []
    for host in ("1", "0"):
        monkeypatch.setenv("PMM_ULTRA_HOST", host)
        s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0.2, 'nu': 0.15})
        s._ensure()
        for t in s.trees:
            for nd in t.traverse_postorder():
                nd.edge_length = max(0.01, nd.edge_length if nd.edge_length is not None else 0.01)
        nllh, iters, status = s.EM_optimization(verbose=-1, ultra_constr=True)
        out.append((nllh, iters, status, s._d(), s.params.nu, s.params.phi, is_ultrametric(s._forest, s._d(), tol=1e-7)))
        s.close()
    host_r, dev_r = out
    assert host_r[2] == dev_r[2] == "optimal" and host_r[6] and dev_r[6]
    assert abs(host_r[1] - dev_r[1]) <= 1                                    # the two solvers stop within their tolerances
    assert abs(host_r[0] - dev_r[0]) <= 1e-8 * abs(host_r[0])
    np.testing.assert_allclose(dev_r[3], host_r[3], rtol=0, atol=2e-5)
    assert abs(host_r[4] - dev_r[4]) <= 1e-5 and abs(host_r[5] - dev_r[5]) <= 1e-5


def test_verbose_ultra_em_prints_progress_from_the_device_loop(capsys):
    from laml_b200.solver import GPU_EM_solver
    case, z = load_npz_case("ref_rand_small")
    res = []
    for verbose in (-1, 1):
        s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0.2, 'nu': 0.15})
        s._ensure()
        for t in s.trees:
            for nd in t.traverse_postorder():
                nd.edge_length = max(0.01, nd.edge_length)
        res.append(s.EM_optimization(verbose=verbose, ultra_constr=True, maxIter=40) + (s._d(),))
        s.close()
    text = capsys.readouterr().out
    assert res[0][1] == res[1][1] and abs(res[0][0] - res[1][0]) <= 1e-10 * abs(res[0][0])
    np.testing.assert_allclose(res[0][3], res[1][3], rtol=0, atol=1e-9)
    assert text.count("Finished EM iter: ") == min(res[1][1], 40) and "Initial phi: 0.2. Initial nu: 0.15." in text
