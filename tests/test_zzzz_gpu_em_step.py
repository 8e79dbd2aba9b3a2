"""``dist.ShardedEngine.em_step`` (one EM iteration, parameters in / parameters out, statistics on the device) on the
GPU at world size 1, against the split protocol on the same engine (``pmm_estep`` -> host S -> ``pmm_mstep``: the
GPU-verified path of tests/test_gpu_parity.py) and against the oracle.

Its host logic is also covered by the world-size-2 gloo test (tests/test_sharded_gloo.py).  First GPU run on
record: the driver's round-1 pass (GPUTEST_r01.json: passed)."""
import numpy as np
import pytest

from laml_b200.capi import Engine
from laml_b200.synth import make_workload

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("seed,n,K,amax,nu,phi", [(3, 300, 24, 40, 0.23, 0.17), (5, 90, 70, 6, 0.0, 0.3)])
def test_em_step_matches_the_split_protocol_and_the_oracle(seed, n, K, amax, nu, phi):
    from laml_b200.dist import ShardedEngine
    from oracle.pmm_oracle import OracleProblem
    w = make_workload(n_leaves=n, n_sites=K, a_lo=1, a_hi=amax, nu=max(nu, 0.1), phi=max(phi, 0.1), seed=seed)
    f = w.forest
    eng = Engine(n, w.qptr, w.qval, w.codes, device=0)
    eng.set_tree(f.parent, f.left, f.right)
    R = float((w.codes != -1).sum())
    d0 = np.maximum(w.d, 0.01)
    est = eng.estep(d0, nu, phi)
    d_split, nu_split, phi_split = eng.mstep(est["S"], est["Rtilde"], R, n, d0, nu, phi)
    sh = ShardedEngine(eng, K)
    cur = (d0, nu, phi)
    for _ in range(2):                                           # twice: the second call starts from downloaded parameters
        first = cur
        cur = sh.em_step(cur[0], cur[1], cur[2], R, n)
        if first[0] is d0:
            d1, nu1, phi1, llh1 = cur
            assert abs(llh1 - est["llh"]) <= 1e-12 * abs(est["llh"])
            np.testing.assert_allclose(d1, d_split, rtol=1e-12, atol=0)
            assert abs(nu1 - nu_split) <= 1e-12 and abs(phi1 - phi_split) <= 1e-12
        cur = cur[:3]
    orc = OracleProblem(f, np.arange(n, dtype=np.int32), w.codes, w.qptr, w.qval)
    ref = orc.em(d0, nu, phi, max_iter=2)                        # the oracle's loop stops after max_iter + 1 E-steps: 2 M-steps
    if ref["iters"] == 3:
        np.testing.assert_allclose(cur[0], ref["d"], rtol=1e-5, atol=1e-7)
        assert abs(cur[1] - ref["nu"]) <= 1e-5 and abs(cur[2] - ref["phi"]) <= 1e-5
    eng.close()


def test_verbose_em_runs_on_em_step_and_equals_the_device_loop(capsys):
    """``GPU_EM_solver.EM_optimization(verbose=1)`` prints the reference's progress lines (EM_solver.py:468-505) from one
    ``pmm_em_step`` per iteration; it must stop at the same iteration, on the same parameters, as the silent run
    (``pmm_em``, everything on the device)"""
    from helpers import load_npz_case
    from laml_b200.solver import GPU_EM_solver
    case, z = load_npz_case("ref_rand_ragged")
    out = []
    for verbose in (-1, 1):
        s = GPU_EM_solver(case.newicks, {'charMtrx': case.msa}, {'Q': case.Q}, {'phi': 0.2, 'nu': 0.15})
        s._ensure()
        for t in s.trees:
            for nd in t.traverse_postorder():
                nd.edge_length = max(0.01, nd.edge_length)
        out.append(s.EM_optimization(verbose=verbose) + (s._d(), s.params.nu, s.params.phi))
        s.close()
    text = capsys.readouterr().out
    quiet, loud = out
    assert quiet[1] == loud[1] and quiet[2] == loud[2] == "optimal"
    assert abs(quiet[0] - loud[0]) <= 1e-12 * abs(quiet[0])
    np.testing.assert_allclose(loud[3], quiet[3], rtol=1e-12, atol=0)
    assert abs(quiet[4] - loud[4]) <= 1e-14 and abs(quiet[5] - loud[5]) <= 1e-14
    assert "Initial phi: 0.2. Initial nu: 0.15. Initial nllh: " in text
    assert text.count("Finished EM iter: ") == loud[1] and ("Finished EM iter: %d. Current nllh: " % loud[1]) in text


def test_em_step_in_place_on_page_locked_memory():
    """``em_step(out=...)``: a page-locked array is read and written in place by the copy engines (no staging copies);
    a pageable one goes through the staging buffers.  Both give the bits of the default call; bad lengths are refused
    on either path."""
    import torch
    from laml_b200.capi import PMMError
    w = make_workload(n_leaves=400, n_sites=50, a_lo=1, a_hi=20, nu=0.2, phi=0.2, seed=9)
    f = w.forest
    eng = Engine(400, w.qptr, w.qval, w.codes, device=0)
    eng.set_tree(f.parent, f.left, f.right)
    R = float((w.codes != -1).sum())
    d0 = np.maximum(w.d, 0.01)
    want = eng.em_step(d0, 0.2, 0.2, R, 400)
    want2 = eng.em_step(want[0], want[1], want[2], R, 400)
    pinned = torch.from_numpy(d0.copy()).pin_memory().numpy()
    pageable = d0.copy()
    for buf in (pinned, pageable):
        a = eng.em_step(buf, 0.2, 0.2, R, 400, out=buf)
        assert a[0] is buf and a[0].tobytes() == want[0].tobytes() and a[1:] == want[1:]
        b = eng.em_step(buf, a[1], a[2], R, 400, out=buf)                # twice: the output of a step is the next one's input
        assert b[0].tobytes() == want2[0].tobytes() and b[1:] == want2[1:]
        buf[3] = np.nan
        with pytest.raises(PMMError):
            eng.em_step(buf, 0.2, 0.2, R, 400, out=buf)
    other = torch.empty(len(d0), dtype=torch.float64).pin_memory().numpy()
    c = eng.em_step(d0, 0.2, 0.2, R, 400, out=other)                     # out of place into a page-locked array
    assert c[0] is other and other.tobytes() == want[0].tobytes() and (d0 == np.maximum(w.d, 0.01)).all()
    eng.close()
