"""Worker of tests/test_zz_gpu_peer.py: the exchange over peer memory with the ranks as handles of ONE process on ONE GPU.

    python tests/peer_local_worker.py WORLD N_LEAVES N_SITES        exit 0 = all checks passed, 3 = an exchange gave up waiting

The ranks' finishing kernels wait for each other ON the device, so they -- and the E-step kernels of the ranks that are
still computing -- must really run concurrently.  Between GPUs (one process per GPU, the production layout) nothing can
prevent that; on one GPU the streams of all the handles share the device's hardware work queues
(CUDA_DEVICE_MAX_CONNECTIONS, 8 by default), and a kernel that lands behind a waiting kernel in the same queue cannot
start: the wait then gives up (PMM_PEER_TIMEOUT_MS; a failed call instead of a hung GPU).  The test therefore runs this
worker with 32 queues and one site slice (one busy stream per handle) and repeats it if it reports exactly that."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class ExchangeTimeout(Exception):
    pass


class _DeviceArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def _stats(eng):
    import torch
    ptr, n = eng.stats_ptr()
    return torch.as_tensor(_DeviceArray(ptr, n), device="cuda:%d" % eng.device).cpu().numpy().copy()


def _shards(w, world):
    from laml_b200.capi import Engine
    from laml_b200.dist import partition_sites
    K = w.codes.shape[1]
    engs = []
    for lo, hi in partition_sites(w.qptr, world):
        qp = (w.qptr[lo:hi + 1] - w.qptr[lo]).astype(np.int32)
        qv = w.qval[w.qptr[lo]:w.qptr[hi]].copy()
        e = Engine(w.forest.n_leaves, qp, qv, np.ascontiguousarray(w.codes[:, lo:hi]))
        e.set_tree(w.forest.parent, w.forest.left, w.forest.right)
        e.set_total_sites(K)
        engs.append(e)
    return engs


def on_every_rank(engs, fn):
    """``fn(engine)`` for all ranks at once, one host thread per rank as in the production layout (one process per rank):
    a call that synchronises with the device inside -- a rank that walks its sites in several chunks does -- must not keep
    the other ranks from being launched (ctypes releases the GIL for the duration of the C call)"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(engs)) as pool:
        for f in [pool.submit(fn, e) for e in engs]:
            f.result()


def run_case(world, n, K):
    from laml_b200.capi import Engine
    from laml_b200.synth import make_workload
    w = make_workload(n, K, 1, 30, 0.25, 0.2, seed=n + world)
    N, nn = w.forest.n_leaves, w.forest.n_nodes
    d0 = np.maximum(w.d, 0.01)
    R = float((w.codes != -1).sum())
    whole = Engine(N, w.qptr, w.qval, w.codes)
    whole.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    ref = whole.estep(d0, w.nu, w.phi)
    engs = _shards(w, world)
    # every rank alone first: its own statistics
    own = [e.estep(d0, w.nu, w.phi) for e in engs]
    for r, e in enumerate(engs):
        e.peer_setup_local(r, engs)
    for e in engs:
        e.upload_params(d0, w.nu, w.phi)
    for rep in range(3):                           # flags and barrier counters carry over from call to call
        on_every_rank(engs, lambda e: e.estep_dev())    # the ranks' kernels wait for each other on the device
        on_every_rank(engs, lambda e: e.sync())
        st = [_stats(e) for e in engs]
        if any(np.isnan(s[5 * nn + 1]) for s in st):
            raise ExchangeTimeout("E-step %d" % rep)
        for s in st[1:]:
            assert s.tobytes() == st[0].tobytes()                                   # the same bits on every rank
        S = st[0][:5 * nn].reshape(nn, 5)
        want = own[0]["S"].copy()
        for o in own[1:]:
            want = want + o["S"]                                                    # rank order, as the kernel adds
        llh_want, rt_want = own[0]["llh"], own[0]["Rtilde"]
        for o in own[1:]:
            llh_want, rt_want = llh_want + o["llh"], rt_want + o["Rtilde"]
        assert np.isfinite(st[0][5 * nn + 1])
        assert st[0][5 * nn + 1] == llh_want, ("llh", rep, st[0][5 * nn + 1], llh_want)
        assert st[0][5 * nn] == rt_want, ("Rtilde", rep, st[0][5 * nn], rt_want)
        if S.tobytes() != want.tobytes():
            bad = np.argwhere(S != want)
            raise AssertionError(("S", rep, len(bad), bad[:5].tolist(), S[S != want][:5].tolist(), want[S != want][:5].tolist()))
        np.testing.assert_allclose(S, ref["S"], rtol=1e-10, atol=1e-12 * K)
        assert abs(st[0][5 * nn + 1] - ref["llh"]) <= 1e-12 * abs(ref["llh"])
        assert abs(st[0][5 * nn] - ref["Rtilde"]) <= 1e-12 * abs(ref["Rtilde"])
    # inside pass only: just the likelihood is exchanged, the counts stay
    on_every_rank(engs, lambda e: e.inside_dev())
    on_every_rank(engs, lambda e: e.sync())
    st2 = [_stats(e) for e in engs]
    if any(np.isnan(s[5 * nn + 1]) for s in st2):
        raise ExchangeTimeout("inside pass")
    for s in st2:
        assert s.tobytes() == st[0].tobytes()
    # EM iterations on the shards (E-step with exchange, replicated M-step) against the single engine
    ref_d, ref_nu, ref_phi = d0, w.nu, w.phi
    for e in engs:
        e.upload_params(d0, w.nu, w.phi)
    for it in range(4):
        est = whole.estep(ref_d, ref_nu, ref_phi)
        ref_d, ref_nu, ref_phi = whole.mstep(est["S"], est["Rtilde"], R, N, ref_d, ref_nu, ref_phi)
        on_every_rank(engs, lambda e: (e.estep_dev(), e.mstep_dev(R, N)))
    outs = [e.download_params() for e in engs]
    if any(np.isnan(_stats(e)[5 * nn + 1]) for e in engs):
        raise ExchangeTimeout("EM iterations")
    for d, nu, phi in outs:
        assert d.tobytes() == outs[0][0].tobytes() and nu == outs[0][1] and phi == outs[0][2]
    np.testing.assert_allclose(outs[0][0], ref_d, rtol=1e-9, atol=1e-12)
    assert abs(outs[0][1] - ref_nu) <= 1e-9 and abs(outs[0][2] - ref_phi) <= 1e-9
    # candidate scoring needs all sites on one handle: refused on a handle with peers (candidates shard instead)
    from laml_b200.capi import PMMError
    from laml_b200.nni import nni_candidates
    cu, ck = nni_candidates(w.forest.parent, w.forest.left, N)
    try:
        engs[0].score_nni_batch(cu[:4], ck[:4], d0, w.nu, w.phi)
        raise AssertionError("score_nni_batch accepted a handle with peers")
    except PMMError as exc:
        assert "shard the candidates" in str(exc)
    # dropping the peers gives back a single-GPU handle
    for e in engs:
        e.peer_close()
    again = engs[0].estep(d0, w.nu, w.phi)
    assert again["S"].tobytes() == own[0]["S"].tobytes() and again["llh"] == own[0]["llh"]
    for e in engs:
        e.close()
    whole.close()


if __name__ == "__main__":
    from laml_b200.capi import PMMError

    def guarded(*a):
        """a wait that gave up surfaces as an error at the next synchronisation (pmm_sync, pmm_download_params, ...)"""
        try:
            run_case(*a)
        except PMMError as exc:
            if "gave up waiting" in str(exc):
                raise ExchangeTimeout(str(exc)[:80])
            raise

    if len(sys.argv) > 4:                                # repetitions: how often does a wait give up on this box?
        import time
        gave_up = 0
        for rep in range(int(sys.argv[4])):
            t0 = time.perf_counter()
            try:
                guarded(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
            except ExchangeTimeout as exc:
                gave_up += 1
                print("run %d: exchange gave up waiting (%s) after %.2f s" % (rep, exc, time.perf_counter() - t0), flush=True)
        print("%d of %s runs gave up" % (gave_up, sys.argv[4]), flush=True)
        sys.exit(0)
    try:
        guarded(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
    except ExchangeTimeout as exc:
        print("exchange gave up waiting:", exc, flush=True)
        sys.exit(3)
    print("ok", flush=True)
