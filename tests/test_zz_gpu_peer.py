"""The exchange over peer memory (laml_b200/csrc/pmm_peer.cuh: count reduction + reduce-scatter + all-gather in the
E-step's finishing kernel) on ONE GPU: the ranks are handles of this process (``pmm_peer_setup_local``), each with its
own stream and a site shard, so that the kernels of all ranks are in flight together and synchronise through their
flags exactly as ranks on different GPUs do.  Checked: every rank ends with the same bits; they equal the sum of the
ranks' own statistics taken in rank order; they agree with one engine holding all sites; the EM loop on shards walks
the single engine's iterations.  (Two GPUs, two processes, CUDA IPC: tests/test_zz_gpu_nccl.py.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("world,n,K", [(2, 700, 300), (3, 90, 50), (4, 2000, 520), (8, 300, 64)])
def test_ranks_in_one_process(world, n, K):
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
        for e in engs:
            e.estep_dev()                           # asynchronous: the ranks' kernels wait for each other on the device
        for e in engs:
            e.sync()
        st = [_stats(e) for e in engs]
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
        assert S.tobytes() == want.tobytes() and st[0][5 * nn + 1] == llh_want and st[0][5 * nn] == rt_want
        np.testing.assert_allclose(S, ref["S"], rtol=1e-10, atol=1e-12 * K)
        assert abs(st[0][5 * nn + 1] - ref["llh"]) <= 1e-12 * abs(ref["llh"])
        assert abs(st[0][5 * nn] - ref["Rtilde"]) <= 1e-12 * abs(ref["Rtilde"])
    # inside pass only: just the likelihood is exchanged, the counts stay
    for e in engs:
        e.inside_dev()
    for e in engs:
        e.sync()
    st2 = [_stats(e) for e in engs]
    for s in st2:
        assert s.tobytes() == st[0].tobytes()
    # EM iterations on the shards (E-step with exchange, replicated M-step) against the single engine
    ref_d, ref_nu, ref_phi = d0, w.nu, w.phi
    for e in engs:
        e.upload_params(d0, w.nu, w.phi)
    for it in range(4):
        est = whole.estep(ref_d, ref_nu, ref_phi)
        ref_d, ref_nu, ref_phi = whole.mstep(est["S"], est["Rtilde"], R, N, ref_d, ref_nu, ref_phi)
        for e in engs:
            e.estep_dev()
            e.mstep_dev(R, N)
    outs = [e.download_params() for e in engs]
    for d, nu, phi in outs:
        assert d.tobytes() == outs[0][0].tobytes() and nu == outs[0][1] and phi == outs[0][2]
    np.testing.assert_allclose(outs[0][0], ref_d, rtol=1e-9, atol=1e-12)
    assert abs(outs[0][1] - ref_nu) <= 1e-9 and abs(outs[0][2] - ref_phi) <= 1e-9
    # candidate scoring needs all sites on one handle: refused on a handle with peers (candidates shard instead)
    from laml_b200.capi import PMMError
    from laml_b200.nni import nni_candidates
    cu, ck = nni_candidates(w.forest.parent, w.forest.left, N)
    with pytest.raises(PMMError, match="shard the candidates"):
        engs[0].score_nni_batch(cu[:4], ck[:4], d0, w.nu, w.phi)
    # dropping the peers gives back a single-GPU handle
    for e in engs:
        e.peer_close()
    again = engs[0].estep(d0, w.nu, w.phi)
    assert again["S"].tobytes() == own[0]["S"].tobytes() and again["llh"] == own[0]["llh"]
    for e in engs:
        e.close()
    whole.close()
