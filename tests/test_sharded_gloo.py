"""World-size-2 `gloo` test of the site-sharded E-step / EM host logic
(laml_b200/dist.py) on CPU.  The CUDA engine cannot run here, so each rank
drives ShardedEngine with a stand-in that answers the device-resident subset of
the Engine API from the oracle; what is under test is the partitioning, the
all-reduce of [S | Rtilde | llh] and the replicated M-step loop."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


class OracleBackedEngine:
    """test double for laml_b200.capi.Engine (device-resident calls only)"""

    def __init__(self, forest, codes, qptr, qval, n_cells):
        from oracle.pmm_oracle import OracleProblem
        N = forest.n_leaves
        self.orc = OracleProblem(forest, np.arange(N, dtype=np.int32), codes, qptr, qval)
        self.n_nodes, self.device, self.K_total = forest.n_nodes, 0, codes.shape[1]
        self.stats = torch.zeros(5 * self.n_nodes + 2, dtype=torch.float64)
        self.n_cells = n_cells

    def set_total_sites(self, k):
        self.K_total = k

    def upload_params(self, d, nu, phi):
        self.d, self.nu, self.phi = np.array(d, np.float64), nu, phi

    def download_params(self):
        return self.d, self.nu, self.phi

    def estep_dev(self):
        out = self.orc.estep(self.d, self.nu, self.phi)
        self.stats[: 5 * self.n_nodes] = torch.from_numpy(out["S"].T.reshape(-1).copy())
        self.stats[5 * self.n_nodes] = out["Rtilde"]
        self.stats[5 * self.n_nodes + 1] = out["llh"]

    def inside_dev(self):
        out = self.orc.inside(self.d, self.nu, self.phi)
        self.stats[5 * self.n_nodes + 1] = out["llh"]

    def mstep_dev(self, R, n_cells, free_mask=None, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0,
                  eps_s=1e-6, eps_nu=1e-5):
        from oracle.pmm_oracle import lib
        n = self.n_nodes
        S = self.stats[: 5 * n].numpy().reshape(n, 5).T.copy()
        fm = np.ones(n, np.uint8)
        Sn = np.zeros_like(S)
        phi_star = lib().oracle_mstep_stats(n, self.K_total, n_cells, S, fm, R, float(self.stats[5 * n]), eps_s, Sn)
        d = self.d.copy()
        lib().oracle_mstep_brlen(n, Sn, fm, self.nu, eps_nu, dmin, dmax, d)
        self.nu = lib().oracle_mstep_nu(n, Sn, fm, d)
        self.d, self.phi = d, phi_star


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from laml_b200.dist import ShardedEngine, partition_sites
    from laml_b200.synth import make_workload
    w = make_workload(60, 18, 2, 12, 0.2, 0.15, seed=21)
    lo, hi = partition_sites(w.qptr, world)[rank]
    qptr = (w.qptr[lo:hi + 1] - w.qptr[lo]).astype(np.int32)
    qval = w.qval[w.qptr[lo]:w.qptr[hi]]
    eng = OracleBackedEngine(w.forest, w.codes[:, lo:hi], qptr, qval, 60)
    sh = ShardedEngine(eng, 18, stats_tensor=eng.stats)
    R = float((w.codes != -1).sum())
    est = sh.estep(w.d, w.nu, w.phi)
    em = sh.em(np.maximum(w.d, 0.01), 0.2, 0.15, R, 60, max_iter=6)
    one = sh.em_step(np.maximum(w.d, 0.01), 0.2, 0.15, R, 60)
    if rank == 0:
        ret["S"], ret["llh"], ret["Rt"] = est["S"].copy(), est["llh"], est["Rtilde"]
        ret["em"] = (em["nllh"], em["nu"], em["phi"], em["d"].copy(), em["iters"])
        ret["one"] = (one[0].copy(), one[1], one[2], one[3])
    dist.destroy_process_group()


def test_sharded_matches_unsharded():
    from laml_b200.synth import make_workload
    from oracle.pmm_oracle import OracleProblem
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    w = make_workload(60, 18, 2, 12, 0.2, 0.15, seed=21)
    orc = OracleProblem(w.forest, np.arange(60, dtype=np.int32), w.codes, w.qptr, w.qval)
    ref = orc.estep(w.d, w.nu, w.phi)
    np.testing.assert_allclose(ret["S"], ref["S"].T, rtol=1e-11, atol=1e-12)
    assert abs(ret["llh"] - ref["llh"]) < 1e-9 and abs(ret["Rt"] - ref["Rtilde"]) < 1e-10
    ref_em = orc.em(np.maximum(w.d, 0.01), 0.2, 0.15, max_iter=6)
    nllh, nu, phi, d, iters = ret["em"]
    assert iters == ref_em["iters"] or iters == 7            # loop exits at max_iter + 1 like the reference
    assert abs(nllh - ref_em["nllh"]) <= 1e-9 * ref_em["nllh"]
    np.testing.assert_allclose(d, ref_em["d"], rtol=1e-9, atol=1e-11)
    assert abs(nu - ref_em["nu"]) < 1e-10 and abs(phi - ref_em["phi"]) < 1e-12
    # one EM iteration, parameters in / parameters out: the first iteration of the oracle's loop
    d1, nu1, phi1, llh1 = ret["one"]
    d0 = np.maximum(w.d, 0.01)
    e0 = orc.estep(d0, 0.2, 0.15)
    assert abs(llh1 - e0["llh"]) <= 1e-9 * abs(e0["llh"])
    ref1 = orc.em(d0, 0.2, 0.15, max_iter=1)
    if ref1["iters"] <= 2:                                  # the oracle's loop also stops after one M-step (+ the check)
        np.testing.assert_allclose(d1, ref1["d"], rtol=1e-9, atol=1e-11)
        assert abs(nu1 - ref1["nu"]) < 1e-10 and abs(phi1 - ref1["phi"]) < 1e-12


def _ultra_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from laml_b200 import ultrametric as um
    from laml_b200.dist import ShardedEngine, partition_sites
    from laml_b200.synth import make_workload
    w = make_workload(48, 20, 2, 10, 0.15, 0.1, seed=33)
    lo, hi = partition_sites(w.qptr, world)[rank]
    qptr = (w.qptr[lo:hi + 1] - w.qptr[lo]).astype(np.int32)
    qval = w.qval[w.qptr[lo]:w.qptr[hi]]
    eng = OracleBackedEngine(w.forest, w.codes[:, lo:hi], qptr, qval, 48)
    sh = ShardedEngine(eng, 20, stats_tensor=eng.stats)
    n = w.forest.n_nodes
    prob = um.UltrametricProblem(w.forest.parent, np.zeros(n, np.int8), np.zeros(n), 0.005, 10.0)
    R = float((w.codes != -1).sum())
    fm = np.ones(n, np.uint8)
    em = sh.em_ultrametric(prob, np.maximum(w.d, 0.01), 0.15, 0.1, R, 48, fm, max_iter=8)
    ret[rank] = (em["nllh"], em["nu"], em["phi"], em["d"].copy(), em["iters"], em["status"])
    dist.destroy_process_group()


def test_sharded_ultrametric_em_matches_unsharded():
    """site-sharded EM under the ultrametric constraint: two ranks end on the same lengths / nu / phi as one rank
    holding all sites (the host M-step is a pure function of the all-reduced statistics), and the lengths are
    ultrametric"""
    mgr = mp.Manager()
    ret2, ret1 = mgr.dict(), mgr.dict()
    port = 29400 + os.getpid() % 90
    mp.spawn(_ultra_worker, args=(2, port, ret2), nprocs=2, join=True)
    mp.spawn(_ultra_worker, args=(1, port + 100, ret1), nprocs=1, join=True)
    for a, b in ((ret2[0], ret2[1]), (ret2[0], ret1[0])):
        assert a[4] == b[4] and a[5] == b[5]
        assert abs(a[0] - b[0]) <= 1e-9 * abs(b[0])
        np.testing.assert_allclose(a[3], b[3], rtol=1e-7, atol=1e-9)
        assert abs(a[1] - b[1]) < 1e-8 and abs(a[2] - b[2]) < 1e-12
    # identical on the two ranks of one run, to the bit: both ran the same M-step on the same reduced statistics
    np.testing.assert_array_equal(ret2[0][3], ret2[1][3])
    from laml_b200.synth import make_workload
    w = make_workload(48, 20, 2, 10, 0.15, 0.1, seed=33)
    d, parent = ret2[0][3], w.forest.parent
    depth = []
    for leaf in range(w.forest.n_leaves):
        x, t = leaf, 0.0
        while parent[x] >= 0:
            t += d[x]
            x = parent[x]
        depth.append(t)
    assert max(depth) - min(depth) < 1e-6


class _FakeNniEngine:
    """answers score_nni_batch with a deterministic function of (u, k) so the candidate sharding /
    all_gather / re-ordering logic can be checked on CPU"""
    device = 0

    def score_nni_batch(self, cu, ck, d, nu, phi, **kw):
        cu = np.asarray(cu); ck = np.asarray(ck)
        nllh = 1000.0 + cu * 2.0 + ck
        d5 = np.stack([cu * 0.001 + j for j in range(5)], axis=1)
        nodes = np.stack([cu + j for j in range(5)], axis=1).astype(np.int32)
        return {"nllh": nllh, "d5": d5, "iters": (cu % 7).astype(np.int32), "nodes": nodes}


def _nni_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from laml_b200.dist import score_nni_sharded
    cu = np.arange(40, 53, dtype=np.int32).repeat(2)[:-1]          # odd count: uneven shards
    ck = np.tile(np.array([0, 1], np.int32), 13)[:-1]
    out = score_nni_sharded(_FakeNniEngine(), cu, ck, None, 0.1, 0.1, device="cpu")
    ret[rank] = (out["nllh"].copy(), out["d5"].copy(), out["iters"].copy(), out["nodes"].copy())
    dist.destroy_process_group()


def test_nni_candidate_sharding():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29900 + os.getpid() % 90
    mp.spawn(_nni_worker, args=(2, port, ret), nprocs=2, join=True)
    cu = np.arange(40, 53, dtype=np.int32).repeat(2)[:-1]
    ck = np.tile(np.array([0, 1], np.int32), 13)[:-1]
    want = _FakeNniEngine().score_nni_batch(cu, ck, None, 0.1, 0.1)
    for r in (0, 1):
        nllh, d5, iters, nodes = ret[r]
        np.testing.assert_array_equal(nllh, want["nllh"]); np.testing.assert_array_equal(d5, want["d5"])
        np.testing.assert_array_equal(iters, want["iters"]); np.testing.assert_array_equal(nodes, want["nodes"])


def _search_worker(rank, world, port, ret, ultra=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from random import seed
    import laml_b200.solver as solver_mod
    from helpers import load_npz_case
    from laml_b200.topology import GPU_Topology_search
    from oracle_engine import OracleEngine
    solver_mod.Engine = OracleEngine                 # test stand-in for the GPU engine (tests/oracle_engine.py)
    case, z = load_npz_case("ref_rand_small")
    seed(a=21)                                       # the same policy stream on every rank
    ts = GPU_Topology_search(case.newicks, solver_mod.GPU_EM_solver, data={'charMtrx': case.msa}, prior={'Q': case.Q},
                             params={'phi': case.phi, 'nu': case.nu}, shard_candidates=world > 1, gather_device="cpu",
                             batch_ultra=ultra)
    strategy = dict(solver_mod.DEFAULT_STRATEGY)
    strategy['ultra_constr'] = ultra
    trees, score, params = ts.search(maxiter=4, nreps=1, strategy=strategy)
    ret[rank] = (trees, score, params, ts.trees_scored)
    dist.destroy_process_group()


def test_candidate_sharded_search_matches_single_process():
    """GPU_Topology_search with shard_candidates: two ranks, same seed, each scoring half of every neighbourhood,
    end on the same trees and score as one process scoring everything"""
    mgr = mp.Manager()
    ret2, ret1 = mgr.dict(), mgr.dict()
    port = 29700 + os.getpid() % 90
    mp.spawn(_search_worker, args=(2, port, ret2), nprocs=2, join=True)
    mp.spawn(_search_worker, args=(1, port + 100, ret1), nprocs=1, join=True)
    assert ret2[0][0] == ret2[1][0] == ret1[0][0]
    assert ret2[0][1] == ret2[1][1] == ret1[0][1]
    assert ret2[0][2] == ret1[0][2] and ret2[0][3] == ret1[0][3]


def test_candidate_sharded_ultrametric_search_matches_single_process():
    """the same under ``ultra_constr`` with ``batch_ultra``: neighbourhoods scored under the constraint, sharded
    ``rank::world``; NaN scores (candidates without room) travel through the all_gather like any other"""
    mgr = mp.Manager()
    ret2, ret1 = mgr.dict(), mgr.dict()
    port = 29500 + os.getpid() % 90
    mp.spawn(_search_worker, args=(2, port, ret2, True), nprocs=2, join=True)
    mp.spawn(_search_worker, args=(1, port + 100, ret1, True), nprocs=1, join=True)
    assert ret2[0][0] == ret2[1][0] == ret1[0][0]
    assert ret2[0][1] == ret2[1][1] == ret1[0][1]
    assert ret2[0][2] == ret1[0][2] and ret2[0][3] == ret1[0][3]


class _PeerFake:
    """what ShardedEngine needs of an engine to set up the exchange over peer memory, with scripted failures"""
    device, n_nodes = 0, 3

    def __init__(self, rank, fail_export=(), fail_setup=()):
        self.rank, self.fail_export, self.fail_setup = rank, fail_export, fail_setup
        self.calls = []

    def set_total_sites(self, k):
        pass

    def peer_export(self):
        self.calls.append("export")
        if self.rank in self.fail_export:
            raise RuntimeError("no exchange buffer")
        return bytes([self.rank + 1]) * 64

    def peer_setup(self, rank, world, handles):
        self.calls.append(("setup", rank, world, handles))
        if self.rank in self.fail_setup:
            raise RuntimeError("cannot map a peer")

    def peer_close(self):
        self.calls.append("close")


def _peer_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from laml_b200.dist import ShardedEngine
    out = {}
    for name, kw in (("ok", {}), ("export_fails_on_1", {"fail_export": (1,)}), ("setup_fails_on_0", {"fail_setup": (0,)})):
        eng = _PeerFake(rank, **kw)
        sh = ShardedEngine(eng, 10, stats_tensor=torch.zeros(17, dtype=torch.float64), peer=True)
        out[name] = (sh.peer, [c if isinstance(c, str) else c[0] for c in eng.calls],
                     [c[3] for c in eng.calls if not isinstance(c, str)])
        sh.close()
        out[name + "_after_close"] = [c if isinstance(c, str) else c[0] for c in eng.calls]
    os.environ.pop("PMM_PEER", None)
    os.environ["PMM_PEER"] = "0"
    eng = _PeerFake(rank)
    out["env_off"] = (ShardedEngine(eng, 10, stats_tensor=torch.zeros(17, dtype=torch.float64), peer=True).peer, list(eng.calls))
    ret[rank] = out
    dist.destroy_process_group()


def test_peer_setup_is_all_or_none():
    """ShardedEngine._setup_peers: handles are all-gathered in rank order and every rank maps them; if ANY rank cannot
    export or map, EVERY rank drops the peers and keeps the NCCL all-reduce (no rank may exchange alone)"""
    port = 29800 + os.getpid() % 150
    ret = mp.Manager().dict()
    mp.spawn(_peer_worker, args=(2, port, ret), nprocs=2, join=True)
    want_handles = bytes([1]) * 64 + bytes([2]) * 64
    for rank in (0, 1):
        r = ret[rank]
        assert r["ok"][0] is True and r["ok"][1] == ["export", "setup"] and r["ok"][2] == [want_handles]
        assert r["ok_after_close"] == ["export", "setup", "close"]
        assert r["export_fails_on_1"][0] is False and "setup" not in r["export_fails_on_1"][1] and "close" in r["export_fails_on_1"][1]
        assert r["setup_fails_on_0"][0] is False and r["setup_fails_on_0"][1].count("close") == 1
        assert r["setup_fails_on_0_after_close"].count("close") == 1            # close() of a ShardedEngine without peers does nothing
        assert r["env_off"] == (False, [])
