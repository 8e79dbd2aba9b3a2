"""Site-sharded E-step / EM across GPUs: one process per GPU, NCCL over NVLink.

Sites are conditionally independent given the tree and the parameters (every
loop of the reference is ``for site in range(numsites)``, ``EM_solver.py:101,177,263``),
so the path shards on the site axis with ONE exchange per E-step: the sum over
ranks of ``[S(n,5) | Rtilde | llh]`` -- (5n+2)*8 bytes, 8 MB at 100k cells.  On
GPUs that can map each other's memory the engine's own finishing kernel does it
(``laml_b200/csrc/pmm_peer.cuh``: count reduction, reduce-scatter and all-gather
over NVLink peer memory in one cooperative launch; handles exchanged once through
``torch.distributed``); otherwise it is an NCCL ``all_reduce(sum, fp64)`` on the
engine's statistics buffer (wrapped, not copied).  The M-step is deterministic
given the reduced statistics and is replicated on every rank (no second
exchange).  SURVEY.md 8(e).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def partition_sites(qptr: np.ndarray, world_size: int, balance: str = "sites") -> List[Tuple[int, int]]:
    """Contiguous site ranges of (nearly) equal cost.  ``balance="sites"``: equal site counts -- the
    lumped-state engine's work and bytes are per (node, site), independent of the alphabet;
    ``balance="slots"``: equal numbers of state slots ``sum(A_s + 2)`` -- equal bytes streamed per GPU
    by the full-state row engine (``PMM_ENGINE=fullstate``)."""
    K = len(qptr) - 1
    if K < world_size:
        raise ValueError("cannot shard %d sites over %d ranks" % (K, world_size))
    if balance == "sites":
        cost = np.ones(K, np.int64)
    elif balance == "slots":
        cost = np.diff(np.asarray(qptr, np.int64)) + 2
    else:
        raise ValueError("balance must be 'sites' or 'slots'")
    cum = np.concatenate([[0], np.cumsum(cost)])
    bounds = [0]
    for r in range(1, world_size):
        s = int(np.argmin(np.abs(cum - cum[-1] * r / world_size)))
        s = min(max(s, bounds[-1] + 1), K - (world_size - r))      # every rank keeps at least one site
        bounds.append(s)
    bounds.append(K)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


class _DeviceArray:
    """Expose a raw device pointer through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class ShardedEngine:
    """An :class:`laml_b200.capi.Engine` on this rank's site shard plus the
    collective that combines the shards.  ``engine`` only needs the
    device-resident subset of the Engine API (``upload_params``, ``estep_dev``,
    ``inside_dev``, ``mstep_dev``, ``download_params``, ``stats`` access), which
    is what lets the world_size-2 gloo tests drive this class on CPU."""

    def __init__(self, engine, n_sites_total: int, group=None, stats_tensor=None, peer=None):
        import torch.distributed as dist
        self.eng = engine
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.K_total = int(n_sites_total)
        engine.set_total_sites(self.K_total)
        self.stream = None
        if stats_tensor is None:
            import torch
            torch.cuda.set_device(engine.device)
            # kernels and the collective are ordered on ONE stream: the engine adopts a torch
            # stream, and the all-reduce is issued with that stream current
            self.stream = torch.cuda.Stream(device=engine.device)
            engine.set_stream(self.stream.cuda_stream)
            ptr, n = engine.stats_ptr()
            stats_tensor = torch.as_tensor(_DeviceArray(ptr, n), device="cuda:%d" % engine.device)
        self.stats = stats_tensor
        # the exchange over peer memory (laml_b200/csrc/pmm_peer.cuh): the engine's finishing kernel reduces the
        # statistics over the ranks itself, through NVLink loads / stores on the peers' buffers -- no NCCL call on the
        # path.  Set up when every rank can map every other rank's buffer; otherwise the NCCL all-reduce stays.
        self.peer = False
        if self.world > 1 and (peer if peer is not None else self.stream is not None):
            self.peer = self._setup_peers()
        self._host_stats = None
        self._llh_pin = None
        self.n_nodes = engine.n_nodes

    def _setup_peers(self) -> bool:
        import os
        import torch
        if os.environ.get("PMM_PEER", "1") == "0" or not hasattr(self.eng, "peer_export"):
            return False
        dist = self.dist
        dev = "cuda:%d" % self.eng.device if self.stream is not None else "cpu"      # ("cpu": the gloo tests of this protocol)
        rank = dist.get_rank(self.group)
        try:
            mine, ok = self.eng.peer_export(), 1.0
        except Exception:                                           # noqa: BLE001  (e.g. PMM_ENGINE=fullstate)
            mine, ok = bytes(64), 0.0
        local = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        parts = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather(parts, local, group=self.group)
        handles = b"".join(bytes(p.cpu().tolist()) for p in parts)
        flag = torch.tensor([ok], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if float(flag.item()) > 0.5:
            try:
                self.eng.peer_setup(rank, self.world, handles)
            except Exception:                                       # noqa: BLE001
                ok = 0.0
        flag = torch.tensor([ok], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if float(flag.item()) < 0.5:                                # all or none: a rank that cannot map a peer sends everybody to NCCL
            self.eng.peer_close()
            dist.barrier(group=self.group)
            return False
        dist.barrier(group=self.group)                              # nobody starts exchanging before everybody has cleared its flags
        return True

    def close(self):
        """unmap the peers on every rank before any rank frees its buffers"""
        if self.peer:
            self.eng.peer_close()
            self.dist.barrier(group=self.group)
            self.peer = False

    def _on_stream(self):
        import contextlib
        if self.stream is None:
            return contextlib.nullcontext()
        import torch
        return torch.cuda.stream(self.stream)

    def _allreduce(self):
        if self.world > 1 and not self.peer:                       # with peers the E-step's finishing kernel has done it
            with self._on_stream():
                self.dist.all_reduce(self.stats, op=self.dist.ReduceOp.SUM, group=self.group)

    def estep_dev(self):
        """E-step on the shard, then the all-reduce; statistics stay on the device."""
        self.eng.estep_dev()
        self._allreduce()

    def inside_dev(self):
        self.eng.inside_dev()
        self._allreduce()

    def estep(self, d, nu, phi):
        """Host-buffer E-step (the e2e path): upload parameters, E-step + all-reduce,
        download ``S``, ``Rtilde``, ``llh``."""
        self.eng.upload_params(d, nu, phi)
        self.estep_dev()
        if self.stream is None:                       # CPU stand-in (gloo tests)
            st = np.array(self.stats.cpu().numpy(), copy=True)
        else:
            import torch
            if self._host_stats is None:              # pinned staging buffer, reused: the returned S is a view of it
                self._host_stats = torch.empty(self.stats.shape, dtype=self.stats.dtype).pin_memory()
            with self._on_stream():
                self._host_stats.copy_(self.stats, non_blocking=True)
            self.stream.synchronize()
            st = self._host_stats.numpy()
        n5 = 5 * self.n_nodes
        return {"S": st[:n5].reshape(self.n_nodes, 5), "Rtilde": float(st[n5]), "llh": float(st[n5 + 1])}

    def llh(self) -> float:
        with self._on_stream():
            return float(self.stats[5 * self.n_nodes + 1].item())

    def em_step(self, d, nu, phi, R, n_cells, free_mask=None, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0, out=None):
        """ONE EM iteration from host parameters to host parameters (the body of the loop at EM_solver.py:483-499:
        ``Estep`` then ``Mstep``) with the statistics left on the device: parameters up, E-step + all-reduce,
        M-step, parameters down.  Returns ``(d, nu, phi, llh)``; ``llh`` is the log-likelihood AT THE INPUT point
        (what the E-step's inside pass evaluates)."""
        if (self.world == 1 or self.peer) and hasattr(self.eng, "em_step"):
            kw = {} if out is None else {"out": out}         # (a page-locked ``out`` is read and written in place by the copy engines)
            return self.eng.em_step(d, nu, phi, R, n_cells, free_mask, optimize_nu, optimize_phi, dmin, dmax, **kw)   # pmm_em_step
        self.eng.upload_params(d, nu, phi)
        self.estep_dev()
        if self.stream is not None:
            # the log-likelihood rides along with the parameters: queued on the stream before the M-step (which does not
            # touch that slot), read after the one synchronisation of the step
            if self._llh_pin is None:
                import torch
                self._llh_pin = torch.empty(1, dtype=self.stats.dtype).pin_memory()
            with self._on_stream():
                self._llh_pin.copy_(self.stats[5 * self.n_nodes + 1:5 * self.n_nodes + 2], non_blocking=True)
        self.eng.mstep_dev(R, n_cells, free_mask, optimize_nu, optimize_phi, dmin, dmax)
        d_out, nu_out, phi_out = self.eng.download_params()        # the one synchronisation of the step
        llh = float(self._llh_pin[0]) if self.stream is not None else self.llh()
        return d_out, nu_out, phi_out, llh

    def em(self, d, nu, phi, R, n_cells, free_mask=None, optimize_nu=True, optimize_phi=True, dmin=0.005, dmax=10.0,
           conv_eps=1e-8, max_iter=1000):
        """``EM_solver.EM_optimization`` (EM_solver.py:461-512, ultra_constr=False) on
        site shards: every rank runs the same loop; one all-reduce per E-step."""
        if self.peer:
            # the exchange is part of the engine's E-step: the device-resident loop of the library (pmm_em) runs as on one
            # GPU, every rank reading the same log-likelihood bits and taking the same decisions
            return self.eng.em(d, nu, phi, R, n_cells, free_mask, optimize_nu, optimize_phi, dmin, dmax, conv_eps, max_iter)
        self.eng.upload_params(d, nu, phi)
        self.estep_dev()
        pre = cur = self.llh()
        it = 1
        while it <= max_iter:
            self.eng.mstep_dev(R, n_cells, free_mask, optimize_nu, optimize_phi, dmin, dmax)   # an unchanged mask is not re-uploaded
            self.estep_dev()
            cur = self.llh()
            if abs((cur - pre) / pre) < conv_eps:
                break
            pre = cur
            it += 1
        d_out, nu_out, phi_out = self.eng.download_params()
        return {"nllh": -cur, "d": d_out, "nu": nu_out, "phi": phi_out, "iters": it, "status": "optimal"}


    def em_ultrametric(self, problem, d, nu, phi, R, n_cells, free_mask, optimize_nu=True, optimize_phi=True,
                       conv_eps=1e-8, max_iter=1000, eps_nu=1e-5, eps_s=1e-6):
        """``EM_solver.EM_optimization`` with ``ultra_constr=True`` (EM_solver.py:461-512) on site shards.
        The branch-length program under the ultrametric equalities runs on the host
        (``laml_b200.ultrametric.mstep``), a pure function of the all-reduced statistics, so every rank
        computes the same lengths and no second exchange is needed: per iteration one E-step +
        all-reduce, one replicated host M-step, one inside pass + all-reduce for the convergence test
        (the same engine calls as ``GPU_EM_solver.EM_optimization``).  ``problem`` is the
        topology-dependent :class:`laml_b200.ultrametric.UltrametricProblem`."""
        from . import ultrametric as um
        fm = np.asarray(free_mask, np.uint8)
        free = fm.astype(bool)
        d = np.array(d, np.float64, copy=True)
        if hasattr(self.eng, "mstep_ultra_dev") and self.eng.info()["inside_levels"] <= 1024:
            # the constrained M-step on the device (pmm_mstep_ultra_dev: one cooperative launch, fixed-order reductions --
            # the same bits on every rank from the same all-reduced statistics): parameters and statistics never leave
            # the GPUs, one all-reduce and one log-likelihood read per iteration
            self.eng.upload_params(d, nu, phi)
            self.estep_dev()
            pre = cur = self.llh()
            it, status = 1, "optimal"
            while it <= max_iter:
                status_d, _ = self.eng.mstep_ultra_dev(R, n_cells, problem.kind, problem.fixed_len, optimize_nu, optimize_phi,
                                                       problem.lo, problem.hi, eps_s, eps_nu, warm=it > 1)
                if status_d == "infeasible":
                    d, nu, phi = self.eng.download_params()
                    return {"nllh": -pre, "d": d, "nu": nu, "phi": phi, "iters": it, "status": "d_infeasible"}
                status = "optimal" if status_d == "optimal" else "failed_d"
                self.estep_dev()
                cur = self.llh()
                if abs((cur - pre) / pre) < conv_eps:
                    break
                pre = cur
                it += 1
            d, nu, phi = self.eng.download_params()
            return {"nllh": -cur, "d": d, "nu": nu, "phi": phi, "iters": it, "status": status, "converged": it <= max_iter}
        self.eng.upload_params(d, nu, phi)
        self.inside_dev()
        pre = cur = self.llh()
        it, H, status, converged = 1, None, "optimal", False
        while it <= max_iter:
            est = self.estep(d, nu, phi)
            d_new, nu, phi, H_new, status_d = um.mstep(problem, est["S"], est["Rtilde"], R, n_cells, self.K_total, fm, d,
                                                       nu, phi, H_prev=H, optimize_phi=optimize_phi,
                                                       optimize_nu=optimize_nu, eps_nu=eps_nu, eps_s=eps_s)
            if status_d == "infeasible":                          # EM_solver.py:425-426, :497-499
                return {"nllh": -pre, "d": d, "nu": nu, "phi": phi, "iters": it, "status": "d_infeasible"}
            H = H_new
            status = "optimal" if status_d == "optimal" else "failed_d"
            d = np.where(free, d_new, d)
            self.eng.upload_params(d, nu, phi)
            self.inside_dev()
            cur = self.llh()
            if abs((cur - pre) / pre) < conv_eps:
                converged = True
                break
            pre = cur
            it += 1
        return {"nllh": -cur, "d": d, "nu": nu, "phi": phi, "iters": it, "status": status, "converged": converged}


def score_nni_sharded(engine, cand_u, cand_k, d, nu, phi, group=None, device=None, **kw):
    """Candidate-sharded NNI scoring (BASELINE config 5): candidate trees are independent units, so
    every rank -- each holding ALL sites of the data set -- scores the slice ``rank::world`` with
    ``Engine.score_nni_batch`` and the per-candidate results are combined with one ``all_gather``
    (2 + 5 fp64 and 6 int32 per candidate).  Returns the full result on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    cand_u, cand_k = np.asarray(cand_u, np.int32), np.asarray(cand_k, np.int32)
    m = len(cand_u)
    mine = np.arange(rank, m, world)
    if len(mine):
        res = engine.score_nni_batch(cand_u[mine], cand_k[mine], d, nu, phi, **kw)
    else:
        res = {"nllh": np.empty(0), "d5": np.empty((0, 5)), "iters": np.empty(0, np.int32), "nodes": np.empty((0, 5), np.int32)}
    if world == 1:
        return res
    per = (m + world - 1) // world
    pack = np.zeros((per, 12))
    k = len(mine)
    pack[:k, 0] = res["nllh"]; pack[:k, 1:6] = res["d5"]; pack[:k, 6] = res["iters"]; pack[:k, 7:12] = res["nodes"]
    dev = device if device is not None else ("cuda:%d" % engine.device if torch.cuda.is_available() else "cpu")
    local = torch.as_tensor(pack, device=dev)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local, group=group)
    full = np.zeros((m, 12))
    for r, p in enumerate(parts):
        idx = np.arange(r, m, world)
        full[idx] = p.cpu().numpy()[: len(idx)]
    return {"nllh": full[:, 0], "d5": full[:, 1:6], "iters": full[:, 6].astype(np.int32), "nodes": full[:, 7:12].astype(np.int32)}
