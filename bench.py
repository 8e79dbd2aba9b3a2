#!/usr/bin/env python
"""Benchmark of the PMM EM hot path (BASELINE.json metric) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads
  config4_100k_x_2k (default), config3_10k_x_500, config2_1k_x_30, config5_5k_x_300
      A *step* is one EM iteration on the whole synthetic workload: E-step (inside pass, outside pass + expected
      counts), the all-reduce of the statistics when N > 1, and the M-step.  Under torchrun every rank holds a
      contiguous site shard of the SAME workload (strong scaling); rank 0 prints ONE JSON line.
  config3_em      BASELINE config 3: EM to convergence (EM_solver.EM_optimization) on site shards; reports iterations,
                  EM iterations/s and the converged nllh / nu / phi (equal across N); ``--oracle-em`` adds the oracle's EM.
  config5_nni     BASELINE config 5: one NNI neighbourhood of the 5k x 300 tree (all 2(N-2) candidates), candidate
                  trees sharded across the ranks (``dist.score_nni_sharded``); reports candidates/s per scoring mode.

``--impl reference`` times the reference's own algorithm on the host cores: the reference is pure Python and cannot
travel to the GPU box, so this arm runs its C restatement (oracle/, pinned against the reference's golden vectors)
with every host thread this process may use, on a bounded site sample of the same workload.  It loads nothing of the
GPU engine.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PEAKS_FILE = os.path.join(ROOT, "MEASURED_PEAKS.json")
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback
EM_WORKLOADS = ("config4_100k_x_2k", "config3_10k_x_500", "config2_1k_x_30", "config5_5k_x_300")


def hbm_peak():
    try:
        return float(json.load(open(PEAKS_FILE))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """the cores this process may run on -- NOT ``OMP_NUM_THREADS`` (torchrun exports 1 for every rank)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def time_em_step(sh, d, nu, phi, R, n_cells, steps, warmup, barrier, all_ok=lambda ok: ok, in_place=False):
    """seconds per ``ShardedEngine.em_step`` (host parameters in, host parameters out) over ``steps`` chained calls,
    or the ``repr`` of the exception that stopped it.  ``all_ok(ok)`` is a collective AND over the ranks, taken after
    the warm-up and after the timed loop, so that a rank that failed cannot leave the others waiting in a barrier:
    every rank reaches both exchanges whatever happened to it, and nobody enters ``barrier()`` unless all are well.
    ``in_place``: the new lengths come back in the array that held the old ones (``d``, page-locked in the bench)."""
    err = None
    cur = (d, nu, phi)
    kw = (lambda c: {"out": c[0]}) if in_place else (lambda c: {})
    try:
        for _ in range(max(1, warmup // 2)):
            cur = sh.em_step(cur[0], cur[1], cur[2], R, n_cells, **kw(cur))[:3]
    except Exception as exc:                                    # noqa: BLE001
        err = repr(exc)
    if not all_ok(err is None):
        return err or "failed on another rank"
    barrier()
    t0 = time.perf_counter()
    try:
        for _ in range(steps):
            cur = sh.em_step(cur[0], cur[1], cur[2], R, n_cells, **kw(cur))[:3]
    except Exception as exc:                                    # noqa: BLE001
        err = repr(exc)
    if not all_ok(err is None):
        return err or "failed on another rank"
    barrier()
    return (time.perf_counter() - t0) / steps


def workload_spec(name):
    from laml_b200.synth import CONFIGS
    if name not in CONFIGS:
        raise SystemExit("unknown workload %r; choose from %s" % (name, sorted(CONFIGS) + ["config3_em", "config5_nni"]))
    return dict(CONFIGS[name])


def build_inputs(spec, sites, device):
    """tree, branch lengths and priors of the whole workload + simulated codes of ``sites``"""
    from laml_b200.synth import lognormal_branch_lengths, random_priors, random_tree, simulate_codes
    seed = 1
    forest = random_tree(spec["n_leaves"], seed=seed)
    d = lognormal_branch_lengths(forest, seed=seed + 1)
    qptr, qval = random_priors(spec["n_sites"], spec["a_lo"], spec["a_hi"], seed=seed + 4,
                               uniform=spec.get("uniform_prior", False))
    codes = simulate_codes(forest, d, qptr, qval, spec["nu"], spec["phi"], seed=seed + 6, device=device, sites=sites)
    return forest, d, qptr, qval, codes


def shard_prior(qptr, qval, lo, hi):
    return (qptr[lo:hi + 1] - qptr[lo]).astype(np.int32), qval[qptr[lo]:qptr[hi]].copy()


def cpu_reference_step(orc, d, nu, phi, sites, nthreads, with_mstep=True):
    """one EM iteration of the reference's algorithm (oracle port) on a site sample: ``(seconds, E-step result)``"""
    t0 = time.perf_counter()
    est = orc.estep(d, nu, phi, sites=sites, nthreads=nthreads)
    if with_mstep:
        orc.mstep(est["S"], est["Rtilde"], d, nu)
    return time.perf_counter() - t0, est


def config_dict(name, spec, qptr, gpus):
    """the workload, and nothing but the workload: both arms print exactly this"""
    W = int(qptr[-1]) + 2 * (len(qptr) - 1)
    return {"workload": name, "cells": spec["n_leaves"], "sites": spec["n_sites"],
            "alphabet": "%d..%d states/site" % (spec["a_lo"], spec["a_hi"]), "nu": spec["nu"], "phi": spec["phi"],
            "state_slots_per_node_W": W, "nodes": 2 * spec["n_leaves"] - 1,
            "parallelism": "site-shard x%d" % gpus}


def run_reference(args, rank, world):
    """the CPU arm: oracle only -- neither the GPU library nor ``__graft_entry__.build`` is touched"""
    if rank != 0:
        return
    if args.workload not in EM_WORKLOADS:
        print(json.dumps({"impl": "reference", "unavailable": "the reference arm covers the EM-iteration workloads %s" %
                          (EM_WORKLOADS,)}), flush=True)
        return
    from oracle import pmm_oracle
    pmm_oracle.build()
    from oracle.pmm_oracle import OracleProblem
    spec = workload_spec(args.workload)
    threads = host_threads()
    n_sample = min(spec["n_sites"], args.cpu_sample_sites or max(8, 2 * threads))
    forest, d, qptr, qval, codes = build_inputs(spec, (0, n_sample), "cpu")
    q2, v2 = shard_prior(qptr, qval, 0, n_sample)
    N = forest.n_leaves
    orc = OracleProblem(forest, np.arange(N, dtype=np.int32), codes, q2, v2)
    for _ in range(args.warmup):
        cpu_reference_step(orc, d, spec["nu"], spec["phi"], (0, n_sample), threads)
    t = [cpu_reference_step(orc, d, spec["nu"], spec["phi"], (0, n_sample), threads)[0] for _ in range(args.steps)]
    dt = float(np.sum(t)) / args.steps
    value = N * n_sample / dt
    sample = "%d of %d sites, full %d-cell tree, %d threads; cell*site/s scales linearly in sites" % (
        n_sample, spec["n_sites"], N, threads)
    line = {
        "impl": "reference", "metric": "em_iteration_cell_site_per_s", "value": value, "unit": "cell*site/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload, spec, qptr, args.gpus),
        "cpu_baseline": {"value": value, "unit": "cell*site/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell*site/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "em_iter_per_s_full_workload": value / (N * spec["n_sites"]),
        "note": "reference is pure Python (cannot travel); this arm is its C restatement oracle/pmm_oracle.c, "
                "pinned to the reference's golden vectors (tests/test_oracle_golden.py); the Python reference itself "
                "(EM_solver.Estep, single-threaded) runs 5.6e4 cell*site/s on one core at 1k x 30, 52x slower than one "
                "thread of this port (measured in the build container, DESIGN.md section 7)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Ranks:
    """torch.distributed plumbing shared by the GPU workloads"""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the PMM engine has no CPU path")
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = rank, local_rank, world
        torch.cuda.set_device(local_rank)
        self.dev = "cuda:%d" % local_rank
        if world > 1:
            # the one collective of the path is an 8 MB fp64 all-reduce (5 n + 2 doubles): on 8 B200s over NVSwitch ring + LL128
            # takes 54.5 us, NCCL's own choice (NVLS) 68 us (profiles/r2b_nccl_variants_n8.log); an explicit setting wins
            os.environ.setdefault("NCCL_ALGO", "Ring")
            os.environ.setdefault("NCCL_PROTO", "LL128")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return float(t.item())

    def all_ok(self, ok: bool) -> bool:
        t = self.torch.tensor([1.0 if ok else 0.0], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def grab_stdout():
    """keep stdout for the ONE JSON line: anything a library prints there (e.g. "NCCL version ...") goes to stderr"""
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def stats_checksum(stats_host, n):
    """what two runs of the same workload must agree on, whatever the number of ranks: the all-reduced E-step
    statistics [S(n,5) | Rtilde | llh] boiled down to a few numbers (summation order differs between shardings, so
    agreement is to ~1e-12 relative, not to the bit)"""
    S = stats_host[:5 * n]
    w = np.cos(np.arange(5 * n, dtype=np.float64) * 0.001)          # position-dependent weights: a permuted S changes it
    return {"llh": float(stats_host[5 * n + 1]), "Rtilde": float(stats_host[5 * n]), "S_sum": float(S.sum()),
            "S_l2": float(np.sqrt((S * S).sum())), "S_weighted": float((S * w).sum())}


def run_ours(args, rank, local_rank, world):
    real_stdout = grab_stdout()
    rk = Ranks(rank, local_rank, world)
    torch, dist = rk.torch, rk.dist
    from laml_b200.capi import Engine
    from laml_b200.dist import ShardedEngine, partition_sites
    spec = workload_spec(args.workload)
    Ktot, N = spec["n_sites"], spec["n_leaves"]
    # priors are cheap: build them first to partition by state slots
    from laml_b200.synth import random_priors
    qptr_all, _ = random_priors(Ktot, spec["a_lo"], spec["a_hi"], seed=5, uniform=spec.get("uniform_prior", False))
    lo, hi = partition_sites(qptr_all, world)[rank]
    forest, d, qptr, qval, codes = build_inputs(spec, (lo, hi), rk.dev)
    assert (qptr == qptr_all).all()
    q2, v2 = shard_prior(qptr, qval, lo, hi)
    n = forest.n_nodes
    eng = Engine(N, q2, v2, codes, device=local_rank)
    eng.set_tree(forest.parent, forest.left, forest.right)
    sh = ShardedEngine(eng, Ktot)
    info = eng.info()
    # R = number of non-'?' leaf entries (EM_solver.py:322) over ALL sites
    R = rk.sum(float((codes != -1).sum()))
    nu0, phi0 = spec["nu"], spec["phi"]
    barrier = rk.barrier

    def reset():
        eng.upload_params(d, nu0, phi0)

    def step_device():
        sh.estep_dev()
        eng.mstep_dev(R, N)

    def read_stats():
        with torch.cuda.stream(sh.stream):
            st = sh.stats.cpu().numpy()
        return st

    # ---- the all-reduced statistics at the workload's own parameters (a pure E-step; comparable across N) --------
    reset()
    sh.estep_dev()
    check_start = stats_checksum(read_stats(), n)

    # ---- device-resident throughput ("value") --------------------------------------
    reset()
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = eng.info()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(sh.stream):
        ev0.record()
        for _ in range(args.steps):
            step_device()                   # enqueued back to back: no host synchronisation inside the timed region
        ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = eng.info()["kernel_launches"] - launches0
    ms_step = rk.max(ev0.elapsed_time(ev1)) / args.steps
    # state after warmup + steps EM iterations from the same start: equal across N for equal (W, K)
    d_end, nu_end, phi_end = eng.download_params()
    check_end = {"em_iterations": args.warmup + args.steps, "llh_at_last_estep": float(read_stats()[5 * n + 1]),
                 "nu": nu_end, "phi": phi_end, "d_sum": float(d_end.sum())}

    # phase breakdown (untimed extra steps; CUDA events on the engine's stream, one sync per step)
    phase = np.zeros(5)
    nb = max(3, min(10, args.steps))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    eng.set_option("slices", 1)             # passes back to back on one stream: inside / outside can be timed apart
    barrier()
    for it in range(nb + 1):
        with torch.cuda.stream(sh.stream):
            eng.estep_dev()
            evs[0].record()
            sh._allreduce()
            evs[1].record()
            eng.mstep_dev(R, N)
            evs[2].record()
        evs[2].synchronize()
        tm = eng.last_timing()
        if it > 0:                          # the first pass absorbs the ranks' skew
            phase += (tm["inside_ms"], tm["outside_ms"], tm["total_ms"], evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]))
    eng.set_option("slices", int(os.environ.get("PMM_SLICES", "2")))
    ph = torch.tensor(phase / nb, dtype=torch.float64, device=rk.dev)
    if world > 1:
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    inside_ms, outside_ms, estep_ms, allreduce_ms, mstep_ms = [float(x) for x in ph.tolist()]

    # ---- end to end ("e2e"): one EM iteration through the host-buffer call a user of the engine makes per iteration,
    # ShardedEngine.em_step (what GPU_EM_solver.EM_optimization runs when it reports progress): d / nu / phi go up from
    # host memory, d / nu / phi / llh come back to host memory, every step; the statistics stay on the device.
    d_pin = torch.from_numpy(d.copy()).pin_memory().numpy()
    em_step = time_em_step(sh, d_pin, nu0, phi0, R, N, args.steps, args.warmup, barrier, rk.all_ok, in_place=True)
    if not isinstance(em_step, str):
        em_step = rk.max(em_step)

    # ---- the reference's split protocol beside it ("e2e_split"): Estep() hands S to the host, Mstep() takes it back
    S_pin = torch.empty((n, 5), dtype=torch.float64).pin_memory().numpy()

    def step_split(dcur, nu, phi):
        if world == 1:
            est = eng.estep(dcur, nu, phi, out_S=S_pin)
        else:
            est = sh.estep(dcur, nu, phi)
        return eng.mstep(est["S"], est["Rtilde"], R, N, dcur, nu, phi)

    cur = (d_pin, nu0, phi0)
    for _ in range(max(1, args.warmup // 2)):
        cur = step_split(*cur)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cur = step_split(*cur)
    barrier()
    split_s = rk.max((time.perf_counter() - t0) / args.steps)
    split_h2d = 8 * n + 16 + 8 * (5 * n + 1) + 8 * n + 16        # estep: d,nu,phi; mstep: S,Rtilde,d,nu,phi
    split_d2h = 8 * (5 * n + 2) + 8 * n + 16                      # estep: S,Rtilde,llh; mstep: d,nu,phi

    # ---- parity on the benchmarked workload itself (N = 1): the oracle on a site sample of THIS tree against the
    # engine on the same sites -- and the CPU baseline timed on the same sample
    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        from oracle.pmm_oracle import OracleProblem
        threads = host_threads()
        n_sample = min(Ktot, args.cpu_sample_sites or max(8, 2 * threads))
        qs, vs = shard_prior(qptr, qval, 0, n_sample)
        orc = OracleProblem(forest, np.arange(N, dtype=np.int32), codes[:, :n_sample], qs, vs)
        _, ref = cpu_reference_step(orc, d, nu0, phi0, (0, n_sample), threads)
        reps, t_acc = 0, 0.0
        while reps < 3 or (t_acc < 10.0 and reps < 50):
            t_acc += cpu_reference_step(orc, d, nu0, phi0, (0, n_sample), threads)[0]
            reps += 1
        cpu = {"value": N * n_sample / (t_acc / reps), "unit": "cell*site/s", "cores": threads, "kind": "port",
               "sample": "%d of %d sites, full %d-cell tree, %d reps, %d threads (oracle/pmm_oracle.c: the "
                         "reference's algorithm in C)" % (n_sample, Ktot, N, reps, threads)}
        whole = eng.estep(d, nu0, phi0, want_site_llh=True)
        small = Engine(N, qs, vs, np.ascontiguousarray(codes[:, :n_sample]), device=local_rank)
        small.set_tree(forest.parent, forest.left, forest.right)
        got = small.estep(d, nu0, phi0, want_site_llh=True)
        small.close()
        Sref = ref["S"].T
        tolS = 1e-8 * np.abs(Sref) + 1e-10 * n_sample
        parity = {
            "what": "engine vs oracle (the reference's algorithm) on sites [0, %d) of this workload's tree" % n_sample,
            "llh_rel": abs(got["llh"] - ref["llh"]) / abs(ref["llh"]),
            "S_max_err_over_tol": float((np.abs(got["S"] - Sref) / tolS).max()),
            "S_max_rel_where_S_gt_1e-3": float((np.abs(got["S"] - Sref) / np.maximum(np.abs(Sref), 1e-300))[np.abs(Sref) > 1e-3].max()),
            "Rtilde_rel": abs(got["Rtilde"] - ref["Rtilde"]) / max(1.0, abs(ref["Rtilde"])),
            "site_llh_max_rel_whole_run_vs_oracle": float(np.max(np.abs(whole["site_llh"][:n_sample] - ref["site_llh"][:n_sample]) /
                                                                 np.abs(ref["site_llh"][:n_sample]))),
            "tolerances": {"llh_rel": 1e-9, "S": "1e-8 rel + 1e-10 K", "Rtilde_rel": 1e-8},
        }
        parity["ok"] = bool(parity["llh_rel"] <= 1e-9 and parity["S_max_err_over_tol"] <= 1.0 and parity["Rtilde_rel"] <= 1e-8
                            and parity["site_llh_max_rel_whole_run_vs_oracle"] <= 1e-9)

    exchange = ("peer memory: count reduction + reduce-scatter + all-gather over NVLink inside the E-step's finishing kernel "
                "(pmm_finish_peer; estep_ms includes it, allreduce_ms is what is left: nothing)" if sh.peer else
                ("NCCL all_reduce on the engine's statistics buffer" if world > 1 else "none (one GPU)"))
    sh.close()
    if rank != 0:
        rk.close()
        return

    # ---- roofline of the dominant kernel (the outside pass) ----------------------------
    peak, peak_src = hbm_peak()
    W_shard_max = info["slots_per_node"]
    n_roots = int((forest.parent < 0).sum())
    # 20 bytes per (child node, state slot) for outside+counts, 16 for inside
    out_bytes = 20.0 * W_shard_max * (n - n_roots)
    in_bytes = 16.0 * W_shard_max * n
    n_int = n - N
    lumped = info["engine"] == 0
    if lumped:
        # lumped-state engine: 32 / 40-byte (node, site) records written once and read where
        # they are not chained in registers, 4-byte leaf codes, 40 B of counts per (branch, 32 sites);
        # the engine derives the exact per-site byte counts from its clade schedule
        K_shard = hi - lo
        out_moved = float(info["outside_bytes_per_site"]) * K_shard
        in_moved = float(info["inside_bytes_per_site"]) * K_shard
        out_launches = info["launches_per_pass"] * info["n_chunks"]
        k_out, k_in = "pmm_lumped_outside", "pmm_lumped_inside"
        note = ("rank 0's site shard; moved bytes = what the lumped-state engine must move (32 / 40-byte records per "
                "(internal node, site), leaf codes, count partials); the SURVEY 8(d) figure charges 8 B per state slot "
                "of every node and is shown beside it (SURVEY 8d: 'a compressed implementation ... must still report "
                "against this figure'); traffic = ncu dram bytes per launch, profiles/")
    else:
        # full-state rows (PMM_ENGINE=fullstate): leaf rows are never materialised
        Wrow = info["padded_slots_per_node"]     # slots per row incl. even-length padding, summed over site chunks
        out_moved = 8.0 * Wrow * (n_int + 2 * (n_int - n_roots))      # read out[u]; read x[c], write out[c] for internal children
        in_moved = 8.0 * Wrow * (n_int + (n_int - n_roots))           # write x[u]; read x[c] for internal children
        out_launches = info["outside_levels"] * info["n_chunks"]     # at least; big levels take one launch per node class
        k_out, k_in = "pmm_level<outside>", "pmm_level<inside>"
        note = ("rank 0's site shard; moved bytes = internal-node rows only (leaf vectors are closed forms of "
                "4-byte codes and are never stored); traffic: see profiles/ for the ncu dram__bytes of one launch")
    # ncu `dram__bytes_read.sum + dram__bytes_write.sum` per launch of the dominant kernel, measured once per
    # build on this workload (profiles/*_traffic.json, written by scripts/summarize_traffic.py)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s/n%d/%s" % (args.workload, world, "lumped" if lumped else "fullstate")
        traffic = tj.get(key, {}).get(k_out)
        if traffic is not None:             # profiled launches cover 1/slices of the sites; `achieved` is per whole-tier launch
            traffic *= tj[key].get("_slices", 1)
    except Exception:
        pass
    # `achieved` counts the bytes this design must move: a true HBM utilisation.  The SURVEY 8(d) per-unit
    # figure (16 B inside / 20 B outside per node*slot, every vector stored) is reported next to it.
    ach = out_moved / (outside_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": k_out, "achieved": ach, "peak": peak, "unit": "GB/s",
        "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
        "bytes_per_step_moved": out_moved, "launches_per_step": out_launches,
        "avg_launch_ms": outside_ms / max(out_launches, 1), "ms_per_step": outside_ms,
        "survey_figure": {"bytes_per_step": out_bytes, "achieved": out_bytes / (outside_ms * 1e-3) / 1e9,
                          "frac": out_bytes / (outside_ms * 1e-3) / 1e9 / peak},
        "inside": {"kernel": k_in, "achieved": in_moved / (inside_ms * 1e-3) / 1e9,
                   "frac": in_moved / (inside_ms * 1e-3) / 1e9 / peak, "ms_per_step": inside_ms,
                   "bytes_per_step_moved": in_moved,
                   "survey_figure": {"bytes_per_step": in_bytes, "achieved": in_bytes / (inside_ms * 1e-3) / 1e9,
                                     "frac": in_bytes / (inside_ms * 1e-3) / 1e9 / peak}},
        "estep": {"ms_per_step": estep_ms, "frac": (in_moved + out_moved) / (estep_ms * 1e-3) / 1e9 / peak,
                  "survey_figure_frac": (in_bytes + out_bytes) / (estep_ms * 1e-3) / 1e9 / peak},
        "note": note,
    }

    value = N * Ktot / (ms_step * 1e-3)
    if isinstance(em_step, str):
        e2e = {"value": None, "unit": "cell*site/s", "error": em_step, "h2d_bytes_per_step": 8 * n + 16, "d2h_bytes_per_step": 8 * n + 24}
    else:
        e2e = {"value": N * Ktot / em_step, "unit": "cell*site/s", "h2d_bytes_per_step": 8 * n + 16,
               "d2h_bytes_per_step": 8 * n + 24, "ms_per_step": em_step * 1e3,
               "call": "ShardedEngine.em_step(out=d): page-locked host d/nu/phi in, d/nu/phi/llh out to the same host memory every step; "
                       "statistics stay on the device"}
    line = {
        "metric": "em_iteration_cell_site_per_s", "value": value, "unit": "cell*site/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload, spec, qptr, world),
        "engine": {"kind": "lumped-state records" if lumped else "full-state rows",
                   "l2": ("per-node state far exceeds L2: %.1f GB per pass per GPU" %
                          ((40e-9 * (hi - lo) if lumped else 8e-9 * info["row_stride"]) * n_int)),
                   "site_chunks": info["n_chunks"], "clades": info["clades"], "launches_per_pass": info["launches_per_pass"],
                   "exchange": exchange},
        "em_iter_per_s": 1e3 / ms_step, "estep_ms": estep_ms, "inside_ms": inside_ms, "outside_ms": outside_ms,
        "allreduce_ms": allreduce_ms, "mstep_ms": mstep_ms,
        "estep_cell_site_per_s": N * Ktot / (estep_ms * 1e-3),
        "node_site_per_s": n * Ktot / (estep_ms * 1e-3),
        "e2e": e2e,
        "e2e_split": {"value": N * Ktot / split_s, "unit": "cell*site/s", "h2d_bytes_per_step": split_h2d,
                      "d2h_bytes_per_step": split_d2h, "ms_per_step": split_s * 1e3,
                      "call": "the reference's split protocol: pmm_estep hands S(n,5) to the host, pmm_mstep takes it back"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "checksum": {"estep_at_start": check_start, "after_steps": check_end,
                     "note": "the all-reduced statistics of one E-step at the workload's parameters, and the state after "
                             "warmup + steps EM iterations: lines of the same workload agree to ~1e-12 whatever n_gpus"},
        "device_bytes": info["device_bytes"],
    }
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
    rk.close()


# ------------------------------------------------------------------------------------------------ config 3: EM to convergence
def run_em_convergence(args, rank, local_rank, world):
    """BASELINE config 3 (10k cells x 500 sites): ``EM_optimization`` to convergence on site shards.  One *step* is
    one whole EM run from the reference's starting point (every length raised to 2 dmin, the simulation's nu / phi)."""
    real_stdout = grab_stdout()
    rk = Ranks(rank, local_rank, world)
    torch = rk.torch
    from laml_b200.capi import Engine
    from laml_b200.dist import ShardedEngine, partition_sites
    from laml_b200.synth import random_priors
    name = "config3_10k_x_500"
    spec = workload_spec(name)
    Ktot, N = spec["n_sites"], spec["n_leaves"]
    qptr_all, _ = random_priors(Ktot, spec["a_lo"], spec["a_hi"], seed=5)
    lo, hi = partition_sites(qptr_all, world)[rank]
    forest, d, qptr, qval, codes = build_inputs(spec, (lo, hi), rk.dev)
    q2, v2 = shard_prior(qptr, qval, lo, hi)
    eng = Engine(N, q2, v2, codes, device=local_rank)
    eng.set_tree(forest.parent, forest.left, forest.right)
    sh = ShardedEngine(eng, Ktot)
    R = rk.sum(float((codes != -1).sum()))
    d0 = np.maximum(d, 0.01)
    nu0, phi0 = spec["nu"], spec["phi"]
    res = None
    for _ in range(max(1, args.warmup)):
        res = sh.em(d0, nu0, phi0, R, N, max_iter=args.max_iter)
    rk.barrier()
    t0 = time.perf_counter()
    runs = max(1, args.steps)
    for _ in range(runs):
        res = sh.em(d0, nu0, phi0, R, N, max_iter=args.max_iter)
    rk.barrier()
    dt = rk.max((time.perf_counter() - t0) / runs)
    oracle = None
    if world == 1 and args.oracle_em:
        from oracle.pmm_oracle import OracleProblem
        threads = host_threads()
        orc = OracleProblem(forest, np.arange(N, dtype=np.int32), codes, q2, v2)
        t1 = time.perf_counter()
        ref = orc.em(d0, nu0, phi0, max_iter=args.max_iter, nthreads=threads)
        t_ref = time.perf_counter() - t1
        oracle = {"iters": ref["iters"], "nllh": ref["nllh"], "nu": ref["nu"], "phi": ref["phi"], "seconds": t_ref, "threads": threads,
                  "nllh_rel": abs(res["nllh"] - ref["nllh"]) / abs(ref["nllh"]),
                  "d_max_abs": float(np.max(np.abs(res["d"] - ref["d"]))),
                  "d_max_rel": float(np.max(np.abs(res["d"] - ref["d"]) / np.maximum(np.abs(ref["d"]), 1e-300))),
                  "nu_abs": abs(res["nu"] - ref["nu"]), "phi_abs": abs(res["phi"] - ref["phi"])}
        oracle["ok"] = bool(oracle["iters"] == res["iters"] and oracle["nllh_rel"] <= 1e-9 and oracle["nu_abs"] <= 1e-5 and
                            oracle["phi_abs"] <= 1e-5 and np.allclose(res["d"], ref["d"], rtol=1e-5, atol=1e-7))
    if rank == 0:
        line = {"metric": "em_iterations_per_s", "value": res["iters"] / dt, "unit": "EM iterations/s", "n_gpus": world,
                "steps": runs, "warmup": max(1, args.warmup), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config_dict(name, spec, qptr, world), workload="config3_em", conv_eps=1e-8, max_iter=args.max_iter),
                "em": {"iterations": res["iters"], "nllh": res["nllh"], "nu": res["nu"], "phi": res["phi"],
                       "d_sum": float(res["d"].sum()), "d_min": float(res["d"].min()), "d_max": float(res["d"].max()),
                       "status": res["status"], "seconds_per_run": dt,
                       "cell_site_per_s": N * Ktot * res["iters"] / dt},
                "oracle_em": oracle,
                "note": "a step = one whole EM run (ShardedEngine.em: EM_solver.EM_optimization on site shards, one llh read-back "
                        "per iteration); iterations / nllh / nu / phi / d_sum must be equal across n_gpus",
                "exchange": "peer memory (pmm_finish_peer), device-resident loop pmm_em on every rank" if sh.peer else ("NCCL all_reduce" if world > 1 else "none")}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    sh.close()
    rk.close()


# ------------------------------------------------------------------------------------------------ config 5: NNI neighbourhood
def ultrametric_from(forest, d):
    """make ``d`` ultrametric without changing the topology: node heights = the longest leaf path below, every
    branch = parent's height - own height (>= its old length, so >= dmin stays true)"""
    n = forest.n_nodes
    H = np.zeros(n)
    for v in forest.order:                              # children first
        v = int(v)
        if forest.left[v] >= 0:
            a, b = int(forest.left[v]), int(forest.right[v])
            H[v] = max(H[a] + d[a], H[b] + d[b])
    out = np.array(d, np.float64, copy=True)
    for v in range(n):
        p = int(forest.parent[v])
        if p >= 0:
            out[v] = H[p] - H[v]
    return out


def run_nni(args, rank, local_rank, world):
    """BASELINE config 5: every candidate of one NNI round of the 5k x 300 tree, scored in one batch per rank
    (``pmm_score_nni_batch``), candidate trees sharded ``rank::world`` (``dist.score_nni_sharded``).  A *step* scores
    the whole neighbourhood: cache E-step on every rank, the batch kernel on the rank's candidates, one all_gather."""
    real_stdout = grab_stdout()
    rk = Ranks(rank, local_rank, world)
    from laml_b200.capi import Engine
    from laml_b200.dist import score_nni_sharded
    from laml_b200.nni import nni_candidates
    name = "config5_5k_x_300"
    spec = workload_spec(name)
    Ktot, N = spec["n_sites"], spec["n_leaves"]
    forest, d, qptr, qval, codes = build_inputs(spec, (0, Ktot), rk.dev)
    eng = Engine(N, qptr, qval, codes, device=local_rank)
    eng.set_tree(forest.parent, forest.left, forest.right)
    cu, ck = nni_candidates(forest.parent, forest.left, N)
    nu, phi = spec["nu"], spec["phi"]
    d_free = np.maximum(d, 0.005)
    modes = {
        "unclamped": dict(d=d_free, kw={}),                                  # GPU_Topology_search default (clamp_start=False)
        "clamped_2dmin": dict(d=np.maximum(d, 0.01), kw={}),                 # the reference's starting point (ini_brlens)
        "ultrametric": dict(d=ultrametric_from(forest, d_free), kw={"ultra": True}),   # strategy['ultra_constr'] (batch_ultra)
    }
    out = {}
    for mname, m in modes.items():
        res = None
        for _ in range(max(1, args.warmup)):
            res = score_nni_sharded(eng, cu, ck, m["d"], nu, phi, device=rk.dev, **m["kw"])
        rk.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = score_nni_sharded(eng, cu, ck, m["d"], nu, phi, device=rk.dev, **m["kw"])
        rk.barrier()
        dt = rk.max((time.perf_counter() - t0) / args.steps)
        ok = ~np.isnan(res["nllh"])
        out[mname] = {"ms_per_neighbourhood": dt * 1e3, "candidates_per_s": len(cu) / dt, "scored": int(ok.sum()),
                      "no_room": int((~ok).sum()), "mean_em_iters": float(res["iters"][ok].mean()), "max_em_iters": int(res["iters"][ok].max()),
                      "best_nllh": float(np.nanmin(res["nllh"])), "sum_nllh": float(np.nansum(res["nllh"])),
                      "sum_d5": float(np.nansum(res["d5"][ok]))}
    if rank == 0:
        main_mode = out["unclamped"]
        line = {"metric": "nni_candidates_per_s", "value": main_mode["candidates_per_s"], "unit": "candidates/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": main_mode["ms_per_neighbourhood"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config_dict(name, spec, qptr, world), workload="config5_nni", candidates=int(len(cu)),
                               parallelism="candidate-shard x%d (every rank holds all sites)" % world),
                "modes": out,
                "note": "value = the search's default mode (lengths outside the five free branches stay where they are); "
                        "best_nllh / sum_nllh / sum_d5 / scored must be equal across n_gpus; host wall-clock incl. the all_gather"}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    rk.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("PMM_BENCH_WORKLOAD", "config4_100k_x_2k"))
    ap.add_argument("--cpu-sample-sites", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--oracle-em", action="store_true", help="config3_em, N=1: also run the oracle's EM (minutes of CPU)")
    ap.add_argument("--max-iter", type=int, default=1000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "config3_em":
        run_em_convergence(args, rank, local_rank, world)
    elif args.workload == "config5_nni":
        run_nni(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
