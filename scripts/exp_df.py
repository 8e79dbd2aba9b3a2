"""Dataflow launches (PMM_DF) against the launch-per-tier schedule on one GPU: E-step time (CUDA events), EM iterations/s of the
device-resident estep_dev + mstep_dev pair, and bit-equality of the results, over clade capacities.
    python scripts/exp_df.py [workload ...]      workloads: config2, config3, config4, shard<K> (K sites of config 4's tree)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from laml_b200.capi import Engine
from laml_b200.synth import CONFIGS, make_workload

GRID = {
    "config2": [dict(PMM_DF="0"), dict(PMM_DF="0", PMM_CLADE="8", PMM_CLADE_TOP="8")] +
               [dict(PMM_DF="2", PMM_CLADE=str(c), PMM_CLADE_TOP=str(t)) for c, t in ((16, 6), (8, 4), (8, 6), (6, 6), (4, 4), (12, 6), (8, 8), (16, 16), (4, 2), (6, 3))],
    "config3": [dict(PMM_DF="0")] + [dict(PMM_DF="2", PMM_CLADE=str(c), PMM_CLADE_TOP=str(t)) for c, t in ((0, 4), (0, 6), (0, 8), (0, 16), (24, 6), (32, 6), (12, 6), (8, 4))] +
               [dict(PMM_DF="1", PMM_CLADE_TOP="6")],
    "shard": [dict(PMM_DF="0", PMM_SLICES="1"), dict(PMM_DF="0", PMM_SLICES="2")] +
             [dict(PMM_DF="2", PMM_CLADE=str(c), PMM_CLADE_TOP=str(t)) for c, t in ((0, 4), (0, 6), (0, 8), (0, 16), (60, 6), (30, 6), (24, 4))] +
             [dict(PMM_DF="1", PMM_CLADE_TOP="6", PMM_SLICES="2"), dict(PMM_DF="1", PMM_CLADE_TOP="6", PMM_SLICES="1")],
    "config4": [dict(PMM_DF="0"), dict(PMM_DF="1", PMM_CLADE_TOP="6"), dict(PMM_DF="1", PMM_CLADE_TOP="16"), dict(PMM_DF="2", PMM_CLADE_TOP="6")],
}
KEYS = ("PMM_DF", "PMM_CLADE", "PMM_CLADE_TOP", "PMM_SLICES")

for wl in sys.argv[1:] or ["config2", "config3", "shard250"]:
    spec = dict(CONFIGS[{"config2": "config2_1k_x_30", "config3": "config3_10k_x_500"}.get(wl, "config4_100k_x_2k")])
    if wl.startswith("shard"):
        spec["n_sites"] = int(wl[5:])
    w = make_workload(seed=1, device="cuda", **spec)
    N = w.forest.n_leaves
    R = float((w.codes != -1).sum())
    base = None
    for env in GRID["shard" if wl.startswith("shard") else wl]:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        eng = Engine(N, w.qptr, w.qval, w.codes)
        eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
        if "PMM_SLICES" in env:
            eng.set_option("slices", int(env["PMM_SLICES"]))
        st = eng.estep(w.d, w.nu, w.phi)
        sig = (st["llh"], float(st["S"].sum()), st["Rtilde"], st["S"].tobytes())
        if base is None:
            base = sig
        same = "same bits" if sig[3] == base[3] and sig[:3] == base[:3] else "DIFFERENT llh %r vs %r S %r vs %r" % (sig[0], base[0], sig[1], base[1])
        eng.upload_params(w.d, w.nu, w.phi)
        for _ in range(5):
            eng.estep_dev()
        ts = []
        for _ in range(20):
            eng.estep_dev()
            ts.append(eng.last_timing()["total_ms"])
        reps = 200 if spec["n_leaves"] <= 10000 else 30
        for _ in range(10):
            eng.estep_dev(); eng.mstep_dev(R, N)
        eng.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.estep_dev(); eng.mstep_dev(R, N)
        eng.sync()
        dt = (time.perf_counter() - t0) / reps
        info = eng.info()
        print("%-9s %-62s estep %8.1f us (min %8.1f)  em-iteration %8.1f us = %7.0f it/s  clades %5d launches/pass %d  %s" % (
            wl, " ".join("%s=%s" % (k[4:], env[k]) for k in KEYS if k in env), 1e3 * float(np.median(ts)), 1e3 * min(ts), dt * 1e6, 1 / dt,
            info["clades"], info["launches_per_pass"], same), flush=True)
        eng.close()
