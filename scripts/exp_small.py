"""EM iterations/s on the small BASELINE configs (device-resident pmm_em loop and the estep_dev + mstep_dev pair)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laml_b200.capi import Engine
from laml_b200.synth import CONFIGS, make_workload
for cfg in sys.argv[1:] or ["config2_1k_x_30", "config3_10k_x_500"]:
    w = make_workload(seed=1, device="cuda", **CONFIGS[cfg])
    N = w.forest.n_leaves
    eng = Engine(N, w.qptr, w.qval, w.codes); eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    R = float((w.codes != -1).sum())
    d0 = np.maximum(w.d, 0.01)
    eng.upload_params(d0, w.nu, w.phi)
    for _ in range(20):
        eng.estep_dev(); eng.mstep_dev(R, N)
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(200):
        eng.estep_dev(); eng.mstep_dev(R, N)
    eng.sync()
    dt = (time.perf_counter() - t0) / 200
    r = eng.em(d0, w.nu, w.phi, R, N, max_iter=50)
    t0 = time.perf_counter()
    r = eng.em(d0, w.nu, w.phi, R, N, max_iter=50)
    t_em = (time.perf_counter() - t0) / r["iters"]
    info = eng.info()
    print("%s PMM_CLADE=%s TOP=%s GRAPH=%s: %.1f us per estep_dev+mstep_dev (%.0f it/s), pmm_em %.1f us per iteration (%d iterations), clades %d launches/pass %d, nllh %.6f" % (
        cfg, os.environ.get("PMM_CLADE", "-"), os.environ.get("PMM_CLADE_TOP", "-"), os.environ.get("PMM_GRAPH", "-"), dt * 1e6, 1 / dt, t_em * 1e6, r["iters"], info["clades"], info["launches_per_pass"], r["nllh"]), flush=True)
    eng.close()
