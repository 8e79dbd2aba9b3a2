"""Time the ultrametric-constrained M-step at BASELINE sizes: the device solver (pmm_mstep_ultra_dev: one cooperative
launch) against the library's host solver on the same statistics, cold and warm, beside the E-step of the same tree."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from laml_b200 import ultrametric as um
from laml_b200.capi import Engine
from laml_b200.synth import CONFIGS, make_workload
out = {}
for cfg in sys.argv[1:] or ["config3_10k_x_500", "config4_100k_x_2k"]:
    w = make_workload(seed=1, device="cuda", **CONFIGS[cfg])
    f, N, K = w.forest, w.forest.n_leaves, w.codes.shape[1]
    n = f.n_nodes
    # an ultrametric start: heights = longest leaf path below
    H = np.zeros(n)
    for v in f.order:
        v = int(v)
        if f.left[v] >= 0:
            H[v] = max(H[f.left[v]] + w.d[f.left[v]], H[f.right[v]] + w.d[f.right[v]])
    d0 = np.maximum(np.where(f.parent >= 0, H[np.maximum(f.parent, 0)] - H, w.d), 0.01)
    kind = np.zeros(n, np.int8); fixed_len = np.zeros(n)
    R = float((w.codes != -1).sum())
    eng = Engine(N, w.qptr, w.qval, w.codes)
    eng.set_tree(f.parent, f.left, f.right)
    res = {"levels": eng.info()["inside_levels"]}
    cur = (d0, w.nu, w.phi)
    problem = um.UltrametricProblem(f.parent, kind, fixed_len, 0.005, 10.0)
    fm = np.ones(n, np.uint8)
    Hh = None
    for it in range(4):
        eng.upload_params(*cur)
        eng.sync(); t0 = time.perf_counter(); eng.estep_dev(); eng.sync(); t_e = time.perf_counter() - t0
        est = eng.estep(*cur)
        t0 = time.perf_counter()
        st, iters = eng.mstep_ultra_dev(R, N, kind, fixed_len, warm=it > 0)
        eng.sync(); t_dev = time.perf_counter() - t0
        d_dev, nu_dev, phi_dev = eng.download_params()
        t0 = time.perf_counter()
        d_h, nu_h, phi_h, Hh, st_h = um.mstep(problem, est["S"], est["Rtilde"], R, N, K, fm, cur[0], cur[1], cur[2], H_prev=Hh)
        t_host = time.perf_counter() - t0
        res["iter%d" % it] = {"estep_ms": t_e * 1e3, "device_mstep_ms": t_dev * 1e3, "host_mstep_ms": t_host * 1e3, "newton_steps": iters,
                              "status": st, "host_status": st_h, "max_abs_d_diff": float(np.abs(d_dev - d_h).max()),
                              "nu_diff": abs(nu_dev - nu_h)}
        cur = (d_dev, nu_dev, phi_dev)
    out[cfg] = res
    eng.close()
print(json.dumps(out))
