"""quick A/B timing of the two E-step engines on synthetic workloads (tuning aid, not the benchmark)"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laml_b200.capi import Engine
from laml_b200.synth import make_workload, CONFIGS
cfgs = sys.argv[1:] or ["config3_10k_x_500"]
for cfg in cfgs:
    w = make_workload(seed=1, device="cuda", **CONFIGS[cfg])
    N = w.forest.n_leaves
    res = {}
    for mode in os.environ.get("MODES", "lumped,fullstate").split(","):
        os.environ["PMM_ENGINE"] = mode
        eng = Engine(N, w.qptr, w.qval, w.codes); eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
        eng.upload_params(w.d, w.nu, w.phi)
        for _ in range(3): eng.estep_dev()
        eng.sync(); ts = []
        for _ in range(5):
            eng.estep_dev(); ts.append(eng.last_timing())
        info = eng.info()
        ins = np.mean([t["inside_ms"] for t in ts]); out = np.mean([t["outside_ms"] for t in ts]); tot = np.mean([t["total_ms"] for t in ts])
        st = eng.estep(w.d, w.nu, w.phi)
        res[mode] = st
        K = w.codes.shape[1]
        extra = ""
        if mode == "lumped":
            bi, bo = info["inside_bytes_per_site"] * K, info["outside_bytes_per_site"] * K
            extra = " clades %d target %d launches/pass %d  inside %.0f GB/s outside %.0f GB/s" % (
                info["clades"], info["clade_target"], info["launches_per_pass"], bi / ins / 1e6, bo / out / 1e6)
        print("%s %-9s inside %.3f outside %.3f total %.3f ms  cell*site/s %.3e  dev %.2f GB%s" % (
            cfg, mode, ins, out, tot, N * K / tot * 1e3, info["device_bytes"] / 1e9, extra), flush=True)
        eng.close()
    if len(res) == 2:
        a, b = res["lumped"], res["fullstate"]
        print("  llh rel diff %.2e  S max abs diff %.2e  Rt rel %.2e" % (
            abs(a["llh"] - b["llh"]) / abs(b["llh"]), np.abs(a["S"] - b["S"]).max(), abs(a["Rtilde"] - b["Rtilde"]) / b["Rtilde"]))
