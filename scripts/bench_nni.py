#!/usr/bin/env python
"""Throughput of batched NNI candidate scoring on BASELINE config 5 (5k cells x 300 sites):
all 2(N-2) candidates of one NNI round in one call, next to (a) the engine's own full-tree EM per
candidate (what a per-candidate score_tree costs on the GPU) and (b) the CPU oracle per candidate."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from laml_b200.capi import Engine
from laml_b200.nni import apply_nni, nni_candidates
from laml_b200.synth import CONFIGS, make_workload

name = sys.argv[1] if len(sys.argv) > 1 else "config5_5k_x_300"
w = make_workload(device="cuda", **CONFIGS[name])
N, f = w.forest.n_leaves, w.forest
d = np.maximum(w.d, 0.01)
eng = Engine(N, w.qptr, w.qval, w.codes)
eng.set_tree(f.parent, f.left, f.right)
cu, ck = nni_candidates(f.parent, f.left, N)
eng.score_nni_batch(cu[:64], ck[:64], d, w.nu, w.phi)          # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
res = eng.score_nni_batch(cu, ck, d, w.nu, w.phi)
t_batch = time.perf_counter() - t0
# (a) per-candidate full-tree EM on the GPU for a sample
R = float((w.codes != -1).sum())
ref = Engine(N, w.qptr, w.qval, w.codes)
sample = np.linspace(0, len(cu) - 1, 16).astype(int)
t0 = time.perf_counter()
worst = 0.0
for i in sample:
    p2, l2, r2 = apply_nni(f.parent, f.left, f.right, int(cu[i]), int(ck[i]))
    ref.set_tree(p2, l2, r2)
    free = np.zeros(f.n_nodes, np.uint8); free[res["nodes"][i]] = 1
    out = ref.em(d, w.nu, w.phi, R, N, free_mask=free, optimize_nu=False, optimize_phi=False)
    worst = max(worst, abs(out["nllh"] - res["nllh"][i]) / abs(out["nllh"]))
t_full = (time.perf_counter() - t0) / len(sample)
line = {"workload": name, "candidates": int(len(cu)), "batch_s": t_batch, "candidates_per_s": len(cu) / t_batch,
        "mean_em_iters": float(res["iters"].mean()), "max_em_iters": int(res["iters"].max()),
        "full_tree_em_per_candidate_s_gpu": t_full, "speedup_vs_per_candidate_gpu_em": t_full * len(cu) / t_batch,
        "max_rel_nllh_diff_vs_full_tree_em": worst,
        "reference_per_candidate_s": "0.16-0.30 (64 cells x 30 sites, compare_EMsolvers/README.txt:35-60)"}
print(json.dumps(line))
