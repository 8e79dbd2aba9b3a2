#!/usr/bin/env python
"""Batched NNI scoring under the ultrametric constraint (pmm_score_nni_batch_ultra) on BASELINE config 5
(5k cells x 300 sites): all 2(N-2) candidates of one NNI round in one call from ultrametric lengths, next to the
unconstrained batch on the same lengths and, for a sample, the one-at-a-time definition it replaces (full-tree
constrained EM per candidate: GPU E-step + host M-step with everything but five branches fixed)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from laml_b200 import ultrametric as um
from laml_b200.capi import Engine
from laml_b200.nni import apply_nni, nni_candidates
from laml_b200.synth import CONFIGS, make_workload
from test_nni_ultra_host import ultrametric_lengths

name = sys.argv[1] if len(sys.argv) > 1 else "config5_5k_x_300"
w = make_workload(device="cuda", **CONFIGS[name])
N, f = w.forest.n_leaves, w.forest
d, _ = ultrametric_lengths(f, np.random.default_rng(5), step=(0.006, 0.08))
eng = Engine(N, w.qptr, w.qval, w.codes)
eng.set_tree(f.parent, f.left, f.right)
cu, ck = nni_candidates(f.parent, f.left, N)
eng.score_nni_batch(cu[:64], ck[:64], d, w.nu, w.phi, ultra=True)          # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter(); res = eng.score_nni_batch(cu, ck, d, w.nu, w.phi, ultra=True); t_ultra = time.perf_counter() - t0
t0 = time.perf_counter(); free = eng.score_nni_batch(cu, ck, d, w.nu, w.phi); t_free = time.perf_counter() - t0
ok = ~np.isnan(res["nllh"])
# the definition, for a sample of feasible candidates: EM on the re-wired tree, host M-step with five free branches
ref = Engine(N, w.qptr, w.qval, w.codes)
sample = np.nonzero(ok)[0][np.linspace(0, ok.sum() - 1, 8).astype(int)]
worst, t_def = 0.0, 0.0
for i in sample:
    p2, l2, r2 = apply_nni(f.parent, f.left, f.right, int(cu[i]), int(ck[i]))
    ref.set_tree(p2, l2, r2)
    kind = np.full(f.n_nodes, um.FIXED, np.int8); kind[res["nodes"][i]] = um.FREE
    fm = (kind == um.FREE).astype(np.uint8)
    P = um.UltrametricProblem(p2, kind, d, 0.005, 10.0)
    t0 = time.perf_counter()
    dd, H, pre = d.copy(), None, ref.inside(d, w.nu, w.phi)
    for it in range(1, 1001):
        est = ref.estep(dd, w.nu, w.phi)
        dn, _, _, H, status = um.mstep(P, est["S"], est["Rtilde"], 1.0, 1, w.codes.shape[1], fm, dd, w.nu, w.phi, H_prev=H,
                                       optimize_phi=False, optimize_nu=False)
        dd = np.where(fm.astype(bool), dn, dd)
        cur = ref.inside(dd, w.nu, w.phi)
        if abs((cur - pre) / pre) < 1e-8:
            break
        pre = cur
    t_def += time.perf_counter() - t0
    worst = max(worst, abs(-cur - res["nllh"][i]) / abs(cur))
line = {"workload": name, "candidates": int(len(cu)), "no_room": int((~ok).sum()), "batch_ultra_s": t_ultra,
        "candidates_per_s": len(cu) / t_ultra, "batch_unconstrained_s_same_lengths": t_free,
        "mean_em_iters": float(res["iters"][ok].mean()), "max_em_iters": int(res["iters"][ok].max()),
        "one_at_a_time_s_per_candidate": t_def / len(sample), "speedup_vs_one_at_a_time": t_def / len(sample) * len(cu) / t_ultra,
        "max_rel_nllh_diff_vs_definition": worst}
print(json.dumps(line))
