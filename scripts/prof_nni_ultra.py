"""one pmm_score_nni_batch_ultra call on config 5 (ncu target)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from laml_b200.capi import Engine
from laml_b200.nni import nni_candidates
from laml_b200.synth import CONFIGS, make_workload
from bench import ultrametric_from
w = make_workload(seed=1, device="cuda", **CONFIGS["config5_5k_x_300"])
N, f = w.forest.n_leaves, w.forest
d = ultrametric_from(f, np.maximum(w.d, 0.005))
eng = Engine(N, w.qptr, w.qval, w.codes); eng.set_tree(f.parent, f.left, f.right)
cu, ck = nni_candidates(f.parent, f.left, N)
n = int(sys.argv[1]) if len(sys.argv) > 1 else len(cu)
for _ in range(2):
    t0 = time.perf_counter(); r = eng.score_nni_batch(cu[:n], ck[:n], d, w.nu, w.phi, ultra=True); dt = time.perf_counter() - t0
print("%d candidates %.2f ms, sum nllh %.6f" % (n, dt * 1e3, np.nansum(r["nllh"])))
