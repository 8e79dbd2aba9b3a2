"""one warm-up E-step + N E-steps of a synthetic workload on the default engine (ncu target)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laml_b200.capi import Engine
from laml_b200.synth import make_workload, CONFIGS
cfg = sys.argv[1] if len(sys.argv) > 1 else "config4_100k_x_2k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = make_workload(seed=1, device="cuda", **CONFIGS[cfg])
eng = Engine(w.forest.n_leaves, w.qptr, w.qval, w.codes); eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
eng.upload_params(w.d, w.nu, w.phi)
for _ in range(1 + reps): eng.estep_dev()
eng.sync()
print(eng.last_timing(), eng.info())
eng.close()
