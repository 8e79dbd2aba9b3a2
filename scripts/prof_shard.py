"""E-step + M-step on a site shard of config 4 (100k cells x K sites) on one GPU: ncu target for the multi-GPU per-rank work"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from laml_b200.capi import Engine
from laml_b200.synth import make_workload
K = int(sys.argv[1]) if len(sys.argv) > 1 else 250
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = make_workload(n_leaves=100000, n_sites=K, a_lo=2, a_hi=100, nu=0.3, phi=0.3, seed=1, device="cuda")
N = w.forest.n_leaves
eng = Engine(N, w.qptr, w.qval, w.codes); eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
eng.set_total_sites(2000)
eng.upload_params(w.d, w.nu, w.phi)
R = float((w.codes != -1).sum()) * 2000 / K
import torch
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
for _ in range(reps):
    eng.estep_dev(); eng.mstep_dev(R, N)
eng.sync()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
acc = np.zeros(5)
for _ in range(10):
    with torch.cuda.stream(st):
        ev[0].record(); eng.estep_dev(); ev[1].record(); eng.mstep_dev(R, N); ev[2].record()
    ev[2].synchronize()
    t = eng.last_timing()
    acc += (t["inside_ms"], t["outside_ms"], t["total_ms"], ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
acc /= 10
print("slices=%s K=%d clades %d launches/pass %d: inside %.3f outside %.3f estep(kernel events) %.3f estep(stream) %.3f mstep %.3f ms" % (
    os.environ.get("PMM_SLICES", "2"), K, eng.info()["clades"], eng.info()["launches_per_pass"], *acc))
eng.close()
