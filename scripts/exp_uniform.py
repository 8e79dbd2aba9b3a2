import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laml_b200.capi import Engine
from laml_b200.synth import make_workload
for name, lo, hi in (("ragged 2..100", 2, 100), ("uniform 50", 50, 50), ("uniform 98", 98, 98), ("uniform 14", 14, 14)):
    w = make_workload(10000, 500, lo, hi, 0.3, 0.3, seed=1, device="cuda")
    eng = Engine(10000, w.qptr, w.qval, w.codes); eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    eng.upload_params(w.d, w.nu, w.phi)
    for _ in range(3): eng.estep_dev()
    eng.sync(); ts=[]
    for _ in range(5):
        eng.estep_dev(); ts.append(eng.last_timing())
    W = eng.info()["padded_slots_per_node"]
    ins = np.mean([t["inside_ms"] for t in ts]); out = np.mean([t["outside_ms"] for t in ts])
    print(name, "W", W, "inside %.3f outside %.3f ms" % (ins, out), "ns per kslot-node: in %.2f out %.2f" % (ins*1e6/(W*10000/1e3), out*1e6/(W*10000/1e3)))
    eng.close()
