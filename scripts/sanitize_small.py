"""small runs of the round-2 kernels for compute-sanitizer: E-step (launch per tier and dataflow launches), EM, the NNI batch
kernel in both forms (the constrained one runs its M-step on a warp)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from laml_b200.capi import Engine
from laml_b200.nni import nni_candidates
from laml_b200.synth import make_workload
from bench import ultrametric_from
w = make_workload(300, 140, 1, 20, 0.2, 0.2, seed=4)
f, N = w.forest, 300
R = float((w.codes != -1).sum())
for df in ("0", "1", "2"):
    os.environ["PMM_DF"] = df; os.environ["PMM_CLADE"] = "12"; os.environ["PMM_CLADE_TOP"] = "4"
    eng = Engine(N, w.qptr, w.qval, w.codes); eng.set_tree(f.parent, f.left, f.right)
    st = eng.estep(w.d, w.nu, w.phi)
    em = eng.em(np.maximum(w.d, 0.01), w.nu, w.phi, R, N, max_iter=3)
    print("PMM_DF=%s llh %.6f em nllh %.6f launches/pass %d" % (df, st["llh"], em["nllh"], eng.info()["launches_per_pass"]), flush=True)
    if df == "0":
        cu, ck = nni_candidates(f.parent, f.left, N)
        a = eng.score_nni_batch(cu[:80], ck[:80], np.maximum(w.d, 0.005), w.nu, w.phi)
        b = eng.score_nni_batch(cu[:80], ck[:80], ultrametric_from(f, np.maximum(w.d, 0.005)), w.nu, w.phi, ultra=True)
        print("nni %.6f ultra %.6f (%d scored)" % (a["nllh"].sum(), np.nansum(b["nllh"]), int((~np.isnan(b["nllh"])).sum())), flush=True)
    eng.close()
