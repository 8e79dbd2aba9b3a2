#!/usr/bin/env python
"""One NNI iteration through the TREE-LEVEL interface (laml_b200/topology.py) on BASELINE config 5
(5k cells x 300 sites): solver construction from newick / dict inputs, scoring the whole neighbourhood
(2(N-2) candidates, one batched launch), applying the best move on the live handle, re-scoring.  Host
wall-clock, since what is measured here is the host path around the kernels."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laml_b200.solver import GPU_EM_solver
from laml_b200.synth import CONFIGS, make_workload, reference_inputs
from laml_b200.solver import DEFAULT_STRATEGY as DEFAULT_STRATEGY_
from laml_b200.topology import apply_move, score_neighbourhood

name = sys.argv[1] if len(sys.argv) > 1 else "config5_5k_x_300"
t = time.perf_counter(); w = make_workload(**CONFIGS[name]); t_gen = time.perf_counter() - t
t = time.perf_counter(); newicks, msa, Q = reference_inputs(w); t_conv = time.perf_counter() - t
t = time.perf_counter()
s = GPU_EM_solver(newicks, {'charMtrx': msa}, {'Q': Q}, {'phi': w.phi, 'nu': w.nu})
t_ctor = time.perf_counter() - t
t = time.perf_counter(); base = -s.negative_llh(); t_bind = time.perf_counter() - t      # first use: flatten + pmm_create + set_tree
rounds = []
for it in range(3):
    t = time.perf_counter(); nb = score_neighbourhood(s); t_score = time.perf_counter() - t
    i = int(np.argmax(nb.score))
    t = time.perf_counter(); apply_move(s, nb, i); t_apply = time.perf_counter() - t
    t = time.perf_counter(); now = -s.negative_llh(); t_llh = time.perf_counter() - t
    rounds.append({"candidates": int(len(nb.score)), "score_all_s": t_score, "apply_move_s": t_apply, "llh_s": t_llh,
                   "best_score": float(nb.score[i]), "llh_after_move": now, "improved": int((nb.score > base).sum())})
    base = now
sc_true, _ = s.score_tree(dict(DEFAULT_STRATEGY_)); base_opt = sc_true
s.close()
# a search through GPU_Topology_search from a PERTURBED topology (100 random NNI moves away from the generating
# tree): initial full EM, NNI iterations (each scores the whole neighbourhood in one launch), no final EM
from random import seed
from laml_b200.flatten import forest_from_parent
from laml_b200.nni import apply_nni
from laml_b200.solver import DEFAULT_STRATEGY
from laml_b200.synth import Workload
from laml_b200.topology import GPU_Topology_search
rng = np.random.default_rng(3)
f = w.forest
p, l, r = f.parent, f.left, f.right
moved = 0
while moved < 100:
    u = int(rng.integers(f.n_leaves, f.n_nodes))
    if p[u] < 0:
        continue
    p, l, r = apply_nni(p, l, r, u, int(rng.integers(2)))
    moved += 1
w2 = Workload(forest_from_parent(p, l, r), w.d, w.qptr, w.qval, w.codes, w.nu, w.phi)
newicks2, _, _ = reference_inputs(w2)
seed(a=11)
iters = int(os.environ.get("NNI_ITERS", "150"))
ts = GPU_Topology_search(newicks2, GPU_EM_solver, data={'charMtrx': msa}, prior={'Q': Q}, params={'phi': w.phi, 'nu': w.nu})
ts._open(True)
t = time.perf_counter()
trees, sc, params = ts.__search_one__(dict(DEFAULT_STRATEGY), maxiter=iters)
t_search = time.perf_counter() - t
n_cand = 2 * (f.n_leaves - 2)
search = {"start": "generating tree + 100 random NNI moves", "maxiter": iters, "search_one_s": t_search,
          "neighbourhoods_scored": (ts.trees_scored - 1) // n_cand, "trees_scored": ts.trees_scored, "best_score": sc,
          "score_of_generating_topology_after_full_em": base_opt}
ts.close()
print(json.dumps({"workload": name, "search": search, "generate_s": t_gen, "to_reference_inputs_s": t_conv, "solver_ctor_s": t_ctor,
                  "first_bind_and_llh_s": t_bind, "rounds": rounds,
                  "reference_per_candidate_s": "0.16-0.30 at 64 cells x 30 sites (compare_EMsolvers/README.txt:35-60)"}))
