"""Worker of tests/test_zz_gpu_nccl.py, launched by ``torch.distributed.run`` with one process per GPU: the N > 1 data
plane on real hardware -- site shards + the exchange of ``[S | R~ | llh]`` (over peer memory inside the engine's
finishing kernel, or with ``PMM_PEER=0`` as an NCCL all-reduce), the replicated M-step, EM to
convergence on shards, candidate-sharded NNI scoring -- against ONE engine holding the whole problem.  Every rank
checks its own results; rank 0 writes the summary the test reads."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    import torch
    import torch.distributed as dist
    from laml_b200.capi import Engine
    from laml_b200.dist import ShardedEngine, partition_sites, score_nni_sharded
    from laml_b200.nni import nni_candidates
    from laml_b200.synth import make_workload
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = make_workload(1500, 200, 2, 60, 0.25, 0.2, seed=11)
    N, K, n = w.forest.n_leaves, w.codes.shape[1], w.forest.n_nodes
    R = float((w.codes != -1).sum())
    d0 = np.maximum(w.d, 0.01)
    # the whole problem on this rank's GPU: the single-GPU answer
    whole = Engine(N, w.qptr, w.qval, w.codes, device=local)
    whole.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    ref = whole.estep(d0, w.nu, w.phi)
    ref_em = whole.em(d0, w.nu, w.phi, R, N)
    # this rank's site shard
    lo, hi = partition_sites(w.qptr, world)[rank]
    qp = (w.qptr[lo:hi + 1] - w.qptr[lo]).astype(np.int32)
    qv = w.qval[w.qptr[lo]:w.qptr[hi]].copy()
    eng = Engine(N, qp, qv, np.ascontiguousarray(w.codes[:, lo:hi]), device=local)
    eng.set_tree(w.forest.parent, w.forest.left, w.forest.right)
    sh = ShardedEngine(eng, K)
    est = sh.estep(d0, w.nu, w.phi)
    S = np.array(est["S"], copy=True)
    res = {"rank": rank, "world": world, "peer": bool(sh.peer)}
    res["llh_rel"] = abs(est["llh"] - ref["llh"]) / abs(ref["llh"])
    res["S_max_err"] = float(np.max(np.abs(S - ref["S"]) / (1e-10 * np.abs(ref["S"]) + 1e-12 * K)))
    res["Rtilde_rel"] = abs(est["Rtilde"] - ref["Rtilde"]) / abs(ref["Rtilde"])
    em = sh.em(d0, w.nu, w.phi, R, N)
    res["em_iters"] = [int(em["iters"]), int(ref_em["iters"])]
    res["em_nllh_rel"] = abs(em["nllh"] - ref_em["nllh"]) / abs(ref_em["nllh"])
    res["em_d_max_abs"] = float(np.max(np.abs(em["d"] - ref_em["d"])))
    res["em_nu_abs"], res["em_phi_abs"] = abs(em["nu"] - ref_em["nu"]), abs(em["phi"] - ref_em["phi"])
    one = sh.em_step(d0, w.nu, w.phi, R, N)
    one_ref = whole.mstep(ref["S"], ref["Rtilde"], R, N, d0, w.nu, w.phi)
    res["em_step_d_max_abs"] = float(np.max(np.abs(one[0] - one_ref[0])))
    res["em_step_llh_rel"] = abs(one[3] - ref["llh"]) / abs(ref["llh"])
    # EM under the ultrametric constraint (the runner's default mode) on shards: the device M-step (pmm_mstep_ultra_dev) is a
    # pure function of the exchanged statistics, so the shards must walk the single engine's iterations (pmm_em_ultra)
    from laml_b200 import ultrametric as um
    kind, fixed_len = np.zeros(n, np.int8), np.zeros(n)
    prob = um.UltrametricProblem(w.forest.parent, kind, fixed_len, 0.005, 10.0)
    ref_u = whole.em_ultra(d0, w.nu, w.phi, R, N, kind, fixed_len, max_iter=12)
    em_u = sh.em_ultrametric(prob, d0, w.nu, w.phi, R, N, np.ones(n, np.uint8), max_iter=12)
    res["ultra_iters"] = [int(em_u["iters"]), int(ref_u["iters"])]
    res["ultra_status"] = [em_u["status"], ref_u["status"]]
    res["ultra_nllh_rel"] = abs(em_u["nllh"] - ref_u["nllh"]) / abs(ref_u["nllh"])
    res["ultra_d_max_abs"] = float(np.max(np.abs(em_u["d"] - ref_u["d"])))
    # every rank must hold bit-identical results (replicated M-step on identical all-reduced statistics)
    mine = torch.tensor(np.concatenate([em["d"], [em["nu"], em["phi"], em["nllh"]], S.reshape(-1), em_u["d"], [em_u["nu"], em_u["phi"], em_u["nllh"]]]),
                        device="cuda:%d" % local)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    res["ranks_bit_identical"] = bool(all(torch.equal(parts[0], p) for p in parts))
    # candidate-sharded NNI scoring: every rank holds all sites, scores rank::world, one all_gather
    cu, ck = nni_candidates(w.forest.parent, w.forest.left, N)
    cu, ck = cu[:400], ck[:400]
    single = whole.score_nni_batch(cu, ck, d0, w.nu, w.phi)
    shard = score_nni_sharded(whole, cu, ck, d0, w.nu, w.phi, device="cuda:%d" % local)
    res["nni_scores_equal"] = bool(np.array_equal(single["nllh"], shard["nllh"]) and np.array_equal(single["iters"], shard["iters"])
                                   and np.array_equal(single["nodes"], shard["nodes"]) and np.array_equal(single["d5"], shard["d5"]))
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        with open(out_path, "w") as fh:
            json.dump(allres, fh)
    sh.close()
    eng.close(); whole.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
