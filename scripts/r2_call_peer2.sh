#!/bin/bash
# two GPUs: the 2-rank hardware test in both exchange modes (peer memory over CUDA IPC / NCCL), then the bench at N = 2 both ways
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_nccl.py -m gpu -x -q > gpurun_out/r2_peer_tests_2gpu.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2_peer_tests_2gpu.log
for mode in 1 0; do
  PMM_PEER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2d_bench_n2_peer$mode.json 2> gpurun_out/r2d_bench_n2_peer$mode.err
  echo "bench peer=$mode rc=$?"; python - <<PY
import json
l=json.loads(open('gpurun_out/r2d_bench_n2_peer$mode.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','allreduce_ms','mstep_ms','gpu_launches')}, l['e2e']['ms_per_step'], l['engine']['exchange'][:40])
print(l['checksum']['estep_at_start'], l['checksum']['after_steps'])
PY
done
