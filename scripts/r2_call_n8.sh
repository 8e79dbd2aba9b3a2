#!/bin/bash
# one 8-GPU call: the headline bench (strong scaling), NCCL all-reduce variants, config 3 EM and config 5 NNI at N = 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29501 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench n8 rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','allreduce_ms','mstep_ms','gpu_launches')})
print('e2e',l['e2e']['ms_per_step'],'split',l['e2e_split']['ms_per_step']); print('checksum',l['checksum'])
PY
for v in "" "NCCL_ALGO=Tree" "NCCL_ALGO=Ring" "NCCL_ALGO=NVLS" "NCCL_PROTO=LL128" "NCCL_ALGO=Ring NCCL_PROTO=LL128" "NCCL_NVLS_ENABLE=0"; do
  env $v $TR --master-port 29502 scripts/nccl_allreduce_bench.py 2>/dev/null | tail -1
done | tee gpurun_out/r2_nccl_variants_n8.log
NCCL_DEBUG=INFO $TR --master-port 29503 scripts/nccl_allreduce_bench.py 2>&1 | grep -i -E "nvls|algo|channel|Connected" | head -12 > gpurun_out/r2_nccl_info_n8.log; head -8 gpurun_out/r2_nccl_info_n8.log
$TR --master-port 29504 bench.py --gpus 8 --workload config3_em --steps 3 --warmup 1 > gpurun_out/r2_config3_em_n8.json 2> gpurun_out/r2_config3_em_n8.err; echo "config3_em rc=$?"; cut -c1-900 gpurun_out/r2_config3_em_n8.json
$TR --master-port 29505 bench.py --gpus 8 --workload config5_nni --steps 5 --warmup 2 > gpurun_out/r2_config5_nni_n8.json 2> gpurun_out/r2_config5_nni_n8.err; echo "config5_nni rc=$?"; cut -c1-1500 gpurun_out/r2_config5_nni_n8.json
