#!/bin/bash
# one 8-GPU call: the headline bench with the exchange over peer memory (default) and with NCCL, config 3 EM to convergence
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for mode in 1 0; do
  PMM_PEER=$mode $TR --master-port 2951$mode bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/r2e_bench_n8_peer$mode.json 2> gpurun_out/r2e_bench_n8_peer$mode.err
  echo "bench n8 peer=$mode rc=$?"; python - <<PY
import json
l=json.loads(open('gpurun_out/r2e_bench_n8_peer$mode.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','allreduce_ms','mstep_ms','gpu_launches')})
print('e2e',l['e2e']['ms_per_step'],'split',l['e2e_split']['ms_per_step'], l['engine']['exchange'][:30]); print('checksum',l['checksum'])
PY
done
$TR --master-port 29504 bench.py --gpus 8 --workload config3_em --steps 3 --warmup 1 > gpurun_out/r2e_config3_em_n8.json 2> gpurun_out/r2e_config3_em_n8.err; echo "config3_em rc=$?"; cut -c1-1000 gpurun_out/r2e_config3_em_n8.json
