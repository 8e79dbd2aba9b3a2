#!/bin/bash
# final build on N GPUs (N = $1): [N = 2: the 2-rank hardware test in both exchange modes], the headline bench, [N = 8: config 3 EM, config 5 NNI]
N=$1; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = 2 ]; then
  timeout 900 python -m pytest tests/test_zz_gpu_nccl.py -m gpu -x -q > gpurun_out/r2j_tests_2gpu.log 2>&1; echo "2-rank tests rc=$?"; tail -3 gpurun_out/r2j_tests_2gpu.log
fi
$TR --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2j_bench_n$N.json 2> gpurun_out/r2j_bench_n$N.err
echo "bench n$N rc=$?"; python - <<PY
import json
l=json.loads(open('gpurun_out/r2j_bench_n$N.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','allreduce_ms','mstep_ms','gpu_launches')})
print('e2e',l['e2e']['ms_per_step'],'split',l['e2e_split']['ms_per_step'], l['engine']['exchange'][:30]); print('checksum',l['checksum']['estep_at_start'], l['checksum']['after_steps'])
PY
if [ "$N" = 8 ]; then
  $TR --master-port 29524 bench.py --gpus 8 --workload config3_em --steps 3 --warmup 1 > gpurun_out/r2j_config3_em_n8.json 2> gpurun_out/r2j_config3_em_n8.err; echo "config3_em rc=$?"; cut -c1-160 gpurun_out/r2j_config3_em_n8.json
  $TR --master-port 29525 bench.py --gpus 8 --workload config5_nni --steps 5 --warmup 2 > gpurun_out/r2j_config5_nni_n8.json 2> gpurun_out/r2j_config5_nni_n8.err; echo "config5_nni rc=$?"
  python -c "
import json;l=json.loads(open('gpurun_out/r2j_config5_nni_n8.json').read().strip().splitlines()[-1])
for k,v in l['modes'].items(): print(k, v['ms_per_neighbourhood'], v['sum_nllh'])"
fi
