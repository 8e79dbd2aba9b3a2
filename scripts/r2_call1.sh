#!/bin/bash
# Round 2, GPU call 1: the whole GPU suite (with the new at-size parity tests), the headline bench, the reference arm,
# config 3 EM to convergence (with the oracle's EM beside it), config 5 NNI neighbourhood, per-rank shard timing.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2c1_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/r2c1_gpu_tests.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2c1_bench_n1.json 2> gpurun_out/r2c1_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2c1_bench_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','mstep_ms','gpu_launches')})
print('e2e',l['e2e']); print('split',l['e2e_split']['ms_per_step']); print('parity',l['parity']); print('checksum',l['checksum']['estep_at_start'])
print('frac',l['roofline']['frac'],'inside',l['roofline']['inside']['frac'],'cpu',l['cpu_baseline'])
PY
OMP_NUM_THREADS=1 timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c1_reference_arm.json 2> gpurun_out/r2c1_reference_arm.err
echo "reference rc=$?"; cut -c1-400 gpurun_out/r2c1_reference_arm.json
timeout 900 python bench.py --workload config3_em --steps 3 --warmup 1 --oracle-em > gpurun_out/r2c1_config3_em_n1.json 2> gpurun_out/r2c1_config3_em_n1.err
echo "config3_em rc=$?"; cat gpurun_out/r2c1_config3_em_n1.json
timeout 600 python bench.py --workload config5_nni --steps 5 --warmup 2 > gpurun_out/r2c1_config5_nni_n1.json 2> gpurun_out/r2c1_config5_nni_n1.err
echo "config5_nni rc=$?"; cat gpurun_out/r2c1_config5_nni_n1.json
for w in config3_10k_x_500 config2_1k_x_30; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 > gpurun_out/r2c1_bench_$w.json 2> gpurun_out/r2c1_bench_$w.err
  python -c "
import json;l=json.loads(open('gpurun_out/r2c1_bench_$w.json').read().strip().splitlines()[-1])
print('$w',{k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','mstep_ms','gpu_launches')}, l['roofline']['frac'], l['parity'] and l['parity']['ok'])"
done
for s in 1 2; do PMM_SLICES=$s timeout 300 python scripts/prof_shard.py 250 3; done 2>&1 | tail -2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
