#!/bin/bash
# the whole GPU suite and the headline bench on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2f_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r2f_gpu_tests.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2f_bench_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','mstep_ms','gpu_launches')})
print('e2e',l['e2e']['ms_per_step'],'split',l['e2e_split']['ms_per_step'],'frac',l['roofline']['frac'],l['roofline']['inside']['frac'], l['parity']['ok'])
PY
