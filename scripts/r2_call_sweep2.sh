#!/bin/bash
# 250-site shard: top-tier clade capacity below 16 with the launch-per-tier schedule; then the headline bench at N = 1
mkdir -p gpurun_out
for top in 16 12 8 6 4; do for c in 0 60; do
  echo "PMM_CLADE=$c PMM_CLADE_TOP=$top"; PMM_CLADE=$c PMM_CLADE_TOP=$top timeout 300 python scripts/exp_variants.py --only cur --workloads shard250
done; done > gpurun_out/r2_clade_sweep2.log 2>&1
cat gpurun_out/r2_clade_sweep2.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2e_bench_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','mstep_ms','gpu_launches')})
print('e2e',l['e2e']['ms_per_step'],'split',l['e2e_split']['ms_per_step'],'frac',l['roofline']['frac'],l['roofline']['inside']['frac'], l['parity']['ok'])
PY
