#!/bin/bash
# warp-cooperative constrained M-step of the NNI batch kernel: tests, then config 5 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzz_gpu_nni_ultra.py tests/test_gpu_nni.py tests/test_zz_gpu_topology.py tests/test_zz_gpu_parity_sizes.py -m gpu -x -q > gpurun_out/r2g_nni_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2g_nni_tests.log
timeout 600 python bench.py --workload config5_nni --steps 5 --warmup 2 > gpurun_out/r2g_config5_nni_n1.json 2> gpurun_out/r2g_config5_nni_n1.err
echo "config5_nni rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2g_config5_nni_n1.json').read().strip().splitlines()[-1])
for k,v in l['modes'].items(): print(k, {a:v[a] for a in ('ms_per_neighbourhood','scored','no_room','mean_em_iters','best_nllh','sum_nllh','sum_d5')})
PY
