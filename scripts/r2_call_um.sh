#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gpu_ultrametric_dev.py tests/test_gpu_solver.py tests/test_zz_gpu_case_tests.py tests/test_zzzz_gpu_em_step.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python scripts/bench_ultra_mstep.py > gpurun_out/r2_ultra_mstep.json 2> gpurun_out/r2_ultra_mstep.err; tail -3 gpurun_out/r2_ultra_mstep.err; python - <<'PY'
import json
o=json.load(open('gpurun_out/r2_ultra_mstep.json'))
for c,r in o.items():
    print(c, 'levels', r['levels'])
    for k,v in r.items():
        if k.startswith('iter'): print('  ',k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
