#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 1500 -- 'bash scripts/r2_first_call.sh'
# Writes gpurun_out/r2_*.{log,json,csv}; copy what matters to profiles/.
mkdir -p gpurun_out
# 1. the ULTRA variant of the NNI batch kernel, the fast_em boundary and ShardedEngine.em_step: first GPU runs (XPASS expected)
timeout 900 python -m pytest tests/test_zzz_gpu_nni_ultra.py tests/test_zz_gpu_fast_em.py tests/test_zzzz_gpu_em_step.py -m gpu -q -rxX > gpurun_out/r2_late_tests.log 2>&1
tail -5 gpurun_out/r2_late_tests.log
# 2. memory checker on the new kernel path (small case)
timeout 600 compute-sanitizer --tool memcheck python -m pytest "tests/test_zzz_gpu_nni_ultra.py::test_batch_ultra_matches_the_cpu_model" -m gpu -q -rxX > gpurun_out/r2_memcheck_nni_ultra.log 2>&1
tail -3 gpurun_out/r2_memcheck_nni_ultra.log
# 3. config 5 in this mode, and the launch list / one full capture of the kernel
timeout 600 python scripts/bench_nni_ultra.py > gpurun_out/r2_nni_ultra_config5.json 2> gpurun_out/r2_nni_ultra_config5.err
cat gpurun_out/r2_nni_ultra_config5.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pmm_nni_lumped -c 1 -o gpurun_out/r2_nni_ultra \
    python scripts/bench_nni_ultra.py > /dev/null 2>&1
# 4. the headline bench on this build
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -1 gpurun_out/r2_bench_n1.json
