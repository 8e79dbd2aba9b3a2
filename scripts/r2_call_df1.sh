#!/bin/bash
# dataflow launches, first GPU run: parity tests with every pass as ONE dataflow launch, then timings
mkdir -p gpurun_out
PMM_DF=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_solver.py -m gpu -x -q > gpurun_out/r2_df_tests_mode2.log 2>&1
echo "mode 2 tests rc=$?"; tail -3 gpurun_out/r2_df_tests_mode2.log
PMM_DF=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q > gpurun_out/r2_df_tests_mode1.log 2>&1
echo "mode 1 tests rc=$?"; tail -3 gpurun_out/r2_df_tests_mode1.log
timeout 900 python scripts/exp_df.py config2 config3 shard250 > gpurun_out/r2_df_exp1.log 2>&1
echo "exp rc=$?"; cat gpurun_out/r2_df_exp1.log
