#!/bin/bash
# the exchange over peer memory, first GPU run (one GPU: the ranks are handles of one process), and the dataflow test file
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_peer.py tests/test_zz_gpu_dataflow.py -m gpu -x -q > gpurun_out/r2_peer_tests_1gpu.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r2_peer_tests_1gpu.log
