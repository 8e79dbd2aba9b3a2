#!/bin/bash
echo "== default queues, default slices"; timeout 600 python tests/peer_local_worker.py 4 2000 520 60 2>&1 | tail -4 | cut -c1-220
echo "== 32 queues, one slice"; CUDA_DEVICE_MAX_CONNECTIONS=32 PMM_SLICES=1 timeout 600 python tests/peer_local_worker.py 4 2000 520 60 2>&1 | tail -3 | cut -c1-220
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_gpu_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2j_gpu_tests.log
