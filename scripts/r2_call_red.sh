#!/bin/bash
# count-reduction variants of the outside kernel (the L1 data pipe is its busiest unit: 87 %, two thirds of it shared memory)
mkdir -p gpurun_out
timeout 1200 python scripts/exp_variants.py --workloads config4,shard250 > gpurun_out/r2_variants14.log 2>&1
cat gpurun_out/r2_variants14.log
