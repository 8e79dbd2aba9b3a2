#!/bin/bash
# A/B of the library variants under variants/ on config 4 and the 250-site shard
mkdir -p gpurun_out
timeout 1200 python scripts/exp_variants.py --workloads config4,shard250 > gpurun_out/r2_variants15.log 2>&1
cat gpurun_out/r2_variants15.log
