#!/bin/bash
# A/B of the library variants under variants/
mkdir -p gpurun_out
timeout 1200 python scripts/exp_variants.py --workloads ${1:-config4,shard250} > gpurun_out/${2:-r2_variants16}.log 2>&1
cat gpurun_out/${2:-r2_variants16}.log
