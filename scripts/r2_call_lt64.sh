#!/bin/bash
# 250-site shard of config 4 (one rank of an 8-GPU run) and config 3: CTAs of 64 sites, more site slices, staggered slices
mkdir -p gpurun_out
for v in "EXP_SLICES=2 PMM_STAGGER=0" "EXP_SLICES=2 PMM_STAGGER=1" "EXP_SLICES=4 PMM_STAGGER=0" "EXP_SLICES=4 PMM_STAGGER=1" "EXP_SLICES=3 PMM_STAGGER=1"; do
  echo "== $v"; env $v timeout 600 python scripts/exp_variants.py --workloads shard250,config3
done > gpurun_out/r2_lt64.log 2>&1
cat gpurun_out/r2_lt64.log
