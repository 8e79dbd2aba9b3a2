#!/bin/bash
mkdir -p gpurun_out variants; cp laml_b200/libpmm_b200.so variants/libpmm_cur.so
for c in 84 100 0; do echo "PMM_CLADE=$c"; PMM_CLADE=$c timeout 300 python scripts/exp_variants.py --only cur --workloads shard500,shard250; done > gpurun_out/r2_clade_sweep4.log 2>&1
cat gpurun_out/r2_clade_sweep4.log
