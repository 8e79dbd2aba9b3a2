#!/bin/bash
mkdir -p gpurun_out
mkdir -p variants; cp laml_b200/libpmm_b200.so variants/libpmm_cur.so
for c in 0 52 60 68; do for top in 16 20 24; do
  echo "PMM_CLADE=$c PMM_CLADE_TOP=$top"; PMM_CLADE=$c PMM_CLADE_TOP=$top timeout 300 python scripts/exp_variants.py --only cur --workloads shard250
done; done > gpurun_out/r2_clade_sweep3.log 2>&1
cat gpurun_out/r2_clade_sweep3.log
