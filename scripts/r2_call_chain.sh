#!/bin/bash
# where a chain-bound (top-tier) launch spends its time: tier-2 launches (30 clades x 2 site blocks) of both passes on a 250-site shard
mkdir -p gpurun_out
PMM_SLICES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pmm_lumped_inside -s 7 -c 1 -o gpurun_out/r2i_chain_in \
   python scripts/prof_shard.py 250 1 > gpurun_out/r2i_chain_in.log 2>&1; echo "in rc=$?"
PMM_SLICES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pmm_lumped_outside -s 7 -c 1 -o gpurun_out/r2i_chain_out \
   python scripts/prof_shard.py 250 1 > gpurun_out/r2i_chain_out.log 2>&1; echo "out rc=$?"
ls -la gpurun_out/r2i_*
