#!/bin/bash
# full ncu captures of the two dominant kernels on the final build (tier-0 launches, all 16 site blocks)
mkdir -p gpurun_out
PMM_SLICES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmm_lumped_outside -s 9 -c 1 -o gpurun_out/r2j_lout \
   python scripts/prof_lumped.py config4_100k_x_2k 1 > gpurun_out/r2j_lout.log 2>&1; echo "lout rc=$?"
PMM_SLICES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmm_lumped_inside -s 5 -c 1 -o gpurun_out/r2j_lin \
   python scripts/prof_lumped.py config4_100k_x_2k 1 > gpurun_out/r2j_lin.log 2>&1; echo "lin rc=$?"
