#!/bin/bash
# profiles of the round-2 build on one GPU: the launch list of the benchmark command, a full capture of the dominant
# kernel (tier-0 launch of pmm_lumped_outside, slices off: all 16 site blocks) and of the inside kernel, launch list of a 250-site shard
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pmm -c 330 --csv \
   --log-file gpurun_out/r2f_launches_config4_n1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
echo "launch list rc=$?"; python scripts/summarize_traffic.py gpurun_out/r2f_launches_config4_n1.csv | tail -14
PMM_SLICES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmm_lumped_outside -s 9 -c 1 -o gpurun_out/r2f_lout \
   python scripts/prof_lumped.py config4_100k_x_2k 1 > gpurun_out/r2f_lout.log 2>&1; echo "lout rc=$?"
PMM_SLICES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmm_lumped_inside -s 5 -c 1 -o gpurun_out/r2f_lin \
   python scripts/prof_lumped.py config4_100k_x_2k 1 > gpurun_out/r2f_lin.log 2>&1; echo "lin rc=$?"
PMM_SLICES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pmm -c 120 --csv --log-file gpurun_out/r2f_shard250_launches.csv \
   python scripts/prof_shard.py 250 3 > gpurun_out/r2f_shard250.log 2>&1; echo "shard rc=$?"; python scripts/summarize_launches.py gpurun_out/r2f_shard250_launches.csv | tail -14
ls -la gpurun_out/r2f_*
