#!/bin/bash
for v in v1 v3; do
  PMM_LIB=$PWD/variants/libpmm_$v.so timeout 600 ncu --clock-control none -k regex:pmm_nni_lumped -c 2 --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__grid_size \
    --log-file gpurun_out/r2g_nni9996_$v.csv python scripts/prof_nni_ultra.py 9996 2>&1 | tail -1
  grep -v "^==" gpurun_out/r2g_nni9996_$v.csv | cut -d, -f1,12- | tail -6
  for i in 1 2 3; do echo -n "$v plain run $i: "; PMM_LIB=$PWD/variants/libpmm_$v.so timeout 300 python scripts/prof_nni_ultra.py 9996 2>&1 | tail -1; done
done
