#!/bin/bash
mkdir -p gpurun_out
for v in v1 v3; do
  PMM_LIB=$PWD/variants/libpmm_$v.so timeout 600 ncu --clock-control none -k regex:pmm_nni_lumped -c 2 --csv \
    --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio \
    --log-file gpurun_out/r2g_nni_$v.csv python scripts/prof_nni_ultra.py 2000 > gpurun_out/r2g_nni_$v.log 2>&1
  echo "$v rc=$?"; tail -1 gpurun_out/r2g_nni_$v.log
  python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2g_nni_$v.csv') if not l.startswith('=='))]
hdr=rows[0]
for r in rows[1:]:
    if r[hdr.index('ID')]=='1': print(r[hdr.index('Metric Name')], r[hdr.index('Metric Value')])
PY
done
