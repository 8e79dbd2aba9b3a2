#!/bin/bash
# final build, one GPU: the whole GPU suite, the headline bench, the reference arm, configs 2 / 3, config 3 EM, config 5, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/r2j_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -5 gpurun_out/r2j_gpu_tests.log
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2j_bench_n1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','mstep_ms','gpu_launches')})
print('e2e',l['e2e']['ms_per_step'],'split',l['e2e_split']['ms_per_step'],'frac',l['roofline']['frac'],l['roofline']['inside']['frac'], l['parity']['ok'], l['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_reference_arm.json 2> gpurun_out/r2j_reference_arm.err
echo "reference rc=$?"; cut -c1-300 gpurun_out/r2j_reference_arm.json
for w in config3_10k_x_500 config2_1k_x_30; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 > gpurun_out/r2j_bench_$w.json 2> gpurun_out/r2j_bench_$w.err
  python -c "
import json;l=json.loads(open('gpurun_out/r2j_bench_$w.json').read().strip().splitlines()[-1])
print('$w',{k:l[k] for k in ('ms_per_step','inside_ms','outside_ms','estep_ms','mstep_ms','gpu_launches')}, l['e2e']['ms_per_step'], l['roofline']['frac'], l['parity'] and l['parity']['ok'])"
done
timeout 900 python bench.py --workload config3_em --steps 3 --warmup 1 --oracle-em > gpurun_out/r2j_config3_em_n1.json 2> gpurun_out/r2j_config3_em_n1.err
echo "config3_em rc=$?"; cut -c1-200 gpurun_out/r2j_config3_em_n1.json
timeout 600 python bench.py --workload config5_nni --steps 5 --warmup 2 > gpurun_out/r2j_config5_nni_n1.json 2> gpurun_out/r2j_config5_nni_n1.err
echo "config5_nni rc=$?"; cut -c1-200 gpurun_out/r2j_config5_nni_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pmm -c 330 --csv \
   --log-file gpurun_out/r2j_launches_config4_n1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_ncu_bench.log 2>&1
echo "launch list rc=$?"; python scripts/summarize_traffic.py gpurun_out/r2j_launches_config4_n1.csv | tail -9
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
