run() { timeout 200 python bench.py --workload config3_10k_x_500 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('estep_ms',round(l['estep_ms'],3),'inside',round(l['inside_ms'],3),'outside',round(l['outside_ms'],3))"; }
for ts in 1024 1536 2048; do for st in 2 3; do
echo "== TS=$ts stages=$st"; PMM_TILE_SLOTS=$ts PMM_STAGES_IN=$st PMM_STAGES_OUT=$st run
done; done
for th in 0 256 100000; do echo "== TS=2048 spec_threshold=$th"; PMM_TILE_SLOTS=2048 PMM_SPEC_THRESHOLD=$th run; done
