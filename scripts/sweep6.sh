run() { timeout 200 python bench.py --workload config3_10k_x_500 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('estep_ms',round(l['estep_ms'],3),'inside',round(l['inside_ms'],3),'outside',round(l['outside_ms'],3))"; }
echo "== default (G=4 TS=1920 in3 out2)"; run
for ts in 1152 1280 1408; do echo "== TS=$ts in=4 out=3"; PMM_TILE_SLOTS=$ts PMM_STAGES_IN=4 PMM_STAGES_OUT=3 run; done
echo "== TS=1920 ctas=1 in=4 out=4"; PMM_CTAS=1 PMM_STAGES_IN=4 PMM_STAGES_OUT=4 run
echo "== TS=2560 ctas=1 in=4 out=3"; PMM_TILE_SLOTS=2560 PMM_CTAS=1 PMM_STAGES_IN=4 PMM_STAGES_OUT=3 run
