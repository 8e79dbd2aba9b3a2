run() { timeout 200 python bench.py --workload config3_10k_x_500 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('estep_ms',round(l['estep_ms'],3),'inside',round(l['inside_ms'],3),'outside',round(l['outside_ms'],3))"; }
for ts in 512 1024 2048; do for st in 2 3 4; do
echo "== TS=$ts stages=$st"; PMM_TILE_SLOTS=$ts PMM_STAGES_IN=$st PMM_STAGES_OUT=$st run
done; done
for c in 1 2 3; do echo "== TS=1024 ctas=$c"; PMM_CTAS_IN=$c PMM_CTAS_OUT=$c run; done
