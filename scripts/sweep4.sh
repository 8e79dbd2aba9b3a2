run() { timeout 200 python bench.py --workload config3_10k_x_500 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('estep_ms',round(l['estep_ms'],3),'inside',round(l['inside_ms'],3),'outside',round(l['outside_ms'],3))"; }
for ts in 1024 1536 1920 2048; do
echo "== TS=$ts"; PMM_TILE_SLOTS=$ts run
done
for st in "2 2" "3 3" "4 3"; do set -- $st; echo "== TS=1920 stages in=$1 out=$2"; PMM_TILE_SLOTS=1920 PMM_STAGES_IN=$1 PMM_STAGES_OUT=$2 run; done
echo "== TS=1920 ctas=3"; PMM_TILE_SLOTS=1920 PMM_CTAS=3 run
echo "== TS=1920 thr=100000"; PMM_TILE_SLOTS=1920 PMM_SPEC_THRESHOLD=100000 run
echo "== TS=1920 thr=0"; PMM_TILE_SLOTS=1920 PMM_SPEC_THRESHOLD=0 run
