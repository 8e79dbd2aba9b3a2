for ts in 512 1024 2048 4096; do for g in 8 16 32; do
echo "== TS=$ts G=$g"; PMM_TILE_SLOTS=$ts PMM_GROUP=$g timeout 200 python bench.py --workload config3_10k_x_500 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('estep_ms',round(l['estep_ms'],3),'inside',round(l['inside_ms'],3),'outside',round(l['outside_ms'],3))"
done; done
