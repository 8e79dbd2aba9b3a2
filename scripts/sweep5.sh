run() { timeout 200 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('estep_ms',round(l['estep_ms'],3),'inside',round(l['inside_ms'],3),'outside',round(l['outside_ms'],3))"; }
for W in config3_10k_x_500 config2_1k_x_30 config5_5k_x_300; do
echo "== $W G=8"; run
echo "== $W G=4"; PMM_LIB=/root/repo/laml_b200/libpmm_b200_g4.so run
done
W=config3_10k_x_500
for ts in 1280 1536 1920; do echo "== G=4 TS=$ts"; PMM_LIB=/root/repo/laml_b200/libpmm_b200_g4.so PMM_TILE_SLOTS=$ts run; done
