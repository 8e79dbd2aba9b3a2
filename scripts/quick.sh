run() { timeout 300 python bench.py --workload $1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('$1 estep_ms',round(l['estep_ms'],3),'inside',round(l['inside_ms'],3),'outside',round(l['outside_ms'],3),'ms/step',round(l['ms_per_step'],3))"; }
for w in "$@"; do run $w; done
