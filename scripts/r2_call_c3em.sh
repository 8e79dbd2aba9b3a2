#!/bin/bash
N=$1; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --workload config3_em --steps 3 --warmup 1 > gpurun_out/r2j_config3_em_n$N.json 2> gpurun_out/r2j_config3_em_n$N.err
echo "rc=$?"; cut -c1-170 gpurun_out/r2j_config3_em_n$N.json; python -c "
import json;l=json.loads(open('gpurun_out/r2j_config3_em_n$N.json').read().strip().splitlines()[-1]); print(l['em'])"
