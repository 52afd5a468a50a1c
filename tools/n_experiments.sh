#!/bin/bash
# usage: tools/n_experiments.sh N TAG  -- bench.py --gpus N --profile with boundary row chunks on (default) and off
N=$1; TAG=$2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
$T 29611 bench.py --gpus $N --steps 20 --warmup 3 --profile > gpurun_out/${TAG}_n${N}_bc1.json 2> gpurun_out/${TAG}_n${N}_bc1.err
MGB_BOUNDARY_CHUNKS=0 $T 29612 bench.py --gpus $N --steps 20 --warmup 3 --profile > gpurun_out/${TAG}_n${N}_bc0.json 2> gpurun_out/${TAG}_n${N}_bc0.err
for f in bc1 bc0; do python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_n${N}_$f.json').read().strip().splitlines()[-1]); print('$f N=$N value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))"; done
