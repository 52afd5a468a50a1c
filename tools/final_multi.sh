#!/bin/bash
# usage: tools/final_multi.sh N -- the bench line at N GPUs (reference arm first for the parity record), CG line at N = 8
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/r2f_ref_n$N.json 2>/dev/null
$T 29631 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
if [ "$N" = "8" ]; then
  $T 29632 bench.py --gpus $N --workload weak --steps 20 --warmup 3 > gpurun_out/r2f_weak_n$N.json 2> gpurun_out/r2f_weak_n$N.err
fi
python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single', round(d['e2e']['single_solve_value'],1), 'fp', d['fingerprint']['equals_one_gpu'], 'parity', (d.get('parity') or {}).get('ok'), d['clocks'])"
