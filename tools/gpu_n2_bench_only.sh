#!/bin/bash
# bench at N GPUs, twice (transient-stall check)   usage: gpurun --gpus N -- bash tools/gpu_n2_bench_only.sh TAG N
TAG=${1:-r2y}; N=${2:-2}
mkdir -p gpurun_out
for R in a b; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 5 \
   > gpurun_out/${TAG}${R}_bench_n${N}.json 2> gpurun_out/${TAG}${R}_bench_n${N}.err; echo "bench N=$N exit $?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/${TAG}${R}_bench_n${N}.json') if l.startswith(chr(123))][0])
print(round(d['value']), round(d['ms_per_step'],3), d['config']['steps_in_flight'], d['config']['bracket_attempts_ms_per_step'], 'e2e', round(d['e2e']['value']), 'weak', round(d.get('weak',{}).get('value',0)), 'train', round(d['train']['value']), round(d['train']['ms_per_step'],2))"
done
