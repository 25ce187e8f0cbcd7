#!/bin/bash
# 8-GPU visit: bench at N = 8 and N = 4 (strong-scaling line + weak + train extras).   usage: gpurun --gpus 8 -- bash tools/gpu_n8_round2.sh TAG
TAG=${1:-r2i}
mkdir -p gpurun_out
for N in 8 4; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N bench.py --gpus $N --steps 20 --warmup 5 \
     > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench N=$N exit $?"
  cut -c1-260 gpurun_out/${TAG}_bench_n${N}.json | tail -1
done
