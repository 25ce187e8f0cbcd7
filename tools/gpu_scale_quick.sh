#!/bin/bash
# strong-scaling line only (no extras, no CPU arm) at N GPUs.   usage: gpurun --gpus N -- bash tools/gpu_scale_quick.sh N TAG
N=${1:-2}; TAG=${2:-r3t}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N \
   --steps 20 --warmup 5 --extras none --no-cpu-baseline > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench N=$N exit $?"
cut -c1-300 gpurun_out/${TAG}_bench_n${N}.json | tail -1
