#!/bin/bash
# N-GPU visit: graphed data-parallel probe, the 2-GPU pytest, bench at N GPUs.   usage: gpurun --gpus N -- bash tools/gpu_n2_round2.sh TAG N
TAG=${1:-r2d}; N=${2:-2}
mkdir -p gpurun_out
bash tools/try_ddp_graph.sh
timeout 300 python -m pytest tests/test_training_gpu.py -m gpu -q -k "ddp_graphed or reference_golden" -s 2>&1 | grep -E "noise|worst|train-mode|passed|failed|Error" | tee gpurun_out/${TAG}_pytest_n${N}.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench N=$N exit $?"
cut -c1-300 gpurun_out/${TAG}_bench_n${N}.json; tail -3 gpurun_out/${TAG}_bench_n${N}.err
