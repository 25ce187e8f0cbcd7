#!/bin/bash
# 2-GPU evidence: inference weak scaling and the training step with its gradient all-reduce
TAG=${1:-r1o}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "infer n2 exit $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --mode train --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_train_n2.json 2> gpurun_out/${TAG}_bench_train_n2.err; echo "train n2 exit $?"
tail -3 gpurun_out/${TAG}_bench_n2.err gpurun_out/${TAG}_bench_train_n2.err
cut -c1-400 gpurun_out/${TAG}_bench_n2.json; grep -o '"value": [0-9.]*\|"allreduce_ms": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/${TAG}_bench_train_n2.json | head
