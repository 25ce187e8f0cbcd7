#!/bin/bash
# parity tests + training bench (BASELINE config 5 shape) on one GPU
TAG=${1:-r1f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/${TAG}_pytest.log
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_train_n1.json 2> gpurun_out/${TAG}_bench_train.err; echo "train bench exit $?"
tail -5 gpurun_out/${TAG}_bench_train.err; cut -c1-1200 gpurun_out/${TAG}_bench_train_n1.json
