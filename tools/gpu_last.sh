#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r3w_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r3w_pytest.log
timeout 120 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/r3w_bench_train.json 2> gpurun_out/r3w_bench_train.err; echo "train bench exit $?"
tail -3 gpurun_out/r3w_pytest.log; cut -c1-260 gpurun_out/r3w_bench_train.json
