#!/bin/bash
# One GPU-box visit (round 2): parity tests, bench (strong-scaling line + extra keys), reference-on-GPU baseline.
# usage (under gpurun): bash tools/gpu_round2.sh TAG [quick]
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
if [ "$2" != "quick" ]; then
  timeout 600 python tools/bench_reference_gpu.py > gpurun_out/${TAG}_refgpu.log 2>&1; echo "refgpu exit $?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu exit $?"
fi
tail -8 gpurun_out/${TAG}_pytest.log; cut -c1-400 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench.err
