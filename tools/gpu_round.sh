#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list, full captures of the dominant kernels.
# usage (under gpurun): bash tools/gpu_round.sh TAG
TAG=${1:-r1e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
# full-counter captures: tools/ncu_full_forward.sh (kept separate: ~6 GPU-minutes)
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json | cut -c1-1500
