#!/bin/bash
# final visit: parity tests, default bench, ncu launch list of one eager training step
TAG=${1:-r3u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_train_launches.csv python tools/ncu_train_step.py > gpurun_out/${TAG}_ncu.log 2>&1; echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/${TAG}_train_launches.csv > gpurun_out/${TAG}_train_launches_summary.txt
tail -4 gpurun_out/${TAG}_pytest.log; cut -c1-300 gpurun_out/${TAG}_bench_n1.json; head -12 gpurun_out/${TAG}_train_launches_summary.txt
