#!/bin/bash
# launch list (durations) of one eager training step, final code.   usage: gpurun -- bash tools/ncu_train_list.sh TAG
TAG=${1:-round2_train_final2}
mkdir -p gpurun_out
timeout 300 python tools/ncu_train_step.py > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/ncu_train_step.py > gpurun_out/${TAG}_ncu.log 2>&1; echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt
head -45 gpurun_out/${TAG}_launches_summary.txt
