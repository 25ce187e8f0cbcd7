#!/bin/bash
# Full counter set for every kernel of one whole forward (no source pages: keeps the report small).
TAG=${1:-r1e}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none -k 'regex:fps|knn|tc_mlp|set_conv|pose_head|gather_rows|transpose_cn' -s 264 -c 70 -f \
    -o gpurun_out/${TAG}_full_forward python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -4 gpurun_out/${TAG}_ncu_full.log
