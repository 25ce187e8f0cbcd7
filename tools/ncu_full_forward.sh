#!/bin/bash
# ncu evidence for one whole forward, sized to stay far below gpurun's 64 MiB return limit:
#  (1) DRAM bytes + duration + pipe utilisation of EVERY kernel of one forward (few passes, CSV only);
#  (2) the full counter set of the dominant kernel families, exported to CSV on the box (the .ncu-rep
#      is kept only when it is small).
TAG=${1:-r1e}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
K='regex:fps|knn|tc_mlp|set_conv|pose_head|gather_rows|transpose_cn'
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics $M --clock-control none -k "$K" -s 264 -c 70 --csv --log-file gpurun_out/${TAG}_forward_metrics.csv \
    $CMD > gpurun_out/${TAG}_ncu_metrics.log 2>&1
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:knn_slab_kernel|tc_mlp_kernel|fps_slab_kernel|set_conv_small' \
    -s 140 -c 24 -f -o /tmp/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
if [ -f /tmp/${TAG}_full.ncu-rep ]; then
  ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/${TAG}_full.ncu-rep > gpurun_out/${TAG}_full_summary.txt 2>&1
  sz=$(stat -c %s /tmp/${TAG}_full.ncu-rep); echo "report bytes $sz"
  if [ "$sz" -lt 30000000 ]; then cp /tmp/${TAG}_full.ncu-rep gpurun_out/; fi
fi
du -sh gpurun_out; tail -2 gpurun_out/${TAG}_ncu_full.log
