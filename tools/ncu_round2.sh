#!/bin/bash
# ncu evidence, round 2 (one GPU; each pass only after the plain program exited 0 without ncu):
#  (1) launch list (duration) of one forward, (2) DRAM bytes / pipe utilisation of every kernel of that forward,
#  (3) --set full of the dominant kernel families (exported to CSV on the box).
TAG=${1:-round2}
mkdir -p gpurun_out
CMD="python tools/ncu_forward.py 64"
$CMD > gpurun_out/${TAG}_ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_ncu_plain.log; exit 1; }
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "launch list exit $?"
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_forward_metrics.csv \
    $CMD > gpurun_out/${TAG}_ncu_metrics.log 2>&1; echo "metrics exit $?"
python tools/traffic_from_ncu.py gpurun_out/${TAG}_forward_metrics.csv gpurun_out/${TAG}_traffic.json
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k 'regex:knn_slab_kernel|tc_mlp_kernel|fps_slab_kernel|set_conv_small' -c 40 -f -o /tmp/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full exit $?"
if [ -f /tmp/${TAG}_full.ncu-rep ]; then
  ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/${TAG}_full.ncu-rep > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1
  sz=$(stat -c %s /tmp/${TAG}_full.ncu-rep); echo "report bytes $sz"
  if [ "$sz" -lt 40000000 ]; then cp /tmp/${TAG}_full.ncu-rep gpurun_out/; fi
fi
head -20 gpurun_out/${TAG}_launches_summary.txt; du -sh gpurun_out
