#!/bin/bash
# A/B of the training step: compact Hamilton product (PWCLO_HAMILTON_FEW 0 / 16) x channel-padding threshold
mkdir -p gpurun_out
: > gpurun_out/ab_train_small_ops.txt
for cfg in "0 40000" "16 40000" "16 8192" "16 2048" "0 40000" "16 8192"; do
  set -- $cfg
  PWCLO_HAMILTON_FEW=$1 PWCLO_PAD_CONV_MIN=$2 python bench.py --mode train --steps 20 --warmup 5 2>/dev/null > /tmp/ab.json
  python -c "import sys,json; d=json.loads([l for l in open('/tmp/ab.json') if l.startswith(chr(123))][0]); print('hamilton_few', $1, 'pad_min', $2, round(d['value'],1), round(d['ms_per_step'],3), d['loss'])" | tee -a gpurun_out/ab_train_small_ops.txt
done
python -m pytest tests/test_training_gpu.py tests/test_layers_gpu.py -m gpu -q 2>&1 | tail -3 | tee -a gpurun_out/ab_train_small_ops.txt
