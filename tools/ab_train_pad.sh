#!/bin/bash
# A/B of the training step: library convolutions with odd input-channel counts as they are (PWCLO_PAD_CONV=0) against
# zero-padded channels (1)
mkdir -p gpurun_out
for G in 0 1 0 1; do
  PWCLO_PAD_CONV=$G python bench.py --mode train --steps 20 --warmup 5 2>/dev/null > /tmp/ab_$G.json
  python -c "import sys,json; d=json.loads([l for l in open('/tmp/ab_$G.json') if l.startswith(chr(123))][0]); print('pad_conv', $G, round(d['value'],1), round(d['ms_per_step'],3), d['loss'])" | tee -a gpurun_out/ab_train_pad.txt
done
PWCLO_PAD_CONV=1 python -m pytest tests/test_training_gpu.py -m gpu -q 2>&1 | tail -3 | tee -a gpurun_out/ab_train_pad.txt
