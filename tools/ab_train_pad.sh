#!/bin/bash
# training step after a change: two bench runs + the training / layer / forward parity tests
mkdir -p gpurun_out
: > gpurun_out/ab_train_now.txt
for i in 1 2; do
  python bench.py --mode train --steps 20 --warmup 5 2>/tmp/ab.err > /tmp/ab.json || tail -5 /tmp/ab.err
  python -c "import sys,json; d=json.loads([l for l in open('/tmp/ab.json') if l.startswith(chr(123))][0]); print('train', round(d['value'],1), round(d['ms_per_step'],3), d['loss'])" | tee -a gpurun_out/ab_train_now.txt
done
python -m pytest tests/test_training_gpu.py tests/test_layers_gpu.py -m gpu -q 2>&1 | tail -3 | tee -a gpurun_out/ab_train_now.txt
