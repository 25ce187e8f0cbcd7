#!/bin/bash
# training step A/B: attentive pooling composed by torch (PWCLO_SOFTMAX_POOL=0) against the fused op (1) + parity tests
mkdir -p gpurun_out
: > gpurun_out/ab_train_softmax_pool.txt
python -m pytest tests/test_training_gpu.py tests/test_layers_gpu.py -m gpu -q 2>&1 | tail -4 | tee -a gpurun_out/ab_train_softmax_pool.txt
for G in 0 1 0 1; do
  PWCLO_SOFTMAX_POOL=$G python bench.py --mode train --steps 20 --warmup 5 2>/tmp/ab.err > /tmp/ab.json || tail -5 /tmp/ab.err
  python -c "import sys,json; d=json.loads([l for l in open('/tmp/ab.json') if l.startswith(chr(123))][0]); print('softmax_pool', $G, round(d['value'],1), round(d['ms_per_step'],3), d['loss'])" | tee -a gpurun_out/ab_train_softmax_pool.txt
done
