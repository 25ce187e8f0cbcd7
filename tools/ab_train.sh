#!/bin/bash
# A/B of the training step: PWCLO_GRAD_COLLECT = 0 (accumulate into the arena views) against 1 (multi-tensor gather)
python -m pytest tests/test_training_gpu.py -m gpu -q 2>&1 | tail -3
for G in 0 1; do
  PWCLO_GRAD_COLLECT=$G python bench.py --mode train --steps 10 --warmup 3 2>/dev/null > /tmp/ab_$G.json
  python -c "import sys,json; d=json.loads([l for l in open('/tmp/ab_$G.json') if l.startswith(chr(123))][0]); print('collect', $G, d['value'], d['ms_per_step'], d['loss'])"
done
