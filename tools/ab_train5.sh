#!/bin/bash
python -m pytest tests/test_training_gpu.py -m gpu -q 2>&1 | tail -3
for P in 0 1; do
  PWCLO_COST_GEO=$P python bench.py --mode train --steps 10 --warmup 3 2>/tmp/err_$P.txt > /tmp/tr_$P.json
  python -c "import json; d=json.loads([l for l in open('/tmp/tr_$P.json') if l.startswith(chr(123))][0]); print('own cost geometry', $P, d['value'], d['ms_per_step'], d['loss'])" || tail -5 /tmp/err_$P.txt
done
