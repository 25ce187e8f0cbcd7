#!/bin/bash
python -m pytest tests/test_ops_gpu.py -m gpu -q -k "knn" 2>&1 | tail -3
python -m pytest tests/test_training_gpu.py tests/test_forward_gpu.py -m gpu -q 2>&1 | tail -2
python bench.py --mode train --steps 10 --warmup 3 2>/dev/null > /tmp/tr.json
python -c "import json; d=json.loads([l for l in open('/tmp/tr.json') if l.startswith(chr(123))][0]); print('train', d['value'], d['ms_per_step'])"
python tools/time_forward.py 2 2>/dev/null | grep -E "knn_sorted\[B4 S2048 N8192|total" 
