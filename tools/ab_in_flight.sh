#!/bin/bash
# forwards in flight (PWCLO_STREAMS) x pairs per GPU, one B200; every stream's graph captured before the timed bracket
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out/ab_in_flight.txt
: > $OUT
for pairs in 1 4 8 16 32 64; do
  for nf in 2 3 4 6 8; do
    echo "== pairs=$pairs in_flight=$nf" >> $OUT
    PWCLO_STREAMS=$nf timeout 120 python bench.py --total-pairs $pairs --extras none \
      --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l)
        print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'lat', round(d['config'].get('forward_latency_ms') or 0, 3), 'attempts', d['config'].get('bracket_attempts_ms_per_step'))
" >> $OUT
  done
done
cat $OUT
