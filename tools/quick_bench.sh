#!/bin/bash
# quick kernel-only timings of the three pipelines (ms per 1024x4096 batch); extra env vars are passed through
for p in p1 p2 p3; do
  python bench.py --pipeline $p --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$p', round(d['ms_per_step']*1000,1),'us', d['clocks'].get('sm_mhz'))"
done
