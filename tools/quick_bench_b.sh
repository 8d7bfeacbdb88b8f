#!/bin/bash
for p in p1 p2 p3; do
  python bench.py --pipeline $p --batch ${BATCH:-1024} --steps ${STEPS:-50} --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$p', round(d['ms_per_step']*1000,1),'us', round(d['value']))"
done
