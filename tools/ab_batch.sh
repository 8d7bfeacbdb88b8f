#!/bin/bash
# A/B of batch sizes / env switches on the GPU box: bash tools/ab_batch.sh "<pipelines>" "<tag:batch:ENV=val,ENV=val> ..."
for spec in $2; do
  IFS=: read tag batch envs <<< "$spec"
  for p in $1; do
    ( IFS=,; for kv in $envs; do export "$kv"; done; python bench.py --pipeline $p --batch $batch --steps 60 --no-cpu-baseline --no-e2e $BENCH_EXTRA > gpurun_out/ab_${tag}_$p.json 2>/dev/null )
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${tag}_$p.json")); print("$tag $p B=$batch %.1f us  %s" % (d["ms_per_step"] * 1e3, d["roofline"]["kernel"]))
except Exception as e: print("$tag $p FAILED", e)
PY
  done
done
