#!/bin/bash
# quick A/B on the GPU box: parity tests that cover the resident kernels, then the three pipelines (short runs, no CPU legs)
# usage (inside gpurun): bash tools/quick_bench.sh <tag> [extra env assignments...]
tag=$1; shift
for kv in "$@"; do export "$kv"; done
python -m pytest tests/test_bench_inputs_gpu.py tests/test_lc_gpu.py tests/test_lm_gpu.py tests/test_fused_gpu.py tests/test_edges_gpu.py -x -q 2>&1 | tail -4
for p in p3 p1 p2; do
  python bench.py --pipeline $p --steps 60 --no-cpu-baseline --no-e2e > gpurun_out/qb_${tag}_$p.json 2> gpurun_out/qb_${tag}_$p.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/qb_${tag}_$p.json"))
    print("$tag $p %.1f us  %s" % (d["ms_per_step"] * 1e3, d["roofline"]["kernel"]))
except Exception as e:
    print("$tag $p FAILED", e); print(open("gpurun_out/qb_${tag}_$p.err").read()[-1500:])
PY
done
