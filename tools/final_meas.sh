#!/bin/bash
# Round-end measurement batch (run on the GPU box): tests, smoke, sanitizer, bench lines, launch list, ncu captures, row benches.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
for t in memcheck racecheck initcheck; do echo "== $t"; timeout 600 compute-sanitizer --tool $t python tools/sanitize_smoke.py 2>&1 | grep -E "SUMMARY|smoke done" ; done > gpurun_out/sanitizer.log 2>&1
python bench.py --pipeline p3 > gpurun_out/bench_r1_p3.json 2> gpurun_out/b_p3.err
python bench.py --pipeline p1 > gpurun_out/bench_r1_p1.json 2> gpurun_out/b_p1.err
python bench.py --pipeline p2 > gpurun_out/bench_r1_p2.json 2> gpurun_out/b_p2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/b_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_launch.log 2>&1
if [ "$1" = "ncu" ]; then
for p in p3 p1 p2; do
ncu --set full --clock-control none --import-source on -k regex:lc_resident -s 3 -c 1 -f -o gpurun_out/prof_r1_final_$p python bench.py --pipeline $p --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu_$p.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:lc_decode -c 1 -f -o gpurun_out/prof_r1_decode python tools/bench_producers.py --no-cpu > gpurun_out/b_ncud.log 2>&1
fi
python tools/bench_producers.py --out gpurun_out/bench_producers_r1.json > gpurun_out/b_prod.log 2>&1
python tools/bench_dense.py > gpurun_out/bench_dense_r1.json 2> gpurun_out/b_dense.err
python tools/bench_chain_components.py > gpurun_out/bench_chain_components_r1.txt 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/sanitizer.log
tail -c 300 gpurun_out/bench_r1_p3.json
