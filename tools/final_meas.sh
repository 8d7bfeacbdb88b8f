#!/bin/bash
# Round-end measurement batch (run on the GPU box, 1 GPU): tests, smoke, sanitizer, bench lines, launch list, ncu captures, row benches.
# usage: bash tools/final_meas.sh [ncu]
set -x
R=r2
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
for t in memcheck racecheck initcheck; do echo "== $t"; timeout 900 compute-sanitizer --tool $t python tools/sanitize_smoke.py 2>&1 | grep -E "SUMMARY|smoke done" ; done > gpurun_out/sanitizer_$R.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2> gpurun_out/b_ref.err
python bench.py --pipeline p3 > gpurun_out/bench_${R}_p3.json 2> gpurun_out/b_p3.err
python bench.py --pipeline p1 > gpurun_out/bench_${R}_p1.json 2> gpurun_out/b_p1.err
python bench.py --pipeline p2 > gpurun_out/bench_${R}_p2.json 2> gpurun_out/b_p2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_final_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_launch.log 2>&1
if [ "$1" = "ncu" ]; then
for p in p3 p1 p2; do
ncu --set full --clock-control none --import-source on -k regex:lc_resident -s 6 -c 1 -f -o gpurun_out/prof_${R}_final_$p python bench.py --pipeline $p --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu_$p.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:lc_select -s 2 -c 1 -f -o gpurun_out/prof_${R}_select python tools/bench_producers.py --no-cpu > gpurun_out/b_ncu_select.log 2>&1
fi
python tools/bench_tiny.py --out gpurun_out/bench_tiny_$R.json > gpurun_out/b_tiny.log 2>&1
python tools/bench_reference_gpu.py --out gpurun_out/reference_gpu_$R.json > gpurun_out/b_refgpu.log 2>&1
rm -f gpurun_out/train_step_$R.jsonl
python tools/train_step.py --config glmo --steps 20 --out gpurun_out/train_step_$R.jsonl > gpurun_out/b_train_glmo.log 2>&1
python tools/train_step.py --config zycbv --steps 20 --out gpurun_out/train_step_$R.jsonl > gpurun_out/b_train_zycbv.log 2>&1
python tools/sweep.py --cpu --out gpurun_out/sweep_$R.md > gpurun_out/b_sweep.log 2>&1
python tools/bench_producers.py --out gpurun_out/bench_producers_$R.json > gpurun_out/b_prod.log 2>&1
python tools/bench_dense.py > gpurun_out/bench_dense_$R.json 2> gpurun_out/b_dense.err
python bench.py --pipeline p3 --lm-mixed --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_${R}_p3_mixed.json 2>/dev/null
python bench.py --pipeline p2 --lm-mixed --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/bench_${R}_p2_mixed.json 2>/dev/null
bash tools/ab_batch.sh "p3" "cluster64:64:LC_B200_SPLIT=1 single64:64:LC_B200_SPLIT=0 cluster1024:1024:LC_B200_SPLIT=1 single1024:1024:LC_B200_SPLIT=0" > gpurun_out/cluster_split_${R}.txt 2>&1
LC_B200_PERSIST=1 python bench.py --pipeline p3 --steps 60 --no-cpu-baseline --no-e2e > gpurun_out/bench_${R}_p3_persist.json 2>/dev/null
LC_B200_PERSIST=1 python bench.py --pipeline p1 --steps 60 --no-cpu-baseline --no-e2e > gpurun_out/bench_${R}_p1_persist.json 2>/dev/null
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/sanitizer_$R.txt
tail -c 400 gpurun_out/bench_${R}_p3.json
