set -x
mkdir -p gpurun_out
python bench.py --pipeline p3 > gpurun_out/bench_r1_p3.json 2> gpurun_out/b_p3.err
python bench.py --pipeline p1 > gpurun_out/bench_r1_p1.json 2> gpurun_out/b_p1.err
python bench.py --pipeline p2 > gpurun_out/bench_r1_p2.json 2> gpurun_out/b_p2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/b_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lc_resident -s 3 -c 1 -f -o gpurun_out/prof_r1_final_p3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lc_resident -s 3 -c 1 -f -o gpurun_out/prof_r1_final_p1 python bench.py --pipeline p1 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lc_resident -s 3 -c 1 -f -o gpurun_out/prof_r1_final_p2 python bench.py --pipeline p2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lc_decode -c 1 -f -o gpurun_out/prof_r1_decode python tools/bench_producers.py --no-cpu > gpurun_out/b_ncud.log 2>&1
python tools/bench_producers.py --out gpurun_out/bench_producers_r1.json > gpurun_out/b_prod.log 2>&1
python tools/bench_dense.py > gpurun_out/bench_dense_r1.json 2> gpurun_out/b_dense.err
tail -c 600 gpurun_out/bench_r1_p3.json
