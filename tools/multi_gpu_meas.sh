#!/bin/bash
# Multi-GPU measurement batch (run under `gpurun --gpus N`): weak + strong scaling of the headline op, the B x N sweep with the
# global batch sharded over the ranks (BASELINE.json configs[4]) and the zycbv train step under DDP (configs[3]).
# usage: bash tools/multi_gpu_meas.sh N
N=$1
R=r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
$TR bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/scale_weak_${R}_g$N.json 2> gpurun_out/b_scale_weak_$N.err
$TR bench.py --gpus $N --steps 200 --warmup 5 --scaling strong --no-e2e > gpurun_out/scale_strong_${R}_g$N.json 2> gpurun_out/b_scale_strong_$N.err
$TR tools/sweep.py --points 8,1024,4096 --out gpurun_out/sweep_${R}_g$N.md > gpurun_out/b_sweep_$N.log 2>&1
$TR tools/train_step.py --config zycbv --ddp --steps 20 --arms reference,ours,fused,nolc --no-probe64 --out gpurun_out/train_step_ddp_${R}.jsonl > gpurun_out/b_train_ddp_$N.log 2>&1
tail -c 600 gpurun_out/scale_weak_${R}_g$N.json; echo; tail -c 300 gpurun_out/scale_strong_${R}_g$N.json; echo; tail -3 gpurun_out/b_sweep_$N.log; tail -c 700 gpurun_out/b_train_ddp_$N.log
