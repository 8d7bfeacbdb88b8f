#!/usr/bin/env python
"""Sparse-keypoint regime (N = 8 / 16, configs/gsplmo.yaml): poses/s of the three pipelines at large batch.
usage: python tools/bench_tiny.py [--out f.json]"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lc_b200.synth import make_correspondences
from lc_b200.fused import solve_and_loss
from lc_b200.cov_mixed import loss_fwd_bwd
from lc_b200.pnp.cer_solver import lm_solve
from lc_b200 import _native as nat

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); a = ap.parse_args()
rows = []
for N in (8, 16, 32):
    for B in (4096, 65536):
        c = make_correspondences(B, N, 10).to(torch.float32).to(device="cuda")
        fns = dict(p1=lambda: loss_fwd_bwd(c.K, c.pose, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d, need=(True, True, True)),
                   p2=lambda: lm_solve(c.K, c.pts3d, c.pts2d, c.inv_std, c.start, weight_mode=nat.W_INV_STD),
                   p3=lambda: solve_and_loss(c.K, c.start, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d, need=(True, True, True)))
        for name, fn in fns.items():
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rows.append(dict(N=N, B=B, pipeline=name, us=ms * 1e3, mposes_per_s=B / ms / 1e3, kernel=nat.lib().lc_b200_last_kernels().decode()))
            print(rows[-1], flush=True)
if a.out: json.dump(rows, open(a.out, "w"), indent=1)
