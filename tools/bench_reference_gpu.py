#!/usr/bin/env python
"""The reference's own PyTorch-op graph on the SAME B200: Loss_cov_mixed forward + backward (lib/cov_mixed.py:100-150 with its
functorch vmap / jacfwd sweeps, files staged unmodified under baseline/_ref) timed beside lc_b200's P1 on identical inputs.
This is the apples-to-apples baseline for P1, the only fwd+bwd the reference has (SURVEY.md §8d "reference timed beside it (i)").

    python tools/bench_reference_gpu.py [--out profiles/reference_gpu_r2.json]
Shapes: (B=32, N=1024) = the glmo training call; (B=1024, N=4096) = BASELINE.json configs[1], run in chunks of `--chunk` poses
because the reference materialises (B,N,2,6,6) Hessians and their tangents (tens of GB at B=1024).
"""
import argparse
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path[:0] = [ROOT, REF]
warnings.filterwarnings("ignore")


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "lib", "cov_mixed.py")):
        raise SystemExit("baseline/_ref is not staged: run `python tools/stage_reference.py` in the build container")
    from lib.cov_mixed import Loss_cov_mixed as ref_lc
    from lc_b200.cov_mixed import Loss_cov_mixed as our_lc
    from lc_b200.synth import make_correspondences, planar_view
    rows = []
    for B, N, reps in ((32, 1024, 10), (1024, 4096, 2)):
        c = make_correspondences(B, N, 10).to(torch.float32).to(device="cuda")
        X, s = planar_view(c.pts3d), planar_view(c.inv_std)

        def run(lc, chunk):
            tot = None
            for i in range(0, B, chunk):
                sl = slice(i, i + chunk)
                p3 = X[sl].detach().requires_grad_(True)
                w = s[sl].detach().requires_grad_(True)
                loss = lc(c.K[sl], c.pose[sl], p3, c.pts2d[sl], w, c.valid[sl], bbox_3d=c.bbox_3d[sl], max_err_len=32)
                loss.sum().backward()
                tot = (p3.grad, w.grad, loss.detach())
            return tot
        chunk = min(B, a.chunk)
        ms_ref = timed(lambda: run(ref_lc, chunk), reps)
        ms_our = timed(lambda: run(our_lc, B), max(reps, 20))
        gr, go = run(ref_lc, chunk), run(our_lc, B)
        last = slice(B - chunk if B > chunk else 0, B)
        rel = float((gr[0] - go[0][last]).norm() / gr[0].norm()) if B > chunk else float((gr[0] - go[0]).norm() / gr[0].norm())
        rows.append(dict(B=B, N=N, reference_ms=ms_ref, reference_poses_per_s=B / ms_ref * 1e3, reference_chunk=chunk,
                         lc_b200_ms=ms_our, lc_b200_poses_per_s=B / ms_our * 1e3, speedup=ms_ref / ms_our,
                         grad_pts3d_rel_diff=rel, peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9))
        print(rows[-1], flush=True)
    line = dict(gpu=torch.cuda.get_device_name(0), what="reference Loss_cov_mixed (PyTorch/functorch, fp32) vs lc_b200 Loss_cov_mixed, fwd+bwd through autograd, same B200", rows=rows)
    if a.out:
        json.dump(line, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
