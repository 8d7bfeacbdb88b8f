#!/usr/bin/env python
"""Timings of the rows around the hot path on one B200: the byte-streaming kernels against the HBM
roofline (algorithmic bytes / CUDA-event time / MEASURED_PEAKS.json), the test-time chain against the reference-shaped CPU
path (OpenCV RANSAC-EPnP as lib/pnp/cv2_solver.py calls it + the CPU LM oracle).  Writes one JSON object per line.

    python tools/bench_producers.py [--out profiles/bench_producers_r1.json] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lc_b200 import _native as nat  # noqa: E402
from lc_b200.synth import make_dense_outputs, make_zebra_outputs, make_correspondences  # noqa: E402


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6436.4, "fallback (B200_PROFILING.md)"


def timeit(fn, n_sets, steps=50, warmup=5):
    """CUDA-event time per call, rotating over n_sets input sets (> L2 in total where the caller sized them so)."""
    for i in range(warmup):
        fn(i % n_sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i % n_sets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e-3


def rep(d, k):
    return {kk: (torch.cat([v] * k) if torch.is_tensor(v) else v) for kk, v in d.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    peak, src = hbm_peak()
    lines = []

    def emit(**kw):
        kw.setdefault("gpu", torch.cuda.get_device_name(0))
        lines.append(kw)
        print(json.dumps(kw), flush=True)

    # ---- f3: test-time ZebraPose decode, B x 21 x 128 x 128 logits -> xyz (pure streaming) ----
    from lc_b200.floatbits import nn_out_to_xyz
    B, H, W, bits = 256, 128, 128, (7, 7, 7)
    sets = []
    for s in range(3):
        d = make_zebra_outputs(8, H, W, 50 + s, bits)
        sets.append(dict(lg=torch.cat([d["bin_logits"]] * (B // 8)).cuda(), ns=torch.cat([d["noc_scale"]] * (B // 8)).cuda(),
                         T=torch.cat([d["model_transform"]] * (B // 8)).cuda(), out=torch.empty(B, H, W, 3, device="cuda")))
    t = timeit(lambda i: nn_out_to_xyz(sets[i]["lg"], sets[i]["ns"], model_transform=sets[i]["T"], bit_cnt=bits, out=sets[i]["out"]), 3)
    nbytes = B * H * W * (sum(bits) * 4 + 12)
    emit(kernel="lc_decode_kernel (f3 test-time decode)", workload=f"B={B} x {sum(bits)} x {H} x {W} logits (3 x {nbytes >> 20} MiB sets)",
         us=t * 1e6, algorithmic_bytes=nbytes, GBps=nbytes / t / 1e9, hbm_peak=peak, frac=nbytes / t / 1e9 / peak, peak_source=src)

    # ---- f2a: point selection, B x 128 x 128, sample 1, quantile_in_mask, fused softmax ----
    from lc_b200.select import dense_point_select
    B, H, W = 256, 128, 128
    sets = []
    for s in range(3):
        d = rep(make_dense_outputs(8, H, W, 60 + s), B // 8)
        g = torch.Generator().manual_seed(s)
        sets.append(dict(xyz=d["xyz_noc"].cuda().permute(0, 2, 3, 1), ns=d["noc_scale"].cuda(), lg=d["logits"].cuda(), sc=d["scale"].cuda(),
                         ml=(2 * torch.randn(B, 1, H, W, generator=g) + 0.8).cuda()))
    t = timeit(lambda i: dense_point_select(sets[i]["xyz"], sets[i]["ml"], xyz_weight_logits=sets[i]["lg"], xyz_weights_scale=sets[i]["sc"],
                                            noc_scale=sets[i]["ns"], sample=1, dense_point_select="quantile_in_mask"), 3)
    nbytes = B * H * W * (3 * 4 + 2 * 4 + 4 + 28)       # xyz + logits + mask in; 28 B per output slot (pts3d, pts2d, inv_cov)
    emit(kernel="lc_select_kernel (f2 point selection)", workload=f"B={B} x {H} x {W}, sample 1, quantile_in_mask, softmax fused",
         us=t * 1e6, algorithmic_bytes=nbytes, GBps=nbytes / t / 1e9, hbm_peak=peak, frac=nbytes / t / 1e9 / peak, peak_source=src,
         note="logits are read twice (max; exp + sum + quantile operand in one pass) and the quantile is a 3-pass radix select in shared memory")

    # ---- f3: zebrapose training producer fused with the LC loss (zycbv: B=32, 128x128, sample 3) and a large batch ----
    from lc_b200.dense import dense_loss_fwd_bwd
    for B, H, W, sample, bits in ((32, 128, 128, 3, (7, 7, 6)), (1024, 64, 64, 1, (7, 7, 6))):
        sets = []
        for s in range(2 if B > 32 else 4):
            d = make_zebra_outputs(8, H, W, 70 + s, bits)
            k = B // 8
            sets.append({kk: (torch.cat([v] * k).cuda() if torch.is_tensor(v) else v) for kk, v in d.items()})
        t = timeit(lambda i: dense_loss_fwd_bwd(None, sets[i]["logits"], sets[i]["scale"], sets[i]["noc_scale"], sets[i]["K"], sets[i]["pose"],
                                                sets[i]["bbox_3d"], sample=sample, top_left=(0, 0), noc_bin_logits=sets[i]["bin_logits"],
                                                noc_bin_raw=sets[i]["raw_bits"], msk_noc=sets[i]["msk_noc"], bit_cnt=bits,
                                                model_transform=sets[i]["model_transform"]), len(sets), steps=30)
        C_ = sum(bits)
        nbytes = B * H * W * (2 * 4 * 2 + C_ * 4 * 2) + B * (H // sample) * (W // sample) * C_      # logits in+grad, bit logits in+grad, raw bits
        emit(kernel="lc_dense_kernel<ZEBRA> (f3 training producer + LC loss fwd+bwd)", workload=f"B={B} x {C_} x {H} x {W}, sample {sample}",
             us=t * 1e6, poses_per_s=B / t, algorithmic_bytes=nbytes, GBps=nbytes / t / 1e9, frac=nbytes / t / 1e9 / peak)

    # ---- f4: ADD/ADI/re/te ----
    from lc_b200.evaluate import compute_pose_errors
    from lc_b200.synth import quat_to_matrix
    B, M = 1024, 4096
    c = make_correspondences(B, 4, 3)
    pts = (torch.rand(M, 3, dtype=torch.float64) - 0.5) * 100
    Re, Rg = quat_to_matrix(c.start[:, :4]).cuda(), quat_to_matrix(c.pose[:, :4]).cuda()
    te, tg, ptc = c.start[:, 4:].cuda(), c.pose[:, 4:].cuda(), pts.cuda()
    t = timeit(lambda i: compute_pose_errors(Re, te, Rg, tg, ptc), 1, steps=10, warmup=2)
    line = dict(kernel="lc_pose_errors_kernel (f4 ADD/ADI/re/te)", workload=f"B={B} poses x M={M} model points (brute-force ADI)", us=t * 1e6,
                poses_per_s=B / t)
    if not a.no_cpu:
        sys.path.insert(0, ROOT)
        from oracle import cpu_oracle
        t0 = time.perf_counter()
        cpu_oracle.pose_errors(Re[:4].cpu().numpy(), te[:4].cpu().numpy(), Rg[:4].cpu().numpy(), tg[:4].cpu().numpy(), pts.numpy())
        line["cpu_port_poses_per_s_1core"] = 4 / (time.perf_counter() - t0)
    emit(**line)

    # ---- f2: the whole test-time chain (selection -> initialiser -> weighted LM), B=32 and B=1024, 64x64, sample 2 ----
    from lc_b200.select import solve_pnp_dense
    for B in (32, 1024):
        d = rep(make_dense_outputs(8, 64, 64, 80), B // 8)
        cu = {k: v.cuda() for k, v in d.items()}
        ml = torch.full((B, 1, 64, 64), 3.0, device="cuda")
        xyz = cu["xyz_noc"].permute(0, 2, 3, 1)
        fn = lambda i: solve_pnp_dense(cu["K"], xyz, ml, cu["logits"], cu["scale"], None, noc_scale=cu["noc_scale"], sample=2,
                                       dense_point_select="quantile_in_mask", solvers=("weighted",))
        t = timeit(fn, 1, steps=20, warmup=3)
        line = dict(kernel="test-time chain: select + init + LM (3 launches)", workload=f"B={B}, 64x64, sample 2, quantile_in_mask", us=t * 1e6,
                    poses_per_s=B / t)
        if not a.no_cpu:
            # reference-shaped CPU path on a bounded sample: OpenCV RANSAC-EPnP per pose (cv2_solver.py:69-88) + the CPU LM oracle
            import cv2
            from oracle import cpu_oracle
            res, sel = fn(0)
            nb = min(B, 16)
            n = sel["n_points"][:nb].cpu().numpy()
            X, x, ic = (sel[k][:nb].cpu().numpy() for k in ("pts3d", "pts2d", "inv_cov"))
            Kc = cu["K"][:nb].cpu().numpy()
            t0 = time.perf_counter()
            starts = np.zeros((nb, 7), np.float32)
            for b in range(nb):
                ok, rvec, tvec, inl = cv2.solvePnPRansac(X[b, :n[b]], x[b, :n[b]], Kc[b], None, flags=cv2.SOLVEPNP_EPNP, confidence=0.99,
                                                         iterationsCount=150, reprojectionError=3.0)
                th = np.linalg.norm(rvec)
                starts[b] = np.concatenate(([np.cos(th / 2)], (rvec[:, 0] / max(th, 1e-12)) * np.sin(th / 2), tvec[:, 0]))
            t_cv = time.perf_counter() - t0
            t0 = time.perf_counter()
            for b in range(nb):
                L = np.zeros((1, n[b], 2, 2), np.float32)
                L[0, :, 0, 0], L[0, :, 1, 1] = np.sqrt(ic[b, :n[b], 0]), np.sqrt(ic[b, :n[b], 1])
                cpu_oracle.lm_solve(Kc[b:b + 1], X[b:b + 1, :n[b]], x[b:b + 1, :n[b]], L, starts[b:b + 1], threads=1)
            t_lm = time.perf_counter() - t0
            line.update(cpu_sample_poses=nb, cpu_opencv_ransac_epnp_poses_per_s=nb / t_cv, cpu_lm_port_poses_per_s_1core=nb / t_lm,
                        cpu_chain_poses_per_s_1core=nb / (t_cv + t_lm), cpu_kind="OpenCV as the reference calls it + LM oracle port")
        emit(**line)

    if a.out:
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
