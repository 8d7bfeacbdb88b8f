#!/usr/bin/env python
"""Scaling sweep of the LC operators (BASELINE.json configs[4]): B in {1..65536} x N in {8, 1024, 4096}, pipelines
p1 (loss fwd+bwd), p2 (LM solve), p3 (fused).  One GPU per process.  Under torchrun (N ranks over NCCL) B is the GLOBAL batch:
every rank processes its B/N shard (batch sharding, no data-path collective), the ranks start each point together (barrier)
and rank 0 reports the MAX over ranks of the device time; points with B < N are skipped.

    python tools/sweep.py [--out profiles/sweep_r2.md] [--max-bytes 8e9]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py --out profiles/sweep_r2_g8.md

Per point: poses/s, algorithmic GB/s, fraction of the measured HBM peak.  CUDA events around `reps` back-to-back
launches after 3 warm-ups; inputs rotate over enough distinct batches to exceed L2 when they are small.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from lc_b200 import _native as nat  # noqa: E402
from lc_b200.cov_mixed import loss_fwd_bwd  # noqa: E402
from lc_b200.fused import solve_and_loss  # noqa: E402
from lc_b200.pnp.cer_solver import lm_solve  # noqa: E402
from lc_b200.synth import make_correspondences  # noqa: E402


def bytes_per_pose(p, n):
    return 28 * n + 104 if p == "p2" else 48 * n + (228 if p == "p3" else 164)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--max-bytes", type=float, default=6e9)
    ap.add_argument("--points", default="8,1024,4096")
    ap.add_argument("--bmax", type=int, default=65536)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle port (all host cores) once per N on a bounded sample")
    a = ap.parse_args()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peak = 6436.4
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    rows, cpu_rows = [], []
    for n in [int(x) for x in a.points.split(",")]:
        if a.cpu and rank == 0:
            import time
            from oracle import cpu_oracle
            cores = os.cpu_count() or 1
            sample = max(8, min(4096 if n <= 64 else 256, 16 * cores if n <= 1024 else 4 * cores))
            c = make_correspondences(sample, n, 7).to(torch.float32)
            for pipe, mode, st in (("p1", 2, c.pose), ("p2", 1, c.start), ("p3", 3, c.start)):
                arrs = [t.numpy() for t in (c.K, c.pts3d, c.pts2d, c.inv_std, c.bbox_3d, st)]
                cpu_oracle.p3(*arrs, mode=mode, threads=cores)
                t0 = time.perf_counter()
                for _ in range(3):
                    cpu_oracle.p3(*arrs, mode=mode, threads=cores)
                dt = (time.perf_counter() - t0) / 3
                cpu_rows.append((n, pipe, sample, cores, sample / dt))
                print(f"N={n:5d} CPU port {pipe}: {sample / dt:12.0f} poses/s ({cores} cores, sample {sample})", flush=True)
        Bg = 1
        while Bg <= a.bmax:
            B = Bg // world           # this rank's shard of the global batch
            if B == 0:
                Bg *= 4
                continue
            in_bytes = B * n * 28
            if in_bytes * 2 > a.max_bytes:
                break
            nrot = max(1, min(4, int(3e8 // max(in_bytes, 1)) + 1))   # rotate inputs to defeat L2 when they are small
            sets = []
            base = make_correspondences(min(B, 256), n, 7 + rank).to(torch.float32)
            rep = (B + base.pts3d.shape[0] - 1) // base.pts3d.shape[0]
            for r in range(nrot):
                t = lambda x: x.repeat((rep,) + (1,) * (x.dim() - 1))[:B].roll(r, 0).contiguous().to(dev)
                sets.append(dict(K=t(base.K), pose=t(base.pose), start=t(base.start), bbox=t(base.bbox_3d),
                                 p3=t(base.pts3d.transpose(1, 2)).transpose(1, 2), p2=t(base.pts2d.transpose(1, 2)).transpose(1, 2),
                                 s=t(base.inv_std.transpose(1, 2)).transpose(1, 2)))
            go = torch.full((B,), 1.0 / Bg, device=dev)
            for pipe in ("p1", "p2", "p3"):
                def step(d):
                    if pipe == "p1":
                        loss_fwd_bwd(d["K"], d["pose"], d["p3"], d["p2"], d["s"], None, d["bbox"], need=(True, False, True), grad_out=go)
                    elif pipe == "p2":
                        lm_solve(d["K"], d["p3"], d["p2"], d["s"], d["start"], weight_mode=nat.W_INV_STD)
                    else:
                        solve_and_loss(d["K"], d["start"], d["p3"], d["p2"], d["s"], None, d["bbox"], need=(True, False, True), grad_out=go)
                for i in range(3):
                    step(sets[i % nrot])
                reps = 20 if B * n < 2e7 else 8
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0.record()
                for i in range(reps):
                    step(sets[i % nrot])
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                if world > 1:
                    t = torch.tensor([ms], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t)
                gbs = B * world * bytes_per_pose(pipe, n) / (ms * 1e-3) / 1e9
                rows.append((n, Bg if world > 1 else B, pipe, ms * 1e3, B * world / (ms * 1e-3), gbs, gbs / (peak * world)))
                if rank == 0:
                    print(f"N={n:5d} B={B * world:6d} {pipe}: {ms * 1e3:9.1f} us  {B * world / (ms * 1e-3):12.0f} poses/s  {gbs:8.1f} GB/s  "
                          f"{100 * gbs / (peak * world):5.1f}% of HBM peak", flush=True)
            del sets
            torch.cuda.empty_cache()
            Bg *= 4
    if a.out and rank == 0:
        with open(a.out, "w") as f:
            f.write(f"# LC operator sweep ({world} x B200; CUDA events, max over ranks; launch overhead included; B = global batch, sharded over the GPUs)\n\n")
            f.write(f"HBM peak used for the last column: {world} x {peak} GB/s (MEASURED_PEAKS.json).  p1 = loss fwd+bwd, p2 = LM solve, p3 = fused.\n\n")
            f.write("| N | B | pipeline | us/launch | poses/s | algorithmic GB/s | % of HBM peak |\n|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.1f} | {r[4]:.0f} | {r[5]:.1f} | {100 * r[6]:.1f} |\n")
            if cpu_rows:
                f.write("\n## CPU oracle port on the box's host cores (OpenMP over poses; a baseline, not a target)\n\n")
                f.write("| N | pipeline | sample poses | cores | poses/s |\n|---|---|---|---|---|\n")
                for r in cpu_rows:
                    f.write(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]:.0f} |\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
