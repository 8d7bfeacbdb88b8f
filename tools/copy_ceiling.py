#!/usr/bin/env python
"""Host <-> device copy ceiling of the box with every rank copying at once: what bounds bench.py's `e2e` line at N GPUs.

Each rank (one per GPU, torchrun) binds to the CPUs local to its GPU exactly like bench.py, allocates pinned host buffers of the
sizes one e2e step moves (117.6 MB host -> device, 83.9 MB device -> host at B = 1024 x N = 4096) and times, with CUDA events and
a barrier on both sides, (a) the H2D copies alone, (b) the D2H copies alone, (c) both directions on two streams — no kernel at
all.  Rank 0 prints one JSON line with per-GPU and aggregate GB/s (max-over-ranks time).  If the aggregate of (c) at 8 GPUs is
not ~8x the 1-GPU figure, the shortfall of `e2e` is the box's host memory / PCIe fabric, not this library.

    python tools/copy_ceiling.py                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/copy_ceiling.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bind_to_gpu_numa_node  # noqa: E402

H2D_BYTES, D2H_BYTES, REPS = 117_604_352, 83_918_848, 20


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    numa = bind_to_gpu_numa_node(local)
    h_in = torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(H2D_BYTES, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(D2H_BYTES, dtype=torch.uint8, device=dev)
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(up, down):
        for _ in range(3):   # warm-up
            if up:
                with torch.cuda.stream(s_up):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    h_out.copy_(d_out, non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_event(e0); s_down.wait_event(e0)
        for _ in range(REPS):
            if up:
                with torch.cuda.stream(s_up):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    h_out.copy_(d_out, non_blocking=True)
        cur = torch.cuda.current_stream()
        cur.wait_stream(s_up); cur.wait_stream(s_down)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e-3 / REPS

    t_up, t_down, t_both = timed(True, False), timed(False, True), timed(True, True)
    if rank == 0:
        gbs = lambda nbytes, t: nbytes / t / 1e9
        print(json.dumps({
            "n_gpus": world, "numa": numa, "h2d_bytes": H2D_BYTES, "d2h_bytes": D2H_BYTES, "reps": REPS,
            "h2d_alone_gbs_per_gpu": gbs(H2D_BYTES, t_up), "d2h_alone_gbs_per_gpu": gbs(D2H_BYTES, t_down),
            "both_h2d_gbs_per_gpu": gbs(H2D_BYTES, t_both), "both_d2h_gbs_per_gpu": gbs(D2H_BYTES, t_both),
            "both_aggregate_gbs": world * gbs(H2D_BYTES + D2H_BYTES, t_both),
            "e2e_ceiling_poses_per_s": world * 1024 / t_both,
            "note": "copies only, every rank at once, max-over-ranks time; e2e_ceiling = 1024 poses per (H2D + D2H of one step) per GPU"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
