#!/usr/bin/env python
"""Fused dense producer vs the unfused glue (torch softmax / scale / strided views + lc_b200 Loss_cov_mixed + autograd).
Config C1 of SURVEY.md §8 (glmo train: B=32, 64x64 head, dense_sample=2 -> N=1024) and larger batches."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc_b200.dense import dense_loss_fwd_bwd
from lc_b200.cov_mixed import Loss_cov_mixed
from lc_b200.synth import make_dense_outputs as _inputs

def timeit(fn, reps):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

rows = []
for (B, H, W, sample) in [(32, 64, 64, 2), (32, 64, 64, 1), (32, 128, 128, 3), (1024, 64, 64, 2), (1024, 64, 64, 1)]:
    base = _inputs(min(B, 64), H, W, 1)
    rep = (B + 63) // 64
    d = {k: v.repeat((rep,) + (1,) * (v.dim() - 1))[:B].contiguous().cuda() for k, v in base.items()}
    go = torch.full((B,), 1.0 / B, device="cuda")
    fused = lambda: dense_loss_fwd_bwd(d["xyz_noc"], d["logits"], d["scale"], d["noc_scale"], d["K"], d["pose"], d["bbox_3d"],
                                       sample=sample, top_left=(0, 0), grad_out=go)
    ys, xs = torch.meshgrid(torch.arange(H, device="cuda", dtype=torch.float32), torch.arange(W, device="cuda", dtype=torch.float32), indexing="ij")
    uv = torch.stack((xs, ys), -1)[0::sample, 0::sample].reshape(-1, 2)
    def unfused():
        xyz = d["xyz_noc"].detach().requires_grad_(True); lg = d["logits"].detach().requires_grad_(True); sc = d["scale"].detach().requires_grad_(True)
        w = lg.reshape(B, 1, -1).softmax(-1).reshape_as(lg) * sc
        inv_std = w[..., 0::sample, 0::sample].flatten(-2).mT
        p3 = xyz[..., 0::sample, 0::sample].flatten(-2).mT * d["noc_scale"].unsqueeze(-2)
        loss = Loss_cov_mixed(d["K"], d["pose"], p3, uv.expand(B, -1, -1), inv_std, torch.ones_like(p3[..., 0]), bbox_3d=d["bbox_3d"], max_err_len=32)
        loss.mean().backward()
    tf, tu = timeit(fused, 50), timeit(unfused, 20)
    n = ((H + sample - 1) // sample) * ((W + sample - 1) // sample)
    rows.append(dict(B=B, H=H, W=W, sample=sample, N=n, fused_us=tf, unfused_us=tu, speedup=tu / tf))
    print(rows[-1], flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
