#!/usr/bin/env python
"""Time the UNMODIFIED reference Loss_cov_mixed (forward + backward) on the host cores of the BUILD container.

Only runs where /root/reference is mounted (not on the GPU box); writes profiles/reference_cpu_r1.json.  This is the
reference's own PyTorch/functorch implementation of pipeline P1 (lib/cov_mixed.py:100-150), fp32 like the training loop.

    PYTHONDONTWRITEBYTECODE=1 python tools/time_reference_cpu.py
"""
import json
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

from lib.cov_mixed import Loss_cov_mixed  # noqa: E402  (reference)
from lc_b200.synth import make_correspondences  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    out = []
    for B, N in ((32, 1024), (64, 4096)):
        c = make_correspondences(B, N, 10).to(torch.float32)
        ts = []
        for rep in range(3):
            p3 = c.pts3d.clone().requires_grad_(True)
            s = c.inv_std.clone().requires_grad_(True)
            t0 = time.perf_counter()
            loss = Loss_cov_mixed(c.K, c.pose, p3, c.pts2d, s, c.valid, bbox_3d=c.bbox_3d, max_err_len=32)
            loss.mean().backward()
            ts.append(time.perf_counter() - t0)
        t = min(ts[1:])
        line = dict(what="unmodified reference Loss_cov_mixed fwd+bwd (P1), PyTorch CPU fp32", B=B, N=N, seconds=t, poses_per_s=B / t,
                    cores=os.cpu_count(), torch=torch.__version__, host="build container (no GPU)")
        out.append(line)
        print(json.dumps(line), flush=True)
    with open(os.path.join(ROOT, "profiles", "reference_cpu_r1.json"), "w") as f:
        for ln in out:
            f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
