#!/usr/bin/env python
"""Per-phase cycle breakdown of the resident kernel (thread 0's clock64() deltas), from a -DLC_TIMING build.

    python tools/phase_timing.py build     # here: compiles lc_b200/csrc/build/liblc_b200_timing.so
    LC_B200_LIB=lc_b200/csrc/build/liblc_b200_timing.so python tools/phase_timing.py run [p1|p2|p3]   # on the GPU box

Slots: 0 stage, 1 LM passes, 2 LM advance (serial), 3 LC setup (serial), 4 LC passes 1-3, 5 six forward+backward, 6 pass 4, 7 total.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "lc_b200", "csrc", "build", "liblc_b200_timing.so")


def build():
    from lc_b200 import _native as nat
    os.makedirs(nat.BUILD_DIR, exist_ok=True)
    objs, procs = [], []
    for src in nat.SOURCES:
        obj = os.path.join(nat.BUILD_DIR, os.path.basename(src)[:-3] + ".timing.o")
        objs.append(obj)
        procs.append(subprocess.Popen(["nvcc"] + nat.NVCC_FLAGS + ["-DLC_TIMING", "-c", "-o", obj, src], cwd=ROOT))
    assert all(p.wait() == 0 for p in procs)
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs, check=True)
    print(OUT)


def run(pipeline):
    import torch
    from lc_b200 import _native as nat
    from lc_b200.synth import make_correspondences, planar_view
    B, N = 1024, 4096
    c = make_correspondences(B, N, 10).to(torch.float32, "cuda")
    X, x, s = planar_view(c.pts3d), planar_view(c.pts2d), planar_view(c.inv_std)
    trace = torch.zeros(B, 52, 4, dtype=torch.float64, device="cuda")
    st = torch.empty(B, 7, device="cuda"); rad = torch.empty(B, device="cuda"); loss = torch.empty(B, device="cuda")
    inv = torch.empty(B, dtype=torch.int32, device="cuda"); it = torch.empty(B, dtype=torch.int32, device="cuda")
    g3, gs = nat.empty_like_dense(X), nat.empty_like_dense(s)
    mode = dict(p1="lc_b200_loss_fwd_bwd", p2="lc_b200_lm_solve", p3="lc_b200_solve_loss")[pipeline]
    kw = dict(K=c.K, pose=c.pose if pipeline == "p1" else c.start, pts3d=X, pts2d=x, weights=s, bbox=c.bbox_3d, trace=trace,
              weight_mode=nat.W_INV_STD, flags=nat.FLAG_TOL_NEEDS_SUCCESS | (nat.FLAG_LM_MIXED if os.environ.get("LC_TIMING_MIXED") else 0))
    if pipeline != "p1":
        kw.update(state=st, radius=rad, invalid=inv, iters=it)
    if pipeline != "p2":
        kw.update(loss=loss, g_pts3d=g3, g_weights=gs)
    a = nat.make_args(B, N, torch.float32, **kw)
    for _ in range(3):
        nat.call(mode, a, X.device)
    torch.cuda.synchronize()
    if b"persist" in nat.lib().lc_b200_last_kernels():
        t = trace.reshape(-1)[: 148 * 64].reshape(148, 64).cpu().numpy().mean(0)
        names = ["-", "LM pass", "LC pass 1", "LC pass 2", "LC pass 3", "LC pass 4", "LC pass 4 (general)", "-"]
        print(f"{pipeline} persistent kernel, mean cycles per CTA (= per SM, ~6.9 poses): total {t[26]:.0f}")
        print("  workers (thread 0):  " + ", ".join(f"{n} {t[k]:.0f}" for k, n in enumerate(names) if t[k] > 0) + f", waiting for the serial warp {t[8]:.0f}")
        print("  serial warp (lane 0): after " + ", after ".join(f"{n} {t[16 + k]:.0f}" for k, n in enumerate(names) if t[16 + k] > 0)
              + f", next-pose setup {t[25]:.0f}, waiting for the workers {t[24]:.0f}")
        if t[27] > 0:
            print(f"  6x6 section executed twice in a row (LC_TIMING_SIX2): first {t[27]:.0f}, second {t[28]:.0f} cycles per CTA")
        return
    t = trace.reshape(B, -1)[:, :8].cpu().numpy()
    names = ["stage", "LM passes", "LM advance", "LC setup", "LC pass1-3", "six fwd+bwd", "LC pass 4", "total"]
    tot = t[:, 7].mean()
    print(f"{pipeline}: mean cycles per pose (thread 0), iters mean {it.float().mean().item() if pipeline != 'p1' else 0:.2f}")
    for k, nme in enumerate(names):
        print(f"  {nme:12s} {t[:, k].mean():10.0f}  {100 * t[:, k].mean() / tot:5.1f}%")
    lv = trace.reshape(B, -1)[:, 60].cpu().numpy()
    print(f"  CTAs resident on the SM when a CTA starts (incl. itself): max {lv.max():.0f}, mean {lv.mean():.2f}")
    if pipeline == "p2":
        m = trace.reshape(B, -1)[:, 48:54].cpu().numpy().mean(0)
        print("  inside lm_advance (cycles per pose): normal eq %.0f, Cholesky solve %.0f, model change %.0f, eval point %.0f, rest %.0f" % tuple(m[:5]))
    if pipeline == "p1":
        m = trace.reshape(B, -1)[:, 8:48].cpu().numpy().mean(0)
        print("  6x6 section barrier marks (cycles since entry):", " ".join(f"{v:.0f}" for v in m if v > 0))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        for p in sys.argv[2:] or ["p1", "p2", "p3"]:
            run(p)
