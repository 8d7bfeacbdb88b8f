"""Host-side cost of one operator call (no synchronisation inside the loop): Python + ctypes + allocation."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc_b200.synth import make_correspondences
from lc_b200.cov_mixed import loss_fwd_bwd, Loss_cov_mixed
from lc_b200.fused import solve_and_loss
from lc_b200.pnp.cer_solver import lm_solve
from lc_b200 import _native as nat
c = make_correspondences(4, 128, 0).to(torch.float32).to(device="cuda")
def t(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    dt = time.perf_counter() - t0; torch.cuda.synchronize()
    return dt / n * 1e6
print("loss_fwd_bwd      %.1f us/call" % t(lambda: loss_fwd_bwd(c.K, c.pose, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d)))
print("solve_and_loss    %.1f us/call" % t(lambda: solve_and_loss(c.K, c.start, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d)))
print("lm_solve          %.1f us/call" % t(lambda: lm_solve(c.K, c.pts3d, c.pts2d, c.inv_std, c.start, weight_mode=nat.W_INV_STD)))
a = nat.make_args(4, 128, torch.float32, K=c.K, pose=c.pose, pts3d=c.pts3d, pts2d=c.pts2d, weights=c.inv_std, bbox=c.bbox_3d,
                  loss=torch.empty(4, device="cuda"), g_pts3d=torch.empty_like(c.pts3d), g_weights=torch.empty_like(c.inv_std))
print("raw ABI call      %.1f us/call" % t(lambda: nat.call("lc_b200_loss_fwd_bwd", a, c.K.device)))
print("make_args only    %.1f us/call" % t(lambda: nat.make_args(4, 128, torch.float32, K=c.K, pose=c.pose, pts3d=c.pts3d, pts2d=c.pts2d, weights=c.inv_std, bbox=c.bbox_3d)))
