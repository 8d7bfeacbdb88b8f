"""Small run of every kernel path for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc_b200.synth import make_correspondences, planar_view, full_icov_from_inv_std
from lc_b200.cov_mixed import loss_fwd_bwd
from lc_b200.fused import solve_and_loss
from lc_b200.pnp.cer_solver import lm_solve
from lc_b200.nll.pnp_auto import weighted_pnp_jac_wrt_pts2d
from lc_b200 import _native as nat

for (B, N) in [(3, 8), (2, 70), (2, 700), (2, 1300)]:
    c = make_correspondences(B, N, 1).to(torch.float32).to(device="cuda")
    for stream in (False, True):
        loss_fwd_bwd(c.K, c.pose, planar_view(c.pts3d), c.pts2d, planar_view(c.inv_std), c.valid, c.bbox_3d, want_cov=True, force_streaming=stream)
        solve_and_loss(c.K, c.start, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d, need=(True, True, True), force_streaming=stream)
        lm_solve(c.K, c.pts3d, c.pts2d, c.inv_std ** 2, c.start, weight_mode=nat.W_ICOV_DIAG, filter_input_nan=True, force_streaming=stream, want_trace=True)
    lm_solve(c.K, c.pts3d, c.pts2d, full_icov_from_inv_std(c.inv_std.cpu().double(), 0).float().cuda(), c.start, weight_mode=nat.W_ICOV_FULL)
    w = (c.inv_std ** 2).requires_grad_(True)
    j, cv = weighted_pnp_jac_wrt_pts2d(c.pts2d, c.pose, c.K, c.pts3d, w, with_cov=True)
    (j.sum() + cv.sum()).backward()
torch.cuda.synchronize()
print("sanitize smoke done")
