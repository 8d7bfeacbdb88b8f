"""The north-star operator: weighted-PnP solve -> LC loss at the solution -> input gradients, one launch.

Reference equivalent (two subsystems it never connects): ``cer_solver.solve(K, pts3d, pts2d,
inv_std**2, start)`` (``test.py:127``) followed by ``Loss_cov_mixed(K, states, pts3d, pts2d, inv_std, ...)``
(``losses.py:383``) and ``loss.backward()``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _native as nat


def solve_and_loss(K: Tensor, start: Tensor, pts3d: Tensor, pts2d: Tensor, inv_std: Tensor, valid: Optional[Tensor],
                   bbox_3d: Tensor, *, max_iter_count=50, function_tolerance=1e-6, max_err_len=32.0, rel_thresh=3.0,
                   w_e_thresh=4.0, need=(True, False, True), grad_out: Optional[Tensor] = None, grad_scale=1.0,
                   tol_needs_success=True, out: Optional[dict] = None, force_streaming=False,
                   loss_sum: Optional[Tensor] = None, mixed: bool = False, n_points: Optional[Tensor] = None):
    """Returns dict(states, radius, invalid, iters, loss, g_pts3d, g_pts2d, g_inv_std, flags).

    ``n_points`` (B,) int32: ragged batch padded to N (only the first n_points[b] correspondences of pose b count).
    ``out`` may carry preallocated output tensors from a previous call (same shapes) to avoid allocation.
    ``mixed`` (opt-in) lets the resident kernels form the Jacobian sums of the solve in packed fp32 (``LC_FLAG_LM_MIXED``):
    residuals, cost and every trust-region decision stay fp64; over 10 240 poses the iteration counts, accept / reject sequences
    and invalid flags are identical to the all-fp64 pass and the poses agree to one fp32 ulp (``profiles/lm_mixed_check_r2.md``).
    It is NOT the default: a one-ulp difference of the returned fp32 pose moves the loss evaluated at it by up to 2.4e-5 relative
    (the linear term vanishes at the optimum, so the loss is that sensitive there), which is outside this repo's own 1e-5 bar for
    the fused operator although inside every north-star tolerance.
    """
    dev = nat.check_cuda(K, start, pts3d, pts2d, inv_std, valid, bbox_3d, grad_out, n_points)
    dt = pts3d.dtype
    B, N = pts3d.shape[:2]
    o = out or {}
    if not all(k in o for k in ("states", "radius", "loss", "invalid", "iters", "flags")):
        # two allocations instead of six: the small per-pose outputs are views of one float and one int32 buffer
        fbuf = torch.empty(9 * B, dtype=dt, device=dev)
        ibuf = torch.empty(3 * B, dtype=torch.int32, device=dev)
        o = dict(o, states=fbuf[: 7 * B].view(B, 7), radius=fbuf[7 * B: 8 * B], loss=fbuf[8 * B:],
                 invalid=ibuf[:B], iters=ibuf[B: 2 * B], flags=ibuf[2 * B:])
    dense_like = lambda key, t: o[key] if key in o else nat.empty_like_dense(t)
    res = dict(states=o["states"], radius=o["radius"], invalid=o["invalid"], iters=o["iters"], loss=o["loss"], flags=o["flags"],
               g_pts3d=dense_like("g_pts3d", pts3d) if need[0] else None,
               g_pts2d=dense_like("g_pts2d", pts2d.expand(B, N, 2)) if need[1] else None,
               g_inv_std=dense_like("g_inv_std", inv_std) if need[2] else None)
    flags = ((nat.FLAG_TOL_NEEDS_SUCCESS if tol_needs_success else 0) | (nat.FLAG_FORCE_STREAMING if force_streaming else 0)
             | (nat.FLAG_LM_MIXED if mixed else 0))
    ftol = nat.as_c_float(function_tolerance)
    fit = nat.fit
    args = nat.make_args(B, N, dt, K=fit(K, (B, 3, 3), dt), pose=fit(start, (B, 7), dt), pts3d=pts3d,
                         pts2d=fit(pts2d, (B, N, 2), dt), weights=fit(inv_std, (B, N, 2), dt),
                         valid=fit(valid, (B, N), dt), bbox=fit(bbox_3d, (B, 8, 3), dt), grad_out=fit(grad_out, (B,), dt),
                         loss=res["loss"], g_pts3d=res["g_pts3d"], g_pts2d=res["g_pts2d"], g_weights=res["g_inv_std"],
                         state=res["states"], radius=res["radius"], invalid=res["invalid"], iters=res["iters"],
                         lc_flags=res["flags"], flags=flags, weight_mode=nat.W_INV_STD, max_iter=int(max_iter_count),
                         function_tolerance=ftol, max_err_len=float(max_err_len), rel_thresh=float(rel_thresh),
                         w_e_thresh=float(w_e_thresh), grad_scale=float(grad_scale), loss_sum=loss_sum, n_points=n_points)
    res["launches"] = nat.call("lc_b200_solve_loss", args, dev)
    return res
