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
                   loss_sum: Optional[Tensor] = None):
    """Returns dict(states, radius, invalid, iters, loss, g_pts3d, g_pts2d, g_inv_std, flags).

    ``out`` may carry preallocated output tensors from a previous call (same shapes) to avoid allocation.
    """
    dev = nat.check_cuda(K, start, pts3d, pts2d, inv_std, valid, bbox_3d, grad_out)
    dt = pts3d.dtype
    B, N = pts3d.shape[:2]
    o = out or {}
    new = lambda key, *shape, dtype=dt: o[key] if key in o else torch.empty(*shape, dtype=dtype, device=dev)
    dense_like = lambda key, t: o[key] if key in o else nat.empty_like_dense(t)
    res = dict(states=new("states", B, 7), radius=new("radius", B), invalid=new("invalid", B, dtype=torch.int32),
               iters=new("iters", B, dtype=torch.int32), loss=new("loss", B), flags=new("flags", B, dtype=torch.int32),
               g_pts3d=dense_like("g_pts3d", pts3d) if need[0] else None,
               g_pts2d=dense_like("g_pts2d", pts2d.expand(B, N, 2)) if need[1] else None,
               g_inv_std=dense_like("g_inv_std", inv_std) if need[2] else None)
    flags = (nat.FLAG_TOL_NEEDS_SUCCESS if tol_needs_success else 0) | (nat.FLAG_FORCE_STREAMING if force_streaming else 0)
    ftol = float(torch.tensor(function_tolerance, dtype=torch.float32))
    args = nat.make_args(B, N, dt, K=K.to(dt).expand(B, 3, 3), pose=start.to(dt).expand(B, 7), pts3d=pts3d,
                         pts2d=pts2d.to(dt).expand(B, N, 2), weights=inv_std.to(dt),
                         valid=None if valid is None else valid.to(dt).expand(B, N), bbox=bbox_3d.to(dt).expand(B, 8, 3),
                         grad_out=None if grad_out is None else grad_out.to(dt).expand(B),
                         loss=res["loss"], g_pts3d=res["g_pts3d"], g_pts2d=res["g_pts2d"], g_weights=res["g_inv_std"],
                         state=res["states"], radius=res["radius"], invalid=res["invalid"], iters=res["iters"],
                         lc_flags=res["flags"], flags=flags, weight_mode=nat.W_INV_STD, max_iter=int(max_iter_count),
                         function_tolerance=ftol, max_err_len=float(max_err_len), rel_thresh=float(rel_thresh),
                         w_e_thresh=float(w_e_thresh), grad_scale=float(grad_scale), loss_sum=loss_sum)
    res["launches"] = nat.call("lc_b200_solve_loss", args, dev)
    return res
