"""Dense pose loss with the producer fused in — SURVEY.md §8 row f1.

Reference: ``Loss_fn.dense_pose_loss`` for the gdr-net structure (``losses.py:336-386``) =
joint softmax over the 2*H*W weight logits times ``xyz_weights_scale`` (``:355-356``), strided sub-sampling of
weights / ``xyz_noc * noc_scale`` / the ``gen_uv`` pixel grid (``dense_pnp_matching_from_xyz``, ``:142-161``),
``valid = ones`` (``:366``) and ``Loss_cov_mixed(...)`` (``:383``).  The reference materialises each of those tensors
(and their gradients, plus the softmax backward) in HBM; ``dense_pose_loss`` here is ONE kernel launch that reads the
network outputs in NCHW and writes ``loss (B,)`` and the gradients w.r.t. ``xyz_noc``, the logits and the scale.
The caller takes ``.mean()`` exactly like the reference.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _native as nat


def _plane_contiguous(t: Tensor) -> Tensor:
    """(B,C,H,W) with contiguous (H,W) planes, without copying channel slices of an NCHW tensor."""
    return t if (t.stride(3) == 1 and t.stride(2) == t.shape[3]) else t.contiguous()


def dense_loss_fwd_bwd(xyz_noc: Tensor, weight_logits: Tensor, weights_scale: Tensor, noc_scale: Tensor, K: Tensor,
                       pose: Tensor, bbox_3d: Tensor, *, sample: int, top_left: Tuple[int, int], max_err_len=32.0,
                       rel_thresh=3.0, w_e_thresh=4.0, need_grads=True, grad_out: Optional[Tensor] = None,
                       grad_scale: float = 1.0, loss_sum: Optional[Tensor] = None):
    dev = nat.check_cuda(xyz_noc, weight_logits, weights_scale, noc_scale, K, pose, bbox_3d, grad_out)
    if xyz_noc.dtype != torch.float32:
        raise TypeError("the fused dense producer takes float32 network outputs")
    B, _, H, W = xyz_noc.shape
    f32 = torch.float32
    xyz_noc, weight_logits = _plane_contiguous(xyz_noc), _plane_contiguous(weight_logits.to(f32))
    top, left = int(top_left[0]), int(top_left[1])
    a = nat.lc_dense_args()
    a.abi_version, a.B, a.H, a.W = nat.ABI_VERSION, B, H, W
    a.sample, a.top, a.left = int(sample), top, left
    a.max_err_len, a.rel_thresh, a.w_e_thresh, a.grad_scale = float(max_err_len), float(rel_thresh), float(w_e_thresh), float(grad_scale)
    loss = torch.empty(B, dtype=f32, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev)
    g_xyz = torch.empty(B, 3, H, W, dtype=f32, device=dev) if need_grads else None
    g_log = torch.empty(B, 2, H, W, dtype=f32, device=dev) if need_grads else None
    g_sc = torch.empty(B, dtype=f32, device=dev) if need_grads else None
    keep = dict(xyz_noc=xyz_noc, logits=weight_logits, weights_scale=weights_scale.to(f32).reshape(-1).expand(B),
                noc_scale=noc_scale.to(f32).expand(B, 3), K=K.to(f32).expand(B, 3, 3), pose=pose.to(f32).expand(B, 7),
                bbox=bbox_3d.to(f32).expand(B, 8, 3), grad_out=None if grad_out is None else grad_out.to(f32).expand(B),
                loss=loss, g_xyz_noc=g_xyz, g_logits=g_log, g_scale=g_sc)
    for k, v in keep.items():
        setattr(a, k, nat.view_of(v))
    a.lc_flags = flags.data_ptr()
    a.loss_sum = None if loss_sum is None else loss_sum.data_ptr()
    nat.call("lc_b200_dense_loss_fwd_bwd", a, dev)
    return dict(loss=loss, g_xyz_noc=g_xyz, g_logits=g_log, g_scale=g_sc, flags=flags)


class _DensePoseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz_noc, weight_logits, weights_scale, noc_scale, K, pose, bbox_3d, sample, top, left, max_err_len):
        need = any(ctx.needs_input_grad[:3])
        out = dense_loss_fwd_bwd(xyz_noc, weight_logits, weights_scale, noc_scale, K, pose, bbox_3d, sample=sample,
                                 top_left=(top, left), max_err_len=max_err_len, need_grads=need)
        ctx.grads = (out["g_xyz_noc"], out["g_logits"], out["g_scale"])
        ctx.scale_shape = weights_scale.shape
        return out["loss"]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss):
        gx, gl, gs = ctx.grads
        ctx.grads = None
        need = ctx.needs_input_grad
        go4 = grad_loss.reshape(-1, 1, 1, 1)
        return (gx * go4 if need[0] else None, gl * go4 if need[1] else None,
                (gs * grad_loss).reshape(ctx.scale_shape) if need[2] else None, None, None, None, None, None, None, None, None)


def dense_pose_loss(xyz_noc: Tensor, xyz_weight_logits: Tensor, xyz_weights_scale: Tensor, noc_scale: Tensor, K: Tensor,
                    pose_best: Tensor, bbox_3d: Tensor, *, dense_sample: int = 2, top_left: Optional[Tuple[int, int]] = None,
                    max_err_len: float = 32) -> Tensor:
    """Per-sample LC pose loss ``(B,)`` from the raw network outputs; differentiable w.r.t. ``xyz_noc``,
    ``xyz_weight_logits`` and ``xyz_weights_scale``.  ``top_left=None`` draws the sub-sampling offset from NumPy's
    global RNG exactly like ``dense_pnp_matching_from_xyz`` (``losses.py:152``)."""
    if top_left is None:
        top_left = tuple(int(v) for v in np.random.randint(0, dense_sample, size=2))
    return _DensePoseLoss.apply(xyz_noc, xyz_weight_logits, xyz_weights_scale, noc_scale, K.detach(), pose_best.detach(),
                                bbox_3d.detach(), int(dense_sample), int(top_left[0]), int(top_left[1]), float(max_err_len))
