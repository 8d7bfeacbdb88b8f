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


def _bits_u8(t: Tensor) -> Tensor:
    """bool / uint8 tensor as uint8 without a copy (bool storage is one byte per element)."""
    if t.dtype == torch.bool:
        return t.view(torch.uint8)
    if t.dtype == torch.uint8:
        return t
    return (t != 0).view(torch.uint8)


def dense_loss_fwd_bwd(xyz_noc: Optional[Tensor], weight_logits: Tensor, weights_scale: Tensor, noc_scale: Tensor, K: Tensor,
                       pose: Tensor, bbox_3d: Tensor, *, sample: int, top_left: Tuple[int, int], max_err_len=32.0,
                       rel_thresh=3.0, w_e_thresh=4.0, need_grads=True, grad_out: Optional[Tensor] = None,
                       grad_scale: float = 1.0, loss_sum: Optional[Tensor] = None, noc_bin_logits: Optional[Tensor] = None,
                       noc_bin_raw: Optional[Tensor] = None, msk_noc: Optional[Tensor] = None, bit_cnt=None,
                       model_transform: Optional[Tensor] = None, black_background: bool = True):
    """One launch of ``lc_b200_dense_loss_fwd_bwd``.  ``noc_bin_logits`` selects the zebrapose producer (row f3): pts3d is
    then decoded from the bit logits with the GT raw bits / ``msk_noc`` and ``xyz_noc`` must be None."""
    zebra = noc_bin_logits is not None
    src = noc_bin_logits if zebra else xyz_noc
    dev = nat.check_cuda(src, weight_logits, weights_scale, noc_scale, K, pose, bbox_3d, grad_out, noc_bin_raw, msk_noc, model_transform)
    if src.dtype != torch.float32:
        raise TypeError("the fused dense producer takes float32 network outputs")
    B, C_, H, W = src.shape
    f32 = torch.float32
    src, weight_logits = _plane_contiguous(src), _plane_contiguous(weight_logits.to(f32))
    if zebra:
        if xyz_noc is not None:
            raise ValueError("either xyz_noc (gdr-net structure) or noc_bin_logits (zebrapose structure), not both (losses.py:360)")
        bit_cnt = [int(bit_cnt)] * 3 if isinstance(bit_cnt, int) else [int(v) for v in bit_cnt]
        if len(bit_cnt) != 3 or sum(bit_cnt) != C_:
            raise ValueError(f"bit_cnt {bit_cnt} does not match the {C_} bit channels")
        if tuple(noc_bin_raw.shape) != (B, C_, H, W) or tuple(msk_noc.shape) != (B, H, W):
            raise ValueError("noc_bin_raw must be (B,C,H,W) and msk_noc (B,H,W)")
    xyz_noc = None if zebra else src
    top, left = int(top_left[0]), int(top_left[1])
    a = nat.lc_dense_args()
    a.abi_version, a.B, a.H, a.W = nat.ABI_VERSION, B, H, W
    a.sample, a.top, a.left = int(sample), top, left
    a.max_err_len, a.rel_thresh, a.w_e_thresh, a.grad_scale = float(max_err_len), float(rel_thresh), float(w_e_thresh), float(grad_scale)
    loss = torch.empty(B, dtype=f32, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev)
    g_xyz = torch.empty(B, 3, H, W, dtype=f32, device=dev) if need_grads and not zebra else None
    g_bin = torch.empty(B, C_, H, W, dtype=f32, device=dev) if need_grads and zebra else None
    g_log = torch.empty(B, 2, H, W, dtype=f32, device=dev) if need_grads else None
    g_sc = torch.empty(B, dtype=f32, device=dev) if need_grads else None
    keep = dict(xyz_noc=xyz_noc, logits=weight_logits, weights_scale=weights_scale.to(f32).reshape(-1).expand(B),
                noc_scale=noc_scale.to(f32).expand(B, 3), K=K.to(f32).expand(B, 3, 3), pose=pose.to(f32).expand(B, 7),
                bbox=bbox_3d.to(f32).expand(B, 8, 3), grad_out=None if grad_out is None else grad_out.to(f32).expand(B),
                loss=loss, g_xyz_noc=g_xyz, g_logits=g_log, g_scale=g_sc)
    if zebra:
        keep.update(noc_bin_logits=src, noc_bin_raw=_bits_u8(noc_bin_raw), msk_noc=_bits_u8(msk_noc), g_noc_bin=g_bin,
                    model_transform=None if model_transform is None else model_transform.to(f32).expand(B, 4, 4))
        a.bit_cnt[:] = bit_cnt
        a.black_background = int(bool(black_background))
    for k, v in keep.items():
        setattr(a, k, nat.view_of(v))
    a.lc_flags = flags.data_ptr()
    a.loss_sum = None if loss_sum is None else loss_sum.data_ptr()
    nat.call("lc_b200_dense_loss_fwd_bwd", a, dev)
    return dict(loss=loss, g_xyz_noc=g_xyz, g_noc_bin=g_bin, g_logits=g_log, g_scale=g_sc, flags=flags)


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
        gx, gl, gs = ctx.grads   # kept for the lifetime of the graph: backward may run again under retain_graph=True
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


class _DensePoseLossNocBin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bin_logits, weight_logits, weights_scale, raw_bits, msk_noc, noc_scale, K, pose, bbox_3d, model_transform,
                bit_cnt, sample, top, left, max_err_len, black_background):
        need = any(ctx.needs_input_grad[:3])
        out = dense_loss_fwd_bwd(None, weight_logits, weights_scale, noc_scale, K, pose, bbox_3d, sample=sample, top_left=(top, left),
                                 max_err_len=max_err_len, need_grads=need, noc_bin_logits=bin_logits, noc_bin_raw=raw_bits,
                                 msk_noc=msk_noc, bit_cnt=bit_cnt, model_transform=model_transform, black_background=black_background)
        ctx.grads = (out["g_noc_bin"], out["g_logits"], out["g_scale"])
        ctx.scale_shape = weights_scale.shape
        return out["loss"]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss):
        gb, gl, gs = ctx.grads
        need = ctx.needs_input_grad
        go4 = grad_loss.reshape(-1, 1, 1, 1)
        return (gb * go4 if need[0] else None, gl * go4 if need[1] else None,
                (gs * grad_loss).reshape(ctx.scale_shape) if need[2] else None) + (None,) * 13


def dense_pose_loss_noc_bin(xyz_noc_bin_logits: Tensor, noc_bin_raw_gt: Tensor, xyz_weight_logits: Tensor, xyz_weights_scale: Tensor,
                            msk_noc: Tensor, noc_scale: Tensor, K: Tensor, pose_best: Tensor, bbox_3d: Tensor, *, bit_cnt,
                            model_transform: Optional[Tensor] = None, dense_sample: int = 2,
                            top_left: Optional[Tuple[int, int]] = None, max_err_len: float = 32,
                            black_background: bool = True) -> Tensor:
    """Zebrapose branch of ``Loss_fn.dense_pose_loss`` (``losses.py:368-375, 383``; ``dense_pnp_matching_from_noc_bin``
    ``:163-184``; ``nn_out_to_xyz`` ``:16-45``; ``floatbits.nn_logits2noc_with_gt`` ``floatbits.py:49-69``) as one launch.
    Per-sample loss ``(B,)``, differentiable w.r.t. the bit logits, the weight logits and the weights scale."""
    if top_left is None:
        top_left = tuple(int(v) for v in np.random.randint(0, dense_sample, size=2))
    bits = (int(bit_cnt),) * 3 if isinstance(bit_cnt, int) else tuple(int(v) for v in bit_cnt)
    return _DensePoseLossNocBin.apply(xyz_noc_bin_logits, xyz_weight_logits, xyz_weights_scale, noc_bin_raw_gt, msk_noc, noc_scale,
                                      K.detach(), pose_best.detach(), bbox_3d.detach(), model_transform, bits, int(dense_sample),
                                      int(top_left[0]), int(top_left[1]), float(max_err_len), bool(black_background))
