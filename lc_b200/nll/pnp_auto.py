"""Implicit differentiation of the weighted-PnP optimum — drop-in for ``lib/nll/pnp_auto.py``.

``weighted_pnp_jac_wrt_pts2d`` (reference ``pnp_auto.py:111-135``) returns
``jac = d(update)/d(pts2d) = W_k H^-1 J_k`` of shape ``(*, 6, N, 2)`` and, with ``with_cov``,
``cov = H^-1`` ``(*, 6, 6)``.  The reference builds them from ``(B,N,2,6,6)`` per-coordinate
Hessians and 6 vmapped ``autograd.grad`` sweeps; here one kernel accumulates the 6x6 Hessian,
inverts it and writes ``jac`` directly.  Like the reference it is differentiable w.r.t. ``weights``
(a second kernel implements that backward).

The Hessian is the reference's ``hessian_6d_elem``: ``sum_k W_k (J_k J_k^T + r_k d2r_k)`` with the exact
exponential-map second derivative (closed form, see ``lc_stream.cu``); the second term vanishes when ``pts2d``
is the re-projection of ``pts3d`` at ``state_gt``, which is how ``Loss_cov_mixed`` calls it (``cov_mixed.py:120-121``).
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _native as nat


def _weights_as_bn2(weights: Tensor, B: int, N: int) -> Tensor:
    if weights.dim() >= 2 and weights.shape[-1] == 2 and weights.shape[-2] == 2 and weights.dim() == 4:
        raise NotImplementedError("full 2x2 weights: the reference's branch (pnp_utils.py:89-92) is not a correct "
                                  "full-covariance treatment (SURVEY.md §8a) and is not reproduced")
    if weights.dim() == 2:
        weights = weights.unsqueeze(-1)
    return weights.expand(B, N, 2)


def _jac_cov_forward(pose, K, pts3d, weights, pts2d=None):
    dev = nat.check_cuda(pose, K, pts3d, weights, pts2d)
    dt = pts3d.dtype
    B, N = pts3d.shape[:2]
    jac = torch.empty(B, 6, N, 2, dtype=dt, device=dev)
    cov = torch.empty(B, 6, 6, dtype=dt, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev)
    args = nat.make_args(B, N, dt, K=K.to(dt).expand(B, 3, 3), pose=pose.to(dt).expand(B, 7), pts3d=pts3d,
                         weights=weights.to(dt), pts2d=None if pts2d is None else pts2d.to(dt).expand(B, N, 2),
                         jac=jac, cov=cov, lc_flags=flags, flags=0 if pts2d is None else nat.FLAG_EXACT_HESSIAN)
    nat.call("lc_b200_pnp_jac_cov", args, dev)
    return jac, cov, flags


class _PnPJacCov(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, pose, K, pts3d, pts2d):
        jac, cov, flags = _jac_cov_forward(pose, K, pts3d, weights, pts2d)
        ctx.save_for_backward(weights, pose, K, pts3d, pts2d)
        ctx.mark_non_differentiable(flags)
        return jac, cov, flags

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_jac, g_cov, _g_flags):
        weights, pose, K, pts3d, pts2d = ctx.saved_tensors
        dev, dt = pts3d.device, pts3d.dtype
        B, N = pts3d.shape[:2]
        gw = torch.empty(B, N, 2, dtype=dt, device=dev)
        if g_jac is None:
            g_jac = torch.zeros(B, 6, N, 2, dtype=dt, device=dev)
        args = nat.make_args(B, N, dt, K=K.to(dt).expand(B, 3, 3), pose=pose.to(dt).expand(B, 7), pts3d=pts3d,
                             weights=weights.to(dt), pts2d=pts2d.to(dt).expand(B, N, 2), g_jac=g_jac.to(dt),
                             g_cov=None if g_cov is None else g_cov.to(dt), g_weights=gw, flags=nat.FLAG_EXACT_HESSIAN)
        nat.call("lc_b200_pnp_jac_cov_bwd", args, dev)
        return gw.sum_to_size(weights.shape) if gw.shape != weights.shape else gw, None, None, None, None


def _flatten_inputs(pts2d, state_gt, cam_K, pts3d, weights):
    batched = pts3d.dim() != 2
    p3 = pts3d.detach() if batched else pts3d.detach().unsqueeze(0)
    p3 = p3.reshape((-1,) + tuple(p3.shape[-2:]))
    B, N = p3.shape[:2]
    lead = pts3d.shape[:-2]
    w = weights if batched else weights.unsqueeze(0)
    w = _weights_as_bn2(w.reshape((B,) + tuple(w.shape[len(lead) if batched else 1:])), B, N)
    pose = state_gt.detach().reshape(-1, 7).expand(B, 7)
    K = cam_K.detach().reshape(-1, 3, 3).expand(B, 3, 3)
    p2 = (pts2d if batched else pts2d.unsqueeze(0)).expand(tuple(lead) + (N, 2) if batched else (1, N, 2)).reshape(B, N, 2)
    return w, pose, K, p3, p2, lead, N


def weighted_pnp_jac_wrt_pts2d(pts2d: Tensor, state_gt: Tensor, cam_K: Tensor, pts3d: Tensor, weights: Tensor,
                               with_cov: bool = False):
    """Reference signature (``pnp_auto.py:111``)."""
    w, pose, K, p3, p2, lead, N = _flatten_inputs(pts2d, state_gt, cam_K, pts3d, weights)
    jac, cov, _ = _PnPJacCov.apply(w, pose, K, p3, p2.detach())
    jac = jac.reshape(lead + (6, N, 2))
    cov = cov.reshape(lead + (6, 6))
    return (jac, cov) if with_cov else jac


class _RightUpdate(torch.autograd.Function):
    """``nll_update`` of the reference (``pnp_utils.py:118-131``): zero in value, its vector-Jacobian product w.r.t. the
    measured points is ``g^T d(update)/d(pts2d) = g^T jac``.  The backward is written with differentiable torch ops on
    ``jac``, so ``autograd.grad(update, pts2d, create_graph=True)`` can be differentiated again w.r.t. the weights (through
    ``_PnPJacCov``), which is how the reference builds ``jac`` and its double-backward (``pnp_auto.py:124-134``)."""
    @staticmethod
    def forward(ctx, pts2d, jac):
        ctx.save_for_backward(jac)
        return jac.new_zeros(jac.shape[0], 6)

    @staticmethod
    def backward(ctx, g):
        (jac,) = ctx.saved_tensors
        return torch.einsum("bk,bknc->bnc", g, jac), None


def diff_pnp_perturb(quat_xyz: Tensor, cam_K: Tensor, pts3d: Tensor, pts2d: Tensor, icov2: Tensor, with_cov: bool = True):
    """Reference signature (``pnp_auto.py:86-108``): returns ``(info, right_update, cov)``.

    ``right_update`` is identically zero in value, exactly like the reference's ``nll_update`` (``pnp_utils.py:118-122``),
    and carries the same gradient information w.r.t. ``pts2d``: ``autograd.grad(right_update, pts2d, g)`` returns
    ``g^T jac`` with ``jac = W_k H^-1 J_k`` (``create_graph=True`` keeps it differentiable w.r.t. ``icov2``).  Its gradient
    w.r.t. ``icov2`` itself (``-H^-1 r_k J_k`` in the reference) is not provided: it vanishes at the optimal operating
    point the reference documents this function for (``pnp_auto.py:89``).
    ``info`` is non-zero where the Hessian was not SPD and got replaced by the identity (``safe_cholesky``)."""
    w, pose, K, p3, p2, lead, N = _flatten_inputs(pts2d, quat_xyz, cam_K, pts3d, icov2)
    jac, cov, flags = _PnPJacCov.apply(w, pose, K, p3, p2.detach())
    info = (flags & nat.ST_HESS_NOT_SPD).reshape(lead)
    update = _RightUpdate.apply(p2, jac).reshape(lead + (6,))
    return info, update, (cov.reshape(lead + (6, 6)) if with_cov else None)
