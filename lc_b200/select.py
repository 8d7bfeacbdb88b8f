"""Test-time point selection on the device — SURVEY.md §8 row f2.

Reference: the selection half of ``test.solve_pnp_dense`` (``test.py:67-119``): softmax weights x scale (``:84-88``),
``dense_pnp_matching_from_xyz(..., top_left=(0,0))`` (``losses.py:142-161``), ``den_inv_cov2d = den_inv_std2d ** 2`` (``:95``),
the ``cfg.dense_point_select`` rules ``'mask' | 'quantile' | 'quantile_in_mask'`` (``:97-104``, ``quantile_msk`` ``:36-45``),
then ``nonzero()`` + ragged python lists (``:106-119``) that ``cer_solver._batch_tensors`` pads again (``cer_solver.py:67-87``).
Here ONE launch writes the zero-padded ``(B,Nmax,.)`` correspondences in selection order plus ``n_points (B,)`` — the inputs
of ``lc_b200.pnp.cer_solver.solve`` — with no device->host synchronisation in between.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _native as nat

_MODES = {"mask": nat.SEL_MASK, "quantile": nat.SEL_QUANTILE, "quantile_in_mask": nat.SEL_QUANTILE_IN_MASK}


def dense_point_select(xyz: Tensor, msk_vis_logits: Tensor, *, xyz_weight_logits: Optional[Tensor] = None,
                       xyz_weights_scale: Optional[Tensor] = None, xyz_weights: Optional[Tensor] = None,
                       noc_scale: Optional[Tensor] = None, sample: int = 2, dense_point_select: str = "quantile_in_mask",
                       quantile: float = 0.2, seg_thresh: float = 0.5, min_points: int = 4, want_index: bool = False):
    """``xyz (B,H,W,3)`` (any strides, e.g. ``nn_out.permute(0,2,3,1)``; multiplied by ``noc_scale (B,3)`` when given),
    ``msk_vis_logits (B,1,H,W)``, and either the raw ``xyz_weight_logits (B,2,H,W)`` + ``xyz_weights_scale (B,1|2,1,1)``
    (softmax fused in) or the precomputed ``xyz_weights (B,2,H,W)``.
    Returns ``dict(pts3d (B,Nmax,3), pts2d (B,Nmax,2), inv_cov (B,Nmax,2), n_points (B,) int32[, index (B,Nmax) int32])``."""
    dev = nat.check_cuda(xyz, msk_vis_logits, xyz_weight_logits, xyz_weights_scale, xyz_weights, noc_scale)
    if dense_point_select not in _MODES:
        raise ValueError(f"dense_point_select must be one of {sorted(_MODES)}")
    f32 = torch.float32
    B, H, W, _ = xyz.shape
    Hn, Wn = -(-H // sample), -(-W // sample)
    N = Hn * Wn
    a = nat.lc_select_args()
    a.abi_version, a.B, a.H, a.W = nat.ABI_VERSION, B, H, W
    a.sample, a.mode, a.min_points, a.Nmax = int(sample), _MODES[dense_point_select], int(min_points), N
    a.quantile, a.one_minus_quantile, a.seg_thresh = float(quantile), float(1 - quantile), float(seg_thresh)
    out = dict(pts3d=torch.empty(B, N, 3, dtype=f32, device=dev), pts2d=torch.empty(B, N, 2, dtype=f32, device=dev),
               inv_cov=torch.empty(B, N, 2, dtype=f32, device=dev), n_points=torch.empty(B, dtype=torch.int32, device=dev))
    if want_index:
        out["index"] = torch.empty(B, N, dtype=torch.int32, device=dev)
    keep = dict(xyz=xyz.to(f32), noc_scale=None if noc_scale is None else noc_scale.to(f32).expand(B, 3),
                msk_logits=msk_vis_logits.to(f32).reshape(B, H, W), pts3d=out["pts3d"], pts2d=out["pts2d"], inv_cov=out["inv_cov"])
    if xyz_weights is not None:
        keep["weights"] = xyz_weights.to(f32)
        a.scale_dim = 1
    else:
        lg = xyz_weight_logits.to(f32)
        keep["logits"] = lg if (lg.stride(3) == 1 and lg.stride(2) == W) else lg.contiguous()
        sc = xyz_weights_scale.to(f32)
        a.scale_dim = 2 if (sc.dim() >= 3 and sc.shape[-3] == 2) else 1      # test.py:86: weight_scale_dim = scale.shape[-3]
        keep["weights_scale"] = sc.reshape(B, -1).expand(B, a.scale_dim)
    for k, v in keep.items():
        setattr(a, k, nat.view_of(v))
    a.index = out["index"].data_ptr() if want_index else None
    a.n_points = out["n_points"].data_ptr()
    nat.call("lc_b200_dense_select", a, dev)
    return out


_select_points = dense_point_select   # solve_pnp_dense keeps the reference's cfg name `dense_point_select` as a keyword


def solve_pnp_dense(K: Tensor, xyz: Tensor, msk_vis_logits: Tensor, xyz_weight_logits: Tensor, xyz_weights_scale: Tensor,
                    start: Optional[Tensor] = None, *, noc_scale: Optional[Tensor] = None, sample: int = 2,
                    dense_point_select: str = "quantile_in_mask", quantile: float = 0.2, seg_thresh: float = 0.5,
                    solvers=("weighted",), reprojectionError=3.0):
    """Device-resident ``test.solve_pnp_dense`` (``test.py:67-136``) after the network and the xyz decode:

    selection launch (``:84-119``) -> start pose (``:120``; ``start=None`` uses ``lc_b200.pnp.init_solver`` in place of
    ``cv2_solver.solve``) -> ``'weighted'`` LM solve on the selected points (``:125-128``) and/or ``'weighted_filtered'`` LM
    solve on the initialiser's inliers (``:130-134``).  No device->host synchronisation anywhere.  The inlier restriction is
    applied by zeroing the inverse variances of the other points: a point with zero weight adds nothing to the cost, the
    Jacobian or the normal equations, so the solve equals the solve on the compacted inlier list.
    Returns ``(dict name -> states (B,7), selection dict)`` with the reference's result names (``'weighted'``,
    ``'weighted-filtered'``)."""
    from .pnp import cer_solver, init_solver
    sel = _select_points(xyz, msk_vis_logits, xyz_weight_logits=xyz_weight_logits, xyz_weights_scale=xyz_weights_scale,
                         noc_scale=noc_scale, sample=sample, dense_point_select=dense_point_select, quantile=quantile,
                         seg_thresh=seg_thresh)
    inliers = None
    if start is None:
        invalid0, start, inliers = init_solver.solve(K, sel["pts3d"], sel["pts2d"], weights=sel["inv_cov"], n_points=sel["n_points"],
                                                     reprojectionError=reprojectionError)
        sel["init_invalid"], sel["start"], sel["inliers"] = invalid0, start, inliers
    res = {}
    if "weighted" in solvers:
        res["weighted"] = cer_solver.solve(K, sel["pts3d"], sel["pts2d"], sel["inv_cov"], start, sel["n_points"],
                                           num_workers=4, filter_input_nan=True)[1]                       # test.py:127
    if "weighted_filtered" in solvers:
        if inliers is None:
            raise ValueError("'weighted_filtered' needs the initialiser's inlier set: call with start=None")
        icov_in = sel["inv_cov"] * inliers["mask"].unsqueeze(-1)
        res["weighted-filtered"] = cer_solver.solve(K, sel["pts3d"], sel["pts2d"], icov_in, start, sel["n_points"],
                                                    num_workers=4, filter_input_nan=True)[1]              # test.py:133
    return res, sel

