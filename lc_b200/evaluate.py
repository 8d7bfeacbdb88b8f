"""Pose-error metrics on the device — SURVEY.md §8 row f4.

Reference: ``lib/utils/evaluate.py:333-339`` ``compute_pose_errors`` = ``error6d.adi / add / re / te``
(``lib/utils/error6d.py:87-159``), evaluated per pose in numpy + ``scipy.spatial.cKDTree`` inside a
``multiprocessing.Pool(6)`` (``evaluate.py:193-210``).  Here a whole batch of poses is one launch (one CTA per pose,
brute-force nearest neighbour over shared-memory tiles for ADI); results are fp64 like the reference's.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _native as nat


def compute_pose_errors(R_est: Tensor, t_est: Tensor, R_gt: Tensor, t_gt: Tensor, pts: Tensor, *,
                        pts_offset: Optional[Tensor] = None, pts_count: Optional[Tensor] = None, want_adi: bool = True):
    """``R_* (B,3,3)``, ``t_* (B,3)`` (or ``(B,3,1)``), ``pts (M,3)`` model points shared by the batch — or a concatenation
    ``(P,3)`` of several models with ``pts_offset (B,) int64`` / ``pts_count (B,) int32`` naming each pose's slice.
    Returns ``dict(adi, add, re, te)`` of ``(B,)`` float64 tensors (``re`` in degrees)."""
    dev = nat.check_cuda(R_est, t_est, R_gt, t_gt, pts)
    f64 = torch.float64
    B = R_est.shape[0]
    a = nat.lc_eval_args()
    a.abi_version, a.B, a.M = nat.ABI_VERSION, B, int(pts.shape[0])
    out = {k: torch.empty(B, dtype=f64, device=dev) for k in (("adi", "add", "re", "te") if want_adi else ("add", "re", "te"))}
    keep = dict(R_est=R_est.to(f64).reshape(B, 3, 3), t_est=t_est.to(f64).reshape(B, 3), R_gt=R_gt.to(f64).reshape(B, 3, 3),
                t_gt=t_gt.to(f64).reshape(B, 3), pts=pts.to(f64).reshape(-1, 3), **out)
    for k, v in keep.items():
        setattr(a, k, nat.view_of(v))
    off = None if pts_offset is None else pts_offset.to(device=dev, dtype=torch.int64).contiguous()
    cnt = None if pts_count is None else pts_count.to(device=dev, dtype=torch.int32).contiguous()
    a.pts_offset = None if off is None else off.data_ptr()
    a.pts_count = None if cnt is None else cnt.data_ptr()
    nat.call("lc_b200_pose_errors", a, dev)
    return out


def add(R_est, t_est, R_gt, t_gt, pts):
    """``error6d.add`` (``error6d.py:87-101``), batched."""
    return compute_pose_errors(R_est, t_est, R_gt, t_gt, pts, want_adi=False)["add"]


def adi(R_est, t_est, R_gt, t_gt, pts):
    """``error6d.adi`` (``error6d.py:104-124``), batched."""
    return compute_pose_errors(R_est, t_est, R_gt, t_gt, pts)["adi"]
