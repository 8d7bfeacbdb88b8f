"""Symmetric pose-candidate selection on the device — SURVEY.md §8 row f4.

Reference: ``symmetry.select_pose_2d`` / ``select_pose_3d`` (``symmetry.py:8-56``), called from
``losses.selete_best_pose`` (``losses.py:88-112``).  The reference materialises ``(B,K,N,3)`` transformed point sets; here
one CTA per sample loops over the candidates and keeps only the running minimum.  Same signatures, same return value.
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _native as nat


def _select(mode: int, cam_K: Tensor, pts_a: Tensor, pts_b: Tensor, pose_candi: Tensor, want_err: bool):
    if pose_candi.shape[-3] == 1 and not want_err:
        return pose_candi.squeeze(-3), None, None                          # symmetry.py:15-16, 40-41
    dev = nat.check_cuda(cam_K, pts_a, pts_b, pose_candi)
    f32 = torch.float32
    B, N = pts_a.shape[0], pts_a.shape[1]
    Kc = pose_candi.shape[1]
    a = nat.lc_candi_args()
    a.abi_version, a.B, a.N, a.Kc, a.mode = nat.ABI_VERSION, B, N, Kc, mode
    best = torch.empty(B, 3, 4, dtype=f32, device=dev)
    idx = torch.empty(B, dtype=torch.int32, device=dev)
    err = torch.empty(B, Kc, dtype=f32, device=dev) if want_err else None
    keep = dict(K=cam_K.to(f32).expand(B, 3, 3), pts_a=pts_a.to(f32), pts_b=pts_b.to(f32), candi=pose_candi.to(f32), best=best, err=err)
    for k, v in keep.items():
        setattr(a, k, nat.view_of(v))
    a.best_index = idx.data_ptr()
    nat.call("lc_b200_select_pose", a, dev)
    return best.to(pose_candi.dtype), idx, err


def select_pose_2d(cam_K: Tensor, pts3d: Tensor, pts2d: Tensor, pose_candi: Tensor, *, return_details: bool = False):
    """``cam_K (B,3,3)``, ``pts3d (B,N,3)``, ``pts2d (B,N,2)``, ``pose_candi (B,K,3,4)`` -> best pose ``(B,3,4)``."""
    best, idx, err = _select(0, cam_K, pts3d, pts2d, pose_candi, return_details)
    return (best, idx, err) if return_details else best


def select_pose_3d(cam_K: Tensor, pts3d_out: Tensor, homo_z: Tensor, pose_candi: Tensor, *, return_details: bool = False):
    """``pts3d_out (B,N,3)`` predicted model points, ``homo_z (B,N,3)`` ground-truth homogeneous pixel coordinates times depth."""
    best, idx, err = _select(1, cam_K, pts3d_out, homo_z, pose_candi, return_details)
    return (best, idx, err) if return_details else best
