"""Device-side pose initialiser — SURVEY.md §8 row f2; takes the place of ``lib/pnp/cv2_solver.py`` in the test-time path.

The reference copies every sample to the host, synchronises and runs ``cv2.solvePnPRansac(EPNP, iterationsCount=150)`` per
sample (``cv2_solver.py:33-88``) to obtain the LM start (``test.py:60,120``) and RANSAC's inlier set (``test.py:131-134``).
``solve`` keeps that contract — ``(invalids, states, inliers)`` — but runs one launch on the device (weighted DLT + Cauchy
IRLS, ``lc_b200/csrc/lc_init.cu``).  It is NOT OpenCV's algorithm and does not reproduce its random stream; what is tested
is that the LM solve started here lands on the same optimum as the LM solve started from OpenCV's result.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
from torch import Tensor

from .. import _native as nat
from .cer_solver import _batch_tensors


def solve(cam_mat, coord_3d, coord_2d, *, reprojectionError: Union[float, Tensor] = 3.0, weights: Optional[Tensor] = None,
          n_points: Optional[Tensor] = None, irls_rounds: int = 3, **kwargs):
    """``cam_mat (B,3,3)``, ``coord_3d (B,N,3)`` | list, ``coord_2d (B,N,2)`` | list; ``weights (B,N,2)`` optional inverse
    variances (the network's), ``reprojectionError`` scalar or ``(B,)`` tensor (``cfg.rel_reproj_err``, ``test.py:115-117``).
    Returns ``(invalids (B,) bool, states (B,7) float32 wxyz+t, inliers dict(mask (B,N) bool, count (B,) int32))``."""
    if isinstance(coord_3d, (list, tuple)):
        if n_points is None:
            n_points = torch.tensor([len(p) for p in coord_3d], dtype=torch.int32, device=coord_3d[0].device)
        cam_mat, coord_3d, coord_2d, weights = _batch_tensors(cam_mat, coord_3d, coord_2d, weights)
    dev = nat.check_cuda(cam_mat, coord_3d, coord_2d, weights)
    f32 = torch.float32
    B, N = coord_3d.shape[:2]
    a = nat.lc_init_args()
    a.abi_version, a.B, a.N, a.irls_rounds = nat.ABI_VERSION, B, N, int(irls_rounds)
    state = torch.empty(B, 7, dtype=f32, device=dev)
    invalid = torch.empty(B, dtype=torch.int32, device=dev)
    inlier = torch.empty(B, N, dtype=torch.uint8, device=dev)
    count = torch.empty(B, dtype=torch.int32, device=dev)
    thr_b = None
    if isinstance(reprojectionError, Tensor):
        thr_b = reprojectionError.to(device=dev, dtype=f32).reshape(-1).expand(B)
        a.reproj_thresh = 0.0
    else:
        a.reproj_thresh = float(reprojectionError)
    npts = None if n_points is None else torch.as_tensor(n_points).to(device=dev, dtype=torch.int32).contiguous()
    keep = dict(K=cam_mat.to(f32).expand(B, 3, 3), pts3d=coord_3d.to(f32), pts2d=coord_2d.to(f32).expand(B, N, 2),
                weights=None if weights is None else weights.to(f32), reproj_thresh_b=thr_b, state=state)
    for k, v in keep.items():
        setattr(a, k, nat.view_of(v))
    a.n_points = None if npts is None else npts.data_ptr()
    a.invalid, a.inlier, a.n_inliers = invalid.data_ptr(), inlier.data_ptr(), count.data_ptr()
    nat.call("lc_b200_pnp_init", a, dev)
    return invalid.to(torch.bool), state, dict(mask=inlier.to(torch.bool), count=count)
