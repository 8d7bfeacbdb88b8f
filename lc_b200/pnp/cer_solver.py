"""Weighted-PnP solver front-end — drop-in for ``lib/pnp/cer_solver.py`` of the reference.

``solve`` keeps the reference signature and return value (``cer_solver.py:6-53``): tensors or ragged
lists in, ``(invalid_dict, states)`` out, invalid samples fall back to ``start``, nothing is raised
for numerical failures.  The reference moves every sample to the host, builds ``float*[B]`` tables
and runs one ``ceres::Solve`` per sample on OpenMP threads (``pnp_ceres.py:6-61``,
``ceres.cpp:147-177``); here the whole batch is one kernel launch (one CTA per pose) that stays on
the device, including the host-side prologue the reference does in PyTorch (``nan_to_num``,
``icov -> L``, ``where(invalid, start, state)``).
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
from torch import Tensor
from torch.nn.utils.rnn import pad_sequence

from .. import _native as nat

TensorOrList = Union[Tensor, Sequence[Tensor]]


def _batch_tensors(*tensor_lsts, device=None):
    """Zero-pad ragged per-sample lists to a batch (role of ``cer_solver._batch_tensors``, :67-87)."""
    out = []
    for lst in tensor_lsts:
        if lst is None:
            out.append(None)
        elif isinstance(lst, Tensor):
            out.append(lst)
        elif isinstance(lst[0], Tensor):
            out.append(pad_sequence(list(lst), batch_first=True) if lst[0].dim() > 0 else torch.stack(list(lst)))
        else:
            out.append(torch.tensor(lst, device=device))
    return out


def lm_solve(cam_mat: Tensor, pts3d: Tensor, pts2d: Tensor, weights: Tensor, start: Tensor,
             n_points: Optional[Tensor] = None, *, weight_mode: int = nat.W_ICOV_DIAG, max_iter_count: int = 50,
             function_tolerance: float = 1e-6, filter_input_nan: bool = False, tol_needs_success: bool = True,
             want_trace: bool = False, force_streaming: bool = False, mixed: bool = False):
    """Batched kernel call.  Returns dict(states, radius, invalid (int32), iters[, trace])."""
    dev = nat.check_cuda(cam_mat, pts3d, pts2d, weights, start, n_points)
    dt = pts3d.dtype
    B, N = pts3d.shape[:2]
    state = torch.empty(B, 7, dtype=dt, device=dev)
    radius = torch.empty(B, dtype=dt, device=dev)
    invalid = torch.empty(B, dtype=torch.int32, device=dev)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    trace = torch.full((B, max_iter_count + 2, 4), float("nan"), dtype=torch.float64, device=dev) if want_trace else None
    flags = ((nat.FLAG_NAN_TO_NUM if filter_input_nan else 0) | (nat.FLAG_TOL_NEEDS_SUCCESS if tol_needs_success else 0)
             | (nat.FLAG_FORCE_STREAMING if force_streaming else 0) | (nat.FLAG_LM_MIXED if mixed else 0))
    # the reference ABI carries function_tolerance as a C float (ext.h:10)
    ftol = nat.as_c_float(function_tolerance)
    npts = None if n_points is None else n_points.to(device=dev, dtype=torch.int32).contiguous()
    fit = nat.fit
    args = nat.make_args(B, N, dt, K=fit(cam_mat, (B, 3, 3), dt), pose=fit(start, (B, 7), dt), pts3d=pts3d,
                         pts2d=fit(pts2d, (B, N, 2), dt), weights=weights if weights.dtype == dt else weights.to(dt), n_points=npts, state=state,
                         radius=radius, invalid=invalid, iters=iters, trace=trace, flags=flags,
                         weight_mode=int(weight_mode), max_iter=int(max_iter_count), function_tolerance=ftol)
    nat.call("lc_b200_lm_solve", args, dev)
    out = dict(states=state, radius=radius, invalid=invalid, iters=iters)
    if want_trace:
        out["trace"] = trace
    return out


def solve(cam_mat: TensorOrList, pts3d: TensorOrList, pts2d: TensorOrList, icovs: TensorOrList, start: TensorOrList,
          n_points=None, *, optimal_start=False, max_iter_count=50, num_workers=1, filter_input_nan=False, **kwargs):
    """Reference contract (``cer_solver.py:6-53``).

    cam_mat (*,3,3) | list, pts3d (*,N,3) | list[(N_i,3)], pts2d (*,N,2) | list, icovs (*,N,2) inverse
    variances or (*,N,2,2) inverse covariances | list, start (*,7) wxyz+t | list.  Returns
    ``(invalid_dict, states)``: ``states (B,7)`` on the input device, ``invalid_dict`` with bool
    ``'solver_invalids'`` and ``'invalids'``.  ``num_workers`` is accepted and ignored (the batch is one
    launch); ``print_summary`` is ignored.
    """
    dev = pts3d.device if isinstance(pts3d, Tensor) else pts3d[0].device
    if isinstance(pts3d, (list, tuple)):
        if n_points is None:
            n_points = [len(p) for p in pts3d]
        cam_mat, pts3d, pts2d, icovs, start, n_points = _batch_tensors(cam_mat, pts3d, pts2d, icovs, start, n_points, device=dev)
    elif isinstance(start, (list, tuple)):
        cam_mat, pts2d, icovs, start, n_points = _batch_tensors(cam_mat, pts2d, icovs, start, n_points, device=dev)
    start = start.detach()
    lead = pts3d.shape[:-2]
    B = 1
    for d in lead:
        B *= d
    N = pts3d.shape[-2]
    if optimal_start:
        invalids = torch.zeros(start.shape[:-1], dtype=torch.bool, device=start.device)
        return dict(invalids=invalids), start
    full = icovs.dim() == pts2d.dim() + 1
    with torch.no_grad():
        res = lm_solve(cam_mat.reshape(-1, 3, 3), pts3d.reshape(B, N, 3), pts2d.reshape(-1, N, 2),
                       icovs.reshape((B, N, 2, 2) if full else (B, N, 2)), start.reshape(-1, 7),
                       None if n_points is None else torch.as_tensor(n_points).reshape(-1),
                       weight_mode=nat.W_ICOV_FULL if full else nat.W_ICOV_DIAG, max_iter_count=max_iter_count,
                       function_tolerance=float(kwargs.get("function_tolerance", 1e-6)), filter_input_nan=filter_input_nan)
    solver_invalids = res["invalid"].to(torch.bool).reshape(lead)
    invalid_dict = dict(solver_invalids=solver_invalids, invalids=solver_invalids.clone())
    return invalid_dict, res["states"].reshape(lead + (7,))
