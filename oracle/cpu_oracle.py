"""ctypes front-end of the CPU oracle (oracle/liblc_oracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module.  It restates
the reference's algorithm for the LC hot path on the CPU (see the headers of
lc_oracle.c / lm_oracle.c for the file:line map and the parity-pinning status).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblc_oracle.so")
_lib = None

LM_TOL_NEEDS_SUCCESS = 1
LM_TRACE_COLS = 4
TERM_NAMES = {0: "CONVERGENCE", 1: "NO_CONVERGENCE", 2: "FAILURE"}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("lc_oracle.c", "lm_oracle.c", "p3_oracle.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def _f64(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a), dtype=np.float64)


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def lc_loss(K, pose, pts3d, pts2d, inv_std, valid, bbox_3d, max_err_len=32.0, rel_thresh=3.0, w_e_thresh=4.0,
            want_jac=False, threads=None):
    """fp64 LC loss forward + gradients (for d loss_b = 1).  Returns a dict of numpy arrays."""
    K, pose, X, x, s, v, bb = map(_f64, (K, pose, pts3d, pts2d, inv_std, valid, bbox_3d))
    B, N = X.shape[0], X.shape[1]
    threads = threads or os.cpu_count()
    out = dict(loss=np.zeros(B), g_pts3d=np.zeros((B, N, 3)), g_pts2d=np.zeros((B, N, 2)),
               g_inv_std=np.zeros((B, N, 2)), cov=np.zeros((B, 6, 6)), update_cov=np.zeros((B, 6, 6)),
               W=np.zeros((B, N, 2)), sigma=np.zeros((B, N, 2)), flags=np.zeros(B, np.int32))
    jac = np.zeros((B, 6, N, 2)) if want_jac else None
    scratch = np.zeros(threads * 8 * N)
    D, I = C.c_double, C.c_int
    lib().lc_oracle_batch(
        I(B), I(N), _p(K, D), _p(pose, D), _p(X, D), _p(x, D), _p(s, D), _p(v, D), _p(bb, D),
        D(max_err_len), D(rel_thresh), D(w_e_thresh), _p(out["loss"], D), _p(out["g_pts3d"], D),
        _p(out["g_pts2d"], D), _p(out["g_inv_std"], D), _p(jac, D), _p(out["cov"], D), _p(out["update_cov"], D),
        _p(out["W"], D), _p(out["sigma"], D), _p(out["flags"], I), _p(scratch, D), I(threads))
    if want_jac:
        out["jac"] = jac
    return out


def lm_solve(K, pts3d, pts2d, L, start, n_points=None, max_iter=50, function_tolerance=1e-6,
             flags=LM_TOL_NEEDS_SUCCESS, threads=None, want_trace=False):
    """Ceres-faithful LM (parity unpinned).  fp32 arrays: K (B,3,3), pts3d (B,N,3), pts2d (B,N,2),
    L (B,N,2,2) lower Cholesky factor of the inverse covariance, start (B,7)."""
    K, X, x, L, st = map(_f32, (K, pts3d, pts2d, L, start))
    st = st.copy()
    B, N = X.shape[0], X.shape[1]
    npts = None if n_points is None else np.ascontiguousarray(n_points, dtype=np.int32)
    threads = threads or os.cpu_count()
    radius = np.zeros(B, np.float32)
    invalid = np.zeros(B, np.int32)
    iters = np.zeros(B, np.int32)
    term = np.zeros(B, np.int32)
    x6 = np.zeros((B, 6))
    trace = np.full((B, max_iter + 2, LM_TRACE_COLS), np.nan) if want_trace else None
    F, D, I = C.c_float, C.c_double, C.c_int
    lib().lm_oracle_batch(I(B), I(N), _p(st, F), _p(K, F), _p(x, F), _p(X, F), _p(L, F), _p(npts, I), I(max_iter),
                          F(function_tolerance), I(flags), _p(radius, F), _p(invalid, I), _p(iters, I), _p(term, I),
                          _p(x6, D), _p(trace, D), I(threads))
    out = dict(states=st, radius=radius, invalid=invalid, iters=iters, term=term, x6=x6)
    if want_trace:
        out["trace"] = trace
    return out


def lm_eval(x6, K, pts3d, pts2d, L):
    """Residuals, Jacobian (2N,6), gradient and cost of one problem at x6 (angle-axis + t)."""
    K, X, x, L = map(_f32, (K, pts3d, pts2d, L))
    x6 = _f64(x6)
    N = X.shape[0]
    cost = C.c_double(0)
    r = np.zeros(2 * N); J = np.zeros((2 * N, 6)); g = np.zeros(6)
    F, D, I = C.c_float, C.c_double, C.c_int
    ok = lib().lm_oracle_eval(_p(x6, D), _p(K, F), _p(x, F), _p(X, F), _p(L, F), I(N), C.byref(cost), _p(r, D), _p(J, D), _p(g, D))
    return dict(ok=bool(ok), cost=cost.value, r=r, J=J, g=g)


def p3(K, pts3d, pts2d, inv_std, bbox_3d, states, mode=3, max_iter=50, function_tolerance=1e-6,
       flags=LM_TOL_NEEDS_SUCCESS, max_err_len=32.0, rel_thresh=3.0, w_e_thresh=4.0, threads=None, want_grads=True):
    """P3 = LM solve from `states` then LC loss fwd+bwd at the solution (mode bit0 = LM, bit1 = LC)."""
    K, X, x, s, bb, st = map(_f32, (K, pts3d, pts2d, inv_std, bbox_3d, states))
    st = st.copy()
    B, N = X.shape[0], X.shape[1]
    threads = threads or os.cpu_count()
    radius = np.zeros(B, np.float32); invalid = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
    loss = np.zeros(B); lcf = np.zeros(B, np.int32)
    g3 = np.zeros((B, N, 3), np.float32) if want_grads else None
    g2 = np.zeros((B, N, 2), np.float32) if want_grads else None
    gs = np.zeros((B, N, 2), np.float32) if want_grads else None
    F, D, I = C.c_float, C.c_double, C.c_int
    lib().p3_oracle_batch(I(B), I(N), I(mode), _p(st, F), _p(K, F), _p(X, F), _p(x, F), _p(s, F), _p(bb, F),
                          I(max_iter), F(function_tolerance), I(flags), D(max_err_len), D(rel_thresh), D(w_e_thresh),
                          _p(radius, F), _p(invalid, I), _p(iters, I), _p(loss, D), _p(g3, F), _p(g2, F), _p(gs, F),
                          _p(lcf, I), I(threads))
    return dict(states=st, radius=radius, invalid=invalid, iters=iters, loss=loss, g_pts3d=g3, g_pts2d=g2,
                g_inv_std=gs, lc_flags=lcf)


def dense_pose_loss(xyz_noc, logits, scale, noc_scale, K, pose, bbox_3d, sample, top_left, max_err_len=32.0):
    """CPU restatement (numpy fp64 + lc_oracle.c) of the gdr-net glue of Loss_fn.dense_pose_loss and its backward:
      losses.py:355-356  joint softmax over the 2*H*W logits, times xyz_weights_scale
      losses.py:142-161  dense_pnp_matching_from_xyz: strided sub-sample of weights / xyz_noc*noc_scale / gen_uv grid
      losses.py:366,383  valid = ones, Loss_cov_mixed(...)
    Returns dict(loss (B,), g_xyz_noc (B,3,H,W), g_logits (B,2,H,W), g_scale (B,)) for d loss_b = 1.
    Pinned by tests/golden/densex_*.npz (generated by the unmodified reference)."""
    xyz_noc, logits = np.asarray(xyz_noc, np.float64), np.asarray(logits, np.float64)
    scale = np.asarray(scale, np.float64).reshape(-1)
    noc_scale = np.asarray(noc_scale, np.float64)
    B, _, H, W = xyz_noc.shape
    top, left = int(top_left[0]), int(top_left[1])
    flat = logits.reshape(B, -1)
    p = np.exp(flat - flat.max(1, keepdims=True))
    p = (p / p.sum(1, keepdims=True)).reshape(B, 2, H, W)
    w = p * scale[:, None, None, None]
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    sl = (slice(None), slice(None), slice(top, None, sample), slice(left, None, sample))
    inv_std = w[sl].reshape(B, 2, -1).transpose(0, 2, 1)
    pts3d = xyz_noc[sl].reshape(B, 3, -1).transpose(0, 2, 1) * noc_scale[:, None, :]
    pts2d = np.stack((xs[top::sample, left::sample].reshape(-1), ys[top::sample, left::sample].reshape(-1)), -1)
    pts2d = np.broadcast_to(pts2d, (B,) + pts2d.shape)
    o = lc_loss(K, pose, pts3d, pts2d, inv_std, np.ones(pts3d.shape[:2]), bbox_3d, max_err_len=max_err_len)
    Hn, Wn = len(range(top, H, sample)), len(range(left, W, sample))
    g_w = np.zeros_like(w)
    g_w[sl] = o["g_inv_std"].transpose(0, 2, 1).reshape(B, 2, Hn, Wn)
    g_xyz = np.zeros_like(xyz_noc)
    g_xyz[sl] = (o["g_pts3d"] * noc_scale[:, None, :]).transpose(0, 2, 1).reshape(B, 3, Hn, Wn)
    S = (g_w * p).reshape(B, -1).sum(1)                       # d/d scale
    g_logits = w * (g_w - S[:, None, None, None])             # softmax backward
    return dict(loss=o["loss"], g_xyz_noc=g_xyz, g_logits=g_logits, g_scale=S)
