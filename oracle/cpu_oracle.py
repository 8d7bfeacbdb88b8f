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

# rule switches of oracle/lm_oracle.c (every from-memory Ceres rule is switchable; tools/lm_sensitivity.py flips them)
LM_TOL_NEEDS_SUCCESS = 1
LM_INVALID_DIV_DEC, LM_GRAD_TEST_ALWAYS, LM_REPORT_CURRENT_RADIUS, LM_SOLVE_NORMAL_EQ = 2, 4, 8, 16
LM_TOL_KEEP_CANDIDATE, LM_GRAD_NORM_PLAIN, LM_SCALE_NO_PLUS_ONE = 32, 64, 128
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
            want_jac=False, threads=None, cov_2d=False):
    """fp64 LC loss forward + gradients (for d loss_b = 1).  Returns a dict of numpy arrays.
    ``cov_2d``: the projected-bbox variant of cov_mixed.py:76-80, 91-97."""
    K, pose, X, x, s, v, bb = map(_f64, (K, pose, pts3d, pts2d, inv_std, valid, bbox_3d))
    B, N = X.shape[0], X.shape[1]
    threads = threads or os.cpu_count()
    out = dict(loss=np.zeros(B), g_pts3d=np.zeros((B, N, 3)), g_pts2d=np.zeros((B, N, 2)),
               g_inv_std=np.zeros((B, N, 2)), cov=np.zeros((B, 6, 6)), update_cov=np.zeros((B, 6, 6)),
               W=np.zeros((B, N, 2)), sigma=np.zeros((B, N, 2)), flags=np.zeros(B, np.int32))
    jac = np.zeros((B, 6, N, 2)) if want_jac else None
    scratch = np.zeros(threads * 8 * N)
    D, I = C.c_double, C.c_int
    lib().lc_oracle_batch_ex(
        I(B), I(N), _p(K, D), _p(pose, D), _p(X, D), _p(x, D), _p(s, D), _p(v, D), _p(bb, D),
        D(max_err_len), D(rel_thresh), D(w_e_thresh), _p(out["loss"], D), _p(out["g_pts3d"], D),
        _p(out["g_pts2d"], D), _p(out["g_inv_std"], D), _p(jac, D), _p(out["cov"], D), _p(out["update_cov"], D),
        _p(out["W"], D), _p(out["sigma"], D), _p(out["flags"], I), _p(scratch, D), I(threads), I(1 if cov_2d else 0))
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


# ------------------------------------------------------------------------------------------------
# ZebraPose binary-code producer
# ------------------------------------------------------------------------------------------------
def _axis_slices(bit_cnt):
    bit_cnt = [int(b) for b in ([bit_cnt] * 3 if np.isscalar(bit_cnt) else bit_cnt)]
    off = np.concatenate(([0], np.cumsum(bit_cnt)))
    return bit_cnt, [slice(int(off[a]), int(off[a + 1])) for a in range(3)]


def noc_to_bits(noc, bit_cnt, black_background=True):
    """floatbits.py:76-97 mod_noc2bits_bb / :13-31 nn_noc2target: noc (B,H,W,3) in (-1,1) -> (mod_bits, raw_bits), both
    (B,sum(bit_cnt),H,W) bool.  raw = MSB-first binary of round(clamp((noc+1)*max/2, 0, max)); mod = Gray code of it with
    the two leading bits inverted under a black background."""
    noc = np.asarray(noc)
    if noc.dtype not in (np.float32, np.float64):
        noc = noc.astype(np.float64)
    ft = noc.dtype.type                      # the reference computes in the tensor's dtype (fp32 in training)
    bit_cnt, _ = _axis_slices(bit_cnt)
    mods, raws = [], []
    for a, N in enumerate(bit_cnt):
        mx = 2 ** N - 1
        ints = np.rint(np.clip((noc[..., a] + ft(1)) * ft(mx * 0.5), ft(0), ft(mx))).astype(np.int64)   # torch.round = half-to-even = rint
        raw = ((ints[..., None] >> np.arange(N - 1, -1, -1)) & 1).astype(bool)
        mod = raw.copy()
        mod[..., 1:] ^= raw[..., :-1]
        if black_background:
            mod[..., 0:2] = ~mod[..., 0:2]
        mods.append(mod)
        raws.append(raw)
    return np.concatenate(mods, -1).transpose(0, 3, 1, 2), np.concatenate(raws, -1).transpose(0, 3, 1, 2)


def noc_bin_decode_with_gt(logits, raw_bits, msk, bit_cnt, black_background=True):
    """floatbits.py:49-69 nn_logits2noc_with_gt -> :99-160 mod_logits2float_with_gt_bb_scripted, per axis:
      signed logits l' = l * m, m_j = -1 where the previous GT bit is set, m_0, m_1 *= -1 (black background)   (:139-142)
      pred = l' > 0; idx = first bit where pred != gt (the LSB if none)                                          (:146-152)
      in-mask value  = sum of GT bits except bit idx, weighted 2^(N-1-j), + sigmoid(l'_idx) * 2^(N-1-idx)       (:153-157)
      out-mask value = sum of pred bits weighted                                                                (:147)
      noc = val / (max_val/2) - 1                                                                               (:112)
    logits (B,C,H,W) float, raw_bits (B,C,H,W) bool, msk (B,H,W) bool.
    Returns noc (B,H,W,3) and, for the backward, sel (B,H,W,3) channel index carrying the gradient and
    dnoc (B,H,W,3) = d noc_a / d logits[sel_a] (0 outside the mask)."""
    logits = np.asarray(logits, np.float64)
    raw_bits = np.asarray(raw_bits).astype(bool)
    msk = np.asarray(msk).astype(bool)
    bit_cnt, sls = _axis_slices(bit_cnt)
    B, _, H, W = logits.shape
    noc = np.zeros((B, H, W, 3))
    sel = np.zeros((B, H, W, 3), np.int64)
    dnoc = np.zeros((B, H, W, 3))
    bf = -1.0 if black_background else 1.0
    for a, (N, sl) in enumerate(zip(bit_cnt, sls)):
        l = logits[:, sl].transpose(0, 2, 3, 1)
        g = raw_bits[:, sl].transpose(0, 2, 3, 1)
        m = np.ones_like(l)
        m[..., 1:][g[..., :-1]] = -1.0
        m[..., 0:2] *= bf
        lp = l * m
        wgt = 2.0 ** np.arange(N - 1, -1, -1)
        pred = lp > 0
        out_val = (pred * wgt).sum(-1)
        err = pred ^ g
        err[..., -1] = True
        idx = np.argmax(err, -1)
        g_wo = g.copy()
        np.put_along_axis(g_wo, idx[..., None], False, -1)
        lsel = np.take_along_axis(lp, idx[..., None], -1)[..., 0]
        sg = 1.0 / (1.0 + np.exp(-lsel))
        in_val = (g_wo * wgt).sum(-1) + sg * wgt[idx]
        val = np.where(msk, in_val, out_val)
        half = (2 ** N - 1) * 0.5
        noc[..., a] = val / half - 1
        sel[..., a] = sl.start + idx
        dnoc[..., a] = np.where(msk, sg * (1 - sg) * wgt[idx] * np.take_along_axis(m, idx[..., None], -1)[..., 0] / half, 0.0)
    return noc, sel, dnoc


def noc_bin_decode(logits, bit_cnt, black_background=True):
    """floatbits.py:33-47 nn_logits2noc (inference, no LUT) -> :197-224 mod_logits2float_bb, per axis: hard Gray bits
    (leading two inverted under a black background) -> binary integer; its LSB is replaced by
    sigmoid(l_last * (1 - (val & 2))).  Returns noc (B,H,W,3)."""
    logits = np.asarray(logits, np.float64)
    bit_cnt, sls = _axis_slices(bit_cnt)
    B, _, H, W = logits.shape
    noc = np.zeros((B, H, W, 3))
    for a, (N, sl) in enumerate(zip(bit_cnt, sls)):
        l = logits[:, sl].transpose(0, 2, 3, 1)
        bits = l > 0
        if black_background:
            bits[..., 0:2] = ~bits[..., 0:2]
        binb = np.bitwise_xor.accumulate(bits.astype(np.int64), axis=-1)     # Gray -> binary, MSB first
        val = (binb * (2 ** np.arange(N - 1, -1, -1))).sum(-1)
        lsb_factor = 1 - (val & 2)
        v = (val & -2) + 1.0 / (1.0 + np.exp(-(l[..., -1] * lsb_factor)))
        noc[..., a] = v / ((2 ** N - 1) * 0.5) - 1
    return noc


def dense_pose_loss_noc_bin(bin_logits, raw_bits, msk_noc, logits, scale, noc_scale, K, pose, bbox_3d, bit_cnt, sample,
                            top_left, model_transform=None, max_err_len=32.0, black_background=True):
    """CPU restatement of the ZebraPose branch of Loss_fn.dense_pose_loss and its backward:
      losses.py:355-356  joint softmax * scale (as dense_pose_loss above)
      losses.py:163-184  dense_pnp_matching_from_noc_bin: strided sub-sample, nn_out_to_xyz(raw_bits_gt, noc_mask, ...)
      losses.py:16-45    nn_out_to_xyz: noc * noc_scale, then (xyz - T[:3,3]) @ T[:3,:3] with the model transform
      losses.py:375,383  valid = ones, Loss_cov_mixed(...)
    Returns dict(loss, g_bin_logits (B,C,H,W), g_logits (B,2,H,W), g_scale (B,), pts3d (B,n,3))."""
    logits = np.asarray(logits, np.float64)
    scale = np.asarray(scale, np.float64).reshape(-1)
    noc_scale = np.asarray(noc_scale, np.float64)
    B, C_, H, W = np.asarray(bin_logits).shape
    top, left = int(top_left[0]), int(top_left[1])
    flat = logits.reshape(B, -1)
    p = np.exp(flat - flat.max(1, keepdims=True))
    p = (p / p.sum(1, keepdims=True)).reshape(B, 2, H, W)
    w = p * scale[:, None, None, None]
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    sl = (slice(None), slice(None), slice(top, None, sample), slice(left, None, sample))
    sl3 = (slice(None), slice(top, None, sample), slice(left, None, sample))
    inv_std = w[sl].reshape(B, 2, -1).transpose(0, 2, 1)
    noc, sel, dnoc = noc_bin_decode_with_gt(np.asarray(bin_logits)[sl], np.asarray(raw_bits)[sl], np.asarray(msk_noc)[sl3],
                                            bit_cnt, black_background)
    Hn, Wn = noc.shape[1], noc.shape[2]
    xf = noc * noc_scale[:, None, None, :]
    if model_transform is not None:
        T = np.asarray(model_transform, np.float64)
        M = T[:, :3, :3]
        xyz = np.einsum("bhwa,bak->bhwk", xf - T[:, None, None, :3, 3], M)
    else:
        M = np.broadcast_to(np.eye(3), (B, 3, 3))
        xyz = xf
    pts3d = xyz.reshape(B, -1, 3)
    pts2d = np.stack((xs[top::sample, left::sample].reshape(-1), ys[top::sample, left::sample].reshape(-1)), -1)
    pts2d = np.broadcast_to(pts2d, (B,) + pts2d.shape)
    o = lc_loss(K, pose, pts3d, pts2d, inv_std, np.ones(pts3d.shape[:2]), bbox_3d, max_err_len=max_err_len)
    g_w = np.zeros_like(w)
    g_w[sl] = o["g_inv_std"].transpose(0, 2, 1).reshape(B, 2, Hn, Wn)
    S = (g_w * p).reshape(B, -1).sum(1)
    g_logits = w * (g_w - S[:, None, None, None])
    g_xf = np.einsum("bhwk,bak->bhwa", o["g_pts3d"].reshape(B, Hn, Wn, 3), M)
    g_sel = g_xf * noc_scale[:, None, None, :] * dnoc                 # d/d logits[sel]
    g_bin_s = np.zeros((B, C_, Hn, Wn))
    for a in range(3):
        np.put_along_axis(g_bin_s, sel[..., a][:, None], g_sel[..., a][:, None], 1)
    g_bin = np.zeros((B, C_, H, W))
    g_bin[sl] = g_bin_s
    return dict(loss=o["loss"], g_bin_logits=g_bin, g_logits=g_logits, g_scale=S, pts3d=pts3d)


# ------------------------------------------------------------------------------------------------
# Test-time point selection: test.py:67-119
# ------------------------------------------------------------------------------------------------
def _quantile_f32(vals, q):
    """torch.quantile(vals (N,) fp32, q fp32 scalar), default 'linear' interpolation, restated in fp32:
    rank = q*(N-1); lerp(sorted[floor], sorted[ceil], rank-floor) with torch's lerp (w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w))."""
    f = np.float32
    srt = np.sort(vals.astype(f))
    rank = f(q) * f(len(srt) - 1)
    lo = np.floor(rank)
    w = f(rank - lo)
    a, b = srt[int(lo)], srt[min(int(np.ceil(rank)), len(srt) - 1)]
    d = f(b - a)
    return f(a + f(w * d)) if w < f(0.5) else f(b - f(d * f(f(1) - w)))


def dense_point_select(xyz, weights, msk_vis_logits, sample, mode, quantile=0.2, seg_thresh=0.5):
    """CPU restatement (numpy, fp32 where the comparison outcome depends on it) of the selection in test.solve_pnp_dense:
      test.py:70      seg_msk = sigmoid(msk_vis_logits) > seg_thresh
      losses.py:142-161 (top_left=(0,0)): strided sub-sample of the gen_uv grid / weights / xyz / mask
      test.py:95      den_inv_cov2d = den_inv_std2d ** 2
      test.py:97-104  dense_point_select = 'mask' | 'quantile' | 'quantile_in_mask' (test.py:36-45 quantile_msk)
      test.py:106     valid_index_lst = [v.nonzero()[:,0] ...]  (ordered indices)
    xyz (B,H,W,3), weights (B,2,H,W) = softmax * scale, msk_vis_logits (B,1,H,W).
    Returns dict(valid (B,N) bool, pts3d (B,N,3), pts2d (N,2), inv_std (B,N,2), inv_cov (B,N,2), thr (B,))."""
    f = np.float32
    xyz, weights, ml = np.asarray(xyz, f), np.asarray(weights, f), np.asarray(msk_vis_logits, f)
    B, H, W, _ = xyz.shape
    seg = (f(1) / (f(1) + np.exp(-ml[:, 0]).astype(f))) > f(seg_thresh)
    ys, xs = np.meshgrid(np.arange(H, dtype=f), np.arange(W, dtype=f), indexing="ij")
    pts2d = np.stack((xs[::sample, ::sample].reshape(-1), ys[::sample, ::sample].reshape(-1)), -1)
    inv_std = weights[:, :, ::sample, ::sample].reshape(B, 2, -1).transpose(0, 2, 1)
    pts3d = xyz[:, ::sample, ::sample].reshape(B, -1, 3)
    m = seg[:, ::sample, ::sample].reshape(B, -1)
    N = m.shape[1]
    thr = np.zeros(B, f)
    if mode == "mask":
        valid = m.copy()
    elif mode == "quantile":
        wsum = (inv_std[..., 0] + inv_std[..., 1]).astype(f)
        thr = np.array([_quantile_f32(wsum[b], quantile) for b in range(B)], f)
        valid = wsum >= thr[:, None]
    elif mode == "quantile_in_mask":
        vis_ratio = (m.astype(f).sum(1, dtype=f) / f(N)).astype(f)       # seg_valid_mask.float().mean(-1)
        qb = (f(1) - (f(1 - quantile) * vis_ratio).astype(f)).astype(f)  # 1 - (1-cfg.quantile) * vis_ratio
        mf = m.astype(f)
        wsum = ((inv_std[..., 0] * mf).astype(f) + (inv_std[..., 1] * mf).astype(f)).astype(f)
        thr = np.array([_quantile_f32(wsum[b], qb[b]) for b in range(B)], f)
        valid = (wsum >= thr[:, None]) & m
    else:
        raise ValueError(mode)
    return dict(valid=valid, pts3d=pts3d, pts2d=pts2d, inv_std=inv_std, inv_cov=(inv_std * inv_std).astype(f), thr=thr)


# ------------------------------------------------------------------------------------------------
# Pose-error metrics and symmetric candidate selection
# ------------------------------------------------------------------------------------------------
def pose_errors(R_est, t_est, R_gt, t_gt, pts):
    """lib/utils/evaluate.py:333-339 compute_pose_errors for a batch: error6d.add (:87-101), adi (:104-124, nearest neighbour by
    brute force instead of cKDTree), re (:127-140, degrees), te (:143-152).  fp64."""
    R_est, t_est, R_gt, t_gt, pts = (np.asarray(v, np.float64) for v in (R_est, t_est, R_gt, t_gt, pts))
    B = len(R_est)
    out = {k: np.zeros(B) for k in ("adi", "add", "re", "te")}
    for b in range(B):
        pe = pts @ R_est[b].T + t_est[b].reshape(1, 3)
        pg = pts @ R_gt[b].T + t_gt[b].reshape(1, 3)
        out["add"][b] = np.linalg.norm(pe - pg, axis=1).mean()
        nn = np.empty(len(pts))
        for j0 in range(0, len(pts), 512):
            d2 = ((pg[j0:j0 + 512, None, :] - pe[None, :, :]) ** 2).sum(-1)
            nn[j0:j0 + 512] = np.sqrt(d2.min(1))
        out["adi"][b] = nn.mean()
        cs = min(1.0, max(-1.0, 0.5 * (np.trace(R_est[b] @ np.linalg.inv(R_gt[b])) - 1.0)))
        out["re"][b] = 180.0 * np.arccos(cs) / np.pi
        out["te"][b] = np.linalg.norm(t_gt[b].reshape(3) - t_est[b].reshape(3))
    return out


def select_pose(mode, cam_K, pts_a, pts_b, pose_candi):
    """symmetry.py:8-31 select_pose_2d (mode 0: pts_a = pts3d, pts_b = pts2d) / :33-56 select_pose_3d (mode 1: pts_a = pts3d_out,
    pts_b = homo_z).  Returns (best (B,3,4), index (B,), err (B,K)), fp64 arithmetic."""
    K, A, Bp, C_ = (np.asarray(v, np.float64) for v in (cam_K, pts_a, pts_b, pose_candi))
    R, t = C_[..., :3, :3], C_[..., :3, 3]
    if mode == 0:
        x = np.einsum("bkij,bnj->bkni", R, A) + t[:, :, None, :]
        h = np.einsum("bij,bknj->bkni", K, x)
        err = np.linalg.norm(h[..., :2] / h[..., 2:3] - Bp[:, None], axis=-1).mean(-1)
    else:
        cam = np.einsum("bij,bnj->bni", np.linalg.inv(K), Bp)
        ref = np.einsum("bkji,bknj->bkni", R, cam[:, None] - t[:, :, None, :])
        err = np.linalg.norm(A[:, None] - ref, axis=-1).mean(-1)
    idx = err.argmin(1)
    return C_[np.arange(len(idx)), idx], idx, err


# ------------------------------------------------------------------------------------------------
# Device-side initialiser: numpy restatement of lc_init.cu (our algorithm; the reference uses OpenCV)
# ------------------------------------------------------------------------------------------------
def pnp_init(K, pts3d, pts2d, weights=None, reproj_thresh=3.0, irls_rounds=3, max_points=1024):
    """Weighted DLT reduced to a 4x4 eigenproblem + Cauchy IRLS, one pose.  K (3,3), pts3d (n,3), pts2d (n,2), weights (n,2)
    inverse variances or None.  Returns (ok, R (3,3), t (3,), inlier (n,) bool).  Same steps and formulas as lc_init.cu."""
    K, X, x = np.asarray(K, np.float64), np.asarray(pts3d, np.float64), np.asarray(pts2d, np.float64)
    n = len(X)
    n_all = n
    if n < 6:
        return False, np.eye(3), np.zeros(3), np.zeros(n_all, bool)
    base = np.ones(n) if weights is None else 0.5 * np.asarray(weights, np.float64).sum(1)
    base = np.where((base > 0) & np.isfinite(base), base, 0.0)
    Ki = np.linalg.inv(K)
    X_all, x_all = X, x
    step = -(-n // max_points) if n > max_points else 1            # the sums use every step-th correspondence (lc_init.cu)
    X, x, base = X[::step], x[::step], base[::step]
    n = len(X)
    cen = X.mean(0)
    var = (X * X).mean(0) - cen * cen
    sc = np.sqrt(var.sum() / 3.0) if var.sum() > 0 else 1.0
    Y = np.concatenate(((X - cen) / sc, np.ones((n, 1))), 1)
    h = np.concatenate((x, np.ones((n, 1))), 1) @ Ki.T
    xh, yh = h[:, 0] / h[:, 2], h[:, 1] / h[:, 2]
    R, t = np.eye(3), np.zeros(3)

    def reproj_err2(R, t):
        P = X @ R.T + t
        hh = P @ K.T
        e = hh[:, :2] / hh[:, 2:3] - x
        return (e * e).sum(1), hh[:, 2]

    for rnd in range(irls_rounds + 1):
        wt = base.copy()
        if rnd > 0:
            e2, h2 = reproj_err2(R, t)
            e2 = e2 / reproj_thresh ** 2
            wt = wt * np.where((h2 > 0) & np.isfinite(e2), 1.0 / (1.0 + e2), 0.0)
        YY = Y[:, :, None] * Y[:, None, :]
        S = (wt[:, None, None] * YY).sum(0)
        Sx = ((wt * xh)[:, None, None] * YY).sum(0)
        Sy = ((wt * yh)[:, None, None] * YY).sum(0)
        Sq = ((wt * (xh * xh + yh * yh))[:, None, None] * YY).sum(0)
        try:
            np.linalg.cholesky(S)
        except np.linalg.LinAlgError:
            return False, np.eye(3), np.zeros(3), np.zeros(n_all, bool)
        Zx, Zy = np.linalg.solve(S, Sx), np.linalg.solve(S, Sy)
        D = Sq - Sx @ Zx - Sy @ Zy
        D = 0.5 * (D + D.T)
        tr = np.trace(D)
        if not (tr > 0):
            return False, np.eye(3), np.zeros(3), np.zeros(n_all, bool)
        try:
            Lc = np.linalg.cholesky(D + 1e-12 * tr * np.eye(4))      # inverse iteration, as lc_init.cu::sym4_min_eigvec
        except np.linalg.LinAlgError:
            return False, np.eye(3), np.zeros(3), np.zeros(n_all, bool)
        p3 = np.array([0.1, 0.1, 0.1, 1.0])
        for _ in range(5):
            p3 = np.linalg.solve(Lc.T, np.linalg.solve(Lc, p3))
            p3 = p3 / np.linalg.norm(p3)
        P = np.stack((Zx @ p3, Zy @ p3, p3))
        Q = np.concatenate((P[:, :3] / sc, (P[:, 3] - P[:, :3] @ cen / sc)[:, None]), 1)
        if Q[2, :3] @ cen + Q[2, 3] < 0:
            Q = -Q
        n1, n2 = np.linalg.norm(Q[0, :3]), np.linalg.norm(Q[1, :3])
        lam = 0.5 * (n1 + n2)
        r1 = Q[0, :3] / n1
        r2 = Q[1, :3] - (r1 @ Q[1, :3]) * r1
        r2 = r2 / np.linalg.norm(r2)
        R = np.stack((r1, r2, np.cross(r1, r2)))
        t = Q[:, 3] / lam
        if not (np.isfinite(R).all() and np.isfinite(t).all() and lam > 0):
            return False, np.eye(3), np.zeros(3), np.zeros(n_all, bool)
    X, x = X_all, x_all                                            # the inlier mask covers every correspondence
    e2, h2 = reproj_err2(R, t)
    return True, R, t, (h2 > 0) & (e2 < reproj_thresh ** 2)
