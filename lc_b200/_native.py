"""ctypes binding of the C ABI in ``include/lc_b200.h`` (``liblc_b200.so``).

PyTorch is used here only as plumbing: device memory, the current CUDA stream and
dtype/stride bookkeeping.  Every call hands raw device pointers + element strides to
the hand-written sm_100a kernels.  There is NO CPU or PyTorch fallback: if the shared
library is missing, or the tensors are not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("LC_B200_LIB") or os.path.join(_HERE, "liblc_b200.so")   # env override: instrumented builds (tools/)
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("lc_abi.cu", "lc_stream.cu", "lc_resident.cu", "lc_resident_vec.cu", "lc_resident_cluster.cu", "lc_resident_lm3.cu", "lc_persist.cu", "lc_tiny.cu", "lc_dense.cu", "lc_select.cu", "lc_eval.cu", "lc_init.cu", "lc_compat.cu")]
HEADERS = [os.path.join(_HERE, "csrc", "lc_device.cuh"), os.path.join(_HERE, "csrc", "lc_pose.cuh"),
           os.path.join(_HERE, "csrc", "lc_resident.cuh"), os.path.join(_HERE, "csrc", "lc_vec.cuh"),
           os.path.join(_HERE, "csrc", "lc_resident_kernel.cuh"), os.path.join(_HERE, "csrc", "lc_point.cuh"),
           os.path.join(_ROOT, "include", "lc_b200.h")]
BUILD_DIR = os.path.join(_HERE, "csrc", "build")

ABI_VERSION = 2
LC_F32, LC_F64 = 0, 1
W_ICOV_DIAG, W_ICOV_FULL, W_INV_STD, W_SQRT_L = 0, 1, 2, 3
FLAG_NAN_TO_NUM, FLAG_TOL_NEEDS_SUCCESS, FLAG_EXACT_HESSIAN, FLAG_FORCE_STREAMING, FLAG_LM_MIXED, FLAG_COV_2D = 1, 2, 4, 8, 16, 32
ST_HESS_NOT_SPD, ST_PRIOR_NOT_GOOD, ST_COV_NOT_GOOD = 1, 2, 4

EXPORTS = ("lc_b200_abi_version", "lc_b200_last_error", "lc_b200_last_launch_count", "lc_b200_lm_solve",
           "lc_b200_loss_fwd_bwd", "lc_b200_solve_loss", "lc_b200_pnp_jac_cov", "lc_b200_pnp_jac_cov_bwd",
           "lc_b200_dense_loss_fwd_bwd", "lc_b200_noc_bin_decode", "lc_b200_dense_select", "lc_b200_pose_errors",
           "lc_b200_select_pose", "lc_b200_pnp_init", "lc_b200_noc_bin_encode")
# exports with their own signatures (not the (args*, stream) pattern)
EXTRA_EXPORTS = ("lc_b200_last_kernels", "pnp_ceres_f32_omp", "pnp_ceres_f32", "lc_b200_compat_release")


class NativeLibraryError(RuntimeError):
    pass


class lc_view(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride", C.c_int64 * 4)]


_VIEW_FIELDS = ("K", "pose", "pts3d", "pts2d", "weights", "valid", "bbox", "grad_out")
_OUT_VIEW_FIELDS = ("loss", "g_pts3d", "g_pts2d", "g_weights", "cov", "update_cov", "jac", "g_jac", "g_cov", "state", "radius")


class lc_args(C.Structure):
    _fields_ = ([("abi_version", C.c_int32), ("B", C.c_int32), ("N", C.c_int32), ("dtype", C.c_int32),
                 ("flags", C.c_int32), ("weight_mode", C.c_int32), ("max_iter", C.c_int32), ("reserved0", C.c_int32),
                 ("function_tolerance", C.c_double), ("max_err_len", C.c_double), ("rel_thresh", C.c_double),
                 ("w_e_thresh", C.c_double), ("grad_scale", C.c_double)]
                + [(f, lc_view) for f in _VIEW_FIELDS]
                + [("n_points", C.c_void_p)]
                + [(f, lc_view) for f in _OUT_VIEW_FIELDS]
                + [("invalid", C.c_void_p), ("iters", C.c_void_p), ("lc_flags", C.c_void_p), ("trace", C.c_void_p),
                   ("loss_sum", C.c_void_p)])


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


_DENSE_VIEWS = ("xyz_noc", "logits", "weights_scale", "noc_scale", "K", "pose", "bbox", "grad_out", "loss", "g_xyz_noc",
                "g_logits", "g_scale", "cov", "update_cov")


_ZEBRA_VIEWS = ("noc_bin_logits", "noc_bin_raw", "msk_noc", "model_transform", "g_noc_bin")


class lc_dense_args(C.Structure):
    _fields_ = ([("abi_version", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                 ("sample", C.c_int32), ("top", C.c_int32), ("left", C.c_int32), ("reserved0", C.c_int32),
                 ("max_err_len", C.c_double), ("rel_thresh", C.c_double), ("w_e_thresh", C.c_double), ("grad_scale", C.c_double)]
                + [(f, lc_view) for f in _DENSE_VIEWS]
                + [("lc_flags", C.c_void_p), ("loss_sum", C.c_void_p)]
                + [(f, lc_view) for f in _ZEBRA_VIEWS]
                + [("bit_cnt", C.c_int32 * 3), ("black_background", C.c_int32)])


class lc_decode_args(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("bit_cnt", C.c_int32 * 3), ("black_background", C.c_int32),
                ("noc_bin_logits", lc_view), ("noc_scale", lc_view), ("model_transform", lc_view), ("xyz", lc_view)]


SEL_MASK, SEL_QUANTILE, SEL_QUANTILE_IN_MASK = 0, 1, 2


class lc_select_args(C.Structure):
    _fields_ = ([("abi_version", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                 ("sample", C.c_int32), ("mode", C.c_int32), ("scale_dim", C.c_int32), ("min_points", C.c_int32),
                 ("Nmax", C.c_int32), ("reserved0", C.c_int32), ("quantile", C.c_float), ("one_minus_quantile", C.c_float),
                 ("seg_thresh", C.c_float), ("reserved1", C.c_float)]
                + [(f, lc_view) for f in ("xyz", "noc_scale", "weights", "logits", "weights_scale", "msk_logits", "pts3d", "pts2d", "inv_cov")]
                + [("index", C.c_void_p), ("n_points", C.c_void_p)])


class lc_encode_args(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("bit_cnt", C.c_int32 * 3), ("black_background", C.c_int32), ("noc", lc_view),
                ("mod_bits", C.c_void_p), ("raw_bits", C.c_void_p)]


class lc_init_args(C.Structure):
    _fields_ = ([("abi_version", C.c_int32), ("B", C.c_int32), ("N", C.c_int32), ("irls_rounds", C.c_int32),
                 ("reproj_thresh", C.c_float), ("reserved0", C.c_float)]
                + [(f, lc_view) for f in ("K", "pts3d", "pts2d", "weights", "reproj_thresh_b")]
                + [("n_points", C.c_void_p), ("state", lc_view), ("invalid", C.c_void_p), ("inlier", C.c_void_p),
                   ("n_inliers", C.c_void_p)])


class lc_eval_args(C.Structure):
    _fields_ = ([("abi_version", C.c_int32), ("B", C.c_int32), ("M", C.c_int32), ("reserved0", C.c_int32)]
                + [(f, lc_view) for f in ("R_est", "t_est", "R_gt", "t_gt", "pts")]
                + [("pts_offset", C.c_void_p), ("pts_count", C.c_void_p)]
                + [(f, lc_view) for f in ("add", "adi", "re", "te")])


class lc_candi_args(C.Structure):
    _fields_ = ([("abi_version", C.c_int32), ("B", C.c_int32), ("N", C.c_int32), ("Kc", C.c_int32), ("mode", C.c_int32),
                 ("reserved0", C.c_int32)]
                + [(f, lc_view) for f in ("K", "pts_a", "pts_b", "candi", "best", "err")]
                + [("best_index", C.c_void_p)])


def nvcc_commands(out: str = LIB_PATH):
    """One `nvcc -c` per translation unit (run in parallel) and the final link."""
    objs = [os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o") for src in SOURCES]
    compiles = [["nvcc"] + NVCC_FLAGS + ["-c", "-o", obj, src] for src, obj in zip(SOURCES, objs)]
    link = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs
    return compiles, link


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the kernels in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    deps = SOURCES + HEADERS
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if force or stale:
        os.makedirs(BUILD_DIR, exist_ok=True)
        compiles, link = nvcc_commands()
        procs = []
        for cmd in compiles:
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append(subprocess.Popen(cmd, cwd=_ROOT))
        for cmd, pr in zip(compiles, procs):
            if pr.wait() != 0:
                raise subprocess.CalledProcessError(pr.returncode, cmd)
        if verbose:
            print(" ".join(link), flush=True)
        subprocess.run(link, check=True, cwd=_ROOT)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load liblc_b200.so; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(lc_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name in EXPORTS + EXTRA_EXPORTS:
            if not hasattr(handle, name):
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}")
        handle.lc_b200_last_error.restype = C.c_char_p
        handle.lc_b200_last_kernels.restype = C.c_char_p
        for name in EXPORTS[3:]:
            argt = {"lc_b200_dense_loss_fwd_bwd": lc_dense_args, "lc_b200_noc_bin_decode": lc_decode_args,
                    "lc_b200_dense_select": lc_select_args, "lc_b200_pose_errors": lc_eval_args,
                    "lc_b200_select_pose": lc_candi_args, "lc_b200_pnp_init": lc_init_args,
                    "lc_b200_noc_bin_encode": lc_encode_args}.get(name, lc_args)
            getattr(handle, name).argtypes = [C.POINTER(argt), C.c_void_p]
            getattr(handle, name).restype = C.c_int
        if handle.lc_b200_abi_version() != ABI_VERSION:
            raise NativeLibraryError("liblc_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return LC_F32
    if dt == torch.float64:
        return LC_F64
    raise TypeError(f"lc_b200 kernels take float32 or float64 tensors, got {dt}")


def view_of(t: Optional[torch.Tensor]) -> lc_view:
    v = lc_view()
    if t is None:
        v.ptr = None
        return v
    st = t.stride()
    n = len(st)
    if n > 4:
        raise ValueError("lc_view supports at most 4 dimensions")
    v.ptr = t.data_ptr()
    v.stride[:n] = st
    return v


def check_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise NativeLibraryError(
                "lc_b200 operators run only on CUDA tensors (sm_100a kernels; there is no CPU fallback); "
                f"got a tensor on {t.device}")
        dev = dev or t.device
        if t.device != dev:
            raise ValueError("all tensors must be on the same CUDA device")
    return dev


def empty_like_dense(t: torch.Tensor) -> torch.Tensor:
    """Uninitialised tensor of t's shape; keeps t's strides when t is dense and non-overlapping (e.g. the
    planar (B,N,C) views), contiguous otherwise (e.g. stride-0 broadcasts)."""
    expected = 1
    for size, stride in sorted(zip(t.shape, t.stride()), key=lambda p: p[1]):
        if size == 1:
            continue
        if stride != expected:
            return torch.empty(t.shape, dtype=t.dtype, device=t.device)
        expected *= size
    return torch.empty_like(t)


def fit(t: Optional[torch.Tensor], shape: tuple, dtype: torch.dtype) -> Optional[torch.Tensor]:
    """t as `dtype` broadcast to `shape`; returns t itself when nothing has to change (the common case: this is
    on the per-call critical path of small batches)."""
    if t is None:
        return None
    if t.dtype != dtype:
        t = t.to(dtype)
    if tuple(t.shape) != shape:
        t = t.expand(shape)
    return t


def make_args(B: int, N: int, dtype: torch.dtype, **kw) -> lc_args:
    """Fill an lc_args; tensor-valued keywords become views, int/float keywords are copied."""
    a = lc_args()
    a.abi_version = ABI_VERSION
    a.B, a.N, a.dtype = int(B), int(N), _dtype_code(dtype)
    a.max_iter = 50
    a.function_tolerance, a.max_err_len, a.rel_thresh, a.w_e_thresh, a.grad_scale = 1e-6, 32.0, 3.0, 4.0, 1.0
    for k, v in kw.items():
        if k in _VIEW_FIELDS or k in _OUT_VIEW_FIELDS:
            if v is not None and v.dtype != dtype:
                raise TypeError(f"{k}: expected dtype {dtype}, got {v.dtype}")
            setattr(a, k, view_of(v))
        elif k in ("n_points", "invalid", "iters", "lc_flags"):
            if v is not None and v.dtype != torch.int32:
                raise TypeError(f"{k} must be int32")
            setattr(a, k, None if v is None else v.data_ptr())
        elif k in ("trace", "loss_sum"):
            if v is not None and v.dtype != torch.float64:
                raise TypeError(f"{k} must be float64")
            setattr(a, k, None if v is None else v.data_ptr())
        else:
            setattr(a, k, v)
    return a


def as_c_float(x: float) -> float:
    """x rounded to a C float and back: the reference ABI carries function_tolerance as `float` (ext.h:10)."""
    return C.c_float(x).value


def call(name: str, args, device: torch.device) -> int:
    """Enqueue one entry point on torch's current stream of `device`; returns the launch count."""
    handle = lib()
    if torch.cuda.current_device() == device.index:
        rc = getattr(handle, name)(C.byref(args), C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    else:
        with torch.cuda.device(device):
            rc = getattr(handle, name)(C.byref(args), C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {handle.lc_b200_last_error().decode()}")
    return handle.lc_b200_last_launch_count()
