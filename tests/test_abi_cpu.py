"""CPU: the C-ABI library loads, exports what include/lc_b200.h declares, and the operators fail loudly off-GPU."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT
from lc_b200 import _native as nat


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lc_b200.h")).read()
    return sorted(set(re.findall(r"^\s*(?:int|void|const char\*)\s+((?:lc_b200_|pnp_ceres_)\w+)\s*\(", src, flags=re.M)))


def test_library_builds_and_exports_every_declared_symbol():
    nat.build()
    handle = ctypes.CDLL(nat.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/lc_b200.h but not exported"
    assert set(declared) == set(nat.EXPORTS) | set(nat.EXTRA_EXPORTS)
    # the symbol the reference's cffi module binds (lib/pnp/cxx/ext.h:1-14) is exported under its own name
    assert "pnp_ceres_f32_omp" in declared
    assert handle.lc_b200_abi_version() == nat.ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirror of lc_args == the C struct (size and a few offsets), compiled with the host gcc."""
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu",'
                 'sizeof(lc_args),offsetof(lc_args,K),offsetof(lc_args,n_points),offsetof(lc_args,loss),'
                 'offsetof(lc_args,invalid),offsetof(lc_args,loss_sum));return 0;}\n' % os.path.join(ROOT, "include", "lc_b200.h"))
    exe = tmp_path / "sz"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([gcc, "-o", str(exe), str(c)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    A = nat.lc_args
    assert got == [ctypes.sizeof(A), A.K.offset, A.n_points.offset, A.loss.offset, A.invalid.offset, A.loss_sum.offset]


def test_dense_struct_layout_matches_header(tmp_path):
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu",'
                 'sizeof(lc_dense_args),offsetof(lc_dense_args,xyz_noc),offsetof(lc_dense_args,loss),'
                 'offsetof(lc_dense_args,loss_sum),offsetof(lc_dense_args,noc_bin_logits),offsetof(lc_dense_args,g_noc_bin),'
                 'offsetof(lc_dense_args,bit_cnt),offsetof(lc_dense_args,black_background),'
                 'sizeof(lc_decode_args),offsetof(lc_decode_args,xyz),sizeof(lc_select_args),offsetof(lc_select_args,xyz),'
                 'offsetof(lc_select_args,n_points));return 0;}\n' % os.path.join(ROOT, "include", "lc_b200.h"))
    exe = tmp_path / "sz"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([gcc, "-o", str(exe), str(c)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    A, D, S = nat.lc_dense_args, nat.lc_decode_args, nat.lc_select_args
    assert got == [ctypes.sizeof(A), A.xyz_noc.offset, A.loss.offset, A.loss_sum.offset, A.noc_bin_logits.offset, A.g_noc_bin.offset,
                   A.bit_cnt.offset, A.black_background.offset, ctypes.sizeof(D), D.xyz.offset, ctypes.sizeof(S), S.xyz.offset,
                   S.n_points.offset]


def test_bad_arguments_are_rejected_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable here."""
    handle = nat.lib()
    a = nat.lc_args()
    a.abi_version = 999
    assert handle.lc_b200_loss_fwd_bwd(ctypes.byref(a), None) == -1
    assert b"abi_version" in handle.lc_b200_last_error()
    assert handle.lc_b200_lm_solve(None, None) == -3
    a.abi_version = nat.ABI_VERSION
    a.B, a.N, a.dtype = 4, 8, 7
    assert handle.lc_b200_loss_fwd_bwd(ctypes.byref(a), None) == -2      # bad dtype
    a.dtype = nat.LC_F32
    assert handle.lc_b200_loss_fwd_bwd(ctypes.byref(a), None) == -3      # required pointers missing
    a.B = 0
    assert handle.lc_b200_loss_fwd_bwd(ctypes.byref(a), None) == 0       # empty batch is a no-op
    assert handle.lc_b200_last_launch_count() == 0


def test_operators_refuse_cpu_tensors():
    """No CPU / PyTorch fallback: CPU tensors raise instead of silently computing something else."""
    from lc_b200.cov_mixed import Loss_cov_mixed
    from lc_b200.pnp import cer_solver
    from lc_b200.nll import pnp_auto
    from lc_b200.synth import make_correspondences
    c = make_correspondences(2, 8, 0).to(torch.float32)
    with pytest.raises(nat.NativeLibraryError):
        Loss_cov_mixed(c.K, c.pose, c.pts3d, c.pts2d, c.inv_std, None, bbox_3d=c.bbox_3d)
    with pytest.raises(nat.NativeLibraryError):
        cer_solver.solve(c.K, c.pts3d, c.pts2d, c.inv_std ** 2, c.start)
    with pytest.raises(nat.NativeLibraryError):
        pnp_auto.weighted_pnp_jac_wrt_pts2d(c.pts2d, c.pose, c.K, c.pts3d, c.inv_std ** 2)


def test_row_operators_refuse_cpu_tensors_and_validate_arguments():
    """Rows f1-f4: same rule (no fallback), and the C ABI rejects malformed argument structs before touching a device."""
    from lc_b200.dense import dense_pose_loss, dense_pose_loss_noc_bin
    from lc_b200.floatbits import nn_out_to_xyz, nn_noc2target
    from lc_b200.select import dense_point_select
    from lc_b200.pnp import init_solver
    from lc_b200.evaluate import compute_pose_errors
    from lc_b200.symmetry import select_pose_2d
    from lc_b200.synth import make_dense_outputs, make_zebra_outputs, make_correspondences
    d = make_dense_outputs(2, 8, 8, 0)
    z = make_zebra_outputs(2, 8, 8, 0, (3, 3, 3))
    c = make_correspondences(2, 8, 0).to(torch.float32)
    calls = [
        lambda: dense_pose_loss(d["xyz_noc"], d["logits"], d["scale"], d["noc_scale"], d["K"], d["pose"], d["bbox_3d"], top_left=(0, 0)),
        lambda: dense_pose_loss_noc_bin(z["bin_logits"], z["raw_bits"], z["logits"], z["scale"], z["msk_noc"], z["noc_scale"], z["K"], z["pose"],
                                        z["bbox_3d"], bit_cnt=3, top_left=(0, 0)),
        lambda: nn_out_to_xyz(z["bin_logits"], z["noc_scale"], bit_cnt=3),
        lambda: nn_noc2target(torch.zeros(1, 4, 4, 3), 3),
        lambda: dense_point_select(torch.zeros(1, 8, 8, 3), torch.zeros(1, 1, 8, 8), xyz_weights=torch.ones(1, 2, 8, 8)),
        lambda: init_solver.solve(c.K, c.pts3d, c.pts2d),
        lambda: compute_pose_errors(torch.eye(3)[None].double(), torch.zeros(1, 3).double(), torch.eye(3)[None].double(),
                                    torch.zeros(1, 3).double(), torch.zeros(5, 3).double()),
        lambda: select_pose_2d(c.K, c.pts3d, c.pts2d, torch.zeros(2, 3, 3, 4)),
    ]
    for fn in calls:
        with pytest.raises(nat.NativeLibraryError):
            fn()
    handle = nat.lib()
    for name, ty in (("lc_b200_dense_loss_fwd_bwd", nat.lc_dense_args), ("lc_b200_noc_bin_decode", nat.lc_decode_args),
                     ("lc_b200_noc_bin_encode", nat.lc_encode_args), ("lc_b200_dense_select", nat.lc_select_args),
                     ("lc_b200_pnp_init", nat.lc_init_args), ("lc_b200_pose_errors", nat.lc_eval_args),
                     ("lc_b200_select_pose", nat.lc_candi_args)):
        a = ty()
        assert getattr(handle, name)(None, None) < 0                      # NULL struct
        assert getattr(handle, name)(ctypes.byref(a), None) < 0           # abi_version 0
        assert b"abi_version" in handle.lc_b200_last_error()
        a.abi_version = nat.ABI_VERSION
        a.B = 1
        assert getattr(handle, name)(ctypes.byref(a), None) < 0           # sizes / required pointers missing
    assert len(_declared_symbols()) == len(nat.EXPORTS) + len(nat.EXTRA_EXPORTS) == 19


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "cpu_oracle" not in txt and "liblc_oracle" not in txt and "import oracle" not in txt, f


def test_flag_constants_match_the_header(tmp_path):
    """The Python mirror of the flag / weight-mode / status enums equals the header's values (compiled with the host gcc)."""
    c = tmp_path / "fl.c"
    c.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d %%d",'
                 'LC_FLAG_NAN_TO_NUM,LC_FLAG_TOL_NEEDS_SUCCESS,LC_FLAG_EXACT_HESSIAN,LC_FLAG_FORCE_STREAMING,LC_FLAG_LM_MIXED,LC_FLAG_COV_2D,'
                 'LC_W_ICOV_DIAG,LC_W_ICOV_FULL,LC_W_INV_STD,LC_W_SQRT_L,LC_ST_HESS_NOT_SPD,LC_ST_PRIOR_NOT_GOOD,LC_ST_COV_NOT_GOOD,'
                 'LC_B200_ABI_VERSION);return 0;}\n' % os.path.join(ROOT, "include", "lc_b200.h"))
    exe = tmp_path / "fl"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([gcc, "-o", str(exe), str(c)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [nat.FLAG_NAN_TO_NUM, nat.FLAG_TOL_NEEDS_SUCCESS, nat.FLAG_EXACT_HESSIAN, nat.FLAG_FORCE_STREAMING, nat.FLAG_LM_MIXED,
                   nat.FLAG_COV_2D, nat.W_ICOV_DIAG, nat.W_ICOV_FULL, nat.W_INV_STD, nat.W_SQRT_L, nat.ST_HESS_NOT_SPD,
                   nat.ST_PRIOR_NOT_GOOD, nat.ST_COV_NOT_GOOD, nat.ABI_VERSION]
