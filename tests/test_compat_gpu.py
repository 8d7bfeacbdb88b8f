"""GPU: the reference's own FFI symbol, `pnp_ceres_f32_omp` (lib/pnp/cxx/ext.h:1-14), exported by liblc_b200.so.

(1) driven through cffi with host pointer tables marshalled the way lib/pnp/pnp_ceres.py:74-140 does (one ragged job per
    entry), compared with the CPU oracle;
(2) when the reference files are staged under baseline/_ref (tools/stage_reference.py; git-ignored, travels with gpurun),
    the reference's UNMODIFIED lib/pnp/cer_solver.py + pnp_ceres.py run on top of it with only `_ext` swapped."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, quat_angle
from lc_b200.synth import make_correspondences, full_icov_from_inv_std

pytestmark = pytest.mark.gpu


def _marshal_and_call(ext, states, Ks, p2, p3, Ls, counts, max_iter=50, ftol=1e-6, threads=4):
    ffi, lib = ext.ffi, ext.lib
    n = len(states)
    st = np.stack(states).astype(np.float32)
    keep = [np.ascontiguousarray(a, np.float32) for a in Ks + p2 + p3 + Ls]       # keep the buffers alive
    tabs = [ffi.new(f"float*[{n}]") for _ in range(5)]
    for i in range(n):
        tabs[0][i] = ffi.cast("float*", st[i].ctypes.data)
        for t, group in zip(tabs[1:], (0, 1, 2, 3)):
            t[i] = ffi.cast("float*", keep[group * n + i].ctypes.data)
    tr = np.zeros(n, np.float32)
    inv = np.zeros(n, np.int32)
    cnt = np.asarray(counts, np.int32)
    rc = lib.pnp_ceres_f32_omp(tabs[0], tabs[1], tabs[2], tabs[3], tabs[4], ffi.cast("int*", cnt.ctypes.data), max_iter, ftol, 0,
                               ffi.cast("float*", tr.ctypes.data), ffi.cast("int*", inv.ctypes.data), n, threads)
    assert rc == 0
    return st, tr, inv


@pytest.mark.parametrize("full_L", [False, True], ids=["diag", "full2x2"])
def test_pnp_ceres_f32_omp_through_cffi_matches_oracle(oracle, full_L):
    from lc_b200.pnp import _ext
    B, N = 7, 300
    c = make_correspondences(B, N, 31).to(torch.float32)
    ns = [300, 120, 2, 77, 300, 3, 211]                     # job 2: < 3 points (ceres.cpp:84-91)
    if full_L:
        L = torch.linalg.cholesky_ex(full_icov_from_inv_std(c.inv_std, 31))[0]
    else:
        L = torch.diag_embed((c.inv_std ** 2).sqrt())
    st, tr, inv = _marshal_and_call(_ext, [s.numpy() for s in c.start], [k.numpy() for k in c.K],
                                    [c.pts2d[i, :n].numpy() for i, n in enumerate(ns)], [c.pts3d[i, :n].numpy() for i, n in enumerate(ns)],
                                    [L[i, :n].numpy() for i, n in enumerate(ns)], ns)
    ref = oracle.lm_solve(c.K, c.pts3d, c.pts2d, L, c.start, n_points=np.asarray(ns, np.int32))
    assert inv.tolist() == ref["invalid"].tolist() and inv[2] == 1 and tr[2] == 1.0
    assert np.array_equal(st[2], c.start[2].numpy())                      # invalid: state untouched (ceres.cpp:137-138)
    ok = inv == 0
    ang = quat_angle(st[ok, :4].astype(np.float64), ref["states"][ok, :4].astype(np.float64))
    trn = np.linalg.norm(st[ok, 4:].astype(np.float64) - ref["states"][ok, 4:], axis=1) / np.linalg.norm(ref["states"][ok, 4:], axis=1)
    assert ang.max() <= 1e-6 and trn.max() <= 1e-6
    assert np.allclose(tr[ok], ref["radius"][ok], rtol=1e-5)
    # max_iter exhausted -> NO_CONVERGENCE is invalid and the state stays (ceres.cpp:134-138)
    st1, _, inv1 = _marshal_and_call(_ext, [s.numpy() for s in c.start], [k.numpy() for k in c.K],
                                     [c.pts2d[i, :n].numpy() for i, n in enumerate(ns)], [c.pts3d[i, :n].numpy() for i, n in enumerate(ns)],
                                     [L[i, :n].numpy() for i, n in enumerate(ns)], ns, max_iter=1)
    assert inv1.all() and np.array_equal(st1, c.start.numpy())


def test_pnp_ceres_f32_omp_large_batch_takes_the_resident_kernel(oracle):
    from lc_b200.pnp import _ext
    from lc_b200 import _native as nat
    B, N = 64, 4096
    c = make_correspondences(B, N, 33).to(torch.float32)
    L = torch.diag_embed((c.inv_std ** 2).sqrt())
    st, tr, inv = _marshal_and_call(_ext, [s.numpy() for s in c.start], [k.numpy() for k in c.K], [x.numpy() for x in c.pts2d],
                                    [x.numpy() for x in c.pts3d], [x.numpy() for x in L], [N] * B, threads=8)
    ref = oracle.lm_solve(c.K, c.pts3d, c.pts2d, L, c.start)
    assert not inv.any() and not ref["invalid"].any()
    ang = quat_angle(st[:, :4].astype(np.float64), ref["states"][:, :4].astype(np.float64))
    assert ang.max() <= 1e-6
    assert np.allclose(tr, ref["radius"], rtol=1e-5)


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "lib", "pnp", "pnp_ceres.py")),
                    reason="reference files not staged (python tools/stage_reference.py in the build container)")
def test_unmodified_reference_cer_solver_runs_on_liblc_b200(oracle):
    """lib/pnp/cer_solver.py + lib/pnp/pnp_ceres.py from the reference, byte for byte, with `lib.pnp._ext` provided by
    lc_b200/pnp/_ext.py: the one-line swap of INTEGRATION.md §2b."""
    from lc_b200.pnp import _ext
    from lc_b200.pnp import cer_solver as ours
    sys.path.insert(0, REF_DIR)
    try:
        sys.modules["lib.pnp._ext"] = _ext
        ref_solver = importlib.import_module("lib.pnp.cer_solver")
        c = make_correspondences(6, 500, 35).to(torch.float32)
        ns = [500, 320, 2, 64, 500, 123]
        d = c.to(device="cuda")
        p3 = [d.pts3d[i, :n] for i, n in enumerate(ns)]
        p2 = [d.pts2d[i, :n] for i, n in enumerate(ns)]
        ic = [d.inv_std[i, :n] ** 2 for i, n in enumerate(ns)]
        stl = [s for s in d.start]
        inv_r, st_r = ref_solver.solve(d.K, p3, p2, ic, stl, num_workers=4, filter_input_nan=True)
        inv_o, st_o = ours.solve(d.K, p3, p2, ic, stl, num_workers=4, filter_input_nan=True)
        assert inv_r["invalids"].cpu().tolist() == inv_o["invalids"].cpu().tolist() == [False, False, True, False, False, False]
        assert st_r.dtype == torch.float32 and st_r.device == st_o.device
        assert torch.equal(st_r.cpu(), st_o.cpu())                       # same kernel underneath: bit-identical
        P3 = torch.zeros(6, 500, 3); P2 = torch.zeros(6, 500, 2); IC = torch.zeros(6, 500, 2)
        for i, n in enumerate(ns):
            P3[i, :n], P2[i, :n], IC[i, :n] = c.pts3d[i, :n], c.pts2d[i, :n], c.inv_std[i, :n] ** 2
        ref = oracle.lm_solve(c.K, P3, P2, torch.diag_embed(IC.sqrt()), c.start, n_points=np.asarray(ns, np.int32))
        ang = quat_angle(st_r.cpu().numpy()[:, :4].astype(np.float64), ref["states"][:, :4].astype(np.float64))
        assert ang.max() <= 1e-6
    finally:
        sys.path.remove(REF_DIR)
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
            del sys.modules[k]
