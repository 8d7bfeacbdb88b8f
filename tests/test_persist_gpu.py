"""GPU: the opt-in persistent two-context kernel (lc_persist.cu, LC_B200_PERSIST=1) against the CPU oracle and against the
default CTA-per-pose kernels: same poses, same iteration schedule, gradients / loss at the north-star tolerances.  Ragged
n_points and a batch that leaves some SMs with one context only are included."""
import numpy as np
import pytest
import torch

from conftest import quat_angle, rel_err
from lc_b200.synth import make_correspondences

pytestmark = pytest.mark.gpu


def _inputs(B, N, seed):
    c = make_correspondences(B, N, seed).to(torch.float32)
    d = dict(K=c.K, start=c.start, pose=c.pose, pts3d=c.pts3d.transpose(1, 2).contiguous(), pts2d=c.pts2d.transpose(1, 2).contiguous(),
             inv_std=c.inv_std.transpose(1, 2).contiguous(), bbox=c.bbox_3d)
    return c, {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("B", [150, 333])
def test_persistent_kernel_matches_oracle_and_default_kernels(oracle, monkeypatch, B):
    from lc_b200.fused import solve_and_loss
    from lc_b200.cov_mixed import loss_fwd_bwd
    from lc_b200.pnp.cer_solver import lm_solve
    from lc_b200 import _native as nat
    N = 2048
    c, d = _inputs(B, N, 77)
    p3, p2, s = d["pts3d"].transpose(1, 2), d["pts2d"].transpose(1, 2), d["inv_std"].transpose(1, 2)
    npts = torch.full((B,), N, dtype=torch.int32)
    npts[5], npts[17], npts[B - 1] = 1999, 2, 1234          # ragged; pose 17 has < 3 correspondences
    npts_d = npts.cuda()
    idx = np.r_[0:24, B - 8:B]
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("LC_B200_PERSIST", mode)
        f = solve_and_loss(d["K"], d["start"], p3, p2, s, None, d["bbox"], need=(True, True, True))
        assert (b"lc_persist_kernel<LM|LC>" in nat.lib().lc_b200_last_kernels()) == (mode == "1")
        l = loss_fwd_bwd(d["K"], d["pose"], p3, p2, s, None, d["bbox"], need=(True, False, True))
        assert (b"lc_persist_kernel<LC>" in nat.lib().lc_b200_last_kernels()) == (mode == "1")
        m = lm_solve(d["K"], p3, p2, s, d["start"], weight_mode=nat.W_INV_STD, n_points=npts_d)
        assert (b"lc_persist_kernel<LM>" in nat.lib().lc_b200_last_kernels()) == (mode == "1")
        torch.cuda.synchronize()
        runs[mode] = (f, l, m)
    (f0, l0, m0), (f1, l1, m1) = runs["0"], runs["1"]
    # the two kernel families agree (same LM code, different summation order in the loss passes)
    assert torch.equal(f0["iters"], f1["iters"]) and torch.equal(m0["iters"], m1["iters"]) and torch.equal(m0["invalid"], m1["invalid"])
    assert (f0["states"] - f1["states"]).abs().max() <= 2e-6 * f0["states"].abs().max()
    assert rel_err(l1["loss"].cpu().numpy()[:, None], l0["loss"].cpu().numpy()[:, None]) <= 2e-6
    assert rel_err(l1["g_pts3d"].cpu().numpy(), l0["g_pts3d"].cpu().numpy()) <= 2e-5
    assert m1["invalid"][17] == 1 and torch.equal(m1["states"][17], d["start"][17])
    # and the oracle
    ref = oracle.p3(c.K[idx], c.pts3d[idx], c.pts2d[idx], c.inv_std[idx], c.bbox_3d[idx], c.start[idx])
    assert np.array_equal(f1["iters"].cpu().numpy()[idx], ref["iters"]) and not ref["invalid"].any()
    ang = quat_angle(f1["states"].cpu().numpy()[idx][:, :4].astype(np.float64), ref["states"][:, :4].astype(np.float64))
    assert ang.max() <= 1e-6
    assert np.abs(f1["loss"].cpu().numpy()[idx] / ref["loss"] - 1).max() <= 1e-5
    for k in ("g_pts3d", "g_pts2d", "g_inv_std"):
        assert rel_err(f1[k].cpu().numpy()[idx], ref[k]) <= 1e-4, k
    L = torch.diag_embed((c.inv_std[idx] ** 2).sqrt())
    refm = oracle.lm_solve(c.K[idx], c.pts3d[idx], c.pts2d[idx], L, c.start[idx], n_points=npts.numpy()[idx])
    assert np.array_equal(m1["iters"].cpu().numpy()[idx], refm["iters"]) and np.array_equal(m1["invalid"].cpu().numpy()[idx], refm["invalid"])
