"""GPU: poses split over a two-CTA thread-block cluster (lc_resident_kernel.cuh, CL = 2; lc_resident.cu: split_tail).

The cluster variant runs the same point loops on half of the points per CTA and completes every reduction through distributed
shared memory, so its results must agree with the one-CTA-per-pose kernel to reduction-order rounding, and with the CPU oracle
at the north-star tolerances.  Covered: batches that run entirely as clusters (B <= half a wave), a batch whose last wave is
split off (two launches), ragged n_points (one CTA of the pair may own no live point), odd group counts (N / 4 odd)."""
import os

import numpy as np
import pytest
import torch

from conftest import quat_angle, rel_err
from lc_b200.synth import make_correspondences, planar_view

pytestmark = pytest.mark.gpu


def _kernels():
    from lc_b200 import _native as nat
    return nat.lib().lc_b200_last_kernels().decode()


def _run(pipeline, d, n_points=None):
    from lc_b200.fused import solve_and_loss
    from lc_b200.cov_mixed import loss_fwd_bwd
    from lc_b200.pnp.cer_solver import lm_solve
    from lc_b200 import _native as nat
    X, x, w = planar_view(d.pts3d), planar_view(d.pts2d), planar_view(d.inv_std)
    if pipeline == "p3":
        o = solve_and_loss(d.K, d.start, X, x, w, None, d.bbox_3d, need=(True, True, True), n_points=n_points)
    elif pipeline == "p1":
        o = loss_fwd_bwd(d.K, d.pose, X, x, w, None, d.bbox_3d, need=(True, True, True), n_points=n_points)
    else:
        o = lm_solve(d.K, X, x, w, d.start, weight_mode=nat.W_INV_STD, n_points=n_points)
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in o.items() if torch.is_tensor(v)}, _kernels()


def _with_split(flag, fn):
    old = os.environ.get("LC_B200_SPLIT")
    os.environ["LC_B200_SPLIT"] = flag
    try:
        return fn()
    finally:
        if old is None:
            del os.environ["LC_B200_SPLIT"]
        else:
            os.environ["LC_B200_SPLIT"] = old


def _close(a, b, pipeline):
    if pipeline != "p1":
        assert torch.equal(a["iters"], b["iters"]) and torch.equal(a["invalid"], b["invalid"])
        assert (a["states"].double() - b["states"].double()).abs().max() <= 2e-6 * b["states"].double().abs().max()
    if pipeline != "p2":
        assert (a["loss"].double() / b["loss"].double() - 1).abs().max() <= 3e-5   # a one-ulp pose difference moves the loss by ~2e-5
        for k in ("g_pts3d", "g_pts2d", "g_inv_std"):
            assert rel_err(a[k].cpu().numpy(), b[k].cpu().numpy()) <= 1e-4, k


@pytest.mark.parametrize("pipeline", ["p1", "p2", "p3"])
@pytest.mark.parametrize("B,N,seed", [(5, 700, 1), (3, 4096, 2), (7, 1028, 3)])
def test_cluster_split_equals_one_cta_per_pose(pipeline, B, N, seed):
    d = make_correspondences(B, N, seed).to(torch.float32).to(device="cuda")
    one, k1 = _with_split("0", lambda: _run(pipeline, d))
    two, k2 = _with_split("1", lambda: _run(pipeline, d))
    assert "cluster2" in k2 and "cluster" not in k1, (k1, k2)
    _close(two, one, pipeline)
    if pipeline == "p1":   # same pose in both runs: the loss agrees to fp32 rounding of the reduction order
        assert (two["loss"].double() / one["loss"].double() - 1).abs().max() <= 2e-6


def test_cluster_split_matches_the_oracle(oracle):
    B, N = 6, 2048
    c = make_correspondences(B, N, 11).to(torch.float32)
    ref = oracle.p3(c.K, c.pts3d, c.pts2d, c.inv_std, c.bbox_3d, c.start)
    o, k = _with_split("1", lambda: _run("p3", c.to(device="cuda")))
    assert "cluster2" in k
    assert np.array_equal(o["invalid"].cpu().numpy(), ref["invalid"]) and np.array_equal(o["iters"].cpu().numpy(), ref["iters"])
    st = o["states"].cpu().numpy().astype(np.float64)
    assert quat_angle(st[:, :4], ref["states"][:, :4].astype(np.float64)).max() <= 1e-6
    assert (np.linalg.norm(st[:, 4:] - ref["states"][:, 4:], axis=1) / np.linalg.norm(ref["states"][:, 4:], axis=1)).max() <= 1e-6
    assert np.abs(o["loss"].cpu().numpy() - ref["loss"]).max() <= 1e-5 * np.abs(ref["loss"]).max()
    for key in ("g_pts3d", "g_pts2d", "g_inv_std"):
        assert rel_err(o[key].cpu().numpy(), ref[key]) <= 1e-4, key


@pytest.mark.parametrize("pipeline", ["p1", "p3"])
def test_cluster_split_with_ragged_n_points(pipeline):
    """n_points below half of N: the second CTA of the pair owns no live point; gradients of the padding are zero."""
    B, N = 4, 1024
    d = make_correspondences(B, N, 5).to(torch.float32).to(device="cuda")
    npts = torch.tensor([1024, 300, 515, 700], dtype=torch.int32, device="cuda")
    one, k1 = _with_split("0", lambda: _run(pipeline, d, npts))
    two, k2 = _with_split("1", lambda: _run(pipeline, d, npts))
    assert "cluster2" in k2 and "cluster" not in k1
    _close(two, one, pipeline)
    for b in range(B):
        assert two["g_pts3d"][b, npts[b]:].abs().max().item() == 0 if npts[b] < N else True
        assert two["g_inv_std"][b, npts[b]:].abs().max().item() == 0 if npts[b] < N else True


def test_last_wave_runs_as_clusters():
    """B = one wave + a small tail: two launches (clusters first, full waves behind them), every pose equal to the unsplit run."""
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    B, N = 4 * sms + 9, 512   # N <= 2048: four 128-thread CTAs per SM
    d = make_correspondences(B, N, 8).to(torch.float32).to(device="cuda")
    one, k1 = _with_split("0", lambda: _run("p3", d))
    two, k2 = _with_split("1", lambda: _run("p3", d))
    assert k2.count("lc_resident_kernel") == 2 and "cluster2" in k2 and "cluster" not in k1, (k1, k2)
    _close(two, one, "p3")
    # the poses of the full waves take the same kernel in both runs
    assert torch.equal(two["loss"][: B - 9], one["loss"][: B - 9]) and torch.equal(two["g_pts3d"][: B - 9], one["g_pts3d"][: B - 9])
