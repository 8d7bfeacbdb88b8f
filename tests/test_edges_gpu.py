"""GPU edge cases of the operator contracts: empty batches, sizes beyond the shared-memory resident limit (streaming
fallback), the largest size a reference config produces (zycbv test: 128x128, dense_sample 1 -> N = 16384), un-batched
and multi-dimensional leading shapes, non-contiguous K / pose, degenerate inputs."""
import numpy as np
import pytest
import torch

from conftest import quat_angle, rel_err
from lc_b200.synth import make_correspondences, planar_view

pytestmark = pytest.mark.gpu


def test_empty_batch_is_a_noop_everywhere():
    from lc_b200.cov_mixed import Loss_cov_mixed, loss_fwd_bwd
    from lc_b200.pnp import cer_solver
    from lc_b200.fused import solve_and_loss
    z = lambda *s: torch.zeros(*s, device="cuda")
    o = loss_fwd_bwd(z(0, 3, 3), z(0, 7), z(0, 16, 3), z(0, 16, 2), z(0, 16, 2), None, z(0, 8, 3))
    assert o["loss"].shape == (0,) and o["g_pts3d"].shape == (0, 16, 3)
    p3 = z(0, 16, 3).requires_grad_(True)
    l = Loss_cov_mixed(z(0, 3, 3), z(0, 7), p3, z(0, 16, 2), z(0, 16, 2), None, bbox_3d=z(0, 8, 3))
    assert l.shape == (0,)
    inv, st = cer_solver.solve(z(0, 3, 3), z(0, 16, 3), z(0, 16, 2), z(0, 16, 2), z(0, 7))
    assert st.shape == (0, 7) and inv["invalids"].shape == (0,)
    f = solve_and_loss(z(0, 3, 3), z(0, 7), z(0, 16, 3), z(0, 16, 2), z(0, 16, 2), None, z(0, 8, 3))
    assert f["loss"].shape == (0,) and f["launches"] == 0


@pytest.mark.parametrize("N", [12000, 16384])
def test_sizes_beyond_the_resident_limit_take_the_streaming_kernel(oracle, N):
    """N = 16384 is the largest size a reference config produces (configs/zycbv.yaml test: 128x128, dense_sample 1)."""
    from lc_b200.fused import solve_and_loss
    c = make_correspondences(2, N, 3).to(torch.float32)
    ref = oracle.p3(c.K, c.pts3d, c.pts2d, c.inv_std, c.bbox_3d, c.start)
    d = c.to(device="cuda")
    o = solve_and_loss(d.K, d.start, planar_view(d.pts3d), d.pts2d, planar_view(d.inv_std), None, d.bbox_3d, need=(True, True, True))
    assert np.array_equal(o["iters"].cpu().numpy(), ref["iters"]) and np.array_equal(o["invalid"].cpu().numpy(), ref["invalid"])
    st = o["states"].cpu().numpy().astype(np.float64)
    assert quat_angle(st[:, :4], ref["states"][:, :4].astype(np.float64)).max() <= 1e-6
    assert np.abs(o["loss"].cpu().numpy() - ref["loss"]).max() <= 1e-5 * np.abs(ref["loss"]).max()
    for k in ("g_pts3d", "g_pts2d", "g_inv_std"):
        assert rel_err(o[k].cpu().numpy(), ref[k]) <= 1e-4, k


def test_leading_shapes_and_noncontiguous_small_tensors(oracle):
    """Reference operators take arbitrary leading dims (*,N,3); K may be an expanded (stride-0) tensor, pose a slice."""
    from lc_b200.cov_mixed import Loss_cov_mixed
    c = make_correspondences(6, 96, 4).to(torch.float32)
    ref = oracle.lc_loss(c.K[:1].expand(6, 3, 3), c.pose, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d)
    d = c.to(device="cuda")
    pose_wide = torch.zeros(6, 9, device="cuda")
    pose_wide[:, 1:8] = d.pose
    loss = Loss_cov_mixed(d.K[:1].expand(2, 3, 3, 3), pose_wide[:, 1:8].reshape(2, 3, 7), d.pts3d.reshape(2, 3, 96, 3),
                          d.pts2d.reshape(2, 3, 96, 2), d.inv_std.reshape(2, 3, 96, 2), None, bbox_3d=d.bbox_3d.reshape(2, 3, 8, 3))
    assert loss.shape == (2, 3)
    assert np.abs(loss.reshape(-1).cpu().numpy() - ref["loss"]).max() <= 2e-6 * np.abs(ref["loss"]).max()


def test_degenerate_inputs_do_not_poison_the_batch():
    """A sample with all-zero weights (non-SPD Hessian -> identity), one with NaN correspondences in the solver (flagged
    invalid, start returned) and one with every point behind the z-clamp: the other samples are unaffected."""
    from lc_b200.cov_mixed import loss_fwd_bwd
    from lc_b200.pnp import cer_solver
    c = make_correspondences(4, 128, 8).to(torch.float32).to(device="cuda")
    base = loss_fwd_bwd(c.K, c.pose, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d)
    s = c.inv_std.clone(); s[1] = 0
    X = c.pts3d.clone(); X[2] = X[2] - 1e5 * torch.tensor([0, 0, 1.0], device="cuda")      # far behind the camera
    o = loss_fwd_bwd(c.K, c.pose, X, c.pts2d, s, None, c.bbox_3d)
    assert torch.isfinite(o["loss"][[0, 1, 3]]).all()
    assert torch.equal(o["loss"][[0, 3]], base["loss"][[0, 3]]) and torch.equal(o["g_pts3d"][[0, 3]], base["g_pts3d"][[0, 3]])
    x = c.pts2d.clone(); x[0, :, :] = float("nan")
    inv, st = cer_solver.solve(c.K, c.pts3d, x, c.inv_std ** 2, c.start)
    assert inv["invalids"].cpu().tolist() == [True, False, False, False]
    assert torch.equal(st[0], c.start[0]) and torch.isfinite(st[1:]).all()


def test_fp64_tensors_through_the_python_operators(oracle):
    from lc_b200.cov_mixed import Loss_cov_mixed
    from lc_b200.pnp import cer_solver
    c = make_correspondences(3, 300, 12)
    ref = oracle.lc_loss(c.K, c.pose, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d)
    d = c.to(device="cuda")
    p3 = d.pts3d.clone().requires_grad_(True)
    loss = Loss_cov_mixed(d.K, d.pose, p3, d.pts2d, d.inv_std, None, bbox_3d=d.bbox_3d)
    assert loss.dtype == torch.float64
    loss.sum().backward()
    assert np.abs(loss.detach().cpu().numpy() - ref["loss"]).max() <= 1e-9 * np.abs(ref["loss"]).max()
    assert rel_err(p3.grad.cpu().numpy(), ref["g_pts3d"]) <= 1e-9
    inv, st = cer_solver.solve(d.K, d.pts3d, d.pts2d, d.inv_std ** 2, d.start)
    assert st.dtype == torch.float64 and not inv["invalids"].any()


def test_ragged_batch_padded_beyond_the_resident_limit_is_split_by_n_points(oracle):
    """The test-time chain pads to the full map (N = 16384 at 128x128) while n_points is a few thousand: poses that fit take
    the shared-memory resident kernel, the rest the streaming kernel, two launches, no host sync.  Every pose must equal the
    CPU oracle run on its own ragged point set, whichever kernel took it."""
    from lc_b200.pnp.cer_solver import lm_solve
    from lc_b200.fused import solve_and_loss
    from lc_b200 import _native as nat
    N = 12000
    npts = [100, 3000, 4900, 5100, 9000, 12000]
    B = len(npts)
    c = make_correspondences(B, N, 5).to(torch.float32)
    n_t = torch.tensor(npts, dtype=torch.int32)
    d = c.to(device="cuda")
    o = lm_solve(d.K, d.pts3d, d.pts2d, d.inv_std ** 2, d.start, n_t.cuda(), weight_mode=nat.W_ICOV_DIAG)
    s = lm_solve(d.K, d.pts3d, d.pts2d, d.inv_std ** 2, d.start, n_t.cuda(), weight_mode=nat.W_ICOV_DIAG, force_streaming=True)
    assert nat.lib().lc_b200_last_launch_count() == 1
    st, ss = o["states"].cpu().numpy().astype(np.float64), s["states"].cpu().numpy().astype(np.float64)
    assert quat_angle(st[:, :4], ss[:, :4]).max() <= 1e-7 and np.array_equal(o["iters"].cpu().numpy(), s["iters"].cpu().numpy())
    for b, n in enumerate(npts):
        L = torch.diag_embed(c.inv_std[b:b + 1, :n]).numpy()
        ref = oracle.lm_solve(c.K[b:b + 1].numpy(), c.pts3d[b:b + 1, :n].numpy(), c.pts2d[b:b + 1, :n].numpy(), L, c.start[b:b + 1].numpy())
        assert quat_angle(st[b:b + 1, :4], ref["states"][:, :4].astype(np.float64)).max() <= 1e-6, n
        assert int(o["iters"][b]) == int(ref["iters"][0]) and int(o["invalid"][b]) == int(ref["invalid"][0])
    f = solve_and_loss(d.K, d.start, planar_view(d.pts3d), d.pts2d, planar_view(d.inv_std), None, d.bbox_3d, need=(True, False, True))
    assert f["launches"] == 1                                   # no n_points: one streaming launch
    args = dict(K=d.K, pose=d.start, pts3d=d.pts3d, pts2d=d.pts2d, weights=d.inv_std ** 2, n_points=n_t.cuda(),
                state=torch.empty(B, 7, device="cuda"), invalid=torch.empty(B, dtype=torch.int32, device="cuda"),
                weight_mode=nat.W_ICOV_DIAG, flags=nat.FLAG_TOL_NEEDS_SUCCESS)
    assert nat.call("lc_b200_lm_solve", nat.make_args(B, N, torch.float32, **args), d.K.device) == 2   # resident + streaming


def test_loss_kernel_with_ragged_n_points_in_the_tensor_memory_range():
    """N = 4096 takes the tensor-memory variant of the loss kernel (model points in TMEM, warp-uniform point loops with a
    `live` predicate); ragged n_points must give what each pose gives alone on its trimmed tensors (shared-memory variant for
    n <= 2048, streaming kernel for tiny n), and LC_B200_TMEM=0 must agree with the default."""
    import os
    from lc_b200 import _native as nat
    from lc_b200.cov_mixed import loss_fwd_bwd
    N = 4096
    npts = [4096, 100, 3000, 2049, 33, 4095]
    B = len(npts)
    c = make_correspondences(B, N, 9).to(torch.float32).to(device="cuda")

    def run(planar=False):
        loss = torch.empty(B, device="cuda")
        # outputs start as NaN: the kernels must define every slot, including the padding beyond n_points (zeros)
        g3, gs = torch.full((B, N, 3), float("nan"), device="cuda"), torch.full((B, N, 2), float("nan"), device="cuda")
        X, x, w = c.pts3d, c.pts2d, c.inv_std
        if planar:   # the layout of the dense call site: takes the vectorised kernels
            X, x, w, g3, gs = planar_view(X), planar_view(x), planar_view(w), planar_view(g3), planar_view(gs)
        a = nat.make_args(B, N, torch.float32, K=c.K, pose=c.pose, pts3d=X, pts2d=x, weights=w, bbox=c.bbox_3d,
                          n_points=torch.tensor(npts, dtype=torch.int32, device="cuda"), loss=loss, g_pts3d=g3, g_weights=gs)
        nat.call("lc_b200_loss_fwd_bwd", a, c.K.device)
        torch.cuda.synchronize()
        return loss, g3, gs

    loss, g3, gs = run()
    for b, n in enumerate(npts):
        o = loss_fwd_bwd(c.K[b:b + 1], c.pose[b:b + 1], c.pts3d[b:b + 1, :n].contiguous(), c.pts2d[b:b + 1, :n].contiguous(),
                         c.inv_std[b:b + 1, :n].contiguous(), None, c.bbox_3d[b:b + 1])
        assert abs(loss[b].item() - o["loss"][0].item()) <= 2e-6 * max(1.0, abs(o["loss"][0].item())), n
        assert rel_err(g3[b:b + 1, :n].cpu().numpy(), o["g_pts3d"].cpu().numpy()) <= 2e-5, n
        assert rel_err(gs[b:b + 1, :n].cpu().numpy(), o["g_inv_std"].cpu().numpy()) <= 2e-5, n
        assert (g3[b, n:] == 0).all() and (gs[b, n:] == 0).all()          # slots beyond n_points are written as zeros
    os.environ["LC_B200_TMEM"] = "0"
    try:
        loss0, g30, gs0 = run()
    finally:
        del os.environ["LC_B200_TMEM"]
    assert torch.allclose(loss, loss0, rtol=2e-6) and rel_err(g3.cpu().numpy(), g30.cpu().numpy()) <= 2e-5
    lossv, g3v, gsv = run(planar=True)
    assert b"vec4" in nat.lib().lc_b200_last_kernels()
    assert torch.allclose(loss, lossv, rtol=2e-6) and rel_err(g3v.cpu().numpy(), g3.cpu().numpy()) <= 2e-5
    assert rel_err(gsv.cpu().numpy(), gs.cpu().numpy()) <= 2e-5 and not torch.isnan(g3v).any() and not torch.isnan(gsv).any()
