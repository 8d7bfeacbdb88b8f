"""GPU parity of the test-time point selection vs fixtures produced by the reference's own functions
(test.py:36-45, 67-106; losses.py:142-161) and vs the CPU oracle; plus the selection -> LM chain without a host sync."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, quat_angle

pytestmark = pytest.mark.gpu

MODES = ["mask", "quantile", "quantile_in_mask"]


def _c(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x)).to(device="cuda", dtype=dtype)


def _unpack(sel, b):
    n = int(sel["n_points"][b])
    return n, sel["pts3d"][b, :n].cpu().numpy(), sel["pts2d"][b, :n].cpu().numpy(), sel["inv_cov"][b, :n].cpu().numpy()


@pytest.mark.parametrize("name", ["select_b3_32x32_s1.npz", "select_b2_64x48_s2.npz"])
@pytest.mark.parametrize("mode", MODES)
def test_selection_is_bit_exact_given_the_reference_weight_map(name, mode):
    """With the reference's own softmax*scale map as input, the selected set, its order and the gathered values are
    bit-identical to nonzero() / indexing in the reference (fp32 order statistics + torch's lerp)."""
    from lc_b200.select import dense_point_select
    z = np.load(os.path.join(GOLDEN_DIR, name))
    xyz = _c(z["in_xyz_noc"]).permute(0, 2, 3, 1)
    sel = dense_point_select(xyz, _c(z["in_msk_logits"]), xyz_weights=_c(z["ref_weights"]), noc_scale=_c(z["in_noc_scale"]),
                             sample=int(z["sample"]), dense_point_select=mode, want_index=True)
    valid = z["valid_" + mode]
    for b in range(valid.shape[0]):
        idx = np.nonzero(valid[b])[0]
        n, p3, p2, ic = _unpack(sel, b)
        assert n == len(idx)
        assert np.array_equal(sel["index"][b, :n].cpu().numpy(), idx)
        assert np.array_equal(p3, z["ref_pts3d"][b][idx]) and np.array_equal(p2, z["ref_pts2d"][b][idx])
        assert np.array_equal(ic, z["ref_inv_cov"][b][idx])
        assert (sel["pts3d"][b, n:] == 0).all() and (sel["inv_cov"][b, n:] == 0).all() and (sel["index"][b, n:] == -1).all()


@pytest.mark.parametrize("name", ["select_b3_32x32_s1.npz", "select_b2_64x48_s2.npz"])
@pytest.mark.parametrize("mode", MODES)
def test_fused_softmax_selection_matches_reference(name, mode):
    """Softmax fused in (logits + scale as inputs): the weight map differs from torch's by rounding (~1e-7), so points whose
    quantile operand sits within 1e-5 relative of the threshold may fall on either side; everything else must agree."""
    from lc_b200.select import dense_point_select
    z = np.load(os.path.join(GOLDEN_DIR, name))
    sel = dense_point_select(_c(z["in_xyz_noc"]).permute(0, 2, 3, 1), _c(z["in_msk_logits"]), xyz_weight_logits=_c(z["in_logits"]),
                             xyz_weights_scale=_c(z["in_scale"]), noc_scale=_c(z["in_noc_scale"]), sample=int(z["sample"]),
                             dense_point_select=mode, want_index=True)
    valid = z["valid_" + mode]
    s = int(z["sample"])
    wsum = z["ref_weights"][:, :, ::s, ::s].reshape(valid.shape[0], 2, -1).sum(1)
    for b in range(valid.shape[0]):
        n = int(sel["n_points"][b])
        got = np.zeros(valid.shape[1], bool)
        got[sel["index"][b, :n].cpu().numpy()] = True
        diff = np.nonzero(got != valid[b])[0]
        if mode != "mask" and len(diff):
            thr = wsum[b][valid[b]].min()
            assert np.all(np.abs(wsum[b][diff] - thr) <= 1e-5 * thr), (len(diff), wsum[b][diff], thr)
        else:
            assert len(diff) == 0
        idx = sel["index"][b, :n].cpu().numpy()
        assert np.all(np.diff(idx) > 0)                                        # nonzero() order
        ic = sel["inv_cov"][b, :n].cpu().numpy()
        assert np.allclose(ic, z["ref_inv_cov"][b][idx], rtol=2e-5, atol=0)


@pytest.mark.parametrize("B,H,W,sample,mode", [(2, 128, 128, 1, "quantile_in_mask"), (3, 40, 24, 3, "quantile"), (2, 16, 16, 2, "mask")])
def test_selection_matches_oracle(oracle, B, H, W, sample, mode):
    from lc_b200.select import dense_point_select
    from lc_b200.synth import make_dense_outputs
    d = make_dense_outputs(B, H, W, 300 + H)
    g = torch.Generator().manual_seed(H)
    ml = 2.0 * torch.randn(B, 1, H, W, generator=g) + 0.5
    lg = d["logits"]
    weights = (lg.reshape(B, 1, -1).softmax(-1).reshape_as(lg) * d["scale"])
    xyz = d["xyz_noc"].permute(0, 2, 3, 1) * d["noc_scale"][:, None, None, :]
    ref = oracle.dense_point_select(xyz.numpy(), weights.numpy(), ml.numpy(), sample, mode)
    sel = dense_point_select(xyz.cuda(), ml.cuda(), xyz_weights=weights.cuda(), sample=sample, dense_point_select=mode, want_index=True)
    for b in range(B):
        idx = np.nonzero(ref["valid"][b])[0]
        n, p3, p2, ic = _unpack(sel, b)
        assert n == len(idx) and np.array_equal(sel["index"][b, :n].cpu().numpy(), idx)
        assert np.array_equal(p3, ref["pts3d"][b][idx]) and np.array_equal(p2, ref["pts2d"][idx]) and np.array_equal(ic, ref["inv_cov"][b][idx])


def test_degenerate_selections_are_padded_to_min_points():
    """No pixel in the mask: the reference pads the empty index list with 4 random indices (test.py:108-113)."""
    from lc_b200.select import dense_point_select
    B, H, W = 2, 16, 16
    xyz = torch.randn(B, H, W, 3, device="cuda")
    ml = torch.full((B, 1, H, W), -5.0, device="cuda")
    ml[1, 0, 3, 4] = 5.0
    w = torch.rand(B, 2, H, W, device="cuda")
    sel = dense_point_select(xyz, ml, xyz_weights=w, sample=1, dense_point_select="mask", want_index=True)
    assert sel["n_points"].tolist() == [4, 4]
    idx = sel["index"].cpu().numpy()
    assert idx[1, 0] == 3 * W + 4 and (idx[:, :4] >= 0).all() and (idx[:, :4] < H * W).all() and (idx[:, 4:] == -1).all()


def test_selection_feeds_the_solver_without_a_host_sync(oracle):
    """select -> cer_solver.solve(n_points=...): only in-mask points are used, and the solved poses equal the CPU LM oracle
    run per sample on the selected (ragged) point sets, as the reference does after nonzero() (test.py:106-127)."""
    from lc_b200.select import solve_pnp_dense
    from lc_b200.synth import make_dense_outputs
    B, H, W = 4, 64, 64
    d = {k: v.cuda() for k, v in make_dense_outputs(B, H, W, 5).items()}
    ml = torch.full((B, 1, H, W), 3.0, device="cuda")
    ml[:, :, :8] = -3.0
    aa = 0.02 * torch.randn(B, 3, generator=torch.Generator().manual_seed(1)).cuda()
    ang = aa.norm(dim=-1, keepdim=True)
    dq = torch.cat((torch.cos(ang / 2), aa / ang * torch.sin(ang / 2)), -1)
    q = d["pose"][:, :4]
    w1, x1, y1, z1 = q.unbind(-1); w2, x2, y2, z2 = dq.unbind(-1)
    q0 = torch.stack((w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                      w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2), -1)
    start = torch.cat((q0, d["pose"][:, 4:] * 1.01), -1)
    res, sel = solve_pnp_dense(d["K"], d["xyz_noc"].permute(0, 2, 3, 1), ml, d["logits"], d["scale"], start,
                               noc_scale=d["noc_scale"], sample=2, dense_point_select="quantile_in_mask")
    states = res["weighted"]
    assert (sel["pts2d"][:, :, 1][sel["pts2d"][:, :, 1] > 0].min() >= 8)          # rows 0..7 are outside the mask
    st = states.cpu().numpy().astype(np.float64)
    assert quat_angle(st[:, :4], d["pose"][:, :4].cpu().numpy().astype(np.float64)).max() < 0.01   # start was 0.02 rad away
    for b in range(B):
        n = int(sel["n_points"][b])
        L = torch.diag_embed(sel["inv_cov"][b:b + 1, :n].sqrt()).cpu().numpy()
        o = oracle.lm_solve(d["K"][b:b + 1].cpu().numpy(), sel["pts3d"][b:b + 1, :n].cpu().numpy(), sel["pts2d"][b:b + 1, :n].cpu().numpy(),
                            L, start[b:b + 1].cpu().numpy())
        assert quat_angle(st[b:b + 1, :4], o["states"][:, :4].astype(np.float64)).max() <= 1e-6
        assert np.abs(st[b, 4:] - o["states"][0, 4:]).max() <= 1e-6 * np.abs(o["states"][0, 4:]).max()


def test_full_test_time_chain_without_a_start_pose(oracle):
    """selection -> device initialiser -> 'weighted' and 'weighted_filtered' LM solves (test.py:84-134), all on the device.
    The filtered solve (weights of non-inliers zeroed) equals the CPU LM oracle run on the compacted inlier list."""
    from lc_b200.select import solve_pnp_dense
    from lc_b200.synth import make_dense_outputs
    B, H, W = 3, 64, 64
    d = {k: v.cuda() for k, v in make_dense_outputs(B, H, W, 9).items()}
    ml = torch.full((B, 1, H, W), 3.0, device="cuda")
    res, sel = solve_pnp_dense(d["K"], d["xyz_noc"].permute(0, 2, 3, 1), ml, d["logits"], d["scale"], None, noc_scale=d["noc_scale"],
                               sample=2, dense_point_select="quantile_in_mask", solvers=("weighted", "weighted_filtered"))
    assert not sel["init_invalid"].any()
    truth = d["pose"][:, :4].cpu().numpy().astype(np.float64)
    for name in ("weighted", "weighted-filtered"):
        assert quat_angle(res[name][:, :4].cpu().numpy().astype(np.float64), truth).max() < 0.02
    st = res["weighted-filtered"].cpu().numpy().astype(np.float64)
    for b in range(B):
        n = int(sel["n_points"][b])
        keep = sel["inliers"]["mask"][b, :n].cpu().numpy()
        assert keep.sum() >= 0.5 * n
        L = torch.diag_embed(sel["inv_cov"][b:b + 1, :n].sqrt()).cpu().numpy()[:, keep]
        o = oracle.lm_solve(d["K"][b:b + 1].cpu().numpy(), sel["pts3d"][b:b + 1, :n].cpu().numpy()[:, keep],
                            sel["pts2d"][b:b + 1, :n].cpu().numpy()[:, keep], L, sel["start"][b:b + 1].cpu().numpy())
        assert quat_angle(st[b:b + 1, :4], o["states"][:, :4].astype(np.float64)).max() <= 1e-6
        assert np.abs(st[b, 4:] - o["states"][0, 4:]).max() <= 1e-6 * np.abs(o["states"][0, 4:]).max()
