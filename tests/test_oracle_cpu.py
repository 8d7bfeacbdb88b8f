"""CPU: the oracle against the reference's golden vectors and against independent checks."""
import numpy as np
import pytest
import torch

import os

from conftest import GOLDEN_DIR, golden_files, load_golden, rel_err, quat_angle
from lc_b200.synth import make_correspondences, quat_to_matrix


@pytest.mark.parametrize("name", golden_files())
def test_lc_oracle_matches_reference_golden(oracle, name):
    """oracle/lc_oracle.c == unmodified reference (fp64) on the committed fixtures."""
    g = load_golden(name)
    o = oracle.lc_loss(g["in_K"], g["in_pose"], g["in_pts3d"], g["in_pts2d"], g["in_inv_std"], g["valid"],
                       g["in_bbox_3d"], *g["params"], want_jac="ref_jac" in g)
    tol = 1e-9 if "init" in name else 1e-11   # the random-init regime has cond(H) ~ 1e8
    assert np.abs(o["loss"] - g["ref_loss"]).max() <= tol * np.abs(g["ref_loss"]).max()
    assert rel_err(o["g_pts3d"], g["ref_g_pts3d"]) <= tol
    assert rel_err(o["g_pts2d"], g["ref_g_pts2d"]) <= tol
    assert rel_err(o["g_inv_std"], g["ref_g_inv_std"]) <= tol
    assert rel_err(o["cov"], g["ref_cov"]) <= tol
    if "ref_jac" in g:
        assert rel_err(o["jac"], g["ref_jac"]) <= tol
        assert rel_err(o["W"], g["ref_W"]) <= 1e-13
        assert rel_err(o["sigma"], g["ref_sigma"]) <= 1e-13
    assert (o["flags"] == 0).all()


def _lm_problem(B, N, seed):
    c = make_correspondences(B, N, seed).to(torch.float32)
    L = torch.diag_embed(c.inv_std)
    return c, L.numpy()


def test_lm_oracle_jacobian_vs_finite_differences(oracle):
    c, L = _lm_problem(2, 64, 0)
    K, X, x = c.K.numpy()[0], c.pts3d.numpy()[0], c.pts2d.numpy()[0]
    for x6 in (np.array([0.3, -0.2, 0.5, 10.0, -20.0, 900.0]), np.array([1e-9, 0.0, 0.0, 1.0, 2.0, 800.0]),
               np.array([2.0, 1.5, -1.0, -30.0, 5.0, 600.0])):
        e = oracle.lm_eval(x6, K, X, x, L[0])
        Jfd = np.zeros_like(e["J"])
        for k in range(6):
            h = 1e-6 * max(1.0, abs(x6[k]))
            xp, xm = x6.copy(), x6.copy()
            xp[k] += h
            xm[k] -= h
            Jfd[:, k] = (oracle.lm_eval(xp, K, X, x, L[0])["r"] - oracle.lm_eval(xm, K, X, x, L[0])["r"]) / (2 * h)
        assert np.abs(Jfd - e["J"]).max() <= 1e-6 * np.abs(e["J"]).max()
        assert np.allclose(e["J"].T @ e["r"], e["g"], rtol=1e-12, atol=1e-9)


def test_lm_oracle_noise_free_known_answer(oracle):
    c = make_correspondences(6, 50, 3)
    R = quat_to_matrix(c.pose[:, :4])
    P = c.pts3d @ R.mT + c.pose[:, None, 4:]
    KP = P @ c.K.mT
    c.pts2d = KP[..., :2] / KP[..., 2:]
    c32 = c.to(torch.float32)
    o = oracle.lm_solve(c32.K.numpy(), c32.pts3d.numpy(), c32.pts2d.numpy(), torch.diag_embed(c32.inv_std).numpy(),
                        c32.start.numpy())
    assert (o["invalid"] == 0).all()
    assert quat_angle(o["states"][:, :4], c32.pose.numpy()[:, :4]).max() < 5e-5   # fp32 inputs
    t_ref = c32.pose.numpy()[:, 4:]
    assert (np.linalg.norm(o["states"][:, 4:] - t_ref, axis=1) / np.linalg.norm(t_ref, axis=1)).max() < 5e-5


@pytest.mark.parametrize("B,N", [(6, 8), (4, 200), (2, 1024)])
def test_lm_oracle_close_to_true_optimum(oracle, B, N):
    """The early-stopped Ceres-style answer sits within the documented gap of the fully converged optimum
    (scipy LM run to machine precision from the oracle's answer), and is a descent from the start."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation as Rot
    c, L = _lm_problem(B, N, 1)
    K, X, x, st = c.K.numpy(), c.pts3d.numpy(), c.pts2d.numpy(), c.start.numpy()
    o = oracle.lm_solve(K, X, x, L, st, want_trace=True)
    assert (o["term"] == 0).all()
    for b in range(B):
        tr = o["trace"][b]
        tr = tr[~np.isnan(tr[:, 0])]          # finalised iterations only (the converging one is not)
        assert np.all(np.diff(tr[tr[:, 2] > 0, 0]) < 0)          # accepted costs strictly decrease
        f = lambda p: oracle.lm_eval(p, K[b], X[b], x[b], L[b])["r"]
        jf = lambda p: oracle.lm_eval(p, K[b], X[b], x[b], L[b])["J"]
        s = least_squares(f, o["x6"][b], jac=jf, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
        ang = (Rot.from_rotvec(o["x6"][b][:3]).inv() * Rot.from_rotvec(s.x[:3])).magnitude()
        assert ang < (2e-3 if N <= 8 else 2e-4)
        assert 0.5 * (s.fun ** 2).sum() <= tr[-1, 0] * (1 + 1e-12)
        assert (tr[-1, 0] - 0.5 * (s.fun ** 2).sum()) <= 1e-4 * tr[-1, 0]


def test_lm_oracle_wrapper_semantics(oracle):
    """ceres.cpp:84-91 (ptCnt<3 -> invalid, tr=1, state untouched) and :134-138 (no write-back when not converged)."""
    c, L = _lm_problem(3, 16, 5)
    K, X, x, st = c.K.numpy(), c.pts3d.numpy(), c.pts2d.numpy(), c.start.numpy()
    o = oracle.lm_solve(K, X, x, L, st, n_points=np.array([2, 16, 16], np.int32))
    assert o["invalid"][0] == 1 and o["radius"][0] == 1.0 and np.array_equal(o["states"][0], st[0])
    assert o["invalid"][1] == 0
    o1 = oracle.lm_solve(K, X, x, L, st, max_iter=1)
    assert (o1["invalid"] == 1).all() and np.array_equal(o1["states"], st) and (o1["term"] == 1).all()


def test_p3_driver_equals_separate_calls(oracle):
    c, L = _lm_problem(3, 128, 7)
    a = oracle.p3(c.K.numpy(), c.pts3d.numpy(), c.pts2d.numpy(), c.inv_std.numpy(), c.bbox_3d.numpy(), c.start.numpy())
    lm = oracle.lm_solve(c.K.numpy(), c.pts3d.numpy(), c.pts2d.numpy(), L, c.start.numpy())
    assert np.array_equal(a["states"], lm["states"])
    lc = oracle.lc_loss(c.K.numpy(), lm["states"], c.pts3d.numpy(), c.pts2d.numpy(), c.inv_std.numpy(), None, c.bbox_3d.numpy())
    assert np.allclose(a["loss"], lc["loss"], rtol=1e-13)
    assert rel_err(a["g_pts3d"], lc["g_pts3d"]) < 1e-6   # p3 stores fp32 gradients


@pytest.mark.parametrize("name", ["densex_b2_16x16_s2.npz", "densex_b2_64x64_s2.npz", "densex_b2_40x56_s3.npz"])
def test_dense_producer_oracle_matches_reference_golden(oracle, name):
    """oracle.dense_pose_loss == the reference's dense_pose_loss glue + Loss_cov_mixed + autograd (fp64)."""
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, name))
    o = oracle.dense_pose_loss(z["in_xyz_noc"], z["in_logits"], z["in_scale"], z["in_noc_scale"], z["in_K"], z["in_pose"],
                               z["in_bbox_3d"], int(z["sample"]), tuple(z["top_left"]))
    assert np.abs(o["loss"] - z["ref_loss"]).max() <= 1e-11 * np.abs(z["ref_loss"]).max()
    assert rel_err(o["g_xyz_noc"], z["ref_g_xyz_noc"]) <= 1e-10
    assert rel_err(o["g_logits"], z["ref_g_logits"]) <= 1e-10
    assert np.abs(o["g_scale"] - z["ref_g_scale"]).max() <= 1e-10 * np.abs(z["ref_g_scale"]).max()


ZEBRA_GOLDEN = ["zebra_b2_16x16_s2.npz", "zebra_b2_48x40_s3.npz", "zebra_b1_72x64_s3.npz"]


@pytest.mark.parametrize("name", ZEBRA_GOLDEN)
def test_zebra_producer_oracle_matches_reference_golden(oracle, name):
    """oracle.dense_pose_loss_noc_bin / noc_bin_decode / noc_to_bits == the reference's ZebraPose branch of
    dense_pose_loss (+ autograd), floatbits.nn_logits2noc and floatbits.nn_noc2target, run unmodified in fp64."""
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, name))
    T = z["in_model_transform"] if z["in_model_transform"].size else None
    o = oracle.dense_pose_loss_noc_bin(z["in_bin_logits"], z["in_raw_bits"], z["in_msk_noc"], z["in_logits"], z["in_scale"],
                                       z["in_noc_scale"], z["in_K"], z["in_pose"], z["in_bbox_3d"], z["bit_cnt"], int(z["sample"]),
                                       tuple(z["top_left"]), T)
    assert np.abs(o["loss"] - z["ref_loss"]).max() <= 1e-11 * np.abs(z["ref_loss"]).max()
    assert rel_err(o["pts3d"], z["ref_pts3d"]) <= 1e-13
    assert rel_err(o["g_bin_logits"], z["ref_g_bin_logits"]) <= 2e-7      # stored as fp32
    assert rel_err(o["g_logits"], z["ref_g_logits"]) <= 1e-10
    assert np.abs(o["g_scale"] - z["ref_g_scale"]).max() <= 1e-10 * np.abs(z["ref_g_scale"]).max()
    noc = oracle.noc_bin_decode(z["in_bin_logits"], z["bit_cnt"])
    assert np.abs(noc - z["ref_noc_inference"]).max() <= 1e-7             # stored as fp32
    mod, raw = oracle.noc_to_bits(z["ref_noc_inference"].astype(np.float64), z["bit_cnt"])
    assert np.array_equal(np.packbits(mod), z["ref_target_mod"]) and np.array_equal(np.packbits(raw), z["ref_target_raw"])


@pytest.mark.parametrize("name", ["select_b3_32x32_s1.npz", "select_b2_64x48_s2.npz"])
def test_selection_oracle_matches_reference_golden(oracle, name):
    """oracle.dense_point_select == the reference's quantile_msk / dense_pnp_matching_from_xyz / nn_out_to_xyz (test.py:36-45,
    67-106) bit for bit: selected sets for the three cfg.dense_point_select rules and the gathered fp32 values."""
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, name))
    xyz = (z["in_xyz_noc"].transpose(0, 2, 3, 1) * z["in_noc_scale"][:, None, None, :]).astype(np.float32)
    for mode in ("mask", "quantile", "quantile_in_mask"):
        r = oracle.dense_point_select(xyz, z["ref_weights"], z["in_msk_logits"], int(z["sample"]), mode)
        assert np.array_equal(r["valid"], z["valid_" + mode])
        assert np.array_equal(r["pts3d"], z["ref_pts3d"]) and np.array_equal(r["inv_cov"], z["ref_inv_cov"])
        assert np.array_equal(r["pts2d"], z["ref_pts2d"][0])


def test_eval_oracle_matches_reference_golden(oracle):
    """oracle.pose_errors / select_pose == error6d.add/adi/re/te (numpy + cKDTree) and symmetry.select_pose_2d/3d."""
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "eval_b6_m2500.npz"))
    o = oracle.pose_errors(z["R_est"], z["t_est"], z["R_gt"], z["t_gt"], z["pts"])
    for k in ("adi", "add", "te"):
        assert np.allclose(o[k], z["ref_" + k], rtol=1e-12, atol=1e-12), k
    assert np.allclose(o["re"], z["ref_re"], rtol=0, atol=1e-5)          # acos near 1 amplifies rounding (1e-16 -> 1e-6 deg)
    b2, i2, _ = oracle.select_pose(0, z["c_K"], z["c_pts3d"], z["c_pts2d"], z["c_candi"])
    b3, i3, _ = oracle.select_pose(1, z["c_K"], z["c_noisy3d"], z["c_homo_z"], z["c_candi"])
    assert np.array_equal(b2.astype(np.float32), z["ref_best2d"]) and np.array_equal(b3.astype(np.float32), z["ref_best3d"])
    assert list(i2) == [0, 5, 7, 11] and list(i3) == [0, 5, 7, 11]       # the candidate lists were rolled by [0, 5, 7, 11] of 12


@pytest.mark.parametrize("tag", ["a", "b"])
def test_init_oracle_lands_in_the_same_lm_basin_as_opencv(oracle, tag):
    """The initialiser restated in oracle.pnp_init is our own algorithm (the reference calls OpenCV's RANSAC-EPnP).  What ties
    it to the reference: the LM solve started from it stops at the same optimum as the LM solve started from the pose the
    reference's cv2_solver.solve returned (fixture init_cv2.npz), within the solver's early-stop gap."""
    import os
    from conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "init_cv2.npz"))
    K, X, x, s = (z[f"{tag}_{k}"] for k in ("K", "pts3d", "pts2d", "inv_std"))
    B = len(K)
    from scipy.spatial.transform import Rotation as Ro
    starts = np.zeros((B, 7), np.float32)
    for b in range(B):
        ok, R, t, inl = oracle.pnp_init(K[b], X[b], x[b], s[b] ** 2, 3.0, 3)
        assert ok
        q = Ro.from_matrix(R).as_quat()
        starts[b] = np.concatenate(([q[3]], q[:3], t))
        # the inlier sets agree with RANSAC's up to points near the 3 px threshold / a slightly different pose
        assert (inl != z[f"{tag}_cv_inliers"][b]).mean() <= (0.15 if X.shape[1] >= 64 else 0.25)
    L = np.zeros(X.shape[:2] + (2, 2), np.float32)
    L[..., 0, 0], L[..., 1, 1] = s[..., 0], s[..., 1]
    ours = oracle.lm_solve(K, X, x, L, starts)
    cv = oracle.lm_solve(K, X, x, L, z[f"{tag}_cv_states"])
    assert not ours["invalid"].any() and not cv["invalid"].any()
    assert quat_angle(ours["states"][:, :4].astype(np.float64), cv["states"][:, :4].astype(np.float64)).max() <= 1e-3
    assert (np.abs(ours["states"][:, 4:] - cv["states"][:, 4:]).max(1) <= 2e-3 * np.abs(cv["states"][:, 4:]).max(1)).all()


def test_zebra_oracle_properties(oracle):
    """Size-independent properties of the restated ZebraPose coding: (i) encode -> saturated logits -> training decode with the
    GT bits returns the quantised coordinate inside the mask; (ii) a single wrong bit moves the decoded value by at most that
    bit's weight and the gradient lands on exactly that bit; (iii) outside the mask the value is the hard decode, no gradient."""
    rng = np.random.default_rng(0)
    bits = (7, 6, 5)
    noc = rng.uniform(-1.1, 1.1, (2, 6, 7, 3))
    mod, raw = oracle.noc_to_bits(noc, bits)
    logits = (mod.astype(np.float64) * 2 - 1) * 30.0
    msk = np.ones((2, 6, 7), bool)
    dec, sel, dnoc = oracle.noc_bin_decode_with_gt(logits, raw, msk, bits)
    off = np.concatenate(([0], np.cumsum(bits)))
    for a, N in enumerate(bits):
        mx = 2 ** N - 1
        q = np.rint(np.clip((noc[..., a] + 1) * (mx * 0.5), 0, mx))
        # no wrong bit: the LSB is the soft one, sigmoid(+-30) restores it to 1e-13
        assert np.abs(dec[..., a] - (q / (mx * 0.5) - 1)).max() <= 1e-12
        assert (sel[..., a] == off[a] + N - 1).all()
    # flip one mid bit of axis 0 at one pixel
    l2 = logits.copy()
    j = 2
    l2[0, j, 3, 4] *= -1
    dec2, sel2, dnoc2 = oracle.noc_bin_decode_with_gt(l2, raw, msk, bits)
    assert sel2[0, 3, 4, 0] == j and abs(dec2[0, 3, 4, 0] - dec[0, 3, 4, 0]) <= 2 ** (bits[0] - 1 - j) / ((2 ** bits[0] - 1) * 0.5) + 1e-9
    assert (sel2[..., 1:] == sel[..., 1:]).all()
    # outside the mask: hard decode, zero gradient
    msk0 = np.zeros_like(msk)
    dec3, _, dnoc3 = oracle.noc_bin_decode_with_gt(l2, raw, msk0, bits)
    assert (dnoc3 == 0).all()
    hard = oracle.noc_bin_decode(logits, bits)
    assert np.abs(oracle.noc_bin_decode_with_gt(logits, raw, msk0, bits)[0] - hard).max() <= 2.0 / (2 ** min(bits) - 1) + 1e-9


def test_selection_oracle_properties(oracle):
    """quantile rule keeps ceil-ish (1-q) of the points, 'quantile_in_mask' is a subset of the mask, 'mask' ignores weights."""
    rng = np.random.default_rng(1)
    B, H, W = 3, 20, 24
    xyz = rng.normal(size=(B, H, W, 3)).astype(np.float32)
    w = rng.uniform(0.1, 2.0, (B, 2, H, W)).astype(np.float32)
    ml = rng.normal(size=(B, 1, H, W)).astype(np.float32)
    for sample in (1, 2):
        m = oracle.dense_point_select(xyz, w, ml, sample, "mask")
        q = oracle.dense_point_select(xyz, w, ml, sample, "quantile", quantile=0.2)
        qm = oracle.dense_point_select(xyz, w, ml, sample, "quantile_in_mask", quantile=0.2)
        N = m["valid"].shape[1]
        assert np.array_equal(m["valid"], (ml[:, 0, ::sample, ::sample] > 0).reshape(B, -1))
        assert (np.abs(q["valid"].sum(1) - 0.8 * N) <= 2).all()
        assert not (qm["valid"] & ~m["valid"]).any()
        assert (np.abs(qm["valid"].sum(1) - 0.8 * m["valid"].sum(1)) <= 2).all()


def test_lm_oracle_rule_switches_are_wired(oracle):
    """The from-memory Ceres rules are flags of the oracle (tools/lm_sensitivity.py, profiles/lm_unpinned_sensitivity.md):
    the guard rule decides whether a restart from a converged pose takes 1 or 2 iterations, keeping the tolerance-triggering
    candidate moves the pose by the early-stop gap, QR vs normal equations agree to 1e-10 on well-posed problems."""
    c = make_correspondences(24, 64, 77).to(torch.float32)
    L = torch.diag_embed((c.inv_std ** 2).sqrt())
    base = oracle.lm_solve(c.K, c.pts3d, c.pts2d, L, c.start)
    ne = oracle.lm_solve(c.K, c.pts3d, c.pts2d, L, c.start, flags=oracle.LM_TOL_NEEDS_SUCCESS | oracle.LM_SOLVE_NORMAL_EQ)
    assert np.array_equal(base["iters"], ne["iters"]) and np.abs(base["x6"] - ne["x6"]).max() < 1e-10
    keep = oracle.lm_solve(c.K, c.pts3d, c.pts2d, L, c.start, flags=oracle.LM_TOL_NEEDS_SUCCESS | oracle.LM_TOL_KEEP_CANDIDATE)
    assert np.array_equal(base["iters"], keep["iters"])
    d = np.abs(base["x6"][:, :3] - keep["x6"][:, :3]).max(axis=1)
    assert (d > 0).all() and d.max() < 5e-3
    again = oracle.lm_solve(c.K, c.pts3d, c.pts2d, L, base["states"])
    unguarded = oracle.lm_solve(c.K, c.pts3d, c.pts2d, L, base["states"], flags=0)
    assert (again["iters"] == 2).all() and (unguarded["iters"] == 1).all()
    # candidate discarded: the start comes back (through quaternion -> angle-axis -> quaternion in fp32)
    assert np.allclose(unguarded["states"], base["states"], rtol=3e-7, atol=1e-7)


@pytest.mark.parametrize("name", ["cov2d_b3_n200_s4.npz", "cov2d_b2_n16_s5.npz", "cov2d_b2_n700_s6_mask.npz"])
def test_oracle_cov_2d_matches_the_reference(oracle, name):
    """Loss_cov_mixed(..., cov_2d=True) (lib/cov_mixed.py:76-80, 91-97): fixtures generated by the unmodified reference
    (tests/golden/make_golden.py --only-cov2d)."""
    z = np.load(os.path.join(GOLDEN_DIR, name))
    v = z["in_valid"] if bool(z["has_valid"]) else None
    o = oracle.lc_loss(z["in_K"], z["in_pose"], z["in_pts3d"], z["in_pts2d"], z["in_inv_std"], v, z["in_bbox_3d"], cov_2d=True)
    assert np.abs(o["loss"] - z["ref_loss"]).max() <= 1e-12 * np.abs(z["ref_loss"]).max()
    for k in ("g_pts3d", "g_pts2d", "g_inv_std"):
        assert np.abs(o[k] - z["ref_" + k]).max() <= 1e-11 * np.abs(z["ref_" + k]).max(), k
    o3 = oracle.lc_loss(z["in_K"], z["in_pose"], z["in_pts3d"], z["in_pts2d"], z["in_inv_std"], v, z["in_bbox_3d"])
    assert np.abs(o3["loss"] - z["ref_loss"]).max() > 0.1   # the 3-D variant is a different loss
