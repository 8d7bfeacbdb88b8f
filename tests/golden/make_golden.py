"""Generate golden vectors for the loss half by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is not on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``lib.cov_mixed.Loss_cov_mixed`` and ``lib.nll.pnp_auto`` from
``/root/reference`` (importable as-is with torch 2.11), feeds them
the seeded synthetic correspondences of ``lc_b200.synth`` in fp64, and stores
inputs (rounded to fp32-representable values so the same fixture serves the fp32
kernel path) plus the reference outputs:

  loss (B,), grads wrt pts3d / pts2d / inv_std, jac (B,6,N,2) = d(update)/d(pts2d),
  cov (B,6,6) = H^-1, robust weights W and variance estimate sigma (B,N,2).

Nothing under tests/ or the product imports the reference at run time; only
this script does.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from lib.cov_mixed import Loss_cov_mixed, clamp_error, robust_weights_cov  # noqa: E402  (reference)
from lib.nll import pnp_auto  # noqa: E402  (reference)
from lib import transforms as xforms  # noqa: E402  (reference)

from lc_b200.synth import make_correspondences  # noqa: E402

# name, B, N, seed, valid mode, regime, store_jac
CASES = [
    ("b3_n8_s0", 3, 8, 0, "none", "nominal", True),
    ("b3_n16_s1", 3, 16, 1, "ones", "nominal", True),
    ("b4_n200_s0", 4, 200, 0, "ones", "nominal", True),
    ("b4_n200_s2_mask", 4, 200, 2, "mask", "nominal", True),
    ("b3_n333_s1_init", 3, 333, 1, "ones", "random_init", True),
    ("b2_n1024_s0", 2, 1024, 0, "ones", "nominal", True),
    ("b2_n1024_s1_heavy", 2, 1024, 1, "ones", "heavy_outliers", False),
    ("b1_n4096_s2", 1, 4096, 2, "ones", "nominal", False),
]


def build_inputs(B, N, seed, valid_mode, regime):
    outlier = 0.30 if regime == "heavy_outliers" else 0.05
    c = make_correspondences(B, N, seed, outlier_frac=outlier)
    if regime == "random_init":
        # what a random-init network emits: near-uniform softmax weights ~1/(2HW) and tiny xyz
        c.inv_std = c.inv_std * 1e-4
        c.pts3d = c.pts3d * 0.02
    if regime == "heavy_outliers":
        c.pts2d = c.pts2d + 30.0 * (torch.rand(c.pts2d.shape, generator=torch.Generator().manual_seed(seed + 1), dtype=torch.float64) < 0.1)
    r32 = lambda t: t.to(torch.float32).to(torch.float64)
    d = dict(K=r32(c.K), pose=r32(c.pose), pts3d=r32(c.pts3d), pts2d=r32(c.pts2d),
             inv_std=r32(c.inv_std), bbox_3d=r32(c.bbox_3d))
    # keep the quaternion unit in fp64 after the fp32 rounding? No: the kernel sees the fp32 values.
    if valid_mode == "none":
        d["valid"] = None
    elif valid_mode == "ones":
        d["valid"] = torch.ones(B, N, dtype=torch.float64)
    else:
        g = torch.Generator().manual_seed(seed + 99)
        d["valid"] = (torch.rand(B, N, generator=g) < 0.7).to(torch.float64)
    return d


def run_reference(d, max_err_len=32, rel_thresh=3, w_e_thresh=4):
    pts3d = d["pts3d"].clone().requires_grad_(True)
    pts2d = d["pts2d"].clone().requires_grad_(True)
    inv_std = d["inv_std"].clone().requires_grad_(True)
    loss = Loss_cov_mixed(d["K"], d["pose"], pts3d, pts2d, inv_std, d["valid"], bbox_3d=d["bbox_3d"],
                          max_err_len=max_err_len, rel_thresh=rel_thresh, w_e_thresh=w_e_thresh)
    g3, g2, gs = torch.autograd.grad(loss.sum(), (pts3d, pts2d, inv_std))

    with torch.no_grad():
        R, t = xforms.quaternion_rep_to_RT(d["pose"])
        proj = xforms.project_apply(d["K"], d["pts3d"], R, t)
        ec = clamp_error(d["pts2d"] - proj, max_err_len)
        W, sigma = robust_weights_cov(d["inv_std"], ec, d["valid"], rel_thresh=rel_thresh, w_e_thresh=w_e_thresh)
    jac, cov = pnp_auto.weighted_pnp_jac_wrt_pts2d(proj, d["pose"], d["K"], d["pts3d"], W, with_cov=True)
    return dict(loss=loss.detach(), g_pts3d=g3, g_pts2d=g2, g_inv_std=gs, W=W, sigma=sigma,
                jac=jac.detach(), cov=cov.detach())


def make_cov2d():
    """Loss_cov_mixed(..., cov_2d=True) (cov_mixed.py:76-80, 91-97, 129-131): the projected-bbox-corner variant no reference config
    enables.  Stand-alone fixtures so the files of the default variant stay untouched."""
    for name, B, N, seed, vm in (("cov2d_b3_n200_s4", 3, 200, 4, "ones"), ("cov2d_b2_n16_s5", 2, 16, 5, "none"), ("cov2d_b2_n700_s6_mask", 2, 700, 6, "mask")):
        d = build_inputs(B, N, seed, vm, "nominal")
        pts3d = d["pts3d"].clone().requires_grad_(True)
        pts2d = d["pts2d"].clone().requires_grad_(True)
        inv_std = d["inv_std"].clone().requires_grad_(True)
        loss = Loss_cov_mixed(d["K"], d["pose"], pts3d, pts2d, inv_std, d["valid"], bbox_3d=d["bbox_3d"], max_err_len=32, cov_2d=True)
        g3, g2, gs = torch.autograd.grad(loss.sum(), (pts3d, pts2d, inv_std))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, in_K=d["K"].numpy().astype(np.float32), in_pose=d["pose"].numpy().astype(np.float32),
                            in_pts3d=d["pts3d"].numpy().astype(np.float32), in_pts2d=d["pts2d"].numpy().astype(np.float32),
                            in_inv_std=d["inv_std"].numpy().astype(np.float32), in_bbox_3d=d["bbox_3d"].numpy().astype(np.float32),
                            has_valid=np.array(d["valid"] is not None),
                            in_valid=(d["valid"].numpy().astype(np.float32) if d["valid"] is not None else np.zeros((B, N), np.float32)),
                            ref_loss=loss.detach().numpy(), ref_g_pts3d=g3.numpy(), ref_g_pts2d=g2.numpy(), ref_g_inv_std=gs.numpy())
        print(name, os.path.getsize(path) // 1024, "KiB", loss.detach().numpy())


def make_jac_exact():
    """weighted_pnp_jac_wrt_pts2d AWAY from the optimum (measured pts2d, residual != 0): exercises the r * d2r term
    of hessian_6d_elem (pnp_auto.py:59-83) and the double-backward w.r.t. the weights (:129-134)."""
    c = make_correspondences(3, 40, 77)
    r32 = lambda t: t.to(torch.float32).to(torch.float64)
    K, pose, X, x = r32(c.K), r32(c.pose), r32(c.pts3d), r32(c.pts2d)
    W = r32(c.inv_std ** 2).requires_grad_(True)
    g = torch.Generator().manual_seed(5)
    Gj = torch.randn(3, 6, 40, 2, generator=g, dtype=torch.float64)
    Gc = torch.randn(3, 6, 6, generator=g, dtype=torch.float64)
    jac, cov = pnp_auto.weighted_pnp_jac_wrt_pts2d(x, pose, K, X, W, with_cov=True)
    gW, = torch.autograd.grad((jac * Gj).sum() + (cov * Gc).sum(), W)
    path = os.path.join(HERE, "jacx_b3_n40.npz")
    np.savez_compressed(path, in_K=K.numpy().astype(np.float32), in_pose=pose.numpy().astype(np.float32),
                        in_pts3d=X.numpy().astype(np.float32), in_pts2d=x.numpy().astype(np.float32),
                        in_W=W.detach().numpy().astype(np.float32), Gj=Gj.numpy(), Gc=Gc.numpy(),
                        ref_jac=jac.detach().numpy(), ref_cov=cov.detach().numpy(), ref_gW=gW.numpy())
    print("jacx_b3_n40:", os.path.getsize(path) // 1024, "KiB")


def dense_inputs(B, H, W, seed):
    """Network-output-shaped inputs whose back-projection roughly matches the pixel grid (plus noise/outliers)."""
    g = torch.Generator().manual_seed(seed)
    c = make_correspondences(B, 4, seed)
    f64 = torch.float64
    K, pose = c.K.clone(), c.pose
    K[:, :2, :] = K[:, :2, :] * (W / 64.0)                       # crop of W x H pixels instead of 64 x 64
    R = __import__("lc_b200.synth", fromlist=["quat_to_matrix"]).quat_to_matrix(pose[:, :4])
    ys, xs = torch.meshgrid(torch.arange(H, dtype=f64), torch.arange(W, dtype=f64), indexing="ij")
    pix = torch.stack((xs, ys, torch.ones_like(xs)), -1).reshape(1, H * W, 3).expand(B, -1, -1)
    z = pose[:, None, 6:7] + 40 * (2 * torch.rand(B, H * W, 1, generator=g, dtype=f64) - 1)
    P = torch.linalg.solve(K, pix.mT).mT * z
    X = (P - pose[:, None, 4:]) @ R                                # R^T (P - t)
    noc_scale = torch.tensor([[40.0, 50.0, 60.0]], dtype=f64).expand(B, 3).clone()
    X = X + 1.5 * torch.randn(B, H * W, 3, generator=g, dtype=f64)
    xyz_noc = (X / noc_scale[:, None, :]).mT.reshape(B, 3, H, W)
    logits = torch.randn(B, 2, H, W, generator=g, dtype=f64)
    scale = (2.0 * H * W) * torch.exp(0.2 * torch.randn(B, 1, 1, 1, generator=g, dtype=f64))
    r32 = lambda t: t.to(torch.float32).to(f64)
    return dict(xyz_noc=r32(xyz_noc), logits=r32(logits), scale=r32(scale), noc_scale=r32(noc_scale), K=r32(K),
                pose=r32(pose), bbox_3d=r32(c.bbox_3d))


def make_dense():
    """dense_pose_loss glue of the reference (losses.py:355-356, 142-161, 366, 383) run unmodified in fp64."""
    import losses as ref_losses  # noqa: E402  (reference)
    for name, B, H, W, sample, tl, seed in (("dense_b2_16x16_s2", 2, 16, 16, 2, (1, 0), 3), ("dense_b2_64x64_s2", 2, 64, 64, 2, (0, 1), 4),
                                           ("dense_b2_40x56_s3", 2, 40, 56, 3, (2, 1), 5)):
        d = dense_inputs(B, H, W, seed)
        xyz = d["xyz_noc"].clone().requires_grad_(True)
        lg = d["logits"].clone().requires_grad_(True)
        sc = d["scale"].clone().requires_grad_(True)
        w_raw = lg.reshape(lg.shape[:-3] + (1, -1)).softmax(dim=-1)
        weights = w_raw.reshape_as(lg) * sc
        p2, inv_std, p3, _ = ref_losses.dense_pnp_matching_from_xyz(xyz, weights, None, d["noc_scale"], sample=sample, top_left=tl)
        valid = torch.ones_like(p3[..., 0])
        loss = Loss_cov_mixed(d["K"], d["pose"], p3, p2, inv_std, valid, bbox_3d=d["bbox_3d"], max_err_len=32)
        gx, gl, gs = torch.autograd.grad(loss.sum(), (xyz, lg, sc))
        path = os.path.join(HERE, "densex_" + name[6:] + ".npz")
        np.savez_compressed(path, **{("in_" + k): v.numpy().astype(np.float32) for k, v in d.items()},
                            sample=np.array(sample), top_left=np.array(tl), ref_loss=loss.detach().numpy(),
                            ref_g_xyz_noc=gx.numpy(), ref_g_logits=gl.numpy(), ref_g_scale=gs.numpy().reshape(B))
        print(name, loss.detach().numpy(), os.path.getsize(path) // 1024, "KiB")


def make_zebra():
    """ZebraPose branch of dense_pose_loss (losses.py:355-356, 163-184, 16-45, 375, 383; floatbits.py:49-69, 99-160) and
    the inference decode (floatbits.py:33-47, 197-224), run unmodified in fp64."""
    import losses as ref_losses  # noqa: E402  (reference)
    import floatbits as ref_fb   # noqa: E402  (reference)
    from lc_b200.synth import make_zebra_outputs
    for name, B, H, W, sample, tl, seed, bits, xf in (("zebra_b2_16x16_s2", 2, 16, 16, 2, (1, 0), 3, (5, 5, 5), False),
                                                     ("zebra_b2_48x40_s3", 2, 48, 40, 3, (2, 1), 5, (7, 6, 5), True),
                                                     ("zebra_b1_72x64_s3", 1, 72, 64, 3, (0, 2), 6, (7, 7, 6), True)):
        d = make_zebra_outputs(B, H, W, seed, bits, with_transform=xf)
        f64 = torch.float64
        bl = d["bin_logits"].to(f64).requires_grad_(True)
        lg = d["logits"].to(f64).requires_grad_(True)
        sc = d["scale"].to(f64).requires_grad_(True)
        T = None if d["model_transform"] is None else d["model_transform"].to(f64)
        gt = dict(bit_cnt=list(bits))
        if T is not None:
            gt["model_transform"] = T
        w_raw = lg.reshape(lg.shape[:-3] + (1, -1)).softmax(dim=-1)
        weights = w_raw.reshape_as(lg) * sc
        msk_vis = torch.ones(B, H, W, dtype=torch.bool)
        ref_fb.set_black_background(True)
        p2, inv_std, p3, _ = ref_losses.dense_pnp_matching_from_noc_bin(bl, d["raw_bits"], weights, msk_vis, d["msk_noc"],
                                                                       d["noc_scale"].to(f64), gt, sample=sample, top_left=tl)
        valid = torch.ones_like(p3[..., 0])
        loss = Loss_cov_mixed(d["K"].to(f64), d["pose"].to(f64), p3, p2, inv_std, valid, bbox_3d=d["bbox_3d"].to(f64), max_err_len=32)
        gb, gl, gs = torch.autograd.grad(loss.sum(), (bl, lg, sc))
        with torch.no_grad():
            noc_inf = ref_fb.nn_logits2noc(d["bin_logits"].to(f64), list(bits))
            tgt_mod, tgt_raw = ref_fb.nn_noc2target(noc_inf, list(bits))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(
            path, in_bin_logits=d["bin_logits"].numpy(), in_raw_bits=d["raw_bits"].numpy(), in_msk_noc=d["msk_noc"].numpy(),
            in_logits=d["logits"].numpy(), in_scale=d["scale"].numpy(), in_noc_scale=d["noc_scale"].numpy(), in_K=d["K"].numpy(),
            in_pose=d["pose"].numpy(), in_bbox_3d=d["bbox_3d"].numpy(), bit_cnt=np.array(bits),
            in_model_transform=np.zeros((0,), np.float32) if T is None else d["model_transform"].numpy(),
            sample=np.array(sample), top_left=np.array(tl), ref_loss=loss.detach().numpy(), ref_pts3d=p3.detach().numpy(),
            ref_g_bin_logits=gb.numpy().astype(np.float32), ref_g_logits=gl.numpy(), ref_g_scale=gs.numpy().reshape(B),
            ref_noc_inference=noc_inf.numpy().astype(np.float32), ref_target_mod=np.packbits(tgt_mod.numpy()),
            ref_target_raw=np.packbits(tgt_raw.numpy()))
        print(name, loss.detach().numpy(), os.path.getsize(path) // 1024, "KiB")


def make_select():
    """Test-time point selection of solve_pnp_dense (test.py:67-106).  test.py itself cannot be imported (mmcv), so its
    quantile_msk (test.py:36-45) is exec'd from the unmodified source text and the glue lines around it are replayed with
    the reference's own dense_pnp_matching_from_xyz / nn_out_to_xyz, in fp32 like the reference runs them."""
    import re
    import losses as ref_losses  # noqa: E402  (reference)
    from torch import Tensor  # noqa: F401  (used by the exec'd source)
    from typing import Union  # noqa: F401
    src = open("/root/reference/test.py").read()
    fn_src = re.search(r"^def quantile_msk\(.*?(?=^\S)", src, flags=re.S | re.M).group(0)
    ns = dict(torch=torch, Tensor=torch.Tensor, Union=Union)
    exec(fn_src, ns)
    quantile_msk = ns["quantile_msk"]
    from lc_b200.synth import make_dense_outputs
    for name, B, H, W, sample, seed, scale_dim in (("select_b3_32x32_s1", 3, 32, 32, 1, 21, 1), ("select_b2_64x48_s2", 2, 64, 48, 2, 22, 2)):
        d = make_dense_outputs(B, H, W, seed)
        g = torch.Generator().manual_seed(seed)
        msk_logits = 2.0 * torch.randn(B, 1, H, W, generator=g) + 0.8
        scale = d["scale"] if scale_dim == 1 else d["scale"].expand(B, 2, 1, 1) * torch.tensor([1.0, 0.7]).reshape(1, 2, 1, 1)
        lg = d["logits"]
        with torch.no_grad():
            seg_msk = torch.sigmoid(msk_logits) > 0.5                                                    # test.py:70
            xyz_out = ref_losses.nn_out_to_xyz(d["xyz_noc"], d["noc_scale"], bit_cnt=None, inference=True)  # test.py:78-82
            w_raw = lg.reshape(lg.shape[:-3] + (scale.shape[-3], -1)).softmax(dim=-1)                    # test.py:86-88
            weights = w_raw.reshape_as(lg) * scale
            p2, inv_std, p3, seg_valid = ref_losses.dense_pnp_matching_from_xyz(xyz_out.permute(0, 3, 1, 2), weights, seg_msk.squeeze(-3),
                                                                               None, sample, top_left=(0, 0))   # test.py:91-93
            out = dict(in_xyz_noc=d["xyz_noc"].numpy(), in_noc_scale=d["noc_scale"].numpy(), in_logits=lg.numpy(), in_scale=scale.numpy(),
                       in_msk_logits=msk_logits.numpy(), sample=np.array(sample), ref_weights=weights.numpy(),
                       ref_pts3d=p3.numpy(), ref_pts2d=p2.numpy(), ref_inv_cov=(inv_std ** 2).numpy())
            out["valid_mask"] = seg_valid.numpy()                                                        # test.py:97-98
            out["valid_quantile"] = quantile_msk(inv_std, 0.2).numpy()                                   # test.py:99-100
            vis_ratio = seg_valid.float().mean(dim=-1)                                                   # test.py:101-104
            quantile = 1 - (1 - 0.2) * vis_ratio
            out["valid_quantile_in_mask"] = (quantile_msk(inv_std * seg_valid[..., None], quantile) * seg_valid).numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, {k: int(out[k].sum()) for k in out if k.startswith("valid_")}, os.path.getsize(path) // 1024, "KiB")


def make_eval():
    """compute_pose_errors (lib/utils/error6d.py add/adi/re/te) and symmetry.select_pose_2d/3d, run unmodified."""
    from lib.utils import error6d  # noqa: E402  (reference)
    import symmetry as ref_sym      # noqa: E402  (reference)
    from lc_b200.synth import make_correspondences, quat_to_matrix
    rng = np.random.default_rng(7)
    B, M = 6, 2500
    pts = rng.uniform(-1, 1, (M, 3)) * np.array([40.0, 55.0, 70.0])
    c = make_correspondences(B, 4, 31)
    R_gt = quat_to_matrix(c.pose[:, :4]).numpy()
    t_gt = c.pose[:, 4:].numpy()
    R_est = quat_to_matrix(c.start[:, :4]).numpy()
    t_est = c.start[:, 4:].numpy()
    R_est[3] = R_gt[3] @ np.diag([-1.0, -1.0, 1.0])            # a 180 degree flip: ADI << ADD
    R_est[4], t_est[4] = R_gt[4], t_gt[4]                        # exact pose: all errors 0
    errs = {k: np.zeros(B) for k in ("adi", "add", "re", "te")}
    for b in range(B):
        errs["adi"][b] = error6d.adi(R_est[b], t_est[b].reshape(3, 1), R_gt[b], t_gt[b].reshape(3, 1), pts)
        errs["add"][b] = error6d.add(R_est[b], t_est[b].reshape(3, 1), R_gt[b], t_gt[b].reshape(3, 1), pts)
        errs["re"][b] = error6d.re(R_est[b], R_gt[b])
        errs["te"][b] = error6d.te(t_est[b], t_gt[b])
    # candidate selection: Kc candidates = the true pose composed with rotations about z, listed from a random offset
    Bc, N, Kc = 4, 300, 12
    cc = make_correspondences(Bc, N, 33).to(torch.float32)
    Rt = quat_to_matrix(cc.pose[:, :4].double()).float()
    ang = torch.arange(Kc, dtype=torch.float32) * (2 * np.pi / Kc)
    Rz = torch.zeros(Kc, 3, 3)
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = torch.cos(ang), -torch.sin(ang), torch.sin(ang), torch.cos(ang), 1
    shift = torch.tensor([0, 5, 7, 11])
    candi = torch.stack([torch.cat((Rt[b] @ Rz.roll(int(shift[b]), 0), cc.pose[b, 4:].float().reshape(1, 3, 1).expand(Kc, 3, 1)), -1) for b in range(Bc)])
    P = cc.pts3d @ Rt.mT + cc.pose[:, None, 4:].float()
    homo_z = P @ cc.K.mT
    noisy3d = cc.pts3d + 0.5 * torch.randn(cc.pts3d.shape, generator=torch.Generator().manual_seed(3))
    best2d = ref_sym.select_pose_2d(cc.K, cc.pts3d, cc.pts2d, candi)
    best3d = ref_sym.select_pose_3d(cc.K, noisy3d, homo_z, candi)
    path = os.path.join(HERE, "eval_b6_m2500.npz")
    np.savez_compressed(path, pts=pts, R_est=R_est, t_est=t_est, R_gt=R_gt, t_gt=t_gt, **{"ref_" + k: v for k, v in errs.items()},
                        c_K=cc.K.numpy(), c_pts3d=cc.pts3d.numpy(), c_pts2d=cc.pts2d.numpy(), c_noisy3d=noisy3d.numpy(),
                        c_homo_z=homo_z.numpy(), c_candi=candi.numpy(), ref_best2d=best2d.numpy(), ref_best3d=best3d.numpy())
    print("eval_b6_m2500", errs, os.path.getsize(path) // 1024, "KiB")


def make_init():
    """Start poses and inlier sets of the reference's cv2_solver.solve (cv2_solver.py:6-88: cv2.solvePnPRansac, EPNP, 150
    iterations), run unmodified on synthetic correspondences with an axis-aligned K (OpenCV ignores the off-diagonal terms)."""
    from lib.pnp import cv2_solver  # noqa: E402  (reference)
    from lc_b200.synth import make_correspondences, quat_to_matrix
    import cv2
    cv2.setRNGSeed(1234)
    out = {}
    for tag, B, N, outl in (("a", 6, 256, 0.1), ("b", 6, 16, 0.0)):
        c = make_correspondences(B, N, 41, outlier_frac=outl)
        Rt = quat_to_matrix(c.pose[:, :4])
        P = c.pts3d @ Rt.mT + c.pose[:, None, 4:]
        h = P @ c.K.mT
        noise = c.pts2d - h[..., :2] / h[..., 2:]
        f = (c.K[:, 0, 0] ** 2 + c.K[:, 0, 1] ** 2).sqrt()
        K0 = torch.zeros(B, 3, 3, dtype=torch.float64)
        K0[:, 0, 0], K0[:, 1, 1], K0[:, 0, 2], K0[:, 1, 2], K0[:, 2, 2] = f, f, 32.0, 32.0, 1.0
        h0 = P @ K0.mT
        x0 = h0[..., :2] / h0[..., 2:] + noise
        K32, X32, x32 = K0.float(), c.pts3d.float(), x0.float()
        invalids, states, inliers = cv2_solver.solve(K32, X32, x32, reprojectionError=3.0)
        mask = np.zeros((B, N), bool)
        for b, idx in enumerate(inliers):
            mask[b, idx.numpy()] = True
        out.update({f"{tag}_K": K32.numpy(), f"{tag}_pts3d": X32.numpy(), f"{tag}_pts2d": x32.numpy(),
                    f"{tag}_inv_std": c.inv_std.float().numpy(), f"{tag}_pose": c.pose.float().numpy(),
                    f"{tag}_cv_states": torch.stack(states).float().numpy(), f"{tag}_cv_invalid": np.array(invalids),
                    f"{tag}_cv_inliers": mask})
        print(tag, "cv2 invalid", list(invalids), "inlier fraction", mask.mean(1).round(2))
    path = os.path.join(HERE, "init_cv2.npz")
    np.savez_compressed(path, **out)
    print("init_cv2", os.path.getsize(path) // 1024, "KiB")


def main():
    torch.set_num_threads(os.cpu_count())
    if "--only-cov2d" in sys.argv:
        make_cov2d()
        return
    if "--only-init" in sys.argv:
        make_init()
        return
    if "--only-eval" in sys.argv:
        make_eval()
        return
    if "--only-zebra" in sys.argv:
        make_zebra()
        return
    if "--only-select" in sys.argv:
        make_select()
        return
    make_jac_exact()
    make_cov2d()
    make_dense()
    make_zebra()
    make_select()
    make_eval()
    make_init()
    for name, B, N, seed, vmode, regime, store_jac in CASES:
        d = build_inputs(B, N, seed, vmode, regime)
        o = run_reference(d)
        out = {}
        for k, v in d.items():
            if v is not None:
                out["in_" + k] = v.numpy().astype(np.float32)
        out["has_valid"] = np.array(d["valid"] is not None)
        for k, v in o.items():
            if k in ("jac", "W", "sigma") and not store_jac:
                continue
            out["ref_" + k] = v.numpy().astype(np.float64)
        out["params"] = np.array([32.0, 3.0, 4.0])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: loss={o['loss'].numpy()} -> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
