"""Small run of every kernel path for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc_b200.synth import make_correspondences, planar_view, full_icov_from_inv_std
from lc_b200.cov_mixed import loss_fwd_bwd
from lc_b200.fused import solve_and_loss
from lc_b200.pnp.cer_solver import lm_solve
from lc_b200.nll.pnp_auto import weighted_pnp_jac_wrt_pts2d
from lc_b200 import _native as nat

for (B, N) in [(3, 8), (2, 70), (2, 700), (2, 1300), (2, 2500)]:   # 8: thread-per-pose kernel; 2500: tensor-memory variant of the loss kernel
    c = make_correspondences(B, N, 1).to(torch.float32).to(device="cuda")
    if N % 4 == 0:   # planar slabs: vectorised point loops, TMA-out gradients, mixed-precision LM pass
        X, x, w = planar_view(c.pts3d), planar_view(c.pts2d), planar_view(c.inv_std)
        loss_fwd_bwd(c.K, c.pose, X, x, w, c.valid, c.bbox_3d, want_cov=True)
        solve_and_loss(c.K, c.start, X, x, w, None, c.bbox_3d, need=(True, True, True))
        lm_solve(c.K, X, x, w, c.start, weight_mode=nat.W_INV_STD, n_points=torch.tensor([N, N - 3][:B] + [N] * (B - 2), dtype=torch.int32, device="cuda"))
    for stream in (False, True):
        loss_fwd_bwd(c.K, c.pose, planar_view(c.pts3d), c.pts2d, planar_view(c.inv_std), c.valid, c.bbox_3d, want_cov=True, force_streaming=stream)
        solve_and_loss(c.K, c.start, c.pts3d, c.pts2d, c.inv_std, None, c.bbox_3d, need=(True, True, True), force_streaming=stream)
        lm_solve(c.K, c.pts3d, c.pts2d, c.inv_std ** 2, c.start, weight_mode=nat.W_ICOV_DIAG, filter_input_nan=True, force_streaming=stream, want_trace=True)
    lm_solve(c.K, c.pts3d, c.pts2d, full_icov_from_inv_std(c.inv_std.cpu().double(), 0).float().cuda(), c.start, weight_mode=nat.W_ICOV_FULL)
    w = (c.inv_std ** 2).requires_grad_(True)
    j, cv = weighted_pnp_jac_wrt_pts2d(c.pts2d, c.pose, c.K, c.pts3d, w, with_cov=True)
    (j.sum() + cv.sum()).backward()

# rows f1-f4: dense / zebrapose producers, test-time decode, selection, initialiser, test-time chain, metrics, candidates
from lc_b200.synth import make_dense_outputs, make_zebra_outputs, quat_to_matrix
from lc_b200.dense import dense_loss_fwd_bwd
from lc_b200.floatbits import nn_out_to_xyz
from lc_b200.select import dense_point_select, solve_pnp_dense
from lc_b200.evaluate import compute_pose_errors
from lc_b200.symmetry import select_pose_2d, select_pose_3d

for (H, W, sample) in [(16, 20, 1), (40, 36, 3)]:
    d = {k: v.cuda() for k, v in make_dense_outputs(2, H, W, 3).items()}
    dense_loss_fwd_bwd(d["xyz_noc"], d["logits"], d["scale"], d["noc_scale"], d["K"], d["pose"], d["bbox_3d"], sample=sample, top_left=(0, 0))
    z = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in make_zebra_outputs(2, H, W, 4, (5, 4, 3)).items()}
    dense_loss_fwd_bwd(None, z["logits"], z["scale"], z["noc_scale"], z["K"], z["pose"], z["bbox_3d"], sample=sample, top_left=(0, 0),
                       noc_bin_logits=z["bin_logits"], noc_bin_raw=z["raw_bits"], msk_noc=z["msk_noc"], bit_cnt=(5, 4, 3),
                       model_transform=z["model_transform"])
    xyz = nn_out_to_xyz(z["bin_logits"], z["noc_scale"], model_transform=z["model_transform"], bit_cnt=(5, 4, 3))
    nn_out_to_xyz(z["bin_logits"][:, :, :H - 1, :W - 1], z["noc_scale"], bit_cnt=(5, 4, 3))          # scalar path
    ml = torch.randn(2, 1, H, W, device="cuda")
    for mode in ("mask", "quantile", "quantile_in_mask"):
        dense_point_select(xyz, ml, xyz_weight_logits=d["logits"], xyz_weights_scale=d["scale"], sample=sample, dense_point_select=mode, want_index=True)
    solve_pnp_dense(d["K"], d["xyz_noc"].permute(0, 2, 3, 1), ml + 3, d["logits"], d["scale"], None, noc_scale=d["noc_scale"], sample=sample,
                    solvers=("weighted", "weighted_filtered"))
c = make_correspondences(3, 300, 2)
Rg, Re = quat_to_matrix(c.pose[:, :4]).cuda(), quat_to_matrix(c.start[:, :4]).cuda()
compute_pose_errors(Re, c.start[:, 4:].cuda(), Rg, c.pose[:, 4:].cuda(), torch.randn(2500, 3, dtype=torch.float64, device="cuda") * 50)
c32 = c.to(torch.float32).to(device="cuda")
candi = torch.cat((Rg.float(), c32.pose[:, 4:, None]), -1)[:, None].repeat(1, 5, 1, 1)
select_pose_2d(c32.K, c32.pts3d, c32.pts2d, candi)
select_pose_3d(c32.K, c32.pts3d, (c32.pts3d @ Rg.float().mT + c32.pose[:, None, 4:]) @ c32.K.mT, candi)
# the opt-in persistent kernel (one CTA per SM, two poses in flight) and the reference-ABI entry point with host pointer tables
os.environ["LC_B200_PERSIST"] = "1"
c = make_correspondences(150, 2048, 3).to(torch.float32).to(device="cuda")
X, x, w = planar_view(c.pts3d), planar_view(c.pts2d), planar_view(c.inv_std)
loss_fwd_bwd(c.K, c.pose, X, x, w, None, c.bbox_3d)
solve_and_loss(c.K, c.start, X, x, w, None, c.bbox_3d, need=(True, True, True))
lm_solve(c.K, X, x, w, c.start, weight_mode=nat.W_INV_STD)
assert b"persist" in nat.lib().lc_b200_last_kernels()
del os.environ["LC_B200_PERSIST"]
# round 2 paths: poses split over two-CTA clusters (DSMEM reduction exchange, PDL-chained launches), the solve-only kernel with three
# poses per SM, the cov_2d variant of the loss
for env, val in (("LC_B200_SPLIT", "1"), ("LC_B200_LM3", "1")):
    os.environ[env] = val
    c = make_correspondences(5, 2048, 4).to(torch.float32).to(device="cuda")
    c2 = make_correspondences(3, 2504, 5).to(torch.float32).to(device="cuda")
    for cc in (c, c2):
        X, x, w = planar_view(cc.pts3d), planar_view(cc.pts2d), planar_view(cc.inv_std)
        npts = torch.tensor([cc.pts3d.shape[1], 900, 1500, 2000, 2048][: cc.pts3d.shape[0]], dtype=torch.int32, device="cuda")
        loss_fwd_bwd(cc.K, cc.pose, X, x, w, None, cc.bbox_3d, n_points=npts)
        solve_and_loss(cc.K, cc.start, X, x, w, None, cc.bbox_3d, need=(True, True, True), n_points=npts)
        lm_solve(cc.K, X, x, w, cc.start, npts, weight_mode=nat.W_INV_STD, filter_input_nan=True)
        lm_solve(cc.K, cc.pts3d, cc.pts2d, cc.inv_std, cc.start, weight_mode=nat.W_INV_STD)
    del os.environ[env]
c = make_correspondences(3, 200, 6).to(torch.float32).to(device="cuda")
loss_fwd_bwd(c.K, c.pose, c.pts3d, c.pts2d, c.inv_std, c.valid, c.bbox_3d, cov_2d=True)
torch.cuda.synchronize()
print("sanitize smoke done")
