import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from conftest import load_golden, rel_err
from lc_b200.cov_mixed import loss_fwd_bwd
for name in ["b3_n333_s1_init.npz", "b2_n1024_s0.npz", "b1_n4096_s2.npz"]:
    g = load_golden(name)
    for stream in (False, True):
        dt = torch.float32
        c = lambda x: None if x is None else torch.as_tensor(x).to(device="cuda", dtype=dt)
        o = loss_fwd_bwd(c(g["in_K"]), c(g["in_pose"]), c(g["in_pts3d"]), c(g["in_pts2d"]), c(g["in_inv_std"]), c(g["valid"]), c(g["in_bbox_3d"]), want_cov=True, force_streaming=stream)
        print(name, "stream" if stream else "resident", "loss", np.abs(o["loss"].cpu().double().numpy()-g["ref_loss"]).max()/np.abs(g["ref_loss"]).max(),
              "g3", rel_err(o["g_pts3d"].cpu().numpy(), g["ref_g_pts3d"]), "g2", rel_err(o["g_pts2d"].cpu().numpy(), g["ref_g_pts2d"]),
              "gs", rel_err(o["g_inv_std"].cpu().numpy(), g["ref_g_inv_std"]), "cov", rel_err(o["cov"].cpu().numpy(), g["ref_cov"]))
