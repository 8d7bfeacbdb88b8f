import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from test_dense_gpu import _inputs
from oracle import cpu_oracle as O
from lc_b200.dense import dense_loss_fwd_bwd
for (B,H,W,sample,tl) in [(4,64,64,1,(0,0)), (2,128,128,3,(0,2)), (2,32,32,2,(1,1))]:
    d = _inputs(B,H,W,100+H+sample)
    ref = O.dense_pose_loss(d["xyz_noc"], d["logits"], d["scale"], d["noc_scale"], d["K"], d["pose"], d["bbox_3d"], sample, tl)
    o = dense_loss_fwd_bwd(*(d[k].cuda() for k in ("xyz_noc","logits","scale","noc_scale","K","pose","bbox_3d")), sample=sample, top_left=tl)
    gs = o["g_scale"].cpu().numpy().astype(np.float64)
    # conditioning of S: sum |gbar_k p_k| / |S|
    print(H,W,sample,"g_scale gpu",gs,"ref",ref["g_scale"],"abs err",np.abs(gs-ref["g_scale"]))
    gl_ref = ref["g_logits"]; 
    print("   g_logits rel err", np.linalg.norm(o["g_logits"].cpu().numpy().reshape(B,-1)-gl_ref.reshape(B,-1),axis=1)/np.linalg.norm(gl_ref.reshape(B,-1),axis=1))
