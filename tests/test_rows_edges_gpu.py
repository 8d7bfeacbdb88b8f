"""GPU edge cases of the rows around the hot path (f1-f4): empty batches, argument errors that must be raised (not silently
handled), too-large maps, and the LC op inside a real training step (BASELINE.json configs[2]: CNN backbone with random init,
256x256 synthetic crops, B=32, gradient-clip hooks, optimizer step)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_empty_batches_are_noops_in_the_new_entry_points():
    from lc_b200.dense import dense_loss_fwd_bwd
    from lc_b200.floatbits import nn_out_to_xyz
    from lc_b200.select import dense_point_select
    from lc_b200.pnp import init_solver
    from lc_b200.evaluate import compute_pose_errors
    z = lambda *s, dtype=torch.float32: torch.zeros(*s, device="cuda", dtype=dtype)
    o = dense_loss_fwd_bwd(z(0, 3, 8, 8), z(0, 2, 8, 8), z(0, 1, 1, 1), z(0, 3), z(0, 3, 3), z(0, 7), z(0, 8, 3), sample=2, top_left=(0, 0))
    assert o["loss"].shape == (0,) and o["g_logits"].shape == (0, 2, 8, 8)
    assert nn_out_to_xyz(z(0, 9, 8, 8), z(0, 3), bit_cnt=3).shape == (0, 8, 8, 3)
    s = dense_point_select(z(0, 8, 8, 3), z(0, 1, 8, 8), xyz_weights=z(0, 2, 8, 8), sample=1)
    assert s["pts3d"].shape == (0, 64, 3) and s["n_points"].shape == (0,)
    inv, st, inl = init_solver.solve(z(0, 3, 3), z(0, 16, 3), z(0, 16, 2))
    assert st.shape == (0, 7) and inl["mask"].shape == (0, 16)
    e = compute_pose_errors(z(0, 3, 3, dtype=torch.float64), z(0, 3, dtype=torch.float64), z(0, 3, 3, dtype=torch.float64),
                            z(0, 3, dtype=torch.float64), z(10, 3, dtype=torch.float64))
    assert e["add"].shape == (0,)


def test_bad_arguments_raise():
    from lc_b200.dense import dense_loss_fwd_bwd
    from lc_b200.floatbits import nn_out_to_xyz
    from lc_b200.select import dense_point_select
    from lc_b200.synth import make_zebra_outputs, make_dense_outputs
    d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in make_zebra_outputs(2, 16, 16, 1, (5, 4, 3)).items()}
    kw = dict(sample=2, top_left=(0, 0), noc_bin_logits=d["bin_logits"], noc_bin_raw=d["raw_bits"], msk_noc=d["msk_noc"])
    args = (None, d["logits"], d["scale"], d["noc_scale"], d["K"], d["pose"], d["bbox_3d"])
    with pytest.raises(ValueError):
        dense_loss_fwd_bwd(*args, bit_cnt=(5, 5, 5), **kw)                      # channel count mismatch
    with pytest.raises(ValueError):
        dense_loss_fwd_bwd(d["bin_logits"][:, :3], *args[1:], bit_cnt=(5, 4, 3), **kw)   # both producers given
    with pytest.raises(RuntimeError):
        dense_loss_fwd_bwd(*args, bit_cnt=(5, 4, 3), **{**kw, "top_left": (16, 0)})   # offset outside the map (C ABI check)
    with pytest.raises(ValueError):
        nn_out_to_xyz(d["bin_logits"], d["noc_scale"], bit_cnt=(7, 7, 7))
    with pytest.raises(ValueError):
        dense_point_select(torch.zeros(1, 8, 8, 3, device="cuda"), torch.zeros(1, 1, 8, 8, device="cuda"),
                           xyz_weights=torch.ones(1, 2, 8, 8, device="cuda"), dense_point_select="median")
    g = make_dense_outputs(1, 256, 256, 2)
    with pytest.raises(RuntimeError, match="shared memory"):                     # 65536 sampled points: more than one SM holds
        dense_point_select(g["xyz_noc"].cuda().permute(0, 2, 3, 1), torch.zeros(1, 1, 256, 256, device="cuda"),
                           xyz_weight_logits=g["logits"].cuda(), xyz_weights_scale=g["scale"].cuda(), sample=1)


def test_lc_op_inside_a_training_step():
    """configs/glmo.yaml-shaped step: CNN (ResNet-34 trunk, random init) -> 64x64 head split like ptnet.py:54-82 ->
    fused dense LC loss with gradient hooks on the weight logits (losses.py:343-352) -> backward -> optimizer step.
    Checks the drop-in behaviour in a real autograd graph: finite gradients everywhere, hooks fire once per step with the
    gradient of the right shape, identical loss and parameter gradients to the unfused pipeline built from torch ops
    around Loss_cov_mixed, and a loss that goes down when the step is repeated on the same batch."""
    import torchvision
    from lc_b200.cov_mixed import Loss_cov_mixed
    from lc_b200.dense import dense_pose_loss
    from lc_b200.synth import make_dense_outputs
    torch.manual_seed(0)
    B, S = 32, 64
    trunk = torchvision.models.resnet34(weights=None)
    net = torch.nn.Sequential(trunk.conv1, trunk.bn1, trunk.relu, trunk.layer1, trunk.layer2,       # 256 -> 64 (stride 4)
                              torch.nn.Conv2d(128, 5, 3, padding=1)).cuda()
    scale_layer = torch.nn.Linear(128, 1).cuda()
    d = {k: v.cuda() for k, v in make_dense_outputs(B, S, S, 3).items()}
    img = torch.rand(B, 3, 256, 256, device="cuda")
    opt = torch.optim.Adam(list(net.parameters()) + list(scale_layer.parameters()), lr=1e-4)
    fired = []

    def forward(fused):
        feat = net[:-1](img)
        out = net[-1](feat)
        # the synthetic "ground truth" geometry plus a small learnable correction keeps the loss in its working regime
        xyz_noc = d["xyz_noc"] + 0.01 * torch.tanh(out[:, :3])
        logits = out[:, 3:5]
        w_scale = (2.0 * S * S) * torch.exp(0.1 * torch.tanh(scale_layer(feat.mean((2, 3))))).reshape(B, 1, 1, 1)
        logits.register_hook(lambda g: fired.append(tuple(g.shape)) or g.clamp(-1e3, 1e3))
        if fused:
            return dense_pose_loss(xyz_noc, logits, w_scale, d["noc_scale"], d["K"], d["pose"], d["bbox_3d"], dense_sample=2,
                                   top_left=(1, 0)).mean()
        w = logits.reshape(B, 1, -1).softmax(-1).reshape_as(logits) * w_scale
        ys, xs = torch.meshgrid(torch.arange(S, device="cuda", dtype=torch.float32), torch.arange(S, device="cuda", dtype=torch.float32), indexing="ij")
        uv = torch.stack((xs, ys), -1)[1::2, 0::2].reshape(-1, 2)
        inv_std = w[..., 1::2, 0::2].flatten(-2).mT
        p3 = xyz_noc[..., 1::2, 0::2].flatten(-2).mT * d["noc_scale"].unsqueeze(-2)
        return Loss_cov_mixed(d["K"], d["pose"], p3, uv.expand(B, -1, -1), inv_std, torch.ones_like(p3[..., 0]), bbox_3d=d["bbox_3d"],
                              max_err_len=32).mean()

    net.eval()          # frozen batch-norm statistics: the two pipelines must see the same network function
    grads = []
    for fused in (True, False):
        opt.zero_grad(set_to_none=True)
        loss = forward(fused)
        loss.backward()
        grads.append((loss.item(), [p.grad.clone() for p in net.parameters() if p.grad is not None], scale_layer.weight.grad.clone()))
    assert fired == [(B, 2, S, S)] * 2
    assert abs(grads[0][0] - grads[1][0]) <= 2e-6 * max(1.0, abs(grads[1][0]))
    num = sum(((a - b) ** 2).sum() for a, b in zip(grads[0][1], grads[1][1])).sqrt()
    den = sum((b ** 2).sum() for b in grads[1][1]).sqrt()
    assert torch.isfinite(den) and den > 0 and (num / den).item() <= 1e-4
    assert torch.allclose(grads[0][2], grads[1][2], rtol=1e-3, atol=1e-6 * grads[1][2].abs().max().item())
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss = forward(True)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_operators_are_cuda_graph_capturable():
    """The C ABI never allocates or synchronises, so a launch-bound inner loop (small batches) can be captured in a CUDA graph:
    capture loss_fwd_bwd + the test-time chain once, replay on new inputs written into the static buffers."""
    from lc_b200.cov_mixed import loss_fwd_bwd
    from lc_b200.select import solve_pnp_dense
    from lc_b200.synth import make_correspondences, make_dense_outputs
    c = make_correspondences(4, 256, 0).to(torch.float32).to(device="cuda")
    c2 = make_correspondences(4, 256, 1).to(torch.float32).to(device="cuda")
    d = {k: v.cuda() for k, v in make_dense_outputs(4, 32, 32, 3).items()}
    ml = torch.full((4, 1, 32, 32), 3.0, device="cuda")
    static = dict(K=c.K.clone(), pose=c.pose.clone(), X=c.pts3d.clone(), x=c.pts2d.clone(), s=c.inv_std.clone(), bb=c.bbox_3d.clone())
    run = lambda: (loss_fwd_bwd(static["K"], static["pose"], static["X"], static["x"], static["s"], None, static["bb"]),
                   solve_pnp_dense(d["K"], d["xyz_noc"].permute(0, 2, 3, 1), ml, d["logits"], d["scale"], None, noc_scale=d["noc_scale"], sample=1))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()                                           # warm-up outside the capture (library load, smem opt-in attributes)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out, (res, sel) = run()
    for src in (c, c2):
        for k, v in zip(("K", "pose", "X", "x", "s", "bb"), (src.K, src.pose, src.pts3d, src.pts2d, src.inv_std, src.bbox_3d)):
            static[k].copy_(v)
        g.replay()
        torch.cuda.synchronize()
        eager = loss_fwd_bwd(src.K, src.pose, src.pts3d, src.pts2d, src.inv_std, None, src.bbox_3d)
        assert torch.equal(out["loss"], eager["loss"]) and torch.equal(out["g_pts3d"], eager["g_pts3d"])
    eager_res, _ = solve_pnp_dense(d["K"], d["xyz_noc"].permute(0, 2, 3, 1), ml, d["logits"], d["scale"], None, noc_scale=d["noc_scale"], sample=1)
    assert torch.equal(res["weighted"], eager_res["weighted"])
