#!/usr/bin/env python
"""The reference's REAL train step (BASELINE.json configs[2] / configs[3]) on synthetic blobs, with the LC op swapped.

Runs `ptnet` + `losses.annots_on_the_fly` + `Loss_fn.forward` + backward + optimizer step exactly as train.py:52-67 does, from
the reference files staged under baseline/_ref (tools/stage_reference.py; they are used byte for byte), with
  arm "reference"  lib.cov_mixed.Loss_cov_mixed (functorch graph) on the same B200,
  arm "ours"       lc_b200.cov_mixed.Loss_cov_mixed swapped in for that one name (the import-line swap of INTEGRATION.md §1),
  arm "fused"      additionally lc_b200.dense.dense_pose_loss* in place of the producer glue of Loss_fn.dense_pose_loss (row f1/f3),
  arm "nopose"     w_loss_pose = 0 (the step without the LC loss: the LC op's share = 1 - t_nopose / t_arm).
  arm "nolc"       the LC op replaced by a zero-valued differentiable stand-in, glue and hooks alive (what DDP runs use: with
                   w_loss_pose = 0 DDP hands None to the reference's tensor hooks).
Random-init networks (the pretrained resnet34 file is replaced by a random state dict in memory), synthetic blobs with the keys
of dataset.py:451-473 built from a consistent pose / camera / surface so that every loss term is finite and active.

    python tools/train_step.py --config glmo|zycbv [--batch 32] [--steps 20] [--arms reference,ours,fused,nopose] [--out f.json]
    torchrun --nproc-per-node N tools/train_step.py --config zycbv --ddp ...     (batch per GPU; DDP over NCCL)
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path[:0] = [ROOT, REF, os.path.join(ROOT, "tools", "ref_shims")]


def load_reference():
    if not os.path.exists(os.path.join(REF, "losses.py")):
        raise SystemExit("baseline/_ref is not staged: run `python tools/stage_reference.py` in the build container")
    import torchvision
    real_load = torch.load

    def fake_load(path, *a, **k):   # cdpn_resnet.py:200 / zebra_resnet.py:150 read assets/resnet34-333f7ec4.pth: random init instead
        if isinstance(path, str) and path.endswith("resnet34-333f7ec4.pth"):
            return torchvision.models.resnet34(weights=None).state_dict()
        return real_load(path, *a, **k)
    torch.load = fake_load
    import ptnet, losses, floatbits  # noqa: E401  (reference modules)
    from mmcv import Config
    return ptnet, losses, floatbits, Config, (lambda: setattr(torch, "load", real_load))


def make_blob(cfg, B, dev, seed, floatbits):
    """Synthetic training blob (dataset.py:451-473): a planar-ish surface patch seen under a random pose fills the crop."""
    from lc_b200.synth import make_correspondences, quat_to_matrix
    W, H = cfg.train_dataset.get("net_output_wh", cfg.get("net_output_wh", [64, 64]))
    Wi, Hi = cfg.train_dataset.get("net_input_wh", cfg.get("net_input_wh", [256, 256]))
    g = torch.Generator().manual_seed(seed)
    c = make_correspondences(B, 4, seed)
    K = c.K.clone()
    K[:, :2, :] *= W / 64.0
    R = quat_to_matrix(c.pose[:, :4])
    t = c.pose[:, 4:]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    z = t[:, None, None, 2:] + 30 * (torch.rand(B, H, W, 1, generator=g, dtype=torch.float64) - 0.5)
    homo_z = torch.cat((xs[None, :, :, None] * z, ys[None, :, :, None] * z, z), -1)          # (B,H,W,3) = K (R X + t)
    h = torch.tensor([[40.0, 50.0, 60.0]], dtype=torch.float64).expand(B, 3)
    blob = {
        "rgb_in": torch.rand(B, 3, Hi, Wi, generator=g),
        "noc_scale": (h * 4).float(), "noc_scale_ori": h.float(), "out_pix_scale": torch.ones(B),
        "msk_vis": torch.ones(B, H, W), "msk_noc": torch.ones(B, H, W, dtype=torch.bool),
        "homo_z_out": homo_z.float(), "K_no_aug": K.float(), "R_no_aug": R.float(), "t_no_aug": t.float(),
        "Rt_candi": [torch.cat((R, t[..., None]), -1).float()[:, None]], "bbox_3d": c.bbox_3d.float(), "out_K": K.float(),
    }
    bit_cnt = None
    if cfg.get("max_bit_cnt", 0) > 0:
        floatbits.set_black_background(cfg.get("black_background", False))
        bit_cnt = floatbits.calc_bit_count(blob["noc_scale"][0].tolist(), max_bits=cfg.max_bit_cnt)
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, :3, :3] = quat_to_matrix(torch.randn(B, 4, generator=g, dtype=torch.float64)).float()
        T[:, :3, 3] = 3 * torch.randn(B, 3, generator=g)
        blob["model_transform"] = T
    out = {k: ([x.to(dev) for x in v] if isinstance(v, list) else v.to(dev)) for k, v in blob.items()}
    if bit_cnt is not None:
        out["bit_cnt"] = bit_cnt
    return out, bit_cnt


def fused_dense_pose_loss(self, cfg, gt_dict, out_dict):
    """Loss_fn.dense_pose_loss (losses.py:336-386) with the producer glue + LC loss in ONE launch (rows f1 / f3)."""
    from lc_b200 import dense
    sample = cfg.get("dense_sample", 2)
    top, left = np.random.randint(0, sample, size=2)
    logits = out_dict["xyz_weight_logits"]
    if self.weight_grad_clipper is not None:
        logits.register_hook(lambda g: self.weight_grad_clipper.clip(g))
    kw = dict(dense_sample=sample, top_left=(int(top), int(left)), max_err_len=cfg.get("max_err_len", 32))
    if "xyz_noc" in out_dict:
        loss = dense.dense_pose_loss(out_dict["xyz_noc"], logits, out_dict["xyz_weights_scale"], gt_dict["noc_scale"], gt_dict["out_K"],
                                     gt_dict["pose_best"], gt_dict["bbox_3d"], **kw)
    else:
        import floatbits as ref_floatbits
        loss = dense.dense_pose_loss_noc_bin(out_dict["xyz_noc_bin"], gt_dict["xyz_noc_bin_raw"], logits, out_dict["xyz_weights_scale"],
                                             gt_dict["msk_noc"], gt_dict["noc_scale"], gt_dict["out_K"], gt_dict["pose_best"],
                                             gt_dict["bbox_3d"], bit_cnt=gt_dict["bit_cnt"], model_transform=gt_dict.get("model_transform"),
                                             black_background=ref_floatbits._black_background, **kw)
    return loss.mean()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="glmo")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--arms", default="reference,ours,fused,nopose")
    ap.add_argument("--ddp", action="store_true")
    ap.add_argument("--no-probe64", action="store_true", help="skip the float64 run of the reference arm used as the yardstick")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    def tf32(on):   # PyTorch's default (TF32 convolutions) for the timed steps; plain fp32 for the agreement probes
        torch.backends.cudnn.allow_tf32 = on
    if a.ddp and world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    ptnet, losses, floatbits, Config, restore = load_reference()
    import torchvision.transforms as transforms
    from lc_b200.cov_mixed import Loss_cov_mixed as ours_lc
    ref_lc = losses.Loss_cov_mixed
    ref_dense = losses.Loss_fn.dense_pose_loss
    cfg = Config.fromfile(os.path.join(REF, "configs", a.config + ".yaml"))
    blob, bit_cnt = make_blob(cfg, a.batch, dev, 1234 + rank, floatbits)
    total_bits = 0 if bit_cnt is None else sum(bit_cnt)
    normalize = transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    EPOCH, STEP, SPE = 10, 10 ** 6, 1000          # far past pose_loss_start_*: the LC loss is fully applied (losses.py:296-308)

    def build():
        torch.manual_seed(7)
        model = ptnet.ptnet(cfg.model, cfg, total_bit_cnt=total_bits).to(dev)
        model.loss_fn = losses.Loss_fn(cfg.loss, cfg, total_bits).to(dev)
        net = model
        if a.ddp and world > 1:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
        opt = torch.optim.Adam(model.parameters(), lr=cfg.optimizer.lr, weight_decay=cfg.optimizer.wd)   # Ranger for glmo in the reference; Adam here for both
        return model, net, opt

    def one_step(model, net, opt, probe=None):
        nonlocal blob
        gt = dict(blob)
        out = net(normalize(gt["rgb_in"]))
        losses.annots_on_the_fly(gt, out, cfg, STEP)
        if probe is not None:   # gradients arriving at the network outputs (what the LC op and its glue hand back), before any hook
            probe["out_grads"] = {}
            for k, v in out.items():
                if v.requires_grad:
                    v.register_hook(lambda g, k=k: probe["out_grads"].__setitem__(k, g.detach().clone()) if g is not None else None)
        loss_dict, w = model.loss_fn(gt, out, EPOCH, STEP, SPE)
        loss = sum(w.values())
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if probe is not None:
            probe.update(loss_pose=float(loss_dict["loss_pose"]), loss=float(loss),
                         grads={n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})
        opt.step()
        return loss

    def nolc(K, pose, pts3d, pts2d, inv_std, valid=None, **kw):
        # arm "nolc": the step with the LC op itself removed but its glue and hooks alive (zero loss, zero gradients).  Under DDP the
        # "nopose" arm (w_loss_pose = 0) leaves head outputs unused and the reference's tensor hooks then receive None.
        return (pts3d.sum((1, 2)) + inv_std.sum((1, 2))) * 0

    def set_arm(arm):
        losses.Loss_cov_mixed = ref_lc if arm == "reference" else (nolc if arm == "nolc" else ours_lc)      # the one-name swap
        losses.Loss_fn.dense_pose_loss = fused_dense_pose_loss if arm == "fused" else ref_dense
        cfg.loss.w_loss_pose = 0 if arm == "nopose" else w_pose

    w_pose = cfg.loss.w_loss_pose
    res, probes = {}, {}
    for arm in a.arms.split(","):
        set_arm(arm)
        model, net, opt = build()
        np.random.seed(0)
        probes[arm] = {}
        tf32(False)
        one_step(model, net, opt, probes[arm])                 # first step from identical weights / offsets: agreement probe
        tf32(True)
        for _ in range(a.warmup):
            one_step(model, net, opt)
        torch.cuda.synchronize()
        if a.ddp and world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            one_step(model, net, opt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        if a.ddp and world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        res[arm] = dict(ms_per_step=ms, samples_per_s=world * a.batch / (ms * 1e-3), loss_pose_first_step=probes[arm].get("loss_pose"))
        del model, net, opt
        torch.cuda.empty_cache()
    # yardstick: the reference arm in float64 (model, blob and loss in double) from the same weights and offsets; the fp32 arms are
    # then compared with IT, so that the reference's own fp32 rounding is not charged to the replacement
    if not a.no_probe64 and rank == 0:
        set_arm("reference")
        model, net, opt = build()
        model.double()
        blob32 = blob
        dbl = lambda x: x.double() if torch.is_tensor(x) and x.is_floating_point() else x
        blob = {k: ([dbl(x) for x in v] if isinstance(v, list) else dbl(v)) for k, v in blob32.items()}
        np.random.seed(0)
        p64 = {}
        tf32(False)
        one_step(model, model, opt, p64)
        blob = blob32
        den = math.sqrt(sum(float(g.pow(2).sum()) for g in p64["grads"].values()))
        res["vs_float64_reference"] = {arm: dict(loss_pose_rel=abs(pr["loss_pose"] - p64["loss_pose"]) / abs(p64["loss_pose"]),
                                                 grad_rel_l2_all_parameters=math.sqrt(sum(float((pr["grads"][k].double() - p64["grads"][k]).pow(2).sum())
                                                                                       for k in p64["grads"])) / den)
                                       for arm, pr in probes.items() if arm not in ("nopose", "nolc") and "grads" in pr}
        del model, net, opt
        torch.cuda.empty_cache()
    restore()
    if "reference" in probes and "ours" in probes:
        pr, po = probes["reference"], probes["ours"]
        num = math.sqrt(sum(float((pr["grads"][k] - po["grads"][k]).double().pow(2).sum()) for k in pr["grads"]))
        den = math.sqrt(sum(float(pr["grads"][k].double().pow(2).sum()) for k in pr["grads"]))
        og = {k: float((pr["out_grads"][k] - po["out_grads"][k]).double().norm() / pr["out_grads"][k].double().norm())
              for k in pr.get("out_grads", {}) if k in po.get("out_grads", {})}
        res["agreement_ours_vs_reference"] = dict(loss_pose_rel=abs(pr["loss_pose"] - po["loss_pose"]) / abs(pr["loss_pose"]),
                                                  grad_rel_l2_at_network_outputs=og, grad_rel_l2_all_parameters=num / den,
                                                  note="same weights, inputs and sub-sampling offsets; the parameter gradients pass through the same "
                                                       "cuDNN backward in both arms, which amplifies the 1e-6-level differences of the op's input gradients")
    for base, key in (("nopose", "lc_share_of_step"), ("nolc", "lc_op_share_of_step")):
        if base in res:
            for arm in ("reference", "ours", "fused"):
                if arm in res:
                    res[arm][key] = 1 - res[base]["ms_per_step"] / res[arm]["ms_per_step"]
    H, W = blob["msk_vis"].shape[-2:]
    sample = cfg.loss.pose_loss_cfg.get("dense_sample", 2)
    line = dict(config=a.config, batch_per_gpu=a.batch, n_gpus=world, ddp=bool(a.ddp and world > 1), steps=a.steps, out_hw=[H, W], dense_sample=sample,
                points_per_pose=math.ceil(H / sample) * math.ceil(W / sample), bit_cnt=bit_cnt, gpu=torch.cuda.get_device_name(dev), arms=res,
                note="reference files from baseline/_ref used unmodified; random-init networks; synthetic blobs; Adam in every arm; timed steps with "
                     "PyTorch's default TF32 convolutions, agreement probes with plain fp32")
    if rank == 0:
        print(json.dumps(line))
        if a.out:
            with open(a.out, "a") as fh:
                fh.write(json.dumps(line) + "\n")
    if a.ddp and world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
