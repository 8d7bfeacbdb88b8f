"""Synthetic 2D-3D correspondences for the LC hot path.

The reference ships no datasets or fixtures for this path, so every
parity test and the benchmark draw their inputs from this seeded generator.  It
is CPU-only, fp64 and deterministic for a given (B, N, seed); the caller casts
and moves the tensors.

Shapes follow the operator contract of the reference (``lib/cov_mixed.py:100``,
``lib/pnp/cer_solver.py:6``): ``K (B,3,3)``, ``pose (B,7)`` = unit quaternion
wxyz + translation, ``pts3d (B,N,3)``, ``pts2d (B,N,2)``, ``inv_std (B,N,2)``,
``bbox_3d (B,8,3)`` in the corner order of ``model_transform.py:6-18``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

_BBOX_SIGNS = torch.tensor(
    [[1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1],
     [-1, 1, 1], [-1, 1, -1], [-1, -1, 1], [-1, -1, -1]], dtype=torch.float64)


@dataclass
class Correspondences:
    """One synthetic batch; every tensor is fp64 on the CPU."""
    K: torch.Tensor          # (B,3,3) crop intrinsics with a rotated 2x2 block (dataset.py:421-423)
    pose: torch.Tensor       # (B,7) ground-truth pose, wxyz + t
    start: torch.Tensor      # (B,7) perturbed pose used as the LM start (EPnP-quality)
    pts3d: torch.Tensor      # (B,N,3) model-frame points
    pts2d: torch.Tensor      # (B,N,2) noisy measurements
    inv_std: torch.Tensor    # (B,N,2) predicted inverse std (loss half input)
    bbox_3d: torch.Tensor    # (B,8,3)
    valid: torch.Tensor      # (B,N) ones, as losses.py:366 passes

    def to(self, dtype=None, device=None) -> "Correspondences":
        f = lambda t: t.to(dtype=dtype, device=device)
        return Correspondences(*(f(getattr(self, k)) for k in self.__dataclass_fields__))


def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw), -1)


def quat_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """Rotation matrix of a *unit* wxyz quaternion."""
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    return torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)),
                       -1).reshape(q.shape[:-1] + (3, 3))


def make_correspondences(B: int, N: int, seed: int, *, outlier_frac: float = 0.05,
                         start_rot_sigma: float = 0.02, start_t_sigma: float = 0.01) -> Correspondences:
    g = torch.Generator().manual_seed(int(seed))
    f64 = torch.float64
    randn = lambda *s: torch.randn(*s, generator=g, dtype=f64)
    rand = lambda *s: torch.rand(*s, generator=g, dtype=f64)

    q = randn(B, 4)
    q = q / q.norm(dim=-1, keepdim=True)
    q = torch.where(q[:, :1] < 0, -q, q)
    z = 400 + 1100 * rand(B)
    t = torch.stack((0.05 * z * randn(B), 0.05 * z * randn(B), z), -1)
    pose = torch.cat((q, t), -1)

    h = torch.tensor([40.0, 50.0, 60.0], dtype=f64)
    bbox = (_BBOX_SIGNS * h).expand(B, 8, 3).clone()
    X = (2 * rand(B, N, 3) - 1) * h

    # crop intrinsics: rotated 2x2 block (in-plane augmentation), object ~43 px wide in a 64 px crop
    th = 2 * math.pi * rand(B)
    f = 42.7 * z / 120.0
    c, s = torch.cos(th), torch.sin(th)
    K = torch.zeros(B, 3, 3, dtype=f64)
    K[:, 0, 0], K[:, 0, 1] = f * c, -f * s
    K[:, 1, 0], K[:, 1, 1] = f * s, f * c
    K[:, 2, 2] = 1
    centre = 32 + 16 * (rand(B, 2) - 0.5)
    tn = t[:, :2] / t[:, 2:]
    K[:, :2, 2] = centre - torch.einsum('bij,bj->bi', K[:, :2, :2], tn)

    R = quat_to_matrix(q)
    P = X @ R.mT + t[:, None, :]
    KP = P @ K.mT
    proj = KP[..., :2] / KP[..., 2:]

    sigma = 0.5 + 1.5 * rand(B, N, 1)
    x = proj + sigma * randn(B, N, 2)
    outl = rand(B, N, 1) < outlier_frac
    x = x + outl * (40 * rand(B, N, 2) - 20)
    inv_std = (1.0 / sigma) * (0.8 + 0.45 * rand(B, N, 2))

    # LM start: truth composed with a small rotation, translation scaled by (1+eps)
    aa = start_rot_sigma * randn(B, 3)
    ang = aa.norm(dim=-1, keepdim=True).clamp_min(1e-30)
    dq = torch.cat((torch.cos(ang / 2), aa / ang * torch.sin(ang / 2)), -1)
    q0 = _quat_mul(q, dq)
    t0 = t * (1 + start_t_sigma * randn(B, 3))
    start = torch.cat((q0, t0), -1)

    return Correspondences(K=K, pose=pose, start=start, pts3d=X, pts2d=x, inv_std=inv_std,
                           bbox_3d=bbox, valid=torch.ones(B, N, dtype=f64))


def planar_view(t: torch.Tensor) -> torch.Tensor:
    """Return the same values as a (B,N,C) view over (B,C,N) storage — the strides
    the dense call site produces (``losses.py:142-161``: ``flatten(-2).mT``)."""
    return t.transpose(-1, -2).contiguous().transpose(-1, -2)


def full_icov_from_inv_std(inv_std: torch.Tensor, seed: int) -> torch.Tensor:
    """(B,N,2,2) SPD inverse covariances: diag(inv_std^2) rotated by a random angle."""
    g = torch.Generator().manual_seed(int(seed) + 7919)
    th = math.pi * torch.rand(inv_std.shape[:-1], generator=g, dtype=inv_std.dtype)
    c, s = torch.cos(th), torch.sin(th)
    Rm = torch.stack((c, -s, s, c), -1).reshape(inv_std.shape[:-1] + (2, 2))
    return Rm @ torch.diag_embed(inv_std ** 2) @ Rm.mT


def make_dense_outputs(B: int, H: int, W: int, seed: int):
    """Network-output-shaped tensors for the dense path (``losses.py:336-386``): ``xyz_noc (B,3,H,W)`` whose
    back-projection roughly matches the pixel grid (plus noise), weight logits ``(B,2,H,W)``, ``xyz_weights_scale
    (B,1,1,1)``, ``noc_scale (B,3)`` and K / pose / bbox_3d.  fp32 on the CPU."""
    g = torch.Generator().manual_seed(int(seed))
    c = make_correspondences(B, 4, seed)
    K = c.K.clone()
    K[:, :2, :] *= W / 64.0                                  # crop of W pixels instead of 64
    R = quat_to_matrix(c.pose[:, :4])
    f64 = torch.float64
    ys, xs = torch.meshgrid(torch.arange(H, dtype=f64), torch.arange(W, dtype=f64), indexing="ij")
    pix = torch.stack((xs, ys, torch.ones_like(xs)), -1).reshape(1, H * W, 3).expand(B, -1, -1)
    zc = c.pose[:, None, 6:7] + 40 * (2 * torch.rand(B, H * W, 1, generator=g, dtype=f64) - 1)
    X = (torch.linalg.solve(K, pix.mT).mT * zc - c.pose[:, None, 4:]) @ R + torch.randn(B, H * W, 3, generator=g, dtype=f64)
    ns = torch.tensor([[40.0, 50.0, 60.0]], dtype=f64).expand(B, 3)
    return dict(xyz_noc=(X / ns[:, None, :]).mT.reshape(B, 3, H, W).float(), logits=torch.randn(B, 2, H, W, generator=g),
                scale=(2.0 * H * W) * torch.exp(0.2 * torch.randn(B, 1, 1, 1, generator=g)), noc_scale=ns.float(), K=K.float(),
                pose=c.pose.float(), bbox_3d=c.bbox_3d.float())


def make_zebra_outputs(B: int, H: int, W: int, seed: int, bit_cnt=(7, 6, 5), *, with_transform: bool = True,
                       flip_prob: float = 0.08, mask_prob: float = 0.8, black_background: bool = True):
    """ZebraPose-shaped network outputs for the dense path (``losses.py:163-184``): Gray-coded bit logits
    ``bin_logits (B,sum(bit_cnt),H,W)`` that mostly agree with the ground-truth code of a surface consistent with the
    pose (a fraction ``flip_prob`` of the bits is wrong), the GT raw bits ``raw_bits`` (bool, channel-last storage like
    ``floatbits.nn_noc2target`` returns), ``msk_noc (B,H,W)`` bool, a model transform ``(B,4,4)`` and the tensors of
    ``make_dense_outputs``.  The bit coding restates ``floatbits.py:76-97``."""
    d = make_dense_outputs(B, H, W, seed)
    g = torch.Generator().manual_seed(int(seed) + 4242)
    f64 = torch.float64
    ns = d["noc_scale"].to(f64)
    X = (d["xyz_noc"].to(f64) * ns[:, :, None, None]).permute(0, 2, 3, 1)          # (B,H,W,3) model frame
    if with_transform:
        q = torch.randn(B, 4, generator=g, dtype=f64)
        M = quat_to_matrix(q)
        tt = 3.0 * torch.randn(B, 3, generator=g, dtype=f64)
        T = torch.eye(4, dtype=f64).repeat(B, 1, 1)
        T[:, :3, :3], T[:, :3, 3] = M, tt
        Xf = X @ M.mT[:, None] + tt[:, None, None, :]                              # losses.py:56 xyz @ T[:3,:3]^T + T[:3,3]
        ns_x = Xf.abs().amax(dim=(1, 2)) * 1.02
    else:
        T, Xf, ns_x = None, X, ns * 1.5
    noc = (Xf / ns_x[:, None, None, :]).clamp(-0.999, 0.999)
    bit_cnt = [int(b) for b in bit_cnt]
    mods, raws = [], []
    for a, N in enumerate(bit_cnt):
        mx = 2 ** N - 1
        ints = torch.clamp((noc[..., a] + 1) * (mx * 0.5), 0, mx).round().to(torch.int64)
        raw = ((ints[..., None] >> torch.arange(N - 1, -1, -1)) & 1).bool()
        mod = raw.clone()
        mod[..., 1:] ^= raw[..., :-1]
        if black_background:
            mod[..., 0:2] = ~mod[..., 0:2]
        mods.append(mod)
        raws.append(raw)
    mod, raw = torch.cat(mods, -1), torch.cat(raws, -1)                             # (B,H,W,C)
    mag = 0.2 + 2.5 * torch.rand(mod.shape, generator=g, dtype=f64)
    # wrong bits are rarer towards the MSB (a half-trained network gets the coarse code right first)
    pos = torch.cat([torch.arange(1, N + 1, dtype=f64) / N for N in bit_cnt])
    flip = torch.rand(mod.shape, generator=g, dtype=f64) < flip_prob * pos * pos
    logits = (mod ^ flip).to(f64).mul(2).sub(1) * mag
    msk = torch.rand(B, H, W, generator=g) < mask_prob
    out = dict(d)
    out.pop("xyz_noc")
    out.update(bin_logits=logits.permute(0, 3, 1, 2).contiguous().float(), raw_bits=raw.permute(0, 3, 1, 2), msk_noc=msk,
               noc_scale=ns_x.float(), model_transform=None if T is None else T.float(), bit_cnt=bit_cnt)
    return out
