"""ZebraPose binary-code decode at test time — SURVEY.md §8 row f3.

Reference: ``losses.nn_out_to_xyz(..., inference=True)`` (``losses.py:16-45``) and ``floatbits.nn_logits2noc`` without
the nearest-neighbour LUT (``floatbits.py:33-47, 197-224``): per axis the hard Gray bits (leading two inverted under a
black background) are turned into a binary integer whose LSB is replaced by ``sigmoid(l_last * (1 - (val & 2)))``;
``noc = val / (max_val / 2) - 1``; ``xyz = (noc * noc_scale - T[:3,3]) @ T[:3,:3]``.  One elementwise sm_100a kernel
(``lc_b200_noc_bin_decode``); the training-time decode with ground truth is fused into
``lc_b200.dense.dense_pose_loss_noc_bin``.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import torch
from torch import Tensor

from . import _native as nat

_black_background = True


def set_black_background(black: bool = True) -> None:
    """``floatbits.set_black_background`` (``floatbits.py:9-11``)."""
    global _black_background
    _black_background = bool(black)


def _bits(bit_cnt: Union[int, Sequence[int]], channels: int):
    bits = [int(bit_cnt)] * 3 if isinstance(bit_cnt, int) else [int(v) for v in bit_cnt]
    if len(bits) != 3 or sum(bits) != channels:
        raise ValueError(f"bit_cnt {bits} does not match the {channels} bit channels")
    return bits


def nn_out_to_xyz(nn_out: Tensor, noc_scale_xfd: Tensor, *, model_transform: Optional[Tensor] = None,
                  bit_cnt: Union[int, Sequence[int]], inference: bool = True, out: Optional[Tensor] = None) -> Tensor:
    """``(B,C,H,W)`` bit logits -> ``xyz (B,H,W,3)`` (``losses.py:16-45`` with ``inference=True``)."""
    if not inference:
        raise NotImplementedError("the training decode (with GT bits) is fused into lc_b200.dense.dense_pose_loss_noc_bin")
    dev = nat.check_cuda(nn_out, noc_scale_xfd, model_transform, out)
    if nn_out.dtype != torch.float32:
        raise TypeError("nn_out_to_xyz takes float32 logits")
    B, C_, H, W = nn_out.shape
    bits = _bits(bit_cnt, C_)
    if not (nn_out.stride(3) == 1 and nn_out.stride(2) == W):
        nn_out = nn_out.contiguous()
    xyz = torch.empty(B, H, W, 3, dtype=torch.float32, device=dev) if out is None else out
    a = nat.lc_decode_args()
    a.abi_version, a.B, a.H, a.W = nat.ABI_VERSION, B, H, W
    a.bit_cnt[:] = bits
    a.black_background = int(_black_background)
    keep = dict(noc_bin_logits=nn_out, noc_scale=noc_scale_xfd.to(torch.float32).expand(B, 3),
                model_transform=None if model_transform is None else model_transform.to(torch.float32).expand(B, 4, 4), xyz=xyz)
    for k, v in keep.items():
        setattr(a, k, nat.view_of(v))
    nat.call("lc_b200_noc_bin_decode", a, dev)
    return xyz


def nn_logits2noc(logits: Tensor, bit_cnt: Union[int, Sequence[int]]) -> Tensor:
    """``floatbits.nn_logits2noc`` without LUT (``floatbits.py:33-47``): ``(B,C,H,W)`` -> ``noc (B,H,W,3)``."""
    ones = torch.ones(1, 3, dtype=torch.float32, device=logits.device)
    return nn_out_to_xyz(logits, ones, bit_cnt=bit_cnt)


def nn_noc2target(noc: Tensor, bit_cnt: Union[int, Sequence[int]]):
    """``floatbits.nn_noc2target`` (``floatbits.py:13-31``): ``noc (B,H,W,3)`` in (-1,1) -> ``(mod_bits, bits)``, both bool
    ``(B,C,H,W)`` views of channel-last storage exactly like the reference returns (``:24-25, 31``): the Gray-coded training
    targets and the raw binary bits ``lc_b200.dense.dense_pose_loss_noc_bin`` consumes."""
    dev = nat.check_cuda(noc)
    if noc.dtype != torch.float32:
        raise TypeError("nn_noc2target takes float32 coordinates")
    B, H, W, _ = noc.shape
    bits = [int(bit_cnt)] * 3 if isinstance(bit_cnt, int) else [int(v) for v in bit_cnt]
    C_ = sum(bits)
    mod = torch.empty(B, H, W, C_, dtype=torch.uint8, device=dev)
    raw = torch.empty(B, H, W, C_, dtype=torch.uint8, device=dev)
    a = nat.lc_encode_args()
    a.abi_version, a.B, a.H, a.W = nat.ABI_VERSION, B, H, W
    a.bit_cnt[:] = bits
    a.black_background = int(_black_background)
    a.noc = nat.view_of(noc)
    a.mod_bits, a.raw_bits = mod.data_ptr(), raw.data_ptr()
    nat.call("lc_b200_noc_bin_encode", a, dev)
    return mod.view(torch.bool).permute(0, 3, 1, 2), raw.view(torch.bool).permute(0, 3, 1, 2)
