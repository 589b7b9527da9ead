"""Elementwise passes of the training step over the latent tables (csrc/train_ops.cu): the two STE bit planes, the
stand-in latents of rows a data-parallel rank does not own, and Adam fused with the plane refresh.

CUDA tensors go through the C-ABI kernels.  CPU tensors take the same arithmetic in torch ops: that branch exists for the
world_size-2 gloo tests of the data-parallel host logic (tests/test_dp_gloo.py) and is never reached by the GPU product
path -- every caller in this package hands over CUDA tensors, and the hot path itself (field, context model, coder) has
no CPU route at all.
"""
from __future__ import annotations

import math

import torch

from ._lib import check, lib, ptr, stream


def _bits_cpu(flags: torch.Tensor) -> torch.Tensor:
    """bool [n] -> uint8 [n/8], bit k of byte b = flags[8b + k] (cnc_sign_pack layout)"""
    w = (1 << torch.arange(8, dtype=torch.int32)).to(torch.uint8)
    return (flags.view(-1, 8).to(torch.uint8) * w).sum(-1).to(torch.uint8)


def _unbits_cpu(bits: torch.Tensor) -> torch.Tensor:
    return ((bits.view(-1, 1).to(torch.int32) >> torch.arange(8, dtype=torch.int32)) & 1).bool().view(-1)


def planes_pack(p: torch.Tensor, sign: torch.Tensor = None, mask: torch.Tensor = None):
    """(sign plane, window plane) of a contiguous fp32 tensor with numel % 32 == 0: uint8 [numel/8] each"""
    n = p.numel()
    assert p.is_contiguous() and p.dtype == torch.float32 and n % 32 == 0
    if sign is None:
        sign = torch.empty(n // 8, dtype=torch.uint8, device=p.device)
    if mask is None:
        mask = torch.empty(n // 8, dtype=torch.uint8, device=p.device)
    if p.is_cuda:
        check(lib().cnc_ste_planes_pack(ptr(p), ptr(sign), ptr(mask), n, stream()))
    else:
        f = p.detach().reshape(-1)
        sign.copy_(_bits_cpu(f >= 0))
        mask.copy_(_bits_cpu((f >= -1) & (f <= 1)))
    return sign, mask


def surrogate_fill(p: torch.Tensor, sign: torch.Tensor, mask: torch.Tensor, keep_lo: int, keep_hi: int) -> None:
    """p.flat[i] = +-0.5 / +-1.5 from the planes for every i outside [keep_lo, keep_hi) (multiples of 32), in place"""
    n = p.numel()
    assert p.is_contiguous() and n % 32 == 0 and keep_lo % 32 == 0 and keep_hi % 32 == 0 and 0 <= keep_lo <= keep_hi <= n
    if p.is_cuda:
        check(lib().cnc_surrogate_fill(ptr(p), ptr(sign), ptr(mask), n, keep_lo, keep_hi, stream()))
    else:
        s, m = _unbits_cpu(sign), _unbits_cpu(mask)
        v = torch.where(m, 0.5, 1.5) * torch.where(s, 1.0, -1.0)
        f = p.detach().view(-1)
        f[:keep_lo] = v[:keep_lo]
        f[keep_hi:] = v[keep_hi:]


def adam_planes(p, g, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-15, weight_decay: float = 0.0,
                grad_scale: float = 1.0, sign: torch.Tensor = None, mask: torch.Tensor = None, ste_window: bool = False) -> None:
    """one torch.optim.Adam step over the flat fp32 tensors (numel % 32 == 0), in place; `sign` / `mask` (uint8
    [numel/8], optional) receive the planes of the updated values in the same pass; `ste_window`: drop the gradient of latents
    outside [-1, 1] first (STE_binary backward folded in)"""
    n = p.numel()
    assert n % 32 == 0 and all(t.is_contiguous() and t.numel() == n for t in (p, g, exp_avg, exp_avg_sq))
    if p.is_cuda:
        check(lib().cnc_adam_planes(ptr(p), ptr(g), ptr(exp_avg), ptr(exp_avg_sq), ptr(sign), ptr(mask), n, lr, betas[0], betas[1],
                                    eps, weight_decay, step, grad_scale, int(ste_window), stream()))
        return
    with torch.no_grad():
        gr = g / grad_scale
        if ste_window:
            gr = gr * ((p >= -1) & (p <= 1))
        gr = gr + weight_decay * p
        exp_avg.lerp_(gr, 1 - betas[0])
        exp_avg_sq.mul_(betas[1]).addcmul_(gr, gr, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
        p.sub_((lr / bc1) * exp_avg / (exp_avg_sq.sqrt() / math.sqrt(bc2) + eps))
        if sign is not None or mask is not None:
            s, m = planes_pack(p.detach().contiguous())
            if sign is not None:
                sign.copy_(s)
            if mask is not None:
                mask.copy_(m)
