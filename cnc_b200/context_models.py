"""Host-side mirror of the reference context model + codec driver (examples/utils_bpp_acc.py).

`CNC_context_models(num_dim, resolutions_list, resolutions_list_2D, log2_hashmap_size,
log2_hashmap_size_2D, n_features, sample_num, max_context_layer_num, ste_binary, ..., Pg_level,
Pg_level_2D, Rb, step_update, skip_levels_3D, skip_levels_2D, use_dimension_wise,
use_overlap_area_pool)` with

    forward_binary_vxl_mixPg_3D2D(Encoding_xyz, Encoding_xy, Encoding_xz, Encoding_yz, binary_vxl, step=...)
        -> (bits_per_param, MB)                                   utils_bpp_acc.py:533-706
    encode_binary_vxl_mixPg_3D2D(..., binary_vxl, filename_prefix) -> (Pgs_dict, est_MB, coded_MB)   :709-865
    decode_binary_vxl_mixPg_3D2D(..., *_rec tables, binary_vxl, Pgs_dict, filename_prefix)          :867-999

Same constructor arguments, method names, return values and file layout (33 streams at the product
config: `<prefix>_xy0..3.b`, `_xz*`, `_yz*`, `_3D0.b`.. `_3D<n>_<chunk>.b`), so the train scripts'
call sites (train_CNC_nerf_synthetic.py:229-247,351-355,434-464) work unchanged.

B200-native internals
  * the chunk body of the 3D coder (mask query, compaction, 3-level masked gather, context MLP,
    padded packing, overlap-weighted mean) is ONE kernel, `cnc_context3d_probs` (csrc/context_fused.cu);
    `fused=False` keeps the reference's op-by-op data flow on the drop-in kernels (K6/K1/K8 + nn.Linear)
    for parity tests;
  * all streams of a level (all 33 at encode time) are range-coded concurrently on the GPU, one warp per
    stream (`cnc_ac_encode/decode`), instead of one CPU thread after a D2H copy (utils_bpp_acc.py:77-110);
  * the hash tables are read as 1-bit sign planes.
There is no CPU path: every method raises if the tensors are not on a CUDA device.
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as nnf
from torch.autograd import Function

from . import _gridencoder as _backend
from . import pack_and_align
from . import torchac as tac
from ._lib import check, lib, ptr, stream
from .gridencoder import STE_binary, STE_multistep

_PRIMES = (1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737)


def get_grid_index(hashmap_size, resolution, pos_grid):
    """Row index of integer grid vertices: examples/utils.py:492-511 (the python twin of
    gridencoder.cu:45-87).  uint32 wrap-around is applied explicitly, so the result equals the CUDA
    kernel's for any table size (the reference relies on power-of-two tables)."""
    D = pos_grid.shape[-1]
    hashmap_size, resolution = int(hashmap_size), int(resolution)
    pos = pos_grid.to(torch.int64)
    if resolution ** D <= hashmap_size:
        idx = torch.zeros(pos.shape[:-1], dtype=torch.int64, device=pos.device)
        stride = 1
        for d in range(D):
            idx += pos[..., d] * stride
            stride *= resolution
    else:
        idx = torch.zeros(pos.shape[:-1], dtype=torch.int64, device=pos.device)
        for d in range(D):
            idx ^= (pos[..., d] * _PRIMES[d]) & 0xFFFFFFFF
    return idx % hashmap_size


def my_meshgrid3D(start, end, device, dtype=torch.int32):
    """[lx, ly, lz, 3] integer lattice, x slowest (utils_bpp_acc.py:140-161)."""
    if isinstance(start, int):
        start, end = (start,) * 3, (end,) * 3
    ax = [torch.arange(s, e, device=device, dtype=dtype) for s, e in zip(start, end)]
    return torch.stack(torch.meshgrid(*ax, indexing="ij"), dim=-1)


class _cnt_np_embed(Function):
    """3D -> 2D vote fractions (utils_bpp_acc.py:27-75) on cnc_vote_planes_fwd/bwd."""

    @staticmethod
    def forward(ctx, inputs, embeddings, resolution, hashmap_size, axis):
        axis = ["xy", "xz", "yz"].index(axis)
        N, F = inputs.shape[0], embeddings.shape[-1]
        inputs = inputs.to(torch.int16).contiguous()
        embeddings = embeddings.contiguous()
        scale = resolution - 2
        pn_embed = torch.zeros(scale, scale, F, 2, device=inputs.device)
        _backend.cnt_np_embed(inputs, embeddings, pn_embed, N, resolution, F, hashmap_size, axis)
        pn_sum = pn_embed.sum(dim=-1, keepdim=True) + 1e-6
        ctx.save_for_backward(inputs, embeddings, pn_sum)
        ctx.dims = [N, resolution, F, hashmap_size, axis]
        return pn_embed / pn_sum

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, pn_sum = ctx.saved_tensors
        N, resolution, F, hashmap_size, axis = ctx.dims
        g = torch.zeros_like(embeddings)
        _backend.cnt_np_embed_backward(inputs, embeddings, pn_sum, grad.contiguous(), g, N, resolution, F,
                                       hashmap_size, axis)
        return None, g, None, None, None


class _cnt_np_embed3(Function):
    """The three vote-fraction planes of the dimension-wise context at once (3 x `_cnt_np_embed` over
    `get_idx_coords2`, utils_bpp_acc.py:27-75,498-530), on `cnc_vote3_fwd / cnc_vote3_bwd`: the voxel list is replaced
    by its closed form in the occupancy grid, the atomics by per-cell / per-row ownership."""

    @staticmethod
    def forward(ctx, embeddings, vx, pts_by_row, seg, resolution, hashmap_size, axes=(0, 1, 2), members_only=False):
        """-> one plane per entry of `axes` (0 = xy, 1 = xz, 2 = yz), in that order: a data-parallel rank builds (and
        differentiates) only the planes its share of the plane terms reads"""
        embeddings = embeddings.contiguous()
        F = embeddings.shape[-1]
        bits = _backend.sign_pack(embeddings.detach())
        s = resolution - 2
        outs = [torch.empty(s, s, F, 2, device=embeddings.device) if a in axes else None for a in range(3)]
        check(lib().cnc_vote3_fwd(ptr(vx), vx.shape[-1], ptr(bits), resolution, F, hashmap_size, *[ptr(o) for o in outs], stream()))
        sums = [outs[a].sum(dim=-1, keepdim=True) + 1e-6 for a in axes]
        ctx.save_for_backward(bits, vx, pts_by_row, seg, *sums)
        ctx.dims = [resolution, F, hashmap_size, tuple(embeddings.shape), tuple(axes)]
        ctx.set_materialize_grads(False)
        ctx.members_only = members_only
        return tuple(outs[a] / sm for a, sm in zip(axes, sums))

    @staticmethod
    def backward(ctx, *grads):
        bits, vx, pts_by_row, seg, *sums = ctx.saved_tensors
        resolution, F, hashmap_size, shape, axes = ctx.dims
        if all(g is None for g in grads):
            return None, None, None, None, None, None, None, None
        g = torch.zeros(shape, device=bits.device)
        gs = [None, None, None]
        for a, gg, sm in zip(axes, grads, sums):
            if gg is not None:
                gs[a] = (gg / sm).contiguous()     # 1 / sum folded in (:1012)
        check(lib().cnc_vote3_bwd(ptr(pts_by_row), ptr(seg), ptr(vx), vx.shape[-1], ptr(bits), resolution, F, hashmap_size,
                                  ptr(gs[0]), ptr(gs[1]), ptr(gs[2]), ptr(g), int(ctx.members_only), stream()))
        return g, None, None, None, None, None, None, None


class align_and_pack(Function):
    """ragged -> padded [N, M, F] (utils_bpp_acc.py:113-139) on cnc_align_pack_fwd/bwd."""

    @staticmethod
    def forward(ctx, voxel_features, unique_cnt, V, dim=3):
        voxel_features = voxel_features.contiguous()
        cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=unique_cnt.device), torch.cumsum(unique_cnt, 0)])
        N, M, F = unique_cnt.numel(), int(unique_cnt.max()) if unique_cnt.numel() else 0, voxel_features.shape[-1]
        ctx.save_for_backward(voxel_features, unique_cnt, cs)
        ctx.dims = [N, M, F, int(cs[-1]), dim]
        return pack_and_align.align_and_pack_forward(voxel_features, unique_cnt, cs, N, M, F, 0.0, dim)

    @staticmethod
    def backward(ctx, dL):
        voxel_features, unique_cnt, cs = ctx.saved_tensors
        N, M, F, T, dim = ctx.dims
        return pack_and_align.align_and_pack_backward(dL.contiguous(), voxel_features, unique_cnt, cs, N, M, F, T, dim), None, None, None


class segment_sum(Function):
    """out[i] = sum_j w[j] * feat[idx[j]] over the rows j of segment i (cumsum [N+1]; w, idx optional), differentiable in
    `feat`: replaces index_select -> align_and_pack -> (* weights) -> sum(dim=1) and its padded [N, M, F] temporaries
    (utils_bpp_acc.py:563-566,741-745) by one kernel each way (cnc_segment_wsum_idx / _bwd)."""

    @staticmethod
    def forward(ctx, feat, cumsum, weights=None, idx=None):
        feat = feat.contiguous()
        N, F = cumsum.numel() - 1, feat.shape[1]
        out = torch.empty(N, F, device=feat.device, dtype=torch.float32)
        check(lib().cnc_segment_wsum_idx(ptr(feat), ptr(idx), ptr(weights), ptr(cumsum), ptr(out), N, F, stream()))
        ctx.save_for_backward(cumsum, weights, idx)
        ctx.rows = feat.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        cumsum, weights, idx = ctx.saved_tensors
        N, F = cumsum.numel() - 1, g.shape[1]
        gf = torch.zeros(ctx.rows, F, device=g.device, dtype=torch.float32)
        check(lib().cnc_segment_wsum_idx_bwd(ptr(g.contiguous()), ptr(idx), ptr(weights), ptr(cumsum), ptr(gf), N, F, stream()))
        return gf, None, None, None


class _CtxMLP3(Function):
    """context_model_3D (Linear(25,32) LeakyReLU Linear(32,32) LeakyReLU Linear(32,8)) over [voxels, 25] as one forward and
    one backward kernel (`cnc_ctx_mlp_fwd / _bwd`, csrc/context_mlp.cu): the hidden activations are recomputed in the
    backward instead of being written to and read from HBM, the weight gradients are per-CTA partials added in index order."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, W3, b3):
        x = x.contiguous().float()
        packed = torch.cat([t.detach().float().reshape(-1) for t in (W1.t(), b1, W2.t(), b2, W3.t(), b3)]).contiguous()
        y = torch.empty(x.shape[0], W3.shape[0], device=x.device, dtype=torch.float32)
        if x.shape[0]:   # (no voxel of the sampled entries touches the occupancy: nothing to evaluate)
            check(lib().cnc_ctx_mlp_fwd(ptr(x), ptr(packed), ptr(y), x.shape[0], stream()))
        ctx.save_for_backward(x, packed)
        ctx.dims = (W1.shape[1], W1.shape[0], W3.shape[0])
        return y

    @staticmethod
    def backward(ctx, gy):
        x, packed = ctx.saved_tensors
        nin, nh, no = ctx.dims
        M = x.shape[0]
        gx = torch.empty_like(x)
        if M == 0:
            z = lambda *shape: torch.zeros(*shape, device=x.device, dtype=torch.float32)
            return gx, z(nh, nin), z(nh), z(nh, nh), z(nh), z(no, nh), z(no)
        G = max(1, min(lib().cnc_ctx_mlp_max_partials(), (M + 255) // 256))
        parts = torch.empty(G, packed.numel(), device=x.device, dtype=torch.float32)
        check(lib().cnc_ctx_mlp_bwd(ptr(x), ptr(packed), ptr(gy.contiguous().float()), ptr(gx), ptr(parts), G, M, stream()))
        g = parts.sum(0)
        o = [0, nin * nh, nin * nh + nh, nin * nh + nh + nh * nh, nin * nh + 2 * nh + nh * nh, nin * nh + 2 * nh + nh * nh + nh * no]
        return (gx, g[o[0]:o[1]].view(nin, nh).t(), g[o[1]:o[2]], g[o[2]:o[3]].view(nh, nh).t(), g[o[3]:o[4]],
                g[o[4]:o[5]].view(nh, no).t(), g[o[5]:o[5] + no])


class _Lin8(Function):
    """nn.Linear(K, 8) over [N, K] rows, K in {9, 17, 25, 33}: the plane context models (utils_bpp_acc.py:386-393) as one
    pass each way (csrc/context_lin8.cu) instead of cuBLAS' tall-skinny SIMT GEMMs"""

    @staticmethod
    def forward(ctx, x, W, b):
        x = x.contiguous().float()
        Wc, bc = W.detach().contiguous().float(), b.detach().contiguous().float()
        y = torch.empty(x.shape[0], 8, device=x.device, dtype=torch.float32)
        check(lib().cnc_lin8_fwd(ptr(x), ptr(Wc), ptr(bc), ptr(y), x.shape[0], x.shape[1], stream()))
        ctx.save_for_backward(x, Wc)
        ctx.need_gx = x.requires_grad or True
        return y

    @staticmethod
    def backward(ctx, gy):
        x, Wc = ctx.saved_tensors
        N, K = x.shape
        if N == 0:
            return torch.zeros_like(x), torch.zeros(8, K, device=x.device), torch.zeros(8, device=x.device)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        nb = (N + lib().cnc_lin8_rows_per_block() - 1) // lib().cnc_lin8_rows_per_block()
        parts = torch.empty(nb, 8 * K + 8, device=x.device, dtype=torch.float32)
        check(lib().cnc_lin8_bwd(ptr(x), ptr(Wc), ptr(gy.contiguous().float()), ptr(gx), ptr(parts), N, K, stream()))
        g = parts.sum(0)
        return gx, g[:8 * K].view(8, K), g[8 * K:]


def _linear8(lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    if x.is_cuda and lin.out_features == 8 and lin.in_features in (9, 17, 25, 33) and x.shape[0] > 0:
        return _Lin8.apply(x, lin.weight, lin.bias)
    return lin(x)


class _Ctx3DGather(Function):
    """[voxels, 25] input of context_model_3D for the training loss: the three coarser levels' features interpolated at each
    voxel (masked gather, per-point start level) | Pg of the voxel's level -- `forward_diff_levels(..., PV=1001)` + `cat` of
    utils_bpp_acc.py:685-686 in one kernel each way (csrc/context_train.cu).  Gradients: to the latent table (K2 scatter-add
    + STE window) and to the level frequencies."""

    @staticmethod
    def forward(ctx, params, Pg_levels, pts, level, enc, vbits, vbit_off):
        M = pts.shape[0]
        x = torch.empty(M, 25, device=pts.device, dtype=torch.float32)
        bits = enc.sign_bits() if params is enc.params else _backend.sign_pack(params.detach().contiguous())
        pg = Pg_levels.detach().contiguous().float()
        check(lib().cnc_ctx3d_gather_fwd(ptr(pts), ptr(level), M, ptr(bits), ptr(enc.offsets_list), ptr(enc.resolutions_list),
                                         ptr(vbits), ptr(vbit_off), ptr(pg), ptr(x), stream()))
        ctx.save_for_backward(params, pts, level, vbits, vbit_off)
        ctx.enc, ctx.L = enc, Pg_levels.numel()
        return x

    @staticmethod
    def backward(ctx, gx):
        params, pts, level, vbits, vbit_off = ctx.saved_tensors
        enc = ctx.enc
        gx = gx.contiguous()
        ge = torch.zeros_like(params)
        g_pg = torch.zeros(ctx.L, device=gx.device, dtype=torch.float32)
        check(lib().cnc_ctx3d_gather_bwd(ptr(pts), ptr(level), pts.shape[0], ptr(enc.offsets_list), ptr(enc.resolutions_list),
                                         ptr(vbits), ptr(vbit_off), ptr(gx), ptr(ge), ptr(g_pg), stream()))
        g_params = _backend.ste_binary_backward(params.contiguous(), ge)
        return g_params, g_pg, None, None, None, None, None


class _Ctx2DGather(Function):
    """[vertices, 8c (+8) + 1] input of a plane context model: c coarser plane levels | vote fraction plane | Pg, and its
    backward (csrc/context_train.cu).  `params` = the plane's latent table when it is the encoder's own (gradient: K2 + STE
    window), else a caller-provided +-1 table (decode: no gradient)."""

    @staticmethod
    def forward(ctx, params, frac, Pg_n, pts, enc, n, c, vxl2, res_frac, own):
        N = pts.shape[0]
        K = 8 * c + (8 if frac is not None else 0) + 1
        x = torch.empty(N, K, device=pts.device, dtype=torch.float32)
        bits = enc.sign_bits() if own else _backend.sign_pack(params.detach().contiguous())
        fr = None if frac is None else frac.detach().contiguous().float()
        pg = Pg_n.detach().reshape(1).contiguous().float()
        check(lib().cnc_ctx2d_gather_fwd(ptr(pts), N, ptr(bits), ptr(enc.offsets_list), ptr(enc.resolutions_list), n, c, ptr(vxl2),
                                         vxl2.shape[-1], ptr(fr), res_frac, ptr(pg), ptr(x), stream()))
        ctx.save_for_backward(params, pts, vxl2)
        ctx.enc, ctx.dims, ctx.frac_shape = enc, (n, c, K, res_frac), None if frac is None else tuple(frac.shape)
        return x

    @staticmethod
    def backward(ctx, gx):
        params, pts, vxl2 = ctx.saved_tensors
        enc, (n, c, K, res_frac) = ctx.enc, ctx.dims
        gx = gx.contiguous()
        ge = torch.zeros_like(params)
        gf = None if ctx.frac_shape is None else torch.zeros(ctx.frac_shape, device=gx.device, dtype=torch.float32)
        gpg = torch.zeros(1, device=gx.device, dtype=torch.float32)
        check(lib().cnc_ctx2d_gather_bwd(ptr(pts), pts.shape[0], ptr(enc.offsets_list), ptr(enc.resolutions_list), n, c, ptr(vxl2),
                                         vxl2.shape[-1], res_frac, ptr(gx), ptr(ge), ptr(gf), ptr(gpg), stream()))
        g_params = _backend.ste_binary_backward(params.contiguous(), ge) if ctx.needs_input_grad[0] else None
        return g_params, gf, gpg.reshape(()), None, None, None, None, None, None, None


_LEVEL_CONST = {}


def _level_const(offsets, F, device):
    """(level index of every table row [rows] int64, entries per level [L] fp32), cached: building them per call costs a
    host->device copy (a sync) each"""
    key = (tuple(int(o) for o in offsets), int(F), str(device))
    if key not in _LEVEL_CONST:
        offs = key[0]
        sizes = torch.tensor([o1 - o0 for o0, o1 in zip(offs[:-1], offs[1:])])
        level = torch.repeat_interleave(torch.arange(len(sizes)), sizes)
        _LEVEL_CONST[key] = (level.to(device), (sizes * F).to(torch.float32).to(device))
    return _LEVEL_CONST[key]


class _RowsGather(Function):
    """table[rows] for DISTINCT rows, differentiable in the table.  Autograd's own backward of an index is an accumulating
    index_put: a radix sort of the indices and a dozen launches, every step, for indices that never repeat (the coded rows of
    a level, a window of a level's entries); here it is a zero fill and one scatter."""

    @staticmethod
    def forward(ctx, table, rows):
        ctx.save_for_backward(rows)
        ctx.shape = tuple(table.shape)
        ctx.fast = table.is_cuda and table.dim() == 2 and table.shape[1] == 8 and table.dtype == torch.float32 and table.is_contiguous()
        if ctx.fast:
            rows = rows.contiguous()
            out = torch.empty(rows.numel(), 8, device=table.device)
            check(lib().cnc_rows8_gather(ptr(table), ptr(rows), rows.numel(), ptr(out), stream()))
            return out
        return table.index_select(0, rows)

    @staticmethod
    def backward(ctx, g):
        (rows,) = ctx.saved_tensors
        out = torch.zeros(ctx.shape, device=g.device, dtype=g.dtype)
        if ctx.fast and g.dtype == torch.float32:
            check(lib().cnc_rows8_scatter(ptr(g.contiguous()), ptr(rows.contiguous()), rows.numel(), ptr(out), stream()))
        else:
            out.index_copy_(0, rows, g.contiguous())
        return out, None


class _LevelSums(Function):
    """sum of all entries of every level of a table [rows, F] -> [L]; the backward hands one full-size gradient tensor
    to autograd instead of one zero-padded tensor per level slice (what slicing under autograd does)."""

    @staticmethod
    def forward(ctx, params_q, offsets, binary=False):
        ctx.offsets, ctx.shape = tuple(int(o) for o in offsets), tuple(params_q.shape)
        F = params_q.shape[-1]
        if binary and params_q.is_cuda and all((o * F) % 8 == 0 for o in ctx.offsets):
            # every entry is +1 or -1: sum = 2 * (#bits set in the sign plane) - (#entries), exact, one launch over 1 bit per
            # entry instead of one fp32 reduction per level (whose result beyond 2^24 entries is rounded)
            key = ("byte_offs",) + ctx.offsets + (F, str(params_q.device))
            if key not in _LEVEL_CONST:
                _LEVEL_CONST[key] = torch.tensor([o * F // 8 for o in ctx.offsets], dtype=torch.int64).to(params_q.device)
            bits = _backend.sign_pack(params_q.detach().contiguous())
            ones = torch.empty(len(ctx.offsets) - 1, dtype=torch.int64, device=params_q.device)
            check(lib().cnc_level_popcount(ptr(bits), ptr(_LEVEL_CONST[key]), ones.numel(), ptr(ones), stream()))
            _, ttl = _level_const(ctx.offsets, F, params_q.device)
            return 2.0 * ones.to(torch.float32) - ttl
        return torch.stack([params_q[o0:o1].sum() for o0, o1 in zip(ctx.offsets[:-1], ctx.offsets[1:])])

    @staticmethod
    def backward(ctx, g):
        offs = ctx.offsets
        level, _ = _level_const(offs, ctx.shape[1], g.device)
        if offs[0] == 0 and offs[-1] == ctx.shape[0]:
            gr = g[level]
        else:
            gr = torch.zeros(ctx.shape[0], device=g.device, dtype=g.dtype)
            gr[offs[0]:offs[-1]] = g[level]
        return gr.unsqueeze(-1).expand(ctx.shape).contiguous(), None, None


class _BernoulliBitsSum(Function):
    """torch.sum(Bernoulli_entropy(x, p)) with its gradients in x and p as one kernel each way (`cnc_bernoulli_bits_fwd /
    _bwd`): under autograd the same expression is ~12 elementwise launches forward and ~30 backward, four times per step."""

    @staticmethod
    def forward(ctx, x, p):
        x, p = x.contiguous().float(), p.contiguous().float()
        n = x.numel()
        ctx.save_for_backward(x, p)
        if n == 0:
            return torch.zeros((), device=x.device)
        parts = torch.empty(lib().cnc_bernoulli_bits_blocks(n), device=x.device)
        check(lib().cnc_bernoulli_bits_fwd(ptr(x), ptr(p), n, ptr(parts), stream()))
        return parts.sum()

    @staticmethod
    def backward(ctx, g):
        x, p = ctx.saved_tensors
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gp = torch.empty_like(p) if ctx.needs_input_grad[1] else None
        if x.numel() and (gx is not None or gp is not None):
            check(lib().cnc_bernoulli_bits_bwd(ptr(x), ptr(p), ptr(g.contiguous().float()), x.numel(), ptr(gx), ptr(gp), stream()))
        return gx, gp


class Bernoulli_entropy(nn.Module):
    """utils_bpp_acc.py:1002-1013: bits of x in {-1,+1} under P(+1) = p (clamped to [1e-6, 1-1e-6])."""

    def forward(self, x, p):
        p = torch.clamp(p, min=1e-6, max=1 - 1e-6)
        return -torch.log2(p) * ((1 + x) / 2.0) + -torch.log2(1 - p) * ((1 - x) / 2.0)


def _z85(b: bytes) -> str:
    """generator state -> text for the json header (zlib: the mt19937 state array is 5 KB, mostly incompressible, but
    the tail of torch's state blob is zeros)"""
    import base64
    import zlib

    return base64.b85encode(zlib.compress(b, 9)).decode()


def _unz85(s: str) -> bytes:
    import base64
    import zlib

    return zlib.decompress(base64.b85decode(s.encode()))


def _layout(resolutions, num_dim, log2T):
    offs, off = [], 0
    for r in resolutions:
        n = int(np.ceil(min(2 ** log2T, int(r) ** num_dim) / 8) * 8)
        offs.append(off)
        off += n
    offs.append(off)
    return offs


class CNC_context_models(nn.Module):
    def __init__(self, num_dim=3, resolutions_list=(16, 22, 31, 42, 57, 78, 106, 146, 199, 273, 374, 512),
                 resolutions_list_2D=(128, 256, 512, 1024), log2_hashmap_size=19, log2_hashmap_size_2D=21,
                 n_features=4, sample_num=20000, max_context_layer_num=3, ste_binary=False, ste_multistep=False,
                 add_noise=False, Q=100, quantize_epoch=1000, Pg_level=-1, Pg_level_2D=-1, Rb=128, step_update=16,
                 skip_levels_3D=(0, 1, 2, 3), skip_levels_2D=(0,), use_dimension_wise=True,
                 use_overlap_area_pool=True, device="cuda", fused=True, shuffle_seed=None, tables="full"):
        """`shuffle_seed`: the symbol order of the dense context-coded levels is a random permutation
        (utils_bpp_acc.py:311-315) that encoder and decoder must share.  None = the reference's behaviour: drawn from
        torch's global CPU generator (same `torch.manual_seed` => same order as the reference's constructor); an int =
        drawn from a private generator seeded with it.  Either way `self.shuffle_seed` ends up holding what a decoder
        in another process needs to rebuild the same tables (an int, or the CPU generator state as bytes); the container
        stores it (container.pack(..., layout=cm.layout())).

        `tables`: "full" = the reference's inverse hash tables (every lattice vertex of every level grouped by table row,
        utils_bpp_acc.py:294-335: 1.34 GB at the product layout, `pos_grid_sorted_list`).  "pruned" (CUDA, codec only) = only
        the row statistics are computed at construction (one histogram pass per level, csrc/table_build.cu); the vertex
        lists are built per occupancy grid at encode / decode time for the vertices that pass the occupancy test -- what
        `mask` / `mask_exist` keep anyway (:811-833).  Same entry numbering, symbol order and chunking; the probabilities of
        an entry are summed over the same vertices in the same order, grouped differently into the kernel's batches, so
        the two modes agree to fp32 rounding, not bit for bit: the mode is part of `layout()` and must match on both
        sides of the codec.  The training loss needs the full lists and builds them on first use."""
        super().__init__()
        if tables not in ("full", "pruned"):
            raise ValueError("tables must be 'full' or 'pruned'")
        self.tables = tables
        dev = torch.device(device)
        self.MAX_POINTS_NUM_TO_OOM = 20000000
        self.use_overlap_area_pool, self.use_dimension_wise, self.fused = use_overlap_area_pool, use_dimension_wise, fused
        self.num_dim, self.n_features = num_dim, n_features
        self.res = [int(r) for r in resolutions_list]
        self.res_2D = [int(r) for r in resolutions_list_2D]
        self.n_levels, self.n_levels_2D = len(self.res), len(self.res_2D)
        self.resolutions_list = torch.tensor(self.res, device=dev)
        self.resolutions_list_2D = torch.tensor(self.res_2D, device=dev)
        self.scales_list = (self.resolutions_list - 2).unsqueeze(-1)
        self.scales_list_2D = (self.resolutions_list_2D - 2).unsqueeze(-1)
        self.log2_hashmap_size, self.log2_hashmap_size_2D = log2_hashmap_size, log2_hashmap_size_2D
        self.ste_binary, self.ste_multistep, self.add_noise, self.Q = ste_binary, ste_multistep, add_noise, Q
        self.quantize_epoch, self.quantize_epoch_cnt = quantize_epoch, 0
        Pg_level = self.n_levels if (Pg_level == -1 or Pg_level >= self.n_levels) else max(Pg_level, 1)
        Pg_level_2D = self.n_levels_2D if (Pg_level_2D == -1 or Pg_level_2D >= self.n_levels_2D) else max(Pg_level_2D, 1)
        self.Pg_level, self.Pg_level_2D = Pg_level, Pg_level_2D
        self.skip_levels_3D, self.skip_levels_2D = tuple(skip_levels_3D), tuple(skip_levels_2D)
        self.sample_num, self.max_context_layer_num, self.step_update = sample_num, max_context_layer_num, step_update

        self.offs = _layout(self.res, num_dim, log2_hashmap_size)
        self.offs_2D = _layout(self.res_2D, 2, log2_hashmap_size_2D)
        self.offsets_list = torch.tensor(self.offs, dtype=torch.int64, device=dev)
        self.offsets_list_2D = torch.tensor(self.offs_2D, dtype=torch.int64, device=dev)
        max_params = 2 ** log2_hashmap_size
        self.n_levels_thresh, self.resolution_thresh = self.n_levels - 1, float(self.res[-1])
        for i in range(self.n_levels - 1):   # last dense level (utils_bpp_acc.py:288-292)
            if self.res[i] ** num_dim <= max_params < self.res[i + 1] ** num_dim:
                self.n_levels_thresh, self.resolution_thresh = i + 1, float(self.res[i])

        # ---- inverse hash tables (utils_bpp_acc.py:294-335): for every level, the voxels grouped by table row
        self.unique_value_list: List[torch.Tensor] = []         # rows that are hit, in bitstream symbol order
        self.unique_count_list: List[torch.Tensor] = []         # voxels per row
        self.unique_count_cumsum_list: List[torch.Tensor] = []  # running sum, leading 0
        self.pos_grid_sorted_list: List[torch.Tensor] = []      # int16 [res^3, 3] voxel coords grouped by row
        if shuffle_seed is None:          # the reference: global CPU generator; remember where it stood
            gen, self.shuffle_seed = None, torch.get_rng_state().numpy().tobytes()
        elif isinstance(shuffle_seed, (bytes, bytearray)):   # a recorded generator state
            gen = torch.Generator()
            gen.set_state(torch.frombuffer(bytearray(shuffle_seed), dtype=torch.uint8).clone())
            self.shuffle_seed = bytes(shuffle_seed)
        else:
            gen, self.shuffle_seed = torch.Generator().manual_seed(int(shuffle_seed)), int(shuffle_seed)
        # finest level first, like the reference (utils_bpp_acc.py:301): the order fixes which randperm a level gets
        self.entry_of_row_list: List[torch.Tensor] = []         # pruned mode: table row -> entry index (int32), None = identity
        for i in reversed(range(Pg_level)):
            r = self.res[i]
            if tables == "full":
                rows, cnt, pts = self._inverse_table(i, dev)
            else:
                rows, cnt, pts = self._row_statistics(i, dev)
            if r <= self.resolution_thresh:   # dense levels: random symbol order (one voxel per row), CPU generator
                shuffle_idx = torch.randperm(rows.nelement(), generator=gen).to(dev)
                rows, cnt = rows[shuffle_idx], cnt[shuffle_idx]
                if pts is not None:
                    pts = pts[shuffle_idx]
            entry_of_row = None
            T_i = self.offs[i + 1] - self.offs[i]
            if not (rows.numel() == T_i and r > self.resolution_thresh) and dev.type == "cuda":
                entry_of_row = torch.full((T_i,), -1, dtype=torch.int32, device=dev)
                entry_of_row[rows] = torch.arange(rows.numel(), dtype=torch.int32, device=dev)
            self.entry_of_row_list.insert(0, entry_of_row)
            self.unique_value_list.insert(0, rows.to(torch.int64))
            self.unique_count_list.insert(0, cnt.to(torch.int64))
            self.unique_count_cumsum_list.insert(0, torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(cnt, 0)]))
            self.pos_grid_sorted_list.insert(0, None if pts is None else pts.contiguous())
        self.hashparams_num_levels = torch.tensor([v.numel() for v in self.unique_value_list], device=dev)
        snl = torch.round(self.hashparams_num_levels * (sample_num / self.hashparams_num_levels.sum())).to(torch.long)
        self.sample_num_levels = self.hashparams_num_levels if snl[-1] > self.hashparams_num_levels[-1] else snl
        coded = [n for n in range(self.n_levels) if n not in self.skip_levels_3D and n < Pg_level]
        self.ttl_hashparams_num_levels = int(self.hashparams_num_levels.sum())
        self.ttl_hashparams_num_valid_levels = int(sum(int(self.hashparams_num_levels[n]) for n in coded))
        self.ttl_sample_num = int(self.sample_num_levels.sum())
        self.ttl_sample_num_valid_levels = int(sum(int(self.sample_num_levels[n]) for n in coded))
        self.utils_rand = torch.rand(Pg_level).to(dev)   # CPU generator, like the reference (utils_bpp_acc.py:367)
        # voxels per row as the reference computes it (int64 ** / int64 -> float32), it fixes the chunking of the streams
        self.utils_points_per_param_levels = [((self.resolutions_list[i] ** num_dim) / self.hashparams_num_levels[i]).item()
                                              for i in range(Pg_level)]
        self.binary_vxl_2D_idx = torch.stack(torch.meshgrid(torch.arange(Rb, device=dev, dtype=torch.int32),
                                                            torch.arange(Rb, device=dev, dtype=torch.int32), indexing="ij"), -1)

        self.context_model_3D = nn.Sequential(nn.Linear(n_features * max_context_layer_num + 1, 32), nn.LeakyReLU(),
                                              nn.Linear(32, 32), nn.LeakyReLU(), nn.Linear(32, n_features)).to(dev)
        self.context_model_2D = nn.Sequential(*[
            nn.Sequential(nn.Linear(n_features * (min(n, max_context_layer_num) + int(use_dimension_wise)) + 1, n_features))
            for n in range(1, Pg_level_2D)]).to(dev)
        self.entropy_model = Bernoulli_entropy()
        self.binary_vxl_len = Rb
        self.init_binary_vxl_coords(self.res[-1] - 2)
        self.idx_coords2_tmp, self.batched_inputs_list = None, None

    def _inverse_table(self, i, dev):
        """level i -> (rows that are hit [E] ascending, voxels per row [E], int16 voxel coordinates grouped by row
        [res^3, 3]; within a row in lattice order x-major): utils_bpp_acc.py:303-309"""
        r = self.res[i]
        pos_grid = my_meshgrid3D(0, r, dev).view(-1, 3)
        indexes = get_grid_index(self.offs[i + 1] - self.offs[i], r, pos_grid)
        indexes_sorted, order = torch.sort(indexes, stable=True)
        del indexes
        pts = pos_grid.to(torch.int16)[order]
        del pos_grid, order
        rows, cnt = torch.unique_consecutive(indexes_sorted, return_counts=True)
        return rows, cnt, pts

    def _row_statistics(self, i, dev):
        """level i -> (rows that are hit ascending, vertices per row, None): what `_inverse_table` returns minus the vertex
        list, from one histogram pass (cnc_level_row_hist) instead of a sort of res^3 keys"""
        if dev.type != "cuda":
            raise RuntimeError("tables='pruned' builds its tables with CUDA kernels: a CUDA device is required")
        T = self.offs[i + 1] - self.offs[i]
        cnt = torch.zeros(T, dtype=torch.int32, device=dev)
        check(lib().cnc_level_row_hist(self.res[i], T, ptr(cnt), stream()))
        rows = cnt.nonzero().squeeze(1)
        return rows, cnt[rows].to(torch.int64), None

    def _ensure_full_tables(self):
        """the vertex lists of every level (training loss, op-by-op paths): built on first use in pruned mode"""
        if all(p is not None for p in self.pos_grid_sorted_list):
            return
        dev = self.resolutions_list.device
        for i in range(self.Pg_level):
            if self.pos_grid_sorted_list[i] is None:
                rows, _, pts = self._inverse_table(i, dev)
                if self.res[i] <= self.resolution_thresh:    # dense: entry j is row unique_value_list[j] (one vertex each)
                    pts = pts[self.unique_value_list[i]]
                self.pos_grid_sorted_list[i] = pts.contiguous()

    def _pruned_level(self, n, vx):
        """vertex list of level n restricted to the vertices that pass the occupancy test of `vx` [Rb,Rb,Rb] (cached per
        grid): (pts int16 [M,3] grouped by entry, ent int64 [Ee] entry indices that exist ascending, seg int64 [Ee+1]
        running vertex counts, ent / seg also as numpy on the host)"""
        key = (vx.data_ptr(), vx._version, tuple(vx.shape))
        cache = getattr(self, "_pruned_cache", None)
        if cache is None or cache[0] != key:
            cache = self._pruned_cache = (key, {}, vx)   # (vx kept alive: see _vertex_bits)
        if n not in cache[1]:
            dev = vx.device
            r, T = self.res[n], self.offs[n + 1] - self.offs[n]
            eor = self.entry_of_row_list[n]
            counter = torch.zeros(1, dtype=torch.int64, device=dev)
            args = (r, T, ptr(vx), vx.shape[-1], ptr(eor))
            check(lib().cnc_level_pruned_keys(*args, None, ptr(counter), 0, stream()))
            M = int(counter.item())
            keys = torch.empty(M, dtype=torch.int64, device=dev)
            counter.zero_()
            check(lib().cnc_level_pruned_keys(*args, ptr(keys), ptr(counter), 0, stream()))
            keys, _ = torch.sort(keys)
            pts = torch.empty(M, 3, dtype=torch.int16, device=dev)
            entry = torch.empty(M, dtype=torch.int32, device=dev)
            check(lib().cnc_keys_to_points(ptr(keys), M, r, ptr(pts), ptr(entry), stream()))
            del keys
            ent, cnt = torch.unique_consecutive(entry, return_counts=True)
            seg = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(cnt, 0)])
            ent = ent.to(torch.int64)     # (also on the host: searchsorted with a python int would convert an int32 array per call)
            cache[1][n] = (pts, ent, seg, ent.cpu().numpy(), seg.cpu().numpy())
        return cache[1][n]

    def _pruned_weights(self, n, vx, binary_vxl):
        """per-vertex weights of the overlap-weighted mean over an entry's vertices (utils_bpp_acc.py:672-682), for the
        pruned list of level n: clamp(overlap, 1) / sum over the entry, or 1 / count.  Cached with the list."""
        pts, ent, seg, _, _ = self._pruned_level(n, vx)
        cache = self._pruned_cache[1]
        if ("w", n) not in cache:
            cnt = seg[1:] - seg[:-1]
            if self.use_overlap_area_pool:
                _, overlap = self.query_binary_vxl(pts, binary_vxl, n, return_overlap_area=True)
                ov = torch.clamp(overlap, min=1).to(torch.float)
                ov_sum = pack_and_align.segment_wsum(ov.unsqueeze(-1), seg).squeeze(-1)
                w = ov / torch.repeat_interleave(ov_sum, cnt, output_size=pts.shape[0])
            else:
                w = torch.repeat_interleave(1.0 / cnt.to(torch.float), cnt, output_size=pts.shape[0])
            cache[("w", n)] = (w.contiguous(), torch.full((pts.shape[0],), n, dtype=torch.int64, device=pts.device))
        return cache[("w", n)]

    def table_bytes(self) -> int:
        """device bytes held by the inverse hash tables right now (vertex lists + per-entry arrays + pruned caches)"""
        ts = [t for lst in (self.unique_value_list, self.unique_count_list, self.unique_count_cumsum_list, self.pos_grid_sorted_list,
                            self.entry_of_row_list) for t in lst if t is not None]
        cache = getattr(self, "_pruned_cache", None)
        if cache is not None:
            ts += [t for v in cache[1].values() for t in v[:3] if isinstance(t, torch.Tensor)]
        return sum(t.numel() * t.element_size() for t in ts)

    def layout(self) -> dict:
        """what a decoder in another process needs to rebuild this object (json-able; goes into the container header)"""
        seed = self.shuffle_seed
        return {"num_dim": self.num_dim, "resolutions_list": self.res, "resolutions_list_2D": self.res_2D,
                "log2_hashmap_size": self.log2_hashmap_size, "log2_hashmap_size_2D": self.log2_hashmap_size_2D,
                "n_features": self.n_features, "max_context_layer_num": self.max_context_layer_num,
                "Pg_level": self.Pg_level, "Pg_level_2D": self.Pg_level_2D, "Rb": self.binary_vxl_len,
                "skip_levels_3D": list(self.skip_levels_3D), "skip_levels_2D": list(self.skip_levels_2D),
                "use_dimension_wise": self.use_dimension_wise, "use_overlap_area_pool": self.use_overlap_area_pool,
                "tables": self.tables, "shuffle_seed": seed if isinstance(seed, int) else {"cpu_rng_state_hex": _z85(seed)}}

    @classmethod
    def from_layout(cls, layout: dict, device="cuda", **kw):
        lay = dict(layout)
        seed = lay.pop("shuffle_seed")
        if isinstance(seed, dict):
            seed = _unz85(seed["cpu_rng_state_hex"])
        return cls(ste_binary=True, device=device, shuffle_seed=seed, **lay, **kw)

    # ------------------------------------------------------------------------------------------ helpers
    def query_binary_vxl(self, points_n_orig, binary_vxl, n, return_overlap_area=False):
        """utils_bpp_acc.py:404-415 (K6)."""
        N = points_n_orig.shape[0]
        mask = torch.zeros(N, dtype=torch.int16, device=points_n_orig.device)
        overlap = torch.zeros(N, dtype=torch.int32, device=points_n_orig.device)
        pack_and_align.query_mask_3D(points_n_orig.contiguous(), binary_vxl.squeeze(0).contiguous(), mask, overlap, self.res[n], N)
        mask = mask.to(torch.bool)
        return (mask, overlap) if return_overlap_area else mask

    def query_binary_vxl_qlist(self, points_n_orig_list, binary_vxl, n_list, return_overlap_area=False):
        """utils_bpp_acc.py:417-428 (K7)."""
        N = points_n_orig_list.shape[0]
        mask = torch.zeros(N, dtype=torch.int16, device=points_n_orig_list.device)
        overlap = torch.zeros(N, dtype=torch.int32, device=points_n_orig_list.device)
        pack_and_align.query_mask_3D_qlist(points_n_orig_list.contiguous(), binary_vxl.squeeze(0).contiguous(), mask, overlap,
                                           self.resolutions_list[n_list].contiguous(), N)
        mask = mask.to(torch.bool)
        return (mask, overlap) if return_overlap_area else mask

    def fetch_2D_batches(self, binary_vxl_2D, n):
        """utils_bpp_acc.py:431-456: every plane vertex inside an occupied 2D cell (+1-vertex halo):
        (row index in the level, normalised coordinates)."""
        Rb = binary_vxl_2D.shape[-1]
        scale = self.res_2D[n] - 2
        assert scale % Rb == 0
        T = scale // Rb
        cells = self.binary_vxl_2D_idx.view(-1, 2)[binary_vxl_2D.reshape(-1)].to(torch.int64) * T      # [bs, 2]
        o = torch.arange(0, T + 2, device=cells.device)
        offsets = torch.stack(torch.meshgrid(o, o, indexing="ij"), -1).view(1, -1, 2)                  # [1, (T+2)^2, 2]
        points_n_orig = (cells[:, None, :] + offsets).view(-1, 2)
        indexes_2D = get_grid_index(self.offs_2D[n + 1] - self.offs_2D[n], self.res_2D[n], points_n_orig)
        points_n = (points_n_orig.to(torch.float32) - 0.5) / float(scale)
        return indexes_2D, points_n

    def get_STE_params(self, Encoding, mode="ste_binary"):
        assert mode in ("ste_binary", "ste_multistep", "add_noise")
        p = Encoding.params
        if mode == "ste_binary":
            return STE_binary.apply(p)
        if mode == "ste_multistep":
            return STE_multistep.apply(p, self.Q)
        return p + (torch.rand_like(p) - 0.5) * (1 / self.Q)

    def get_BiRF_wentropy_leveln(self, params_q, n, offsets=None):
        """utils_bpp_acc.py:472-486: level-wide +1 frequency and its zeroth-order entropy."""
        offsets = self.offs if offsets is None else offsets
        p = params_q[offsets[n]:offsets[n + 1]]
        ttl = p.numel()
        s = torch.sum(p)
        pos, neg = (ttl + s) / 2.0, (ttl - s) / 2.0
        Pg = pos / ttl
        return Pg, pos * (-torch.log2(Pg)) + neg * (-torch.log2(1 - Pg)), ttl

    def _bits_sum(self, x, p):
        """torch.sum(self.entropy_model(x, p)) -- fused on CUDA (`fused_entropy = False`: the op-by-op expression)"""
        if x.is_cuda and getattr(self, "fused_entropy", True) and x.shape == p.shape:
            return _BernoulliBitsSum.apply(x, p)
        return torch.sum(self.entropy_model(x, p))

    def level_entropies(self, params_q, offsets=None):
        """get_BiRF_wentropy_leveln for every level at once (same arithmetic, utils_bpp_acc.py:472-486): ([Pg_n], [bit_n])"""
        offsets = self.offs if offsets is None else offsets
        F = params_q.shape[-1]
        _, ttl = _level_const(offsets, F, params_q.device)
        s = _LevelSums.apply(params_q, offsets, bool(getattr(self, "ste_binary", False)))
        pos, neg = (ttl + s) / 2.0, (ttl - s) / 2.0
        Pg = pos / ttl
        bits = pos * (-torch.log2(Pg)) + neg * (-torch.log2(1 - Pg))
        return list(Pg.unbind(0)), list(bits.unbind(0))

    def init_binary_vxl_coords(self, scale=512):
        t = scale // self.binary_vxl_len
        resolution = scale + 2
        dev = self.resolutions_list.device
        self.idx_coord_base = my_meshgrid3D(-1, t + 1, dev).view(1, -1, 3)
        self.pn_frac_offsets_list = torch.tensor([0, resolution * resolution], device=dev, dtype=torch.int32)
        self.pn_frac_resolutions_list = torch.tensor([resolution], device=dev, dtype=torch.int32)
        self.pn_frac_resolutions_list_host = resolution

    def get_idx_coords2(self, binary_vxl, resolution=None):
        """utils_bpp_acc.py:498-512: all finest-level voxels inside occupied occupancy cells (+halo), unique."""
        resolution = self.res[-1] if resolution is None else resolution
        t = (resolution - 2) // self.binary_vxl_len
        occ = binary_vxl.squeeze(0).nonzero().to(torch.int32).contiguous()                          # [N, 3] occupied cells
        c = (occ[:, None, :] * t + self.idx_coord_base + 1).reshape(-1, 3).to(torch.int64)
        key = torch.unique(c[:, 0] * resolution * resolution + c[:, 1] * resolution + c[:, 2])
        return torch.stack([key // (resolution * resolution), (key // resolution) % resolution, key % resolution], -1)

    def get_pn_embed_frac(self, embeddings_3D_q, idx_coords2, resolution=None, axis="xy"):
        """utils_bpp_acc.py:515-530: +1 vote fraction plane, zero padded to [res*res, F]."""
        resolution = self.res[-1] if resolution is None else resolution
        frac = _cnt_np_embed.apply(idx_coords2, embeddings_3D_q, resolution, 2 ** self.log2_hashmap_size, axis)[..., 0]
        frac = nnf.pad(frac.permute(2, 0, 1).unsqueeze(0), pad=[1, 1, 1, 1]).squeeze(0).permute(1, 2, 0).contiguous()
        return frac.view(-1, self.n_features)

    def _vote3_ready(self):
        """the fused vote path needs F == 8, the finest level's inverse hash table with every row present (row index ==
        position) and an occupancy grid that divides the finest grid"""
        if self.n_features != 8 or self.Pg_level != self.n_levels or (self.res[-1] - 2) % self.binary_vxl_len or self.binary_vxl_len > 128:
            return False
        ok = getattr(self, "_vote3_ok", None)
        if ok is None:
            rows = self.unique_value_list[-1]
            T = self.offs[-1] - self.offs[-2]
            ok = self._vote3_ok = bool(rows.numel() == T and torch.equal(rows, torch.arange(T, device=rows.device)))
        return ok

    def _vote_member_table(self, vx):
        """finest-level voxels of the dimension-wise context's vote list (get_idx_coords2, utils_bpp_acc.py:498-512), grouped
        by table row: (pts int16 [M,3], seg int64 [T+1]).  Built per occupancy grid (the reference rebuilds its list every
        `step_update` steps too) by the pruned key builder with the vote predicate; the backward of the vote planes then
        walks ~15 % of the level's voxels instead of all res^3."""
        key = (vx.data_ptr(), vx._version, tuple(vx.shape))
        cache = getattr(self, "_vote_table", None)
        if cache is None or cache[0] != key:
            dev, r, T = vx.device, self.res[-1], self.offs[-1] - self.offs[-2]
            counter = torch.zeros(1, dtype=torch.int64, device=dev)
            args = (r, T, ptr(vx), vx.shape[-1], None)
            check(lib().cnc_level_pruned_keys(*args, None, ptr(counter), 1, stream()))
            M = int(counter.item())
            keys = torch.empty(M, dtype=torch.int64, device=dev)
            counter.zero_()
            check(lib().cnc_level_pruned_keys(*args, ptr(keys), ptr(counter), 1, stream()))
            keys, _ = torch.sort(keys)
            pts = torch.empty(M, 3, dtype=torch.int16, device=dev)
            row = torch.empty(M, dtype=torch.int32, device=dev)
            check(lib().cnc_keys_to_points(ptr(keys), M, r, ptr(pts), ptr(row), stream()))
            del keys
            seg = torch.searchsorted(row, torch.arange(T + 1, dtype=torch.int32, device=dev)).to(torch.int64)
            cache = self._vote_table = (key, pts, seg.contiguous(), vx)   # (vx kept alive: the key stays sound)
        return cache[1], cache[2]

    def get_pn_embed_frac3(self, embeddings_3D_q, binary_vxl, axes=("xy", "xz", "yz")):
        """{"xy","xz","yz"} -> +1 vote fraction plane, zero padded to [res*res, F]: what `get_idx_coords2` followed by
        three `get_pn_embed_frac` calls compute (utils_bpp_acc.py:498-530), without the voxel list.  `axes`: the planes wanted."""
        names = ("xy", "xz", "yz")
        axes = tuple(a for a in names if a in axes)
        vx = binary_vxl.squeeze(0)
        vx = (vx if vx.dtype in (torch.bool, torch.uint8) else vx != 0).contiguous()
        if not (self._vote3_ready() and vx.shape[-1] == self.binary_vxl_len):
            idx = self.get_idx_coords2(binary_vxl)
            return {a: self.get_pn_embed_frac(embeddings_3D_q, idx, axis=a) for a in axes}
        res = self.res[-1]
        if torch.is_grad_enabled() and embeddings_3D_q.requires_grad:
            pts_by_row, seg = self._vote_member_table(vx)      # the backward's voxel list, pruned to the vote list itself
        else:
            pts_by_row, seg = None, None                       # (no backward: the forward reads the occupancy grid only)
        fr = _cnt_np_embed3.apply(embeddings_3D_q, vx, pts_by_row, seg, res, self.offs[-1] - self.offs[-2],
                                  tuple(names.index(a) for a in axes), pts_by_row is not None)
        out = {}
        for a, f in zip(axes, fr):
            f = nnf.pad(f[..., 0].permute(2, 0, 1).unsqueeze(0), pad=[1, 1, 1, 1]).squeeze(0).permute(1, 2, 0).contiguous()
            out[a] = f.view(-1, self.n_features)
        return out

    @staticmethod
    def _planes(binary_vxl):
        b = binary_vxl.squeeze(0)
        return {"xy": torch.any(b, dim=2), "xz": torch.any(b, dim=1), "yz": torch.any(b, dim=0)}

    def _chunks(self, n):
        """stream chunking of a context-coded level (utils_bpp_acc.py:798-802)."""
        per = int(self.MAX_POINTS_NUM_TO_OOM // self.utils_points_per_param_levels[n])
        total = int(self.hashparams_num_levels[n])
        per = min(per, total)
        return [(s, min(s + per, total)) for s in range(0, total, per)]

    def _mlp3d_packed(self):
        m = self.context_model_3D
        parts = [m[0].weight.t(), m[0].bias, m[2].weight.t(), m[2].bias, m[4].weight.t(), m[4].bias]
        return torch.cat([p.detach().contiguous().float().reshape(-1) for p in parts]).contiguous()

    # ---------------------------------------------------------------------------- probabilities, 3D
    def _probs_3D(self, Encoding_xyz, table, binary_vxl, n, lo, hi, Pg_n):
        """P(+1) of the entries [lo, hi) of level n -> (prob [E,F] of the existing entries, mask_exist [hi-lo])."""
        if self.fused and self.n_features == 8 and self.max_context_layer_num == 3 and n >= 3 and Encoding_xyz.ste_binary:
            if self.tables == "pruned":
                return self._probs_3D_pruned(Encoding_xyz, table, binary_vxl, n, lo, hi, Pg_n)
            return self._probs_3D_fused(Encoding_xyz, table, binary_vxl, n, lo, hi, Pg_n)
        self._ensure_full_tables()
        return self._probs_3D_unfused(Encoding_xyz, table, binary_vxl, n, lo, hi, Pg_n)

    def _probs_3D_pruned(self, Encoding_xyz, table, binary_vxl, n, lo, hi, Pg_n):
        """`_probs_3D_fused` on the occupancy-pruned vertex list: only the entries of [lo, hi) that exist are handed to the
        kernel, each with exactly the vertices the reference keeps after its mask (utils_bpp_acc.py:811-833)"""
        vx = binary_vxl.squeeze(0).contiguous()
        if vx.dtype != torch.bool and vx.dtype != torch.uint8:
            vx = vx != 0
        pts, ent, seg, ent_h, seg_h = self._pruned_level(n, vx)
        a, b = int(np.searchsorted(ent_h, lo)), int(np.searchsorted(ent_h, hi))
        E, dev = b - a, pts.device
        exist = torch.zeros(hi - lo, dtype=torch.bool, device=dev)
        prob = torch.empty(E, 8, device=dev)
        if E == 0:
            return prob, exist
        exist[ent[a:b] - lo] = True
        seg_base = int(seg_h[a])
        ex = torch.empty(E, dtype=torch.uint8, device=dev)
        vbits, vbit_off = self._vertex_bits(Encoding_xyz, vx)
        check(lib().cnc_context3d_probs(ptr(pts[seg_base:int(seg_h[b])]), ptr(seg[a:b + 1].contiguous()), E, ptr(vx), vx.shape[-1],
                                        ptr(self._sign_bits(table)), ptr(Encoding_xyz.offsets_list), ptr(Encoding_xyz.resolutions_list),
                                        n, float(Pg_n), ptr(self._mlp3d_packed()), ptr(prob), None, ptr(ex), seg_base, a,
                                        ptr(vbits), ptr(vbit_off), stream()))
        return prob, exist

    def _vertex_bits(self, Encoding_xyz, vx):
        """per-vertex occupancy predicate of all levels as bitmaps (cnc_vertex_valid_bits); rebuilt when the
        occupancy grid changes"""
        key = (vx.data_ptr(), vx._version, tuple(vx.shape))
        if getattr(self, "_vbits_key", None) != key:
            offs, off = [], 0
            for r in self.res:
                offs.append(off)
                off += (r ** 3 + 31) // 32 * 32
            offs.append(off)
            bit_off = torch.tensor(offs, dtype=torch.int64, device=vx.device)
            words = torch.empty(off // 32, dtype=torch.int32, device=vx.device)
            check(lib().cnc_vertex_valid_bits(ptr(vx), vx.shape[-1], ptr(Encoding_xyz.resolutions_list), self.n_levels,
                                              ptr(bit_off), off, ptr(words), stream()))
            self._vbits, self._vbits_off, self._vbits_key = words, bit_off, key
            self._vbits_keep = vx   # (a live reference: the storage cannot be recycled for another grid while the key is trusted)
        return self._vbits, self._vbits_off

    def _sign_bits(self, table):
        """1-bit plane of the (partially decoded) table; repacked only when the table changed"""
        t = table.detach()
        key = (t.data_ptr(), t._version, t.numel())
        if getattr(self, "_sbits_key", None) != key:
            self._sbits, self._sbits_key, self._sbits_keep = _backend.sign_pack(t.contiguous()), key, t
        return self._sbits

    def _probs_3D_fused(self, Encoding_xyz, table, binary_vxl, n, lo, hi, Pg_n):
        cs = self.unique_count_cumsum_list[n]
        seg = cs[lo:hi + 1].contiguous()
        seg_base, seg_end = self._cs_host(n, lo), self._cs_host(n, hi)
        pts = self.pos_grid_sorted_list[n][seg_base:seg_end]
        E = hi - lo
        dev = pts.device
        bits = self._sign_bits(table)
        prob = torch.empty(E, 8, device=dev)
        exist = torch.empty(E, dtype=torch.uint8, device=dev)
        vx = binary_vxl.squeeze(0).contiguous()
        if vx.dtype != torch.bool and vx.dtype != torch.uint8:
            vx = vx != 0
        vbits, vbit_off = self._vertex_bits(Encoding_xyz, vx)
        check(lib().cnc_context3d_probs(ptr(pts), ptr(seg), E, ptr(vx), vx.shape[-1], ptr(bits),
                                        ptr(Encoding_xyz.offsets_list), ptr(Encoding_xyz.resolutions_list), n, float(Pg_n),
                                        ptr(self._mlp3d_packed()), ptr(prob), None, ptr(exist), seg_base, lo,
                                        ptr(vbits), ptr(vbit_off), stream()))
        ex = exist.bool()
        return prob[ex], ex

    def _cs_host(self, n, i):
        """unique_count_cumsum_list[n][i] as a python int without a device sync (the tables are static)"""
        h = getattr(self, "_cs_host_cache", None)
        if h is None:
            h = self._cs_host_cache = {}
        if n not in h:
            h[n] = self.unique_count_cumsum_list[n].cpu()
        return int(h[n][i])

    def _probs_3D_unfused(self, Encoding_xyz, table, binary_vxl, n, lo, hi, Pg_n):
        """the reference's op-by-op flow (utils_bpp_acc.py:803-852) on the drop-in kernels"""
        cs = self.unique_count_cumsum_list[n]
        points_n_orig = self.pos_grid_sorted_list[n][int(cs[lo]):int(cs[hi])]
        points_n = (points_n_orig - 0.5) / self.scales_list[n, :]
        mask, overlap = self.query_binary_vxl(points_n_orig, binary_vxl, n, return_overlap_area=True)
        points_n = points_n[mask]
        unique_cnt = self.unique_count_list[n][lo:hi]
        mask_packed = align_and_pack.apply(mask.unsqueeze(-1).to(torch.float), unique_cnt, 0)
        mask_cnt = torch.sum(mask_packed[:, :, 0], dim=1).to(torch.long)
        mask_exist = mask_cnt > 0
        mask_cnt = mask_cnt[mask_exist]
        overlap = torch.clamp(overlap[mask], min=1)
        ov_packed = align_and_pack.apply(overlap.unsqueeze(-1).to(torch.float), mask_cnt, 0)
        ov_packed = ov_packed / torch.sum(ov_packed, dim=1, keepdim=True)
        c = min(n, self.max_context_layer_num)
        context = Encoding_xyz(points_n, n - c, n, outspace_params=table, binary_vxl=binary_vxl.squeeze(0), PV=0)
        context = torch.cat([context, Pg_n.reshape(1, 1).expand(context.shape[0], 1)], dim=-1)
        mean = align_and_pack.apply(self.context_model_3D(context), mask_cnt, 0.0)
        if self.use_overlap_area_pool:
            mean = torch.sum(mean * ov_packed, dim=1)
        else:
            mean = torch.sum(mean, dim=1) / mask_cnt.unsqueeze(-1)
        return torch.clamp(mean, min=1e-6, max=1 - 1e-6), mask_exist

    # ---------------------------------------------------------------------------- probabilities, 2D
    def _probs_2D(self, Encoding_2D, table_2D, binary_vxl_2D, n, pn_embed_frac, Pg_n, batch=None, differentiable=False):
        """plane level n >= 1 -> (mean [U,F] unclamped, rows [U] absolute).  utils_bpp_acc.py:724-746 == :895-916"""
        if batch is None:
            indexes_2D, points_n = self.fetch_2D_batches(binary_vxl_2D, n)
            _, indices_2D = torch.sort(indexes_2D)
            unique_value_2D, unique_cnt_2D = torch.unique_consecutive(indexes_2D[indices_2D], return_counts=True)
            unique_value_2D = unique_value_2D + self.offs_2D[n]
        else:
            points_n, indices_2D, unique_value_2D, unique_cnt_2D = batch
        c = min(n, self.max_context_layer_num)
        if (self.n_features == 8 and Encoding_2D.ste_binary and points_n.is_cuda and getattr(self, "fused_gather2d", True)):
            # one kernel: c plane levels | vote fraction plane | Pg  ->  [N, K]   (csrc/context_train.cu)
            vx2 = (binary_vxl_2D if binary_vxl_2D.dtype in (torch.bool, torch.uint8) else binary_vxl_2D != 0).contiguous()
            own = table_2D is None
            frac = pn_embed_frac if self.use_dimension_wise else None
            if frac is not None and not differentiable:
                frac = frac.detach()
            context = _Ctx2DGather.apply(Encoding_2D.params if own else table_2D, frac, Pg_n, points_n.contiguous().float(),
                                         Encoding_2D, n, c, vx2, int(self.pn_frac_resolutions_list_host), own)
        else:
            context = Encoding_2D(points_n, n - c, n, outspace_params=table_2D, binary_vxl=binary_vxl_2D, PV=0)
            Pg_col = Pg_n.reshape(1, 1).expand(context.shape[0], 1)
            if self.use_dimension_wise:
                context_pn = Encoding_2D.forward_given_params(points_n, self.pn_frac_offsets_list, self.pn_frac_resolutions_list,
                                                              pn_embed_frac, binary_vxl_2D)
                if not differentiable:
                    context_pn = context_pn.detach()
                context = torch.cat([context, context_pn, Pg_col], dim=-1)
            else:
                context = torch.cat([context, Pg_col], dim=-1)
        # index_select by the sort permutation + per-row sum (utils_bpp_acc.py:741-745) as one segment reduction
        cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=points_n.device), torch.cumsum(unique_cnt_2D, 0)])
        lin = self.context_model_2D[n - 1][0]
        mean = segment_sum.apply(_linear8(lin, context) if getattr(self, "fused_lin8", True) else lin(context), cs, None,
                                 indices_2D.contiguous())
        return mean / unique_cnt_2D.unsqueeze(-1), unique_value_2D, (points_n, indices_2D, unique_value_2D, unique_cnt_2D)

    # ------------------------------------------------------------------------------------------ loss
    def set_data_parallel(self, rank: int = 0, world: int = 1) -> None:
        """share the rate term among `world` data-parallel ranks whose gradients are AVERAGED afterwards:
        * the sampled 3D entries: every rank takes sample_num / world of them, its random windows shifted by rank / world of
          the level -- together the ranks still sample `sample_num` entries per step, and each rank's scaled estimate of the
          level's bits is unbiased, so the average is;
        * the plane terms (axis, level): exact sums, dealt round-robin; a rank multiplies its terms by `world`, so the average
          over the ranks is the full sum;
        * everything cheap and exact (the terms of uncoded levels) stays replicated.
        rank 0 of world 1 (the default) is the reference's single-process loss."""
        self.dp_rank, self.dp_world = int(rank), int(world)
        if world <= 1:
            self._dp_snl = None
            return
        snl = torch.round(self.hashparams_num_levels * ((self.sample_num / world) / self.hashparams_num_levels.sum())).to(torch.long)
        snl = torch.minimum(torch.clamp(snl, min=1), self.hashparams_num_levels)
        coded = [n for n in range(self.n_levels) if n not in self.skip_levels_3D and n < self.Pg_level]
        self._dp_snl = (snl, int(sum(int(snl[n]) for n in coded)))

    def forward_binary_vxl_mixPg_3D2D(self, Encoding_xyz, Encoding_xy, Encoding_xz, Encoding_yz, binary_vxl=None,
                                      verbose=False, sample_num=None, step=0, mb_as_tensor=False):
        """Rate term of the training loss: (bits per parameter, MB).  utils_bpp_acc.py:533-706
        `mb_as_tensor` (not in the reference): the MB figure stays a device tensor -- as a Python float it is a host read that
        waits for the whole forward."""
        pq = {k: self.get_STE_params(E) for k, E in (("xy", Encoding_xy), ("xz", Encoding_xz), ("yz", Encoding_yz), ("xyz", Encoding_xyz))}
        refresh = step % self.step_update == 0 or self.idx_coords2_tmp is None
        if refresh:   # the reference caches the voxel list for step_update steps (:541-543): keep the occupancy it stands for
            self.idx_coords2_tmp = binary_vxl.clone()
        planes = self._planes(binary_vxl)
        if refresh:
            self.batched_inputs_list = {}
        ttl_bit_sum, ttl_num_sum = 0, 0
        rank, world = getattr(self, "dp_rank", 0), getattr(self, "dp_world", 1)
        coded_2D = [n for n in range(self.n_levels_2D) if not (n in self.skip_levels_2D or n >= self.Pg_level_2D)]
        # (data parallel: the coded plane terms are dealt round-robin over the ranks, see set_data_parallel)
        mine = {(a, n) for t, (a, n) in enumerate((a, n) for a in ("xy", "xz", "yz") for n in coded_2D) if t % world == rank}
        finest = pq["xyz"][self.offs[-2]:self.offs[-1]]
        pns = self.get_pn_embed_frac3(finest, self.idx_coords2_tmp, axes={a for a, _ in mine}) if (self.use_dimension_wise and mine) else {}
        for axis, Enc in (("xy", Encoding_xy), ("xz", Encoding_xz), ("yz", Encoding_yz)):
            pn = pns.get(axis)
            Pgs_2D, bits_2D = self.level_entropies(pq[axis], self.offs_2D)
            means, rows_l = [], []
            for n in range(self.n_levels_2D):
                Pg_n, bit_n = Pgs_2D[n], bits_2D[n]
                if n in coded_2D:
                    if (axis, n) not in mine:
                        continue
                    mean, rows, batch = self._probs_2D(Enc, None, planes[axis], n, pn, Pg_n,
                                                       self.batched_inputs_list.get((axis, n)), differentiable=True)
                    self.batched_inputs_list[(axis, n)] = batch
                    means.append(mean)
                    rows_l.append(rows)
                else:
                    ttl_bit_sum = ttl_bit_sum + bit_n
            if means:   # one gather of the coded rows of the plane (one index backward instead of one per level)
                plane_bits = self._bits_sum(_RowsGather.apply(pq[axis], torch.cat(rows_l)), torch.cat(means, 0))
                ttl_bit_sum = ttl_bit_sum + (plane_bits * world if world > 1 else plane_bits)
            ttl_num_sum += pq[axis].numel()

        if sample_num is not None:
            snl = torch.round(self.hashparams_num_levels * (sample_num / self.hashparams_num_levels.sum())).to(torch.long)
            snl = self.hashparams_num_levels if snl[-1] > self.hashparams_num_levels[-1] else snl
            n_valid = sum(int(snl[n]) for n in range(self.n_levels) if n not in self.skip_levels_3D and n < self.Pg_level)
        elif world > 1 and getattr(self, "_dp_snl", None) is not None:
            snl, n_valid = self._dp_snl
        else:
            snl, n_valid = self.sample_num_levels, self.ttl_sample_num_valid_levels
        u = torch.rand_like(self.utils_rand)
        if world > 1:      # this rank's windows: the same draw on every rank lands rank / world of the level further on
            u = torch.remainder(u + rank / world, 1.0)
        start = torch.round((self.hashparams_num_levels - snl) * u).to(torch.long)
        start, snl_h = torch.stack([start, snl.to(torch.long)]).tolist()   # one device->host read for both
        fast = (self.n_features == 8 and self.max_context_layer_num == 3 and Encoding_xyz.ste_binary
                and getattr(self, "fused_gather_train", True) and min([n for n in range(self.n_levels) if n not in self.skip_levels_3D] + [99]) >= 3)
        Pgs_3D, bits_3D = self.level_entropies(pq["xyz"])
        if fast and pq["xyz"].is_cuda and getattr(self, "cached_selection_train", True):
            # The vertices of the sampled entries that touch the occupancy, and their pooling weights, depend on the occupancy
            # grid only: they are the pruned vertex lists of section 3.7 (built once per grid, i.e. every `step_update` steps in
            # the training loop) -- a window of entries is a SLICE of them.  The reference (and the path below) enumerates,
            # masks (K7) and compacts ~8.6 M candidate vertices every step to arrive at the same lists.
            vx = binary_vxl.squeeze(0)
            vx = (vx if vx.dtype in (torch.bool, torch.uint8) else vx != 0).contiguous()
            pts_l, w_l, n_l, cnt_l, val_l = [], [], [], [], []
            for n in range(self.n_levels):
                if n in self.skip_levels_3D or n >= self.Pg_level:
                    ttl_bit_sum = ttl_bit_sum + bits_3D[n]
                    continue
                lo, hi = start[n], start[n] + int(snl_h[n])
                pts_n, ent, seg, ent_h, seg_h = self._pruned_level(n, vx)
                w_n, lev_n = self._pruned_weights(n, vx, binary_vxl)
                a_, b_ = int(np.searchsorted(ent_h, lo)), int(np.searchsorted(ent_h, hi))
                v0, v1 = int(seg_h[a_]), int(seg_h[b_])
                pts_l.append(pts_n[v0:v1])
                w_l.append(w_n[v0:v1])
                n_l.append(lev_n[v0:v1])
                cnt_l.append(seg[a_ + 1:b_ + 1] - seg[a_:b_])
                val_l.append(self.unique_value_list[n][ent[a_:b_]] + self.offs[n])
            if pts_l and sum(p.shape[0] for p in pts_l):
                pts, w, nl, cnt, rows3 = (torch.cat(t, 0) for t in (pts_l, w_l, n_l, cnt_l, val_l))
                vals = _RowsGather.apply(pq["xyz"], rows3)      # (a window of distinct entries per level)
                cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=pts.device), torch.cumsum(cnt, 0)])
                vbits, vbit_off = self._vertex_bits(Encoding_xyz, vx)
                ctx_in = _Ctx3DGather.apply(Encoding_xyz.params, torch.stack(Pgs_3D), pts, nl, Encoding_xyz, vbits, vbit_off)
                m3 = self.context_model_3D
                if getattr(self, "fused_mlp_train", True):
                    mlp_out = _CtxMLP3.apply(ctx_in, m3[0].weight, m3[0].bias, m3[2].weight, m3[2].bias, m3[4].weight, m3[4].bias)
                else:
                    mlp_out = m3(ctx_in)
                mean = segment_sum.apply(mlp_out, cs, w, None)
                bits = self._bits_sum(vals, mean)
                ttl_bit_sum = ttl_bit_sum + bits / n_valid * self.ttl_hashparams_num_valid_levels
            ttl_num_sum += pq["xyz"].numel()
            mb = ttl_bit_sum.detach() / 8 / 1024 / 1024
            return ttl_bit_sum / ttl_num_sum, (mb if mb_as_tensor else float(mb))
        self._ensure_full_tables()
        pts_l, ptsn_l, Pg_l, n_l, cnt_l, val_l = [], [], [], [], [], []
        for n in range(self.n_levels):
            Pg_n, bit_n = Pgs_3D[n], bits_3D[n]
            if n in self.skip_levels_3D or n >= self.Pg_level:
                ttl_bit_sum = ttl_bit_sum + bit_n
                continue
            lo, hi = start[n], start[n] + int(snl_h[n])
            p = self.pos_grid_sorted_list[n][self._cs_host(n, lo):self._cs_host(n, hi)]
            pts_l.append(p)
            if not fast:
                ptsn_l.append((p - 0.5) / self.scales_list[n, :])
                Pg_l.append(Pg_n.reshape(1, 1).expand(p.shape[0], 1))
            n_l.append(torch.full((p.shape[0],), n, dtype=torch.int64, device=p.device))
            cnt_l.append(self.unique_count_list[n][lo:hi])
            val_l.append(self.unique_value_list[n][lo:hi] + self.offs[n])
        if pts_l:
            pts, nl, cnt, rows3 = (torch.cat(t, 0) for t in (pts_l, n_l, cnt_l, val_l))
            vals = pq["xyz"][rows3]   # one gather of the sampled entries (one index backward instead of one per level)
            mask, overlap = self.query_binary_vxl_qlist(pts, binary_vxl, nl, return_overlap_area=True)
            # voxels per entry that touch the occupancy, overlap weights normalised per entry, weighted mean of the MLP
            # outputs (utils_bpp_acc.py:553-566): segment reductions over the ragged lists instead of padded [E, M, F] tensors
            zero = torch.zeros(1, dtype=torch.int64, device=pts.device)
            cs_all = torch.cat([zero, torch.cumsum(cnt, 0)])
            mask_cnt = pack_and_align.segment_wsum(mask.to(torch.float).unsqueeze(-1), cs_all).squeeze(-1).to(torch.long)
            # boolean masks -> index lists once each (every `x[bool_mask]` is a nonzero() plus a host sync of its own)
            ex_i = (mask_cnt > 0).nonzero().squeeze(1)
            m_i = mask.nonzero().squeeze(1)
            mask_cnt, vals = mask_cnt[ex_i], vals[ex_i]
            cs = torch.cat([zero, torch.cumsum(mask_cnt, 0)])
            if self.use_overlap_area_pool:
                ov = torch.clamp(overlap[m_i], min=1).to(torch.float)
                ov_sum = pack_and_align.segment_wsum(ov.unsqueeze(-1), cs).squeeze(-1)
                w = ov / torch.repeat_interleave(ov_sum, mask_cnt, output_size=m_i.numel())
            else:
                w = torch.repeat_interleave(1.0 / mask_cnt.to(torch.float), mask_cnt, output_size=m_i.numel())
            c = self.max_context_layer_num
            if fast:   # masked 3-level gather + Pg column straight into the [M, 25] MLP input (csrc/context_train.cu)
                vx = binary_vxl.squeeze(0)
                vx = (vx if vx.dtype in (torch.bool, torch.uint8) else vx != 0).contiguous()
                vbits, vbit_off = self._vertex_bits(Encoding_xyz, vx)
                ctx_in = _Ctx3DGather.apply(Encoding_xyz.params, torch.stack(Pgs_3D), pts[m_i].contiguous(), nl[m_i].contiguous(),
                                            Encoding_xyz, vbits, vbit_off)
            else:
                ptsn, Pgc = torch.cat(ptsn_l, 0), torch.cat(Pg_l, 0)
                context = Encoding_xyz.forward_diff_levels(ptsn[m_i], (nl[m_i] - c).to(torch.int), c, binary_vxl=binary_vxl.squeeze(0), PV=1001)
                ctx_in = torch.cat([context, Pgc[m_i]], dim=-1)
            if getattr(self, "fused_mlp_train", True) and ctx_in.shape[1] == 25 and self.n_features == 8:   # the product layout
                m3 = self.context_model_3D
                mlp_out = _CtxMLP3.apply(ctx_in, m3[0].weight, m3[0].bias, m3[2].weight, m3[2].bias, m3[4].weight, m3[4].bias)
            else:
                mlp_out = self.context_model_3D(ctx_in)
            mean = segment_sum.apply(mlp_out, cs, w.contiguous(), None)
            bits = self._bits_sum(vals, mean)
            ttl_bit_sum = ttl_bit_sum + bits / n_valid * self.ttl_hashparams_num_valid_levels
        ttl_num_sum += pq["xyz"].numel()
        mb = ttl_bit_sum.detach() / 8 / 1024 / 1024
        return ttl_bit_sum / ttl_num_sum, (mb if mb_as_tensor else float(mb))

    # ------------------------------------------------------------------------------------------ encode
    @torch.no_grad()
    def encode_binary_vxl_mixPg_3D2D(self, Encoding_xyz, Encoding_xy, Encoding_xz, Encoding_yz, binary_vxl=None,
                                     filename_prefix="b", return_streams=False):
        """utils_bpp_acc.py:709-865.  Writes `<prefix>_<axis><n>.b`, `<prefix>_3D<n>.b`, `<prefix>_3D<n>_<chunk>.b`."""
        self._sbits_key = self._vbits_key = self._pruned_cache = None   # per-call caches (tensor addresses are only unique while alive)
        pq = {k: self.get_STE_params(E) for k, E in (("xy", Encoding_xy), ("xz", Encoding_xz), ("yz", Encoding_yz), ("xyz", Encoding_xyz))}
        Pgs_dict: Dict[str, torch.Tensor] = {}
        names, c1s, syms, jobs = [], [], [], []
        ttl_bit = torch.zeros((), device=pq["xyz"].device)
        main = torch.cuda.current_stream(pq["xyz"].device)
        if getattr(self, "_coder_streams", None) is None:   # one per flush: the jobs must not queue behind each other
            self._coder_streams = [torch.cuda.Stream(pq["xyz"].device) for _ in range(self.n_levels + 2)]

        def emit(name, xs, ps):
            names.append(name)
            c1s.append(tac.cdf_from_p(ps))
            syms.append(((xs + 1) // 2).to(torch.uint8).reshape(-1))

        def flush():
            """hand the streams collected so far to the coder on the side stream: it runs (one SM per stream)
            while this stream goes on computing the probabilities of the next level"""
            done = sum(len(j[0]) for j in jobs)
            if done == len(names):
                return
            side = self._coder_streams[len(jobs) % len(self._coder_streams)]
            side.wait_stream(main)
            with torch.cuda.stream(side):
                jobs.append((names[done:], tac.encode_streams_async(c1s[done:], syms[done:])))

        # The coder is one serial chain per stream, so the streams with the most symbols go first: their probabilities
        # are evaluated and their coder launched before anything else, everything shorter runs beside them.
        def longest_stream(n):
            if n in self.skip_levels_3D or n >= self.Pg_level:
                return self.offs[n + 1] - self.offs[n]
            return max(hi - lo for lo, hi in self._chunks(n))

        for n in sorted(range(self.n_levels), key=lambda n: -longest_stream(n)):
            Pg_n, bit_n, _ = self.get_BiRF_wentropy_leveln(pq["xyz"], n)
            Pgs_dict["3D" + str(n)] = Pg_n
            if n in self.skip_levels_3D or n >= self.Pg_level:
                xs = pq["xyz"][self.offs[n]:self.offs[n + 1]].reshape(-1)
                emit(f"{filename_prefix}_3D{n}.b", xs, Pg_n.expand(xs.numel()))
                ttl_bit += bit_n
                continue
            for sn, (lo, hi) in enumerate(self._chunks(n)):
                ps, mask_exist = self._probs_3D(Encoding_xyz, pq["xyz"], binary_vxl, n, lo, hi, Pg_n)
                values_q = pq["xyz"][self.unique_value_list[n][lo:hi] + self.offs[n]][mask_exist]
                ttl_bit += torch.sum(self.entropy_model(values_q, ps))
                emit(f"{filename_prefix}_3D{n}_{sn}.b", values_q.reshape(-1), ps.reshape(-1))
            flush()
        planes = self._planes(binary_vxl)
        finest = pq["xyz"][self.offs[-2]:self.offs[-1]]
        pns = self.get_pn_embed_frac3(finest, binary_vxl) if self.use_dimension_wise else {}
        for axis, Enc in (("xy", Encoding_xy), ("xz", Encoding_xz), ("yz", Encoding_yz)):
            pn = pns.get(axis)
            for n in range(self.n_levels_2D):
                Pg_n, bit_n, _ = self.get_BiRF_wentropy_leveln(pq[axis], n, self.offs_2D)
                Pgs_dict[axis + str(n)] = Pg_n
                if n in self.skip_levels_2D or n >= self.Pg_level_2D:
                    xs = pq[axis][self.offs_2D[n]:self.offs_2D[n + 1]].reshape(-1)
                    emit(f"{filename_prefix}_{axis}{n}.b", xs, Pg_n.expand(xs.numel()))
                else:
                    mean, rows, _ = self._probs_2D(Enc, None, planes[axis], n, pn, Pg_n)
                    values_q = pq[axis][rows, :]
                    bit_n = torch.sum(self.entropy_model(values_q, mean))
                    emit(f"{filename_prefix}_{axis}{n}.b", values_q.reshape(-1), torch.clamp(mean, 1e-6, 1 - 1e-6).reshape(-1))
                ttl_bit += bit_n
        flush()
        streams = [b for _, job in jobs for b in job.result()]
        coded_bits = 0
        for name, data in zip(names, streams):
            coded_bits += len(data) * 8
            if not return_streams:
                d = os.path.dirname(name)
                if d:
                    os.makedirs(d, exist_ok=True)
                with open(name, "wb") as f:
                    f.write(data)
        out = (Pgs_dict, float(ttl_bit) / 8.0 / 1024 / 1024, coded_bits / 8.0 / 1024 / 1024)
        return out + (dict(zip(names, streams)),) if return_streams else out

    # ------------------------------------------------------------------------------------------ decode
    @torch.no_grad()
    def decode_binary_vxl_mixPg_3D2D(self, Encoding_xyz, Encoding_xy, Encoding_xz, Encoding_yz, params_q_xyz_rec,
                                     params_q_xy_rec, params_q_xz_rec, params_q_yz_rec, binary_vxl=None, Pgs_dict=None,
                                     filename_prefix="b", streams=None):
        """utils_bpp_acc.py:867-999.  3D levels in order (level n is predicted from the decoded n-3..n-1), then the
        three planes (their dimension-wise context needs the decoded finest 3D level)."""
        self._sbits_key = self._vbits_key = None   # per-call caches (tensor addresses are only unique while alive)
        if getattr(self, "_pruned_cache", None) is not None and self._pruned_cache[0][0] != binary_vxl.squeeze(0).data_ptr():
            self._pruned_cache = None

        def read(name):
            if streams is not None:
                return streams[name]
            with open(name, "rb") as f:
                return f.read()

        def decode(names, ps_list):
            """one launch for the independent streams of a step -> +-1 float tensors"""
            outs = tac.decode_streams([tac.cdf_from_p(p) for p in ps_list], [read(nm) for nm in names])
            return [o.to(torch.float32) * 2 - 1 for o in outs]

        F = self.n_features
        skip = [n for n in range(self.n_levels) if n in self.skip_levels_3D or n >= self.Pg_level]
        if skip:   # zeroth-order levels do not depend on anything: one launch
            outs = decode([f"{filename_prefix}_3D{n}.b" for n in skip],
                          [Pgs_dict["3D" + str(n)].expand((self.offs[n + 1] - self.offs[n]) * F) for n in skip])
            for n, sout in zip(skip, outs):
                params_q_xyz_rec[self.offs[n]:self.offs[n + 1]] = sout.view(-1, F)
        for n in range(self.n_levels):
            Pg_n = Pgs_dict["3D" + str(n)]
            if n in skip:
                continue
            names, ps_l, rows_l = [], [], []
            for sn, (lo, hi) in enumerate(self._chunks(n)):   # chunks of a level are independent streams
                ps, mask_exist = self._probs_3D(Encoding_xyz, params_q_xyz_rec, binary_vxl, n, lo, hi, Pg_n)
                names.append(f"{filename_prefix}_3D{n}_{sn}.b")
                ps_l.append(ps.reshape(-1))
                rows_l.append((self.unique_value_list[n][lo:hi] + self.offs[n])[mask_exist])
            for sout, rows in zip(decode(names, ps_l), rows_l):
                params_q_xyz_rec[rows] = sout.view(-1, F)
        planes = self._planes(binary_vxl)
        finest = params_q_xyz_rec[self.offs[-2]:self.offs[-1]]
        recs = {"xy": params_q_xy_rec, "xz": params_q_xz_rec, "yz": params_q_yz_rec}
        axes = (("xy", Encoding_xy), ("xz", Encoding_xz), ("yz", Encoding_yz))
        pns = self.get_pn_embed_frac3(finest, binary_vxl) if self.use_dimension_wise else {a: None for a, _ in axes}
        for n in range(self.n_levels_2D):   # the three planes are independent of each other: one launch per level
            names, ps_l, where = [], [], []
            for axis, Enc in axes:
                Pg_n = Pgs_dict[axis + str(n)]
                names.append(f"{filename_prefix}_{axis}{n}.b")
                if n in self.skip_levels_2D or n >= self.Pg_level_2D:
                    rows = self.offs_2D[n + 1] - self.offs_2D[n]
                    ps_l.append(Pg_n.expand(rows * F))
                    where.append(slice(self.offs_2D[n], self.offs_2D[n + 1]))
                else:
                    mean, rows, _ = self._probs_2D(Enc, recs[axis], planes[axis], n, pns[axis], Pg_n)
                    ps_l.append(torch.clamp(mean, 1e-6, 1 - 1e-6).reshape(-1))
                    where.append(rows)
            for (axis, _), sout, w in zip(axes, decode(names, ps_l), where):
                recs[axis][w] = sout.view(-1, F)
        return params_q_xyz_rec, params_q_xy_rec, params_q_xz_rec, params_q_yz_rec
