"""Drop-in for the reference's compiled module `_gridencoder` (gridencoder/src/bindings.cpp:5-8).

Same four functions, same positional signatures (gridencoder/src/gridencoder.h:12-53), same
caller-allocates / returns-None contract, same RuntimeError behaviour on bad devices, layouts,
dtypes or unsupported (num_dim, n_features).  `import cnc_b200._gridencoder as _backend` is the
one-line change in examples/radiance_fields/ngp.py:10 and examples/utils_bpp_acc.py:8.
"""
from __future__ import annotations

import torch

from ._lib import check, lib, need_cuda, ptr, stream

_FLOATS = (torch.float32,)


def _chk_float(**ts):
    for n, t in ts.items():
        if t.dtype not in _FLOATS:
            # the reference also instantiates Half/Double kernels, but they are dead code on the
            # CNC path (no autocast anywhere); fp32 is the numerics contract.
            raise RuntimeError(f"{n} must be a float32 tensor (got {t.dtype})")


def _chk_int(**ts):
    for n, t in ts.items():
        if t.dtype != torch.int32:
            raise RuntimeError(f"{n} must be an int tensor")


def _vxl(binary_vxl):
    if binary_vxl is None:
        return None
    if binary_vxl.dtype not in (torch.bool, torch.uint8):
        raise RuntimeError("binary_vxl must be a bool tensor")
    need_cuda(binary_vxl=binary_vxl)
    return binary_vxl


def grid_encode_forward(inputs, embeddings, offsets_list, resolutions_list, outputs, N, num_dim,
                        n_features, n_levels, max_level, Rb, PV, dy_dx=None, binary_vxl=None,
                        min_level_id=None):
    """gridencoder.cu:752-806.  outputs [n_levels, N, n_features] is filled in place."""
    need_cuda(inputs=inputs, embeddings=embeddings, offsets_list=offsets_list,
              resolutions_list=resolutions_list, outputs=outputs, min_level_id=min_level_id)
    _chk_float(inputs=inputs, embeddings=embeddings, outputs=outputs)
    _chk_int(offsets_list=offsets_list, resolutions_list=resolutions_list)
    if dy_dx is not None:
        raise RuntimeError("dy_dx (input gradients) is not supported: dead path in the reference (ngp.py:58-60,84)")
    if min_level_id is not None:
        _chk_int(min_level_id=min_level_id)
    v = _vxl(binary_vxl)
    check(lib().cnc_grid_encode_fwd(ptr(inputs), ptr(embeddings), ptr(offsets_list), ptr(resolutions_list),
                                    ptr(outputs), N, num_dim, n_features, n_levels, Rb, ptr(v),
                                    ptr(min_level_id), stream()))


def grid_encode_backward(grad, inputs, embeddings, offsets_list, resolutions_list, grad_embeddings, N,
                         num_dim, n_features, n_levels, max_level, Rb, dy_dx=None, grad_inputs=None,
                         binary_vxl=None, min_level_id=None):
    """gridencoder.cu:808-866.  grad [n_levels, N, n_features]; grad_embeddings accumulated in place."""
    need_cuda(grad=grad, inputs=inputs, embeddings=embeddings, offsets_list=offsets_list,
              resolutions_list=resolutions_list, grad_embeddings=grad_embeddings, min_level_id=min_level_id)
    _chk_float(grad=grad, inputs=inputs, grad_embeddings=grad_embeddings)
    _chk_int(offsets_list=offsets_list, resolutions_list=resolutions_list)
    if dy_dx is not None:
        raise RuntimeError("dy_dx (input gradients) is not supported: dead path in the reference")
    if min_level_id is not None:
        _chk_int(min_level_id=min_level_id)
    v = _vxl(binary_vxl)
    check(lib().cnc_grid_encode_bwd(ptr(grad), ptr(inputs), ptr(offsets_list), ptr(resolutions_list),
                                    ptr(grad_embeddings), N, num_dim, n_features, n_levels, Rb, ptr(v),
                                    ptr(min_level_id), stream()))


def grid_encode_forward_bits(inputs, sign_bits, offsets_list, resolutions_list, outputs, N, num_dim,
                             n_features, n_levels, Rb, binary_vxl=None, min_level_id=None):
    """Extension (not in the reference): same gather from the 1-bit sign table of `sign_pack`."""
    need_cuda(inputs=inputs, sign_bits=sign_bits, offsets_list=offsets_list,
              resolutions_list=resolutions_list, outputs=outputs, min_level_id=min_level_id)
    _chk_float(inputs=inputs, outputs=outputs)
    _chk_int(offsets_list=offsets_list, resolutions_list=resolutions_list)
    if sign_bits.dtype != torch.uint8:
        raise RuntimeError("sign_bits must be a uint8 tensor")
    v = _vxl(binary_vxl)
    check(lib().cnc_grid_encode_fwd_bits(ptr(inputs), ptr(sign_bits), ptr(offsets_list), ptr(resolutions_list),
                                         ptr(outputs), N, num_dim, n_features, n_levels, Rb, ptr(v),
                                         ptr(min_level_id), stream()))


def cnt_np_embed(inputs, embeddings_clip, outputs, N, resolution, n_features, hashmap_size, axis):
    """gridencoder.cu:940-970.  inputs int16 [N,3]; outputs [res-2,res-2,F,2] accumulated in place."""
    need_cuda(inputs=inputs, embeddings_clip=embeddings_clip, outputs=outputs)
    _chk_float(embeddings_clip=embeddings_clip, outputs=outputs)
    if inputs.dtype != torch.int16:
        raise RuntimeError("inputs must be an int16 tensor")
    check(lib().cnc_vote_planes_fwd(ptr(inputs), ptr(embeddings_clip), ptr(outputs), N, int(resolution),
                                    n_features, int(hashmap_size), int(axis), stream()))


def cnt_np_embed_backward(inputs, embeddings_clip, outputs_sum, grad, grad_embeddings, N, resolution,
                          n_features, hashmap_size, axis):
    """gridencoder.cu:1047-1087."""
    need_cuda(inputs=inputs, embeddings_clip=embeddings_clip, outputs_sum=outputs_sum, grad=grad,
              grad_embeddings=grad_embeddings)
    _chk_float(embeddings_clip=embeddings_clip, outputs_sum=outputs_sum, grad=grad, grad_embeddings=grad_embeddings)
    if inputs.dtype != torch.int16:
        raise RuntimeError("inputs must be an int16 tensor")
    check(lib().cnc_vote_planes_bwd(ptr(inputs), ptr(embeddings_clip), ptr(outputs_sum), ptr(grad),
                                    ptr(grad_embeddings), N, int(resolution), n_features, int(hashmap_size),
                                    int(axis), stream()))


# ---- extensions used by the B200 fast paths (STE + 1-bit sign table) -------------------------
def ste_binary_forward(params, out=None):
    need_cuda(params=params)
    _chk_float(params=params)
    out = torch.empty_like(params) if out is None else out
    check(lib().cnc_ste_binary_fwd(ptr(params), ptr(out), params.numel(), stream()))
    return out


def ste_binary_backward(params, grad_out):
    need_cuda(params=params, grad_out=grad_out)
    gin = torch.empty_like(grad_out)
    check(lib().cnc_ste_binary_bwd(ptr(params), ptr(grad_out), ptr(gin), params.numel(), stream()))
    return gin


def sign_pack(params, bits=None):
    """params [rows,F] f32 -> uint8 [rows*F/8] bit-planes (bit = param >= 0)."""
    need_cuda(params=params)
    _chk_float(params=params)
    n = params.numel()
    if bits is None:
        bits = torch.empty((n // 8 + 3) // 4 * 4, dtype=torch.uint8, device=params.device)
    check(lib().cnc_sign_pack(ptr(params), ptr(bits), n, stream()))
    return bits


def sign_unpack(bits, rows, n_features):
    out = torch.empty(rows, n_features, dtype=torch.float32, device=bits.device)
    check(lib().cnc_sign_unpack(ptr(bits), ptr(out), rows * n_features, stream()))
    return out
