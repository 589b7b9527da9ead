"""ctypes loader of libcnc_b200.so -- the only door from Python into the CUDA kernels.

There is no CPU fallback: if the library is missing or a call fails, a RuntimeError is
raised (the reference raises RuntimeError from TORCH_CHECK / std::runtime_error in the same
situations, gridencoder.cu:15-18,641,669).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcnc_b200.so")
_lib = None

_vp, _u32, _i32, _i64, _u64, _f32 = C.c_void_p, C.c_uint32, C.c_int32, C.c_int64, C.c_uint64, C.c_float

# name -> argtypes (every function returns int); must list every symbol of include/cnc_b200.h
SIGNATURES = {
    "cnc_grid_encode_fwd": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp],
    "cnc_grid_encode_bwd": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp],
    "cnc_grid_encode_bwd_rows": [_vp, _u32, _u32, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp],
    "cnc_grid_encode_fwd_bits": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp],
    "cnc_ste_binary_fwd": [_vp, _vp, _u64, _vp],
    "cnc_ste_binary_bwd": [_vp, _vp, _vp, _u64, _vp],
    "cnc_sign_pack": [_vp, _vp, _u64, _vp],
    "cnc_sign_unpack": [_vp, _vp, _u64, _vp],
    "cnc_lin8_fwd": [_vp, _vp, _vp, _vp, _i64, _i32, _vp],
    "cnc_lin8_bwd": [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp],
    "cnc_ctx3d_gather_fwd": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cnc_ctx3d_gather_bwd": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cnc_ctx2d_gather_fwd": [_vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp],
    "cnc_ctx2d_gather_bwd": [_vp, _i64, _vp, _vp, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "cnc_wavefront_begin": [_vp, _u32, _u32, _u32, _vp],
    "cnc_wavefront_march": [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _u32,
                            _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cnc_wavefront_composite": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _u32, _f32, _f32, _vp],
    "cnc_field_fwd_n": [_vp] * 14 + [_vp, _u32, _vp],
    "cnc_level_row_hist": [_u32, _u32, _vp, _vp],
    "cnc_level_pruned_keys": [_u32, _u32, _vp, _i32, _vp, _vp, _vp, _i32, _vp],
    "cnc_keys_to_points": [_vp, _u64, _u32, _vp, _vp, _vp],
    "cnc_bernoulli_bits_fwd": [_vp, _vp, _i64, _vp, _vp],
    "cnc_bernoulli_bits_bwd": [_vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "cnc_rows8_gather": [_vp, _vp, _i64, _vp, _vp],
    "cnc_rows8_scatter": [_vp, _vp, _i64, _vp, _vp],
    "cnc_level_popcount": [_vp, _vp, _i32, _vp, _vp],
    "cnc_ste_planes_pack": [_vp, _vp, _vp, _u64, _vp],
    "cnc_surrogate_fill": [_vp, _vp, _vp, _u64, _u64, _u64, _vp],
    "cnc_adam_planes": [_vp, _vp, _vp, _vp, _vp, _vp, _u64, _f32, _f32, _f32, _f32, _f32, _i64, _f32, _i32, _vp],
    "cnc_vote_planes_fwd": [_vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _vp],
    "cnc_vote_planes_bwd": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _vp],
    "cnc_vote3_fwd": [_vp, _u32, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp],
    "cnc_vote3_bwd": [_vp, _vp, _vp, _u32, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _i32, _vp],
    "cnc_query_mask": [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i64, _i32, _vp],
    "cnc_align_pack_fwd": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _f32, _vp],
    "cnc_align_pack_bwd": [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    "cnc_segment_wsum": [_vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "cnc_segment_wsum_idx": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "cnc_segment_wsum_idx_bwd": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "cnc_cdf_from_p": [_vp, _vp, _u64, _vp],
    "cnc_ac_encode": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "cnc_ac_decode": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp],
    "cnc_sh16": [_vp, _vp, _u64, C.c_int, _vp],
    "cnc_freq_embed": [_vp, _vp, _u64, C.c_int, _vp],
    "cnc_field_pack_weights": [_vp] * 12,
    "cnc_field_fwd": [_vp] * 15 + [_u32, _vp],
    "cnc_field_fwd_train": [_vp] * 19 + [_u32, _vp],
    "cnc_field_fwd_host": [_vp] * 14 + [_u32] + [_vp] * 4 + [_u32, _u32, _vp, _vp, _vp],
    "cnc_field_set_timeline_buffer": [_vp],
    "cnc_ray_aabb_intersect": [_vp, _vp, _i64, _f32, _f32, _vp, _i32, _f32, _vp, _vp, _vp, _vp],
    "cnc_traverse_grids": [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _i32,
                           _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cnc_packed_scan": [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp],
    "cnc_render_from_density": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cnc_context3d_probs": [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _i32, _f32, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp],
    "cnc_wgrad": [_vp, _u32, _u32, _vp, _u32, _u32, C.c_int, _vp, _u32, _u32, _vp],
    "cnc_dgrad_pack": [_vp, _u32, _u32, _i32, _u32, _u32, _u32, _vp, _vp],
    "cnc_dgrad": [_vp, _u32, _u32, _vp, _u32, _vp, _u32, _vp, _u32, _u32, _vp],
    "cnc_vertex_valid_bits": [_vp, _i32, _vp, _i32, _vp, _i64, _vp, _vp],
    "cnc_ctx_mlp_fwd": [_vp, _vp, _vp, _i64, _vp],
    "cnc_ctx_mlp_bwd": [_vp, _vp, _vp, _vp, _vp, _u32, _i64, _vp],
    "cnc_render_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "cnc_sample_points": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    "cnc_pack_ray_chunks": [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp],
    "cnc_compact_samples": [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "cnc_peer_alloc": [_u64, _vp],
    "cnc_peer_free": [_vp],
    "cnc_peer_export": [_vp, _vp],
    "cnc_peer_import": [_vp, _vp],
    "cnc_peer_unmap": [_vp],
    "cnc_peer_barrier": [_vp, _i32, _i32, _i32, _u32, _u32, _vp],
    "cnc_peer_min": [_vp, _i32, _i32, _i32, _u32, _u32, _vp, _u32, _vp],
    "cnc_peer_reduce": [_vp, _i32, _i64, _i64, _f32, _vp, _i32, _vp],
    "cnc_peer_push": [_vp, _i32, _i32, _vp, _vp, _i32, _vp],
}


def lib():
    """Load (once) and return the C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"cnc_b200: {LIB_PATH} is missing -- run `python -m cnc_b200.build` "
                "(there is no CPU or PyTorch fallback for the CUDA path)"
            )
        L = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        L.cnc_last_error.restype = C.c_char_p
        L.cnc_version.restype = C.c_int
        L.cnc_field_blob_floats.restype = C.c_uint32
        L.cnc_context3d_mlp_floats.restype = C.c_uint32
        L.cnc_dgrad_blob_floats.restype = C.c_uint32
        L.cnc_dgrad_blob_floats.argtypes = [_u32, _u32]
        L.cnc_wgrad_max_partials.restype = C.c_int
        L.cnc_wgrad_max_partials.argtypes = []
        L.cnc_ctx_mlp_floats.restype = C.c_uint32
        L.cnc_ctx_mlp_floats.argtypes = []
        L.cnc_ctx_mlp_max_partials.restype = C.c_int
        L.cnc_ctx_mlp_max_partials.argtypes = []
        L.cnc_lin8_rows_per_block.restype = C.c_int
        L.cnc_lin8_rows_per_block.argtypes = []
        L.cnc_bernoulli_bits_blocks.restype = C.c_int
        L.cnc_bernoulli_bits_blocks.argtypes = [_i64]
        for name in ("cnc_peer_handle_bytes", "cnc_peer_pad_bytes"):
            getattr(L, name).restype = C.c_int
            getattr(L, name).argtypes = []
        _lib = L
    return _lib


LAUNCHES = 0  # successful C-ABI compute calls (each launches exactly one kernel)


def check(rc: int) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        raise RuntimeError(f"cnc_b200 [{rc}]: {lib().cnc_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    """The current torch CUDA stream handle (the reference launches on the legacy default
    stream, gridencoder.cu:635; we follow torch's current stream so DP ranks / side streams work)."""
    # (torch.cuda.current_stream() builds a Stream object: ~15 us per call, once per kernel launch -- a millisecond per
    #  training step on the host-bound paths; the raw handle is a C call)
    return _raw_stream(_cur_device())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)
if _raw_stream is None or _cur_device is None:      # other torch builds: the documented road
    _raw_stream = lambda d: torch.cuda.current_stream(d).cuda_stream
    _cur_device = torch.cuda.current_device


def need_cuda(**tensors):
    for name, t in tensors.items():
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous tensor")
