"""Drop-in for the reference's compiled module `pack_and_align` (my_cuda_backen/aligner.cpp:73-78).

`import cnc_b200.pack_and_align as pack_and_align` replaces examples/utils_bpp_acc.py:7.
"""
from __future__ import annotations

import torch

from ._lib import check, lib, need_cuda, ptr, stream


def _i(v):
    # the reference callers pass 0-d CUDA tensors where C++ takes `int` (utils_bpp_acc.py:121-125,406-410)
    return int(v.item()) if isinstance(v, torch.Tensor) else int(v)


def align_and_pack_forward(voxel_features, unique_count, unique_count_cumsum, N, M, F, V, dim):
    """aligner.cpp:4-17 / aligner_kernel.cu:437-495: returns packed [N, M, F]."""
    need_cuda(voxel_features=voxel_features, unique_count=unique_count)
    N, M, F = _i(N), _i(M), _i(F)
    if unique_count.dtype != torch.int64 or unique_count_cumsum.dtype != torch.int64:
        raise RuntimeError("unique_count / unique_count_cumsum must be int64 tensors")
    feat = voxel_features if voxel_features.dtype == torch.float32 else voxel_features.float()
    packed = torch.empty((N, M, F), dtype=torch.float32, device=feat.device)
    check(lib().cnc_align_pack_fwd(ptr(feat), ptr(unique_count), ptr(unique_count_cumsum.contiguous()),
                                   ptr(packed), N, M, F, float(V), stream()))
    return packed.to(voxel_features.dtype)


def align_and_pack_backward(dL_packed_features, voxel_features, unique_count, unique_count_cumsum, N, M, F, T, dim):
    """aligner.cpp:19-35 / aligner_kernel.cu:518-565: returns dL/dvoxel_features [T, F]."""
    need_cuda(dL_packed_features=dL_packed_features, voxel_features=voxel_features,
              unique_count=unique_count, unique_count_cumsum=unique_count_cumsum)
    N, M, F, T = _i(N), _i(M), _i(F), _i(T)
    d = torch.zeros((T, F), dtype=torch.float32, device=dL_packed_features.device)
    check(lib().cnc_align_pack_bwd(ptr(dL_packed_features.float()), ptr(unique_count), ptr(unique_count_cumsum),
                                   ptr(d), N, M, F, stream()))
    return d.to(dL_packed_features.dtype)


def _query(points, binary_vxl, mask, overlap_area_pool, res_list, res, N):
    need_cuda(points_n_orig=points, binary_vxl=binary_vxl, mask=mask, overlap_area_pool=overlap_area_pool,
              resolution_list=res_list)
    if points.dtype != torch.int16:
        raise RuntimeError("points_n_orig must be an int16 tensor")
    if mask.dtype != torch.int16 or overlap_area_pool.dtype != torch.int32:
        raise RuntimeError("mask must be int16 and overlap_area_pool int32")
    if binary_vxl.dtype not in (torch.bool, torch.uint8):
        raise RuntimeError("binary_vxl must be a bool tensor")
    D = points.shape[1]
    if binary_vxl.dim() != D:
        raise RuntimeError("binary_vxl rank must match the point dimension")
    check(lib().cnc_query_mask(ptr(points), ptr(binary_vxl), binary_vxl.shape[0], ptr(mask), ptr(overlap_area_pool),
                               ptr(res_list), res, _i(N), D, stream()))


def query_mask_3D(points_n_orig, binary_vxl, mask, overlap_area_pool, resolution, N):
    """aligner.cpp:37-51 / aligner_kernel.cu:328-367."""
    _query(points_n_orig, binary_vxl, mask, overlap_area_pool, None, _i(resolution), N)


def query_mask_3D_qlist(points_n_orig_list, binary_vxl, mask, overlap_area_pool, resolution_list, N):
    """aligner.cpp:54-70 / aligner_kernel.cu:370-409.  resolution_list: int64 [N]."""
    if resolution_list.dtype != torch.int64:
        raise RuntimeError("resolution_list must be an int64 tensor")
    _query(points_n_orig_list, binary_vxl, mask, overlap_area_pool, resolution_list, 0, N)


def segment_wsum(feat, cumsum, weights=None):
    """Extension: out[i] = sum_j w[j] * feat[j] over j in [cumsum[i], cumsum[i+1]) without the padded tensor."""
    need_cuda(feat=feat, cumsum=cumsum, weights=weights)
    N, F = cumsum.numel() - 1, feat.shape[1]
    out = torch.empty((N, F), dtype=torch.float32, device=feat.device)
    check(lib().cnc_segment_wsum(ptr(feat), ptr(weights), ptr(cumsum), ptr(out), N, F, stream()))
    return out
