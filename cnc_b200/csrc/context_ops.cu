// context_ops.cu -- the context-model side ops: occupancy queries of voxels, ragged<->padded
// packing, segment reductions and the 3D->2D vote planes.
//
// Reference behaviour restated: my_cuda_backen/aligner_kernel.cu:4-326 (query_mask_*),
// :413-565 (align_and_pack_*); gridencoder/src/gridencoder.cu:873-1020 (cnt_np_embed*).
//
// All of these are HBM/L2-bound byte and integer work: int16 coordinates in, a handful of
// byte probes of the 2 MiB occupancy grid (L2 resident), small integer results out.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {

// ------------------------------------------------------------------------------------------
// K6 / K7: query_mask.  One thread per voxel; coordinates are fetched as int16 (6 B / voxel,
// coalesced across the warp), results are a 2 B mask and a 4 B overlap integer.
// Roundings follow aligner_kernel.cu:171-241 (see oracle/cnc_oracle.c cnc_o_query_mask).
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
query_mask_kernel(const int16_t *__restrict__ pts, const uint8_t *__restrict__ vxl, int32_t Rb,
                  int16_t *__restrict__ mask, int32_t *__restrict__ overlap,
                  const int64_t *__restrict__ res_list, int32_t res_scalar, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float r = res_list ? (float)__ldg(res_list + i) : (float)res_scalar;
    int c[D];
#pragma unroll
    for (int d = 0; d < D; d++) c[d] = (int)pts[i * D + d];
    int32_t ov;
    const bool m = voxel_mask_overlap<D>(c, r, Rb, vxl, ov);
    mask[i] = (int16_t)m;
    overlap[i] = ov;
}

// ------------------------------------------------------------------------------------------
// K8 / K9: align_and_pack.  One thread per packed element, F fastest -> coalesced writes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
align_pack_fwd_kernel(const float *__restrict__ feat, const int64_t *__restrict__ cnt,
                      const int64_t *__restrict__ cumsum, float *__restrict__ packed, int64_t N,
                      int64_t M, int64_t F, float V) {
    const int64_t total = N * M * F;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t k = t % F, ij = t / F, j = ij % M, i = ij / M;
        packed[t] = (j + 1 > __ldg(cnt + i)) ? V : __ldg(feat + (__ldg(cumsum + i) + j) * F + k);
    }
}

__global__ void __launch_bounds__(256)
align_pack_bwd_kernel(const float *__restrict__ dpacked, const int64_t *__restrict__ cnt,
                      const int64_t *__restrict__ cumsum, float *__restrict__ dfeat, int64_t N,
                      int64_t M, int64_t F) {
    const int64_t total = N * M * F;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t k = t % F, ij = t / F, j = ij % M, i = ij / M;
        if (j + 1 > __ldg(cnt + i)) continue;
        dfeat[(__ldg(cumsum + i) + j) * F + k] = __ldg(dpacked + t);
    }
}

// out[i,k] = sum_j w[s+j] * feat[s+j,k], j ascending: one thread per (segment, feature).
__global__ void __launch_bounds__(256)
segment_wsum_kernel(const float *__restrict__ feat, const float *__restrict__ w,
                    const int64_t *__restrict__ cumsum, float *__restrict__ out, int64_t N, int64_t F) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * F) return;
    const int64_t i = t / F, k = t % F;
    const int64_t s = __ldg(cumsum + i), e = __ldg(cumsum + i + 1);
    float acc = 0.f;
    for (int64_t j = s; j < e; j++) {
        const float v = __ldg(feat + j * F + k);
        acc = __fadd_rn(acc, w ? __fmul_rn(v, __ldg(w + j)) : v);
    }
    out[t] = acc;
}

// ------------------------------------------------------------------------------------------
// K4 / K5: vote planes (cnt_np_embed).  Counts are small exact integers in fp32, so the
// order of the atomic adds does not change the result.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool vote_slot(const int16_t *__restrict__ p, uint32_t res, uint32_t F,
                                          uint32_t axis, uint32_t T, uint32_t &row, uint32_t &slot) {
    const uint32_t c[3] = {(uint32_t)(int32_t)p[0], (uint32_t)(int32_t)p[1], (uint32_t)(int32_t)p[2]};
    row = grid_row<3>(c, T, res);  // gridencoder.cu:886 (hash of the raw coords, before the border test)
#pragma unroll
    for (int d = 0; d < 3; d++)
        if (c[d] == 0u || c[d] >= res - 1u) return false;  // :895-898
    const uint32_t s = res - 2u;
    const uint32_t u = axis == 2 ? c[1] : c[0];
    const uint32_t v = axis == 0 ? c[1] : c[2];
    slot = (u - 1u) * s * F * 2u + (v - 1u) * F * 2u;  // :902-906
    return true;
}

__global__ void __launch_bounds__(256)
vote_fwd_kernel(const int16_t *__restrict__ pts, const float *__restrict__ table,
                float *__restrict__ out, uint32_t N, uint32_t res, uint32_t F, uint32_t T, uint32_t axis) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= N) return;
    uint32_t row, slot;
    if (!vote_slot(pts + (size_t)b * 3, res, F, axis, T, row, slot)) return;
    for (uint32_t ch = 0; ch < F; ch++) {
        const float v = __ldg(table + (size_t)row * F + ch);
        atomicAdd(out + slot + ch * 2u + (v > 0.9f ? 0u : 1u), 1.0f);
    }
}

__global__ void __launch_bounds__(256)
vote_bwd_kernel(const int16_t *__restrict__ pts, const float *__restrict__ table,
                const float *__restrict__ out_sum, const float *__restrict__ grad,
                float *__restrict__ grad_table, uint32_t N, uint32_t res, uint32_t F, uint32_t T,
                uint32_t axis) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= N) return;
    uint32_t row, slot;
    if (!vote_slot(pts + (size_t)b * 3, res, F, axis, T, row, slot)) return;
    const uint32_t half = slot >> 1;
    for (uint32_t ch = 0; ch < F; ch++) {
        const float gv = __frcp_rn(__ldg(out_sum + half + ch));  // 1 / sum  (:1012)
        const float v = __ldg(table + (size_t)row * F + ch);
        const float g = (v > 0.9f) ? __fmul_rn(gv, __ldg(grad + slot + ch * 2u))
                                   : __fmul_rn(-gv, __ldg(grad + slot + ch * 2u + 1u));
        atomicAdd(grad_table + (size_t)row * F + ch, g);
    }
}

static inline uint32_t gs_blocks(int64_t total) {
    const int64_t b = (total + 255) / 256;
    const int64_t cap = 148ll * 32;
    return (uint32_t)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_query_mask(const int16_t *pts, const uint8_t *binary_vxl, int32_t Rb, int16_t *mask,
                   int32_t *overlap, const int64_t *res_per_point, int32_t resolution, int64_t N,
                   int32_t D, cnc_stream_t stream) {
    if (N == 0) return CNC_OK;
    if (!pts || !binary_vxl || !mask || !overlap || Rb <= 0) { set_error("query_mask: bad argument"); return CNC_EINVAL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t blocks = div_up((uint64_t)N, 256);
    if (D == 3)
        query_mask_kernel<3><<<blocks, 256, 0, s>>>(pts, binary_vxl, Rb, mask, overlap, res_per_point, resolution, N);
    else if (D == 2)
        query_mask_kernel<2><<<blocks, 256, 0, s>>>(pts, binary_vxl, Rb, mask, overlap, res_per_point, resolution, N);
    else { set_error("query_mask: D must be 2 or 3"); return CNC_ENOTSUP; }
    return check_launch("query_mask");
}

int cnc_align_pack_fwd(const float *feat, const int64_t *cnt, const int64_t *cumsum, float *packed,
                       int64_t N, int64_t M, int64_t F, float V, cnc_stream_t stream) {
    if (N * M * F == 0) return CNC_OK;
    if (!cnt || !cumsum || !packed) { set_error("align_pack_fwd: null pointer"); return CNC_EINVAL; }
    align_pack_fwd_kernel<<<gs_blocks(N * M * F), 256, 0, static_cast<cudaStream_t>(stream)>>>(feat, cnt, cumsum, packed, N, M, F, V);
    return check_launch("align_pack_fwd");
}

int cnc_align_pack_bwd(const float *dpacked, const int64_t *cnt, const int64_t *cumsum, float *dfeat,
                       int64_t N, int64_t M, int64_t F, cnc_stream_t stream) {
    if (N * M * F == 0) return CNC_OK;
    if (!dpacked || !cnt || !cumsum || !dfeat) { set_error("align_pack_bwd: null pointer"); return CNC_EINVAL; }
    align_pack_bwd_kernel<<<gs_blocks(N * M * F), 256, 0, static_cast<cudaStream_t>(stream)>>>(dpacked, cnt, cumsum, dfeat, N, M, F);
    return check_launch("align_pack_bwd");
}

int cnc_segment_wsum(const float *feat, const float *w, const int64_t *cumsum, float *out, int64_t N,
                     int64_t F, cnc_stream_t stream) {
    if (N * F == 0) return CNC_OK;
    if (!feat || !cumsum || !out) { set_error("segment_wsum: null pointer"); return CNC_EINVAL; }
    segment_wsum_kernel<<<div_up((uint64_t)(N * F), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(feat, w, cumsum, out, N, F);
    return check_launch("segment_wsum");
}

int cnc_vote_planes_fwd(const int16_t *pts, const float *table, float *out, uint32_t N,
                        uint32_t resolution, uint32_t F, uint32_t hashmap_size, uint32_t axis,
                        cnc_stream_t stream) {
    if (N == 0) return CNC_OK;
    if (!pts || !table || !out || axis > 2 || resolution < 3) { set_error("vote_planes_fwd: bad argument"); return CNC_EINVAL; }
    vote_fwd_kernel<<<div_up(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pts, table, out, N, resolution, F, hashmap_size, axis);
    return check_launch("vote_planes_fwd");
}

int cnc_vote_planes_bwd(const int16_t *pts, const float *table, const float *out_sum,
                        const float *grad, float *grad_table, uint32_t N, uint32_t resolution,
                        uint32_t F, uint32_t hashmap_size, uint32_t axis, cnc_stream_t stream) {
    if (N == 0) return CNC_OK;
    if (!pts || !table || !out_sum || !grad || !grad_table || axis > 2 || resolution < 3) { set_error("vote_planes_bwd: bad argument"); return CNC_EINVAL; }
    vote_bwd_kernel<<<div_up(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pts, table, out_sum, grad, grad_table, N, resolution, F, hashmap_size, axis);
    return check_launch("vote_planes_bwd");
}

}  // extern "C"
