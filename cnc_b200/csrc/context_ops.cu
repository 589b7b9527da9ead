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

// out[i,k] = sum_j w[j] * feat[row(j), k], row(j) = idx ? idx[j] : j, j ascending over [cumsum[i], cumsum[i+1]):
// the segment reduction of the context models with the gather that precedes it (index_select by the sort permutation,
// utils_bpp_acc.py:741) folded in, and its backward (idx is a permutation: every feat row is written at most once).
__global__ void __launch_bounds__(256)
segment_wsum_idx_kernel(const float *__restrict__ feat, const int64_t *__restrict__ idx, const float *__restrict__ w,
                        const int64_t *__restrict__ cumsum, float *__restrict__ out, int64_t N, int64_t F) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * F) return;
    const int64_t i = t / F, k = t % F;
    const int64_t s = __ldg(cumsum + i), e = __ldg(cumsum + i + 1);
    float acc = 0.f;
    for (int64_t j = s; j < e; j++) {
        const float v = __ldg(feat + (idx ? __ldg(idx + j) : j) * F + k);
        acc = __fadd_rn(acc, w ? __fmul_rn(v, __ldg(w + j)) : v);
    }
    out[t] = acc;
}

__global__ void __launch_bounds__(256)
segment_wsum_idx_bwd_kernel(const float *__restrict__ gout, const int64_t *__restrict__ idx, const float *__restrict__ w,
                            const int64_t *__restrict__ cumsum, float *__restrict__ gfeat, int64_t N, int64_t F) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * F) return;
    const int64_t i = t / F, k = t % F;
    const int64_t s = __ldg(cumsum + i), e = __ldg(cumsum + i + 1);
    const float g = __ldg(gout + t);
    for (int64_t j = s; j < e; j++) gfeat[(idx ? __ldg(idx + j) : j) * F + k] = w ? __fmul_rn(g, __ldg(w + j)) : g;
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

// ------------------------------------------------------------------------------------------
// K4 / K5 without the voxel list and without atomics (cnc_vote3_fwd / cnc_vote3_bwd).
// The reference enumerates every finest-level voxel inside an occupied occupancy cell plus a one-voxel halo
// (get_idx_coords2, utils_bpp_acc.py:498-512: c = occ * t + k + 1, k in [-1, t], made unique) and lets each voxel
// atomically add to its (u, v) plane cell.  Membership of voxel c in that list is a closed form of the occupancy grid:
// per dimension the candidate cells are o in [ceil((c - t - 1) / t), floor(c / t)] (one or two), and c is listed iff
// one of the <= 8 candidate cells is occupied.
//   forward:  one thread per plane cell (u, v) walks the third axis, ORs the <= 4 candidate occupancy columns into a
//             128-bit mask once, and counts +1 / -1 votes per feature in registers -> plain stores, exact integers;
//   backward: one warp per table row walks the row's voxels in the inverse hash table (lanes stride, coalesced),
//             gathers d(fraction)/d(vote) of all three planes and reduces with a fixed shuffle tree -> plain stores.
// ------------------------------------------------------------------------------------------
// (vote_cand: common.cuh)

__global__ void __launch_bounds__(128)
vote3_fwd_kernel(const uint8_t *__restrict__ vxl, uint32_t Rb, const uint8_t *__restrict__ bits, uint32_t res, uint32_t T,
                 float *__restrict__ out_xy, float *__restrict__ out_xz, float *__restrict__ out_yz) {
    const uint32_t s = res - 2u, t = s / Rb, axis = blockIdx.z;
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x + 1u, u = blockIdx.y + 1u;   // plane cell, 1-based like the coords
    if (v > s) return;
    if ((axis == 0 ? out_xy : (axis == 1 ? out_xz : out_yz)) == nullptr) return;   // plane not asked for
    // axis 0: (u, v, w) = (x, y, z); axis 1: (x, z, y); axis 2: (y, z, x)   (gridencoder.cu:902-906)
    const uint32_t du = axis == 2 ? 1u : 0u, dv = axis == 0 ? 1u : 2u, dw = 3u - du - dv;
    const uint32_t stride[3] = {Rb * Rb, Rb, 1u};
    uint32_t ulo, uhi, vlo, vhi;
    vote_cand(u, t, Rb, ulo, uhi);
    vote_cand(v, t, Rb, vlo, vhi);
    uint32_t M[4] = {0u, 0u, 0u, 0u};   // occupancy along w of the candidate columns, ORed (Rb <= 128)
    for (uint32_t ou = ulo; ou <= uhi; ou++)
        for (uint32_t ov = vlo; ov <= vhi; ov++) {
            const uint8_t *col = vxl + (size_t)ou * stride[du] + (size_t)ov * stride[dv];
            for (uint32_t ow = 0; ow < Rb; ow++)
                if (col[(size_t)ow * stride[dw]]) M[ow >> 5] |= 1u << (ow & 31u);
        }
    uint32_t pos[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, cnt = 0u;
    // members along w: an occupied column cell o covers w in [o t, o t + t + 1] (the inverse of vote_cand); walking the set
    // bits of M visits exactly those w, in ascending order and once each, instead of testing all s positions
    uint32_t w_done = 0u;
#pragma unroll
    for (int word = 0; word < 4; word++) {
        uint32_t mbits = M[word];
        while (mbits) {
            const uint32_t o = (uint32_t)word * 32u + (uint32_t)__ffs((int)mbits) - 1u;
            mbits &= mbits - 1u;
            uint32_t w0 = o * t, w1 = o * t + t + 1u;
            if (w0 <= w_done) w0 = w_done + 1u;
            if (w1 > s) w1 = s;
            for (uint32_t w = w0; w <= w1; w++) {
                uint32_t c[3];
                c[du] = u; c[dv] = v; c[dw] = w;
                const uint32_t sb = __ldg(bits + grid_row<3>(c, T, res));
#pragma unroll
                for (int ch = 0; ch < 8; ch++) pos[ch] += (sb >> ch) & 1u;
                cnt++;
            }
            if (w1 > w_done) w_done = w1;
        }
    }
    float *out = axis == 0 ? out_xy : (axis == 1 ? out_xz : out_yz);
    float4 *o = reinterpret_cast<float4 *>(out + ((size_t)(u - 1u) * s + (v - 1u)) * 16u);
#pragma unroll
    for (int q = 0; q < 4; q++)
        o[q] = make_float4((float)pos[2 * q], (float)(cnt - pos[2 * q]), (float)pos[2 * q + 1], (float)(cnt - pos[2 * q + 1]));
}

struct Vote3BwdArgs {
    const int16_t *pts;       // inverse hash table of the level: voxel coords grouped by row
    const int64_t *seg;       // [T+1] running voxel count per row
    const uint8_t *vxl;
    const uint8_t *bits;      // sign bytes of the level's rows
    const float *grad[3];     // (d loss / d fraction) / vote sum, [s,s,8,2] per axis (xy, xz, yz)
    float *grad_table;        // [T,8]
    uint32_t Rb, res, T;
    int32_t members_only;     // pts holds vote-list members only (built with the same predicate): no test per voxel
};

__global__ void __launch_bounds__(256) vote3_bwd_kernel(const Vote3BwdArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= a.T) return;
    const uint32_t s = a.res - 2u, t = s / a.Rb;
    const int64_t v0 = __ldg(a.seg + row), v1 = __ldg(a.seg + row + 1);
    float accp[8], accn[8];   // sums of gv * grad over the row's voxels: +1 votes / -1 votes
#pragma unroll
    for (int ch = 0; ch < 8; ch++) accp[ch] = accn[ch] = 0.f;
    for (int64_t i = v0 + lane; i < v1; i += 32) {
        const uint32_t c[3] = {(uint32_t)(int32_t)__ldg(a.pts + i * 3), (uint32_t)(int32_t)__ldg(a.pts + i * 3 + 1), (uint32_t)(int32_t)__ldg(a.pts + i * 3 + 2)};
        if (!a.members_only) {
            bool ok = true;
            uint32_t lo[3], hi[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                ok = ok && c[d] != 0u && c[d] < a.res - 1u;   // gridencoder.cu:895-898
                vote_cand(c[d], t, a.Rb, lo[d], hi[d]);
            }
            if (!ok) continue;
            bool member = false;
            for (uint32_t o0 = lo[0]; o0 <= hi[0]; o0++)
                for (uint32_t o1 = lo[1]; o1 <= hi[1]; o1++)
                    for (uint32_t o2 = lo[2]; o2 <= hi[2]; o2++) member |= a.vxl[((size_t)o0 * a.Rb + o1) * a.Rb + o2] != 0;
            if (!member) continue;
        }
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            if (a.grad[ax] == nullptr) continue;   // (warp-uniform: a plane nobody differentiated)
            const uint32_t u = ax == 2 ? c[1] : c[0], v = ax == 0 ? c[1] : c[2];
            const size_t cell = (size_t)(u - 1u) * s + (v - 1u);
            // grad[ax] holds d loss / d fraction already divided by the cell's vote sum (1 / sum, :1012, folded by the caller:
            // one 64-byte read per voxel and plane instead of 96)
            const float4 *gr = reinterpret_cast<const float4 *>(a.grad[ax] + cell * 16u);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 g = __ldg(gr + q);      // (pos, neg) of channels 2q, 2q+1
                accp[2 * q] = __fadd_rn(accp[2 * q], g.x);
                accn[2 * q] = __fadd_rn(accn[2 * q], g.y);
                accp[2 * q + 1] = __fadd_rn(accp[2 * q + 1], g.z);
                accn[2 * q + 1] = __fadd_rn(accn[2 * q + 1], g.w);
            }
        }
    }
    const uint32_t sb = __ldg(a.bits + row);
    float g[8];
#pragma unroll
    for (int ch = 0; ch < 8; ch++) g[ch] = ((sb >> ch) & 1u) ? accp[ch] : -accn[ch];   // d fraction / d vote, sign of the row's own value
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1)
#pragma unroll
        for (int ch = 0; ch < 8; ch++) g[ch] = __fadd_rn(g[ch], __shfl_xor_sync(0xFFFFFFFFu, g[ch], sh));
    if (lane < 8) {
        float m = 0.f;
#pragma unroll
        for (int ch = 0; ch < 8; ch++) if (ch == (int)lane) m = g[ch];
        a.grad_table[(size_t)row * 8u + lane] = m;
    }
}

// The same reduction for a list that holds vote-list members only (no test per voxel): FOUR lanes per voxel, lane q loading the
// q-th 16 bytes of the voxel's 64-byte cell record of each plane -- a warp instruction then touches 16 sectors instead of 32 (the
// kernel is bound by L1 wavefronts, not by DRAM: the planes are L2 resident) -- and accumulating only the half (+1 votes or -1
// votes) that the row's own sign selects, for its two channels.
__global__ void __launch_bounds__(256) vote3_bwd_members_kernel(const Vote3BwdArgs a) {
    const uint32_t lane = threadIdx.x & 31u, q = lane & 3u, sub = lane >> 2;
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= a.T) return;
    const uint32_t s = a.res - 2u;
    const int64_t v0 = __ldg(a.seg + row), v1 = __ldg(a.seg + row + 1);
    const uint32_t sb = __ldg(a.bits + row);
    const bool p0 = (sb >> (2u * q)) & 1u, p1 = (sb >> (2u * q + 1u)) & 1u;
    float acc0 = 0.f, acc1 = 0.f;
    for (int64_t i = v0 + sub; i < v1; i += 8) {
        const uint32_t c[3] = {(uint32_t)(int32_t)__ldg(a.pts + i * 3), (uint32_t)(int32_t)__ldg(a.pts + i * 3 + 1), (uint32_t)(int32_t)__ldg(a.pts + i * 3 + 2)};
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            if (a.grad[ax] == nullptr) continue;
            const uint32_t u = ax == 2 ? c[1] : c[0], v = ax == 0 ? c[1] : c[2];
            const size_t cell = (size_t)(u - 1u) * s + (v - 1u);
            const float4 g = __ldg(reinterpret_cast<const float4 *>(a.grad[ax] + cell * 16u) + q);   // (pos, neg) of channels 2q, 2q+1
            acc0 = __fadd_rn(acc0, p0 ? g.x : g.y);
            acc1 = __fadd_rn(acc1, p1 ? g.z : g.w);
        }
    }
#pragma unroll
    for (int sh = 4; sh < 32; sh <<= 1) {
        acc0 = __fadd_rn(acc0, __shfl_xor_sync(0xFFFFFFFFu, acc0, sh));
        acc1 = __fadd_rn(acc1, __shfl_xor_sync(0xFFFFFFFFu, acc1, sh));
    }
    if (sub == 0)   // d fraction / d vote carries the sign of the row's own value
        *reinterpret_cast<float2 *>(a.grad_table + (size_t)row * 8u + 2u * q) = make_float2(p0 ? acc0 : -acc0, p1 ? acc1 : -acc1);
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

int cnc_vote3_fwd(const uint8_t *binary_vxl, uint32_t Rb, const uint8_t *sign_bits, uint32_t resolution, uint32_t F,
                  uint32_t hashmap_size, float *out_xy, float *out_xz, float *out_yz, cnc_stream_t stream) {
    if (!binary_vxl || !sign_bits || !(out_xy || out_xz || out_yz)) { set_error("vote3_fwd: null pointer"); return CNC_EINVAL; }
    if (F != 8 || Rb == 0 || Rb > 128 || resolution < 3 || (resolution - 2) % Rb != 0) {
        set_error("vote3_fwd: needs F == 8, Rb <= 128 and (resolution - 2) a multiple of Rb");
        return CNC_ENOTSUP;
    }
    const uint32_t s = resolution - 2;
    dim3 grid(div_up(s, 128), s, 3);
    vote3_fwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(binary_vxl, Rb, sign_bits, resolution, hashmap_size, out_xy, out_xz, out_yz);
    return check_launch("vote3_fwd");
}

int cnc_vote3_bwd(const int16_t *pts_by_row, const int64_t *seg, const uint8_t *binary_vxl, uint32_t Rb, const uint8_t *sign_bits,
                  uint32_t resolution, uint32_t F, uint32_t hashmap_size, const float *grad_xy, const float *grad_xz,
                  const float *grad_yz, float *grad_table, int32_t members_only, cnc_stream_t stream) {
    if (!pts_by_row || !seg || !binary_vxl || !sign_bits || !(grad_xy || grad_xz || grad_yz) || !grad_table) {
        set_error("vote3_bwd: null pointer");
        return CNC_EINVAL;
    }
    if (F != 8 || Rb == 0 || Rb > 128 || resolution < 3 || (resolution - 2) % Rb != 0) {
        set_error("vote3_bwd: needs F == 8, Rb <= 128 and (resolution - 2) a multiple of Rb");
        return CNC_ENOTSUP;
    }
    Vote3BwdArgs a{pts_by_row, seg, binary_vxl, sign_bits, {grad_xy, grad_xz, grad_yz}, grad_table, Rb, resolution, hashmap_size, members_only};
    if (members_only)
        vote3_bwd_members_kernel<<<div_up((uint64_t)hashmap_size * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    else
        vote3_bwd_kernel<<<div_up((uint64_t)hashmap_size * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("vote3_bwd");
}

int cnc_segment_wsum_idx(const float *feat, const int64_t *idx, const float *w, const int64_t *cumsum, float *out, int64_t N,
                         int64_t F, cnc_stream_t stream) {
    if (N * F == 0) return CNC_OK;
    if (!feat || !cumsum || !out) { set_error("segment_wsum_idx: null pointer"); return CNC_EINVAL; }
    segment_wsum_idx_kernel<<<div_up((uint64_t)(N * F), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(feat, idx, w, cumsum, out, N, F);
    return check_launch("segment_wsum_idx");
}

int cnc_segment_wsum_idx_bwd(const float *grad_out, const int64_t *idx, const float *w, const int64_t *cumsum, float *grad_feat,
                             int64_t N, int64_t F, cnc_stream_t stream) {
    if (N * F == 0) return CNC_OK;
    if (!grad_out || !cumsum || !grad_feat) { set_error("segment_wsum_idx_bwd: null pointer"); return CNC_EINVAL; }
    segment_wsum_idx_bwd_kernel<<<div_up((uint64_t)(N * F), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, idx, w, cumsum, grad_feat, N, F);
    return check_launch("segment_wsum_idx_bwd");
}

}  // extern "C"
