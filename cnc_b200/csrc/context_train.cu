// context_train.cu -- the 3D context gather of the rate term, forward and backward (training loss,
// examples/utils_bpp_acc.py:644-687).
//
// Reference flow per step: the voxels of ~150 000 sampled hash entries that touch the occupancy (5.7 M at the product
// layout) are normalised, pushed through GridEncoder.forward_diff_levels (K1 with a per-point start level and the per-corner
// occupancy test, gridencoder.cu:221-276: a box of up to 17^3 occupancy cells per corner), permuted, concatenated with the
// level frequency Pg into the [M, 25] input of context_model_3D; the backward runs K2 with the same per-corner boxes.
//
// Here one kernel each way works straight from the int16 voxel coordinates:
//   * the per-corner occupancy test is one bit of the per-vertex bitmaps (cnc_vertex_valid_bits, rebuilt only when the
//     occupancy grid changes, every 16 steps) instead of a box scan -- the same predicate, evaluated once per vertex;
//   * features come from the 1-bit sign planes (L2 resident) and land directly in the [M, 25] layout, Pg column included:
//     no [3, M, 8] intermediate, no permute, no cat;
//   * the backward reads the [M, 25] input gradient in place and scatter-adds with red.global.add.v4.f32.
// Thread = (voxel, context level); the arithmetic is make_corners_fn of common.cuh, i.e. K1's, so the features are
// bit-identical to the reference flow on +-1 tables.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {
namespace ct {

struct Args {
    const int16_t *pts;        // [M,3] voxel coordinates at their own level
    const int64_t *level;      // [M]   level n of every voxel (>= 3)
    const uint8_t *sign_bits;  // sign plane of the whole 3D table
    const int32_t *offsets, *resolutions;
    const uint32_t *vbits;     // per-vertex validity bitmaps
    const int64_t *vbit_off;   // [L+1]
    const float *Pg;           // [L] level frequencies (column 24)
    float *x;                  // fwd out [M,25]
    const float *gx;           // bwd in  [M,25]
    float *grad_table;         // bwd out [rows,8], accumulated into
    float *g_pg;               // bwd out [L], accumulated into: column 24 summed per level
    int64_t M;
};

__device__ __forceinline__ bool corners_of(const Args &a, int64_t v, uint32_t l, LevelConst &lc, Corners<3> &cs) {
    const uint32_t n = (uint32_t)__ldg(a.level + v);
    const uint32_t lev = n - 3u + l;
    lc = load_level(a.offsets, a.resolutions, lev);
    // normalised coordinate of the voxel at ITS level: (c - 0.5) / (res_n - 2)   (utils_bpp_acc.py:647)
    const float sc = (float)((uint32_t)__ldg(a.resolutions + n) - 2u);
    float xi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) xi[d] = __fdiv_rn(__fsub_rn((float)a.pts[v * 3 + d], 0.5f), sc);
    const uint32_t *vb = a.vbits + (__ldg(a.vbit_off + lev) >> 5);
    const uint32_t res = lc.res;
    return make_corners_fn<3>(xi, lc, [&](const uint32_t (&c)[3]) {
        const uint32_t bit = (c[0] * res + c[1]) * res + c[2];
        return ((__ldg(vb + (bit >> 5)) >> (bit & 31u)) & 1u) != 0u;
    }, cs);
}

__global__ void __launch_bounds__(256) ctx3d_gather_fwd_kernel(const Args a) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.M) return;
    const uint32_t l = blockIdx.y;
    LevelConst lc;
    Corners<3> cs;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.f;
    if (corners_of(a, v, l, lc, cs)) {
        uint32_t sb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) sb[i] = ((cs.valid >> i) & 1u) ? (uint32_t)__ldg(a.sign_bits + (uint64_t)lc.base_row + cs.row[i]) : 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if ((cs.valid >> i) & 1u) {
                const float ww = __fmul_rn(cs.w[i], cs.wn_re);
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] = __fadd_rn(acc[k], ((sb[i] >> k) & 1u) ? ww : -ww);
            }
        }
    }
    float *o = a.x + v * 25 + l * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = acc[k];
    if (l == 0) a.x[v * 25 + 24] = __ldg(a.Pg + __ldg(a.level + v));
}

__global__ void __launch_bounds__(256) ctx3d_gather_bwd_kernel(const Args a) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t l = blockIdx.y;
    if (l == 0) {
        // d/dPg: the voxels arrive grouped by level, so a warp almost always holds one level: shuffle-reduce and add once
        // (5.7 M single atomics onto 12 addresses cost 3.5 ms; this costs nothing)
        const bool in = v < a.M;
        const int lev = in ? (int)__ldg(a.level + v) : -1;
        float g = in ? __ldg(a.gx + v * 25 + 24) : 0.f;
        const int lev0 = __shfl_sync(0xffffffffu, lev, 0);
        if (__all_sync(0xffffffffu, lev == lev0 || lev < 0)) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) g = __fadd_rn(g, __shfl_xor_sync(0xffffffffu, g, d));
            if ((threadIdx.x & 31) == 0 && lev0 >= 0) atomicAdd(a.g_pg + lev0, g);
        } else if (in) {
            atomicAdd(a.g_pg + lev, g);
        }
    }
    if (v >= a.M) return;
    LevelConst lc;
    Corners<3> cs;
    if (!corners_of(a, v, l, lc, cs)) return;
    float g[8];
    const float *gi = a.gx + v * 25 + l * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) g[k] = __ldg(gi + k);
    float *gt = a.grad_table + (size_t)lc.base_row * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if ((cs.valid >> i) & 1u) {
            const float ww = __fmul_rn(cs.w[i], cs.wn_re);
            float4 *p = reinterpret_cast<float4 *>(gt + (size_t)cs.row[i] * 8);
            atomicAdd(p, make_float4(__fmul_rn(ww, g[0]), __fmul_rn(ww, g[1]), __fmul_rn(ww, g[2]), __fmul_rn(ww, g[3])));
            atomicAdd(p + 1, make_float4(__fmul_rn(ww, g[4]), __fmul_rn(ww, g[5]), __fmul_rn(ww, g[6]), __fmul_rn(ww, g[7])));
        }
    }
}

// ------------------------------------------------------------------------------------------
// Plane context of the rate term / codec (utils_bpp_acc.py:551-558, :731-738): for every plane vertex of the occupied
// region, the c coarser levels of the plane encoder (masked bilinear gather, K1 with D = 2) | the +1 vote fraction plane of the
// dimension-wise context sampled at the vertex (Encoding_2D.forward_given_params: K1 over a float table, one dense level)
// | Pg  ->  [N, 8c + 8 + 1] in one kernel, and the matching backward (K2 into the plane table, K2 into the fraction plane,
// column sum for Pg).  Replaces two K1 launches, a permute-reshape each, expand and cat (108 MB per call at level 3) and, in
// the backward, two zero-fills, two K2 launches and the slicing of the [N, K] gradient.  Arithmetic = make_corners<2>.
// ------------------------------------------------------------------------------------------
struct Args2 {
    const float *pts;          // [N,2] normalised vertex coordinates
    const uint8_t *sign_bits;  // sign plane of the whole plane table
    const int32_t *offsets, *resolutions;
    const uint8_t *vxl2;       // [Rb,Rb] occupancy of the plane
    const float *frac;         // [res_f^2, 8] vote fraction plane (nullable: no dimension-wise context)
    const float *Pg;           // 1 float
    float *x;                  // fwd out [N,K]
    const float *gx;           // bwd in  [N,K]
    float *grad_table;         // bwd out [rows,8]   (accumulated into)
    float *grad_frac;          // bwd out [res_f^2,8] (accumulated into; nullable)
    float *g_pg;               // bwd out 1 float (accumulated into)
    int64_t N;
    int32_t level, c, K, Rb, res_f;
};

__device__ __forceinline__ bool corners2_of(const Args2 &a, int64_t v, int slot, LevelConst &lc, Corners<2> &cs) {
    if (slot < a.c) {
        lc = load_level(a.offsets, a.resolutions, (uint32_t)(a.level - a.c + slot));
    } else {   // the fraction plane: one dense level of res_f^2 rows
        lc.base_row = 0;
        lc.res = (uint32_t)a.res_f;
        lc.T = lc.res * lc.res;
        lc.scale = (float)(lc.res - 2u);
        lc.scale_re = __frcp_rn(lc.scale);
    }
    const float xi[2] = {__ldg(a.pts + v * 2), __ldg(a.pts + v * 2 + 1)};
    return make_corners<2>(xi, lc, (uint32_t)a.Rb, a.vxl2, cs);
}

// thread = vertex, all slots in turn; the block's [256, K] tile of x (K odd: conflict-free rows) goes through shared memory, so
// that the global side is one contiguous, fully coalesced copy -- with a (vertex, slot) thread writing its 8 floats at a row
// stride of K the stores touched 8 sectors per thread where 2 hold the data, and the kernel was bound by exactly that.
__global__ void __launch_bounds__(256) ctx2d_gather_fwd_kernel(const Args2 a, int slots) {
    extern __shared__ float tile[];
    const int64_t row0 = (int64_t)blockIdx.x * blockDim.x, v = row0 + threadIdx.x;
    float *mine = tile + (size_t)threadIdx.x * a.K;
    for (int slot = 0; slot < slots; slot++) {
        LevelConst lc;
        Corners<2> cs;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.f;
        if (v < a.N && corners2_of(a, v, slot, lc, cs)) {
            if (slot < a.c) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if ((cs.valid >> i) & 1u) {
                        const uint32_t sb = __ldg(a.sign_bits + (uint64_t)lc.base_row + cs.row[i]);
                        const float ww = __fmul_rn(cs.w[i], cs.wn_re);
#pragma unroll
                        for (int k = 0; k < 8; k++) acc[k] = __fadd_rn(acc[k], ((sb >> k) & 1u) ? ww : -ww);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if ((cs.valid >> i) & 1u) {
                        const float4 *r = reinterpret_cast<const float4 *>(a.frac + (size_t)cs.row[i] * 8);
                        const float4 r0 = __ldg(r), r1 = __ldg(r + 1);
                        const float t[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                        const float ww = __fmul_rn(cs.w[i], cs.wn_re);
#pragma unroll
                        for (int k = 0; k < 8; k++) acc[k] = __fmaf_rn(ww, t[k], acc[k]);   // gridencoder.cu:301
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) mine[slot * 8 + k] = acc[k];
    }
    mine[a.K - 1] = __ldg(a.Pg);
    __syncthreads();
    const int64_t rows = a.N - row0 < (int64_t)blockDim.x ? a.N - row0 : (int64_t)blockDim.x;
    float *dst = a.x + row0 * a.K;
    for (int64_t e = threadIdx.x; e < rows * a.K; e += blockDim.x) dst[e] = tile[e];
}

__global__ void __launch_bounds__(256) ctx2d_gather_bwd_kernel(const Args2 a, int slots) {
    extern __shared__ float tile[];
    const int64_t row0 = (int64_t)blockIdx.x * blockDim.x, v = row0 + threadIdx.x;
    const int64_t rows = a.N - row0 < (int64_t)blockDim.x ? a.N - row0 : (int64_t)blockDim.x;
    const float *src = a.gx + row0 * a.K;
    for (int64_t e = threadIdx.x; e < rows * a.K; e += blockDim.x) tile[e] = __ldg(src + e);
    __syncthreads();
    const float *mine = tile + (size_t)threadIdx.x * a.K;
    {
        float g = v < a.N ? mine[a.K - 1] : 0.f;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) g = __fadd_rn(g, __shfl_xor_sync(0xffffffffu, g, d));
        if ((threadIdx.x & 31) == 0) atomicAdd(a.g_pg, g);
    }
    const bool inside = v < a.N;
    const uint32_t lane = threadIdx.x & 31u;
    for (int slot = 0; slot < slots; slot++) {
        LevelConst lc;
        Corners<2> cs;
        const bool ok = corners2_of(a, inside ? v : a.N - 1, slot, lc, cs) && inside;
        float g[8];
#pragma unroll
        for (int k = 0; k < 8; k++) g[k] = inside ? mine[slot * 8 + k] : 0.f;
        float *gt = slot < a.c ? a.grad_table + (size_t)lc.base_row * 8 : a.grad_frac;
        // the coarser plane levels: neighbouring vertices of the fine level (= neighbouring lanes) share their corners, and a
        // few thousand rows would take every atomic -- contributions are summed over runs of lanes with the same row first
        // (the segmented shuffle reduction of K2, grid_encode.cu); the fraction plane has the fine level's own resolution
        const bool agg = slot < a.c;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const bool on = ok && ((cs.valid >> i) & 1u);
            const float ww = on ? __fmul_rn(cs.w[i], cs.wn_re) : 0.f;
            float w8[8];
#pragma unroll
            for (int k = 0; k < 8; k++) w8[k] = __fmul_rn(ww, g[k]);
            bool head = on;
            if (agg) {
                const uint32_t key = on ? cs.row[i] : 0xFFFFFFFFu;
                const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
                const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
                const uint32_t start = 31u - (uint32_t)__clz(heads & (0xffffffffu >> (31u - lane)));
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t ostart = __shfl_down_sync(0xffffffffu, start, d);
                    const bool take = (lane + d < 32u) && ostart == start;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float o = __shfl_down_sync(0xffffffffu, w8[k], d);
                        if (take) w8[k] = __fadd_rn(w8[k], o);
                    }
                }
                head = on && start == lane;
            }
            if (head) {
                float4 *p = reinterpret_cast<float4 *>(gt + (size_t)cs.row[i] * 8);
                atomicAdd(p, make_float4(w8[0], w8[1], w8[2], w8[3]));
                atomicAdd(p + 1, make_float4(w8[4], w8[5], w8[6], w8[7]));
            }
        }
    }
}

// table[rows[i]] -> out[i] and back (rows of 8 floats = one 32-byte sector; torch's index_select runs this at 160 GB/s)
__global__ void __launch_bounds__(256) rows8_gather_kernel(const float *__restrict__ table, const int64_t *__restrict__ rows, int64_t M,
                                                           float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const float4 *r = reinterpret_cast<const float4 *>(table + rows[i] * 8);
    float4 *o = reinterpret_cast<float4 *>(out + i * 8);
    o[0] = __ldg(r);
    o[1] = __ldg(r + 1);
}
__global__ void __launch_bounds__(256) rows8_scatter_kernel(const float *__restrict__ g, const int64_t *__restrict__ rows, int64_t M,
                                                            float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const float4 *r = reinterpret_cast<const float4 *>(g + i * 8);
    float4 *o = reinterpret_cast<float4 *>(out + rows[i] * 8);
    o[0] = __ldg(r);
    o[1] = __ldg(r + 1);
}

}  // namespace ct
}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_ctx3d_gather_fwd(const int16_t *pts, const int64_t *level, int64_t M, const uint8_t *sign_bits, const int32_t *offsets,
                         const int32_t *resolutions, const uint32_t *vertex_bits, const int64_t *vertex_bit_offsets, const float *Pg,
                         float *x, cnc_stream_t stream) {
    if (M == 0) return CNC_OK;
    if (!pts || !level || !sign_bits || !offsets || !resolutions || !vertex_bits || !vertex_bit_offsets || !Pg || !x) {
        set_error("ctx3d_gather_fwd: null pointer");
        return CNC_EINVAL;
    }
    ct::Args a{pts, level, sign_bits, offsets, resolutions, vertex_bits, vertex_bit_offsets, Pg, x, nullptr, nullptr, nullptr, M};
    ct::ctx3d_gather_fwd_kernel<<<dim3(div_up((uint64_t)M, 256), 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("ctx3d_gather_fwd");
}

int cnc_ctx3d_gather_bwd(const int16_t *pts, const int64_t *level, int64_t M, const int32_t *offsets, const int32_t *resolutions,
                         const uint32_t *vertex_bits, const int64_t *vertex_bit_offsets, const float *gx, float *grad_table,
                         float *grad_pg, cnc_stream_t stream) {
    if (M == 0) return CNC_OK;
    if (!pts || !level || !offsets || !resolutions || !vertex_bits || !vertex_bit_offsets || !gx || !grad_table || !grad_pg) {
        set_error("ctx3d_gather_bwd: null pointer");
        return CNC_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(grad_table) & 15u) { set_error("ctx3d_gather_bwd: grad_table must be 16-byte aligned"); return CNC_EINVAL; }
    ct::Args a{pts, level, nullptr, offsets, resolutions, vertex_bits, vertex_bit_offsets, nullptr, nullptr, gx, grad_table, grad_pg, M};
    ct::ctx3d_gather_bwd_kernel<<<dim3(div_up((uint64_t)M, 256), 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("ctx3d_gather_bwd");
}

int cnc_ctx2d_gather_fwd(const float *pts, int64_t N, const uint8_t *sign_bits, const int32_t *offsets, const int32_t *resolutions,
                         int32_t level, int32_t n_ctx_levels, const uint8_t *binary_vxl_2D, int32_t Rb, const float *frac, int32_t res_frac,
                         const float *Pg, float *x, cnc_stream_t stream) {
    if (N == 0) return CNC_OK;
    if (!pts || !sign_bits || !offsets || !resolutions || !binary_vxl_2D || !Pg || !x || n_ctx_levels < 0 || n_ctx_levels > level || Rb < 1 ||
        (frac && res_frac < 3)) {
        set_error("ctx2d_gather_fwd: bad argument");
        return CNC_EINVAL;
    }
    const int slots = n_ctx_levels + (frac ? 1 : 0);
    if (slots == 0) { set_error("ctx2d_gather_fwd: nothing to gather"); return CNC_EINVAL; }
    if (frac && (reinterpret_cast<uintptr_t>(frac) & 15u)) { set_error("ctx2d_gather_fwd: frac must be 16-byte aligned"); return CNC_EINVAL; }
    ct::Args2 a{pts, sign_bits, offsets, resolutions, binary_vxl_2D, frac, Pg, x, nullptr, nullptr, nullptr, nullptr, N, level, n_ctx_levels,
                8 * slots + 1, Rb, res_frac};
    ct::ctx2d_gather_fwd_kernel<<<div_up((uint64_t)N, 256), 256, 256 * a.K * sizeof(float), static_cast<cudaStream_t>(stream)>>>(a, slots);
    return check_launch("ctx2d_gather_fwd");
}

int cnc_ctx2d_gather_bwd(const float *pts, int64_t N, const int32_t *offsets, const int32_t *resolutions, int32_t level,
                         int32_t n_ctx_levels, const uint8_t *binary_vxl_2D, int32_t Rb, int32_t res_frac, const float *gx, float *grad_table,
                         float *grad_frac, float *grad_pg, cnc_stream_t stream) {
    if (N == 0) return CNC_OK;
    if (!pts || !offsets || !resolutions || !binary_vxl_2D || !gx || !grad_table || !grad_pg || n_ctx_levels < 0 || n_ctx_levels > level || Rb < 1) {
        set_error("ctx2d_gather_bwd: bad argument");
        return CNC_EINVAL;
    }
    const int slots = n_ctx_levels + (grad_frac ? 1 : 0);
    if ((reinterpret_cast<uintptr_t>(grad_table) & 15u) || (grad_frac && (reinterpret_cast<uintptr_t>(grad_frac) & 15u))) {
        set_error("ctx2d_gather_bwd: gradient buffers must be 16-byte aligned");
        return CNC_EINVAL;
    }
    ct::Args2 a{pts, nullptr, offsets, resolutions, binary_vxl_2D, nullptr, nullptr, nullptr, gx, grad_table, grad_frac, grad_pg, N, level,
                n_ctx_levels, 8 * slots + 1, Rb, res_frac};
    ct::ctx2d_gather_bwd_kernel<<<div_up((uint64_t)N, 256), 256, 256 * a.K * sizeof(float), static_cast<cudaStream_t>(stream)>>>(a, slots);
    return check_launch("ctx2d_gather_bwd");
}

int cnc_rows8_gather(const float *table, const int64_t *rows, int64_t M, float *out, cnc_stream_t stream) {
    if (M == 0) return CNC_OK;
    if (!table || !rows || !out || ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15u)) {
        set_error("rows8_gather: null or misaligned pointer");
        return CNC_EINVAL;
    }
    ct::rows8_gather_kernel<<<div_up((uint64_t)M, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(table, rows, M, out);
    return check_launch("rows8_gather");
}

int cnc_rows8_scatter(const float *grad, const int64_t *rows, int64_t M, float *out, cnc_stream_t stream) {
    if (M == 0) return CNC_OK;
    if (!grad || !rows || !out || ((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(out)) & 15u)) {
        set_error("rows8_scatter: null or misaligned pointer");
        return CNC_EINVAL;
    }
    ct::rows8_scatter_kernel<<<div_up((uint64_t)M, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad, rows, M, out);
    return check_launch("rows8_scatter");
}

}  // extern "C"
