// grid_encode.cu -- multiresolution hash-grid encode: forward gather, backward scatter-add,
// STE binarisation and the 1-bit sign table.
//
// Reference behaviour restated (not translated): gridencoder/src/gridencoder.cu:99-316 (K1),
// :400-585 (K2); examples/radiance_fields/ngp.py:22-39 (STE_binary).
//
// B200 notes
//  * The path is a random gather of 32-byte rows (F=8 fp32) -> bound by L2/HBM sector
//    throughput, not by math.  One thread owns one (point, level): all 2^D row fetches are
//    issued as independent 128-bit loads before the first use (MLP ~16 for D=3,F=8), the
//    [L,N,F] output is written with 128-bit stores that a warp coalesces into 1 KB lines.
//  * The *_bits variant reads a 1-bit/parameter sign table (5 MB for the whole product layout,
//    L2-resident on B200's 126 MB L2) instead of the 161 MB fp32 table: same results, because
//    every STE_binary value is exactly +-1.
//  * Backward uses vector reductions (red.global.add.v4.f32, sm_90+) -- 2 per corner at F=8
//    instead of 8 scalar atomics.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {

// ------------------------------------------------------------------------------------------
template <int F>
struct RowVec {
    float v[F];
};

template <int F, bool VEC>
__device__ __forceinline__ void load_row(const float *__restrict__ p, float (&v)[F]) {
    if constexpr (VEC && F >= 4) {
#pragma unroll
        for (int k = 0; k < F / 4; k++) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(p) + k);
            v[4 * k + 0] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        }
    } else if constexpr (VEC && F == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int k = 0; k < F; k++) v[k] = __ldg(p + k);
    }
}

template <int F, bool VEC>
__device__ __forceinline__ void store_row(float *__restrict__ p, const float (&v)[F]) {
    if constexpr (VEC && F >= 4) {
#pragma unroll
        for (int k = 0; k < F / 4; k++)
            reinterpret_cast<float4 *>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    } else if constexpr (VEC && F == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int k = 0; k < F; k++) p[k] = v[k];
    }
}

// sign bits of one row -> +-1.  bit index = row*F + ch  (cnc_sign_pack layout)
template <int F>
__device__ __forceinline__ uint32_t load_sign_bits(const uint8_t *__restrict__ bits, uint64_t row) {
    if constexpr (F == 32) return __ldg(reinterpret_cast<const uint32_t *>(bits) + row);
    else if constexpr (F == 16) return __ldg(reinterpret_cast<const uint16_t *>(bits) + row);
    else if constexpr (F == 8) return __ldg(bits + row);
    else {
        const uint64_t bit = row * F;
        return (uint32_t)(__ldg(bits + (bit >> 3)) >> (bit & 7)) & ((1u << F) - 1u);
    }
}

// ------------------------------------------------------------------------------------------
// K1.  grid (ceil(N/TPB), L_calc); one thread = one (point, level)
// ------------------------------------------------------------------------------------------
constexpr int TPB = 256;
constexpr int32_t AGG_MAX_RES = 256;   // K2: levels up to this resolution aggregate runs of equal rows inside the warp

template <int D, int F, bool BITS, bool VEC>
__global__ void __launch_bounds__(TPB)
grid_fwd_kernel(const float *__restrict__ x, const void *__restrict__ table_,
                const int32_t *__restrict__ offsets, const int32_t *__restrict__ resolutions,
                float *__restrict__ out, uint32_t N, uint32_t Rb, const uint8_t *__restrict__ vxl,
                const int32_t *__restrict__ min_level_id) {
    const uint32_t b = blockIdx.x * TPB + threadIdx.x;
    if (b >= N) return;
    const uint32_t l = blockIdx.y;
    const uint32_t level = (min_level_id ? (uint32_t)__ldg(min_level_id + b) : 0u) + l;  // :118-126
    const LevelConst lc = load_level(offsets, resolutions, level);

    float xi[D];
#pragma unroll
    for (int d = 0; d < D; d++) xi[d] = __ldg(x + (size_t)b * D + d);

    float acc[F];
#pragma unroll
    for (int k = 0; k < F; k++) acc[k] = 0.f;

    Corners<D> cs;
    if (make_corners<D>(xi, lc, Rb, vxl, cs)) {
        if constexpr (BITS) {
            const uint8_t *bits = static_cast<const uint8_t *>(table_);
            uint32_t sb[1 << D];
#pragma unroll
            for (int i = 0; i < (1 << D); i++)
                sb[i] = ((cs.valid >> i) & 1u) ? load_sign_bits<F>(bits, (uint64_t)lc.base_row + cs.row[i]) : 0u;
#pragma unroll
            for (int i = 0; i < (1 << D); i++) {
                if ((cs.valid >> i) & 1u) {
                    const float ww = __fmul_rn(cs.w[i], cs.wn_re);
#pragma unroll
                    for (int k = 0; k < F; k++)
                        acc[k] = __fadd_rn(acc[k], ((sb[i] >> k) & 1u) ? ww : -ww);
                }
            }
        } else {
            const float *table = static_cast<const float *>(table_) + (size_t)lc.base_row * F;
            float rows[1 << D][F];
#pragma unroll
            for (int i = 0; i < (1 << D); i++) {
                if ((cs.valid >> i) & 1u) load_row<F, VEC>(table + (size_t)cs.row[i] * F, rows[i]);
            }
#pragma unroll
            for (int i = 0; i < (1 << D); i++) {
                if ((cs.valid >> i) & 1u) {
                    const float ww = __fmul_rn(cs.w[i], cs.wn_re);  // gridencoder.cu:301
#pragma unroll
                    for (int k = 0; k < F; k++) acc[k] = __fmaf_rn(ww, rows[i][k], acc[k]);
                }
            }
        }
    }
    store_row<F, VEC>(out + ((size_t)l * N + b) * F, acc);
}

// ------------------------------------------------------------------------------------------
// K2.  same decomposition; scatter with vector reductions
// ------------------------------------------------------------------------------------------
template <int F, bool VEC>
__device__ __forceinline__ void red_row(float *__restrict__ p, const float (&v)[F]) {
    if constexpr (VEC && F >= 4) {
#pragma unroll
        for (int k = 0; k < F / 4; k++)
            atomicAdd(reinterpret_cast<float4 *>(p) + k,
                      make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
    } else if constexpr (VEC && F == 2) {
        atomicAdd(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
    } else {
#pragma unroll
        for (int k = 0; k < F; k++) atomicAdd(p + k, v[k]);
    }
}

template <int D, int F, bool VEC>
__global__ void __launch_bounds__(TPB)
grid_bwd_kernel(const float *__restrict__ grad, const float *__restrict__ x,
                const int32_t *__restrict__ offsets, const int32_t *__restrict__ resolutions,
                float *__restrict__ grad_table, uint32_t N, uint32_t Rb,
                const uint8_t *__restrict__ vxl, const int32_t *__restrict__ min_level_id, uint32_t ld, uint32_t col0) {
    const uint32_t b = blockIdx.x * TPB + threadIdx.x;
    const uint32_t l = blockIdx.y;
    // Coarse levels: a few thousand rows receive the updates of every sample -- the atomics serialise in L2 (measured at
    // the product layout, 380 k samples: level 0 (5 832 rows) 0.174 ms, level 1 0.110, level 2 0.062, against 0.03 ms for a
    // hashed level of 524 288 rows).  Samples arrive in ray order, so the lanes of a warp mostly sit in the same cell:
    // each corner's contributions are summed over RUNS of adjacent lanes with the same row (segmented shuffle reduction)
    // and only the head of a run issues the reduction.  Warp-uniform decision (needs one level per warp).  Measured per
    // level group, ray order: levels 0-2 0.346 -> 0.084 ms, 3-5 0.115 -> 0.070, 6-8 0.094 -> 0.074; 9-11 (cells smaller than
    // the sample spacing: few shared corners) 0.086 -> 0.089, hence the resolution limit.  Unordered samples pay 3 %.
    const bool agg = min_level_id == nullptr && __ldg(resolutions + l) <= AGG_MAX_RES;
    if (b >= N && !agg) return;
    const bool inside = b < N;
    const uint32_t level = (min_level_id ? (uint32_t)__ldg(min_level_id + b) : 0u) + l;
    const LevelConst lc = load_level(offsets, resolutions, level);

    float xi[D];
#pragma unroll
    for (int d = 0; d < D; d++) xi[d] = inside ? __ldg(x + (size_t)b * D + d) : -1.f;   // (-1: out of range -> no corner)
    float g[F];
    if (agg) {
        const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
        for (int k = 0; k < F; k++) g[k] = 0.f;
        if (inside) load_row<F, VEC>(ld ? grad + (size_t)b * ld + col0 + (size_t)l * F : grad + ((size_t)l * N + b) * F, g);
        Corners<D> cs;
        const bool ok = make_corners<D>(xi, lc, Rb, vxl, cs);
        float *gt = grad_table + (size_t)lc.base_row * F;
#pragma unroll
        for (int i = 0; i < (1 << D); i++) {
            const bool on = ok && ((cs.valid >> i) & 1u);
            const uint32_t key = on ? cs.row[i] : 0xFFFFFFFFu;
            const float ww = on ? __fmul_rn(cs.w[i], cs.wn_re) : 0.f;
            float v[F];
#pragma unroll
            for (int k = 0; k < F; k++) v[k] = __fmul_rn(ww, g[k]);  // gridencoder.cu:580
            // runs of equal keys: head flags -> start lane of this lane's run
            const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
            const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
            const uint32_t start = 31u - (uint32_t)__clz(heads & (0xffffffffu >> (31u - lane)));
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t ostart = __shfl_down_sync(0xffffffffu, start, d);
                const bool take = (lane + d < 32u) && ostart == start;
#pragma unroll
                for (int k = 0; k < F; k++) {
                    const float o = __shfl_down_sync(0xffffffffu, v[k], d);
                    if (take) v[k] = __fadd_rn(v[k], o);
                }
            }
            if (on && start == lane) red_row<F, VEC>(gt + (size_t)key * F, v);
        }
        return;
    }
    // grad is [L, N, F] (the reference's layout, ld == 0) or a column block of a row-major [N, ld] matrix (level l at
    // columns col0 + l F .. : what a GEMM that produced the feature gradients leaves behind)
    load_row<F, VEC>(ld ? grad + (size_t)b * ld + col0 + (size_t)l * F : grad + ((size_t)l * N + b) * F, g);

    Corners<D> cs;
    if (!make_corners<D>(xi, lc, Rb, vxl, cs)) return;  // gridencoder.cu:435-440
    float *gt = grad_table + (size_t)lc.base_row * F;
#pragma unroll
    for (int i = 0; i < (1 << D); i++) {
        if ((cs.valid >> i) & 1u) {
            const float ww = __fmul_rn(cs.w[i], cs.wn_re);
            float v[F];
#pragma unroll
            for (int k = 0; k < F; k++) v[k] = __fmul_rn(ww, g[k]);  // gridencoder.cu:580
            red_row<F, VEC>(gt + (size_t)cs.row[i] * F, v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// STE_binary + sign planes
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ste_val(float p) {
    // clamp then (>=0)*1 + (<0)*-1  (ngp.py:26-30); NaN -> 0
    return (p >= 0.f) ? 1.f : ((p < 0.f) ? -1.f : 0.f);
}

__global__ void ste_fwd_kernel(const float *__restrict__ p, float *__restrict__ out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p + i));
            *reinterpret_cast<float4 *>(out + i) = make_float4(ste_val(v.x), ste_val(v.y), ste_val(v.z), ste_val(v.w));
        } else {
            for (uint64_t j = i; j < n; j++) out[j] = ste_val(p[j]);
        }
    }
}

__global__ void ste_bwd_kernel(const float *__restrict__ p, const float *__restrict__ go,
                               float *__restrict__ gi, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = __ldg(p + i);
        // mask = (clamp(v,-1,1) == v)  (ngp.py:36-38); NaN -> 0
        gi[i] = (v >= -1.f && v <= 1.f) ? __ldg(go + i) : __fmul_rn(__ldg(go + i), 0.f);
    }
}

// 8 params -> 1 byte; one thread packs 32 params (one uint32 store)
__global__ void sign_pack_kernel(const float *__restrict__ p, uint8_t *__restrict__ bits, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nw = n / 32;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += stride) {
        uint32_t m = 0;
        const float4 *src = reinterpret_cast<const float4 *>(p + w * 32);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 v = __ldg(src + k);
            m |= (uint32_t)(v.x >= 0.f) << (4 * k + 0);
            m |= (uint32_t)(v.y >= 0.f) << (4 * k + 1);
            m |= (uint32_t)(v.z >= 0.f) << (4 * k + 2);
            m |= (uint32_t)(v.w >= 0.f) << (4 * k + 3);
        }
        reinterpret_cast<uint32_t *>(bits)[w] = m;
    }
    // tail (n % 32 != 0, n % 8 == 0): bytes
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (uint64_t byte = nw * 4; byte < n / 8; byte++) {
            uint32_t m = 0;
            for (int k = 0; k < 8; k++) m |= (uint32_t)(p[byte * 8 + k] >= 0.f) << k;
            bits[byte] = (uint8_t)m;
        }
    }
}

__global__ void sign_unpack_kernel(const uint8_t *__restrict__ bits, float *__restrict__ out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = ((__ldg(bits + (i >> 3)) >> (i & 7)) & 1u) ? 1.f : -1.f;
}

// ------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int D, int F>
static int launch_fwd(bool bits, const float *x, const void *table, const int32_t *offsets,
                      const int32_t *resolutions, float *out, uint32_t N, uint32_t L, uint32_t Rb,
                      const uint8_t *vxl, const int32_t *mlid, cudaStream_t s) {
    const dim3 grid(div_up(N, TPB), L);
    if (bits) {
        if (aligned16(out))
            grid_fwd_kernel<D, F, true, true><<<grid, TPB, 0, s>>>(x, table, offsets, resolutions, out, N, Rb, vxl, mlid);
        else
            grid_fwd_kernel<D, F, true, false><<<grid, TPB, 0, s>>>(x, table, offsets, resolutions, out, N, Rb, vxl, mlid);
    } else {
        if (aligned16(out) && aligned16(table))
            grid_fwd_kernel<D, F, false, true><<<grid, TPB, 0, s>>>(x, table, offsets, resolutions, out, N, Rb, vxl, mlid);
        else
            grid_fwd_kernel<D, F, false, false><<<grid, TPB, 0, s>>>(x, table, offsets, resolutions, out, N, Rb, vxl, mlid);
    }
    return check_launch("grid_encode_fwd");
}

template <int D>
static int dispatch_fwd(uint32_t F, bool bits, const float *x, const void *table, const int32_t *offsets,
                        const int32_t *resolutions, float *out, uint32_t N, uint32_t L, uint32_t Rb,
                        const uint8_t *vxl, const int32_t *mlid, cudaStream_t s) {
    switch (F) {
        case 1: return launch_fwd<D, 1>(bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        case 2: return launch_fwd<D, 2>(bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        case 4: return launch_fwd<D, 4>(bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        case 8: return launch_fwd<D, 8>(bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        case 16: return launch_fwd<D, 16>(bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        case 32: return launch_fwd<D, 32>(bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        default: set_error("GridEncoding: n_features must be 1, 2, 4, 8, 16 or 32."); return CNC_ENOTSUP;
    }
}

static int encode_fwd_any(bool bits, const float *x, const void *table, const int32_t *offsets,
                          const int32_t *resolutions, float *out, uint32_t N, uint32_t D, uint32_t F,
                          uint32_t L, uint32_t Rb, const uint8_t *vxl, const int32_t *mlid, cnc_stream_t st) {
    if (N == 0 || L == 0) return CNC_OK;
    if (!x || !table || !offsets || !resolutions || !out) { set_error("grid_encode_fwd: null pointer"); return CNC_EINVAL; }
    if (L > 65535) { set_error("grid_encode_fwd: too many levels"); return CNC_EINVAL; }
    cudaStream_t s = static_cast<cudaStream_t>(st);
    switch (D) {
        case 1: return dispatch_fwd<1>(F, bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        case 2: return dispatch_fwd<2>(F, bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        case 3: return dispatch_fwd<3>(F, bits, x, table, offsets, resolutions, out, N, L, Rb, vxl, mlid, s);
        default: set_error("GridEncoding: num_dim must be 1, 2, 3."); return CNC_ENOTSUP;
    }
}

template <int D, int F>
static int launch_bwd(const float *grad, const float *x, const int32_t *offsets, const int32_t *resolutions,
                      float *gt, uint32_t N, uint32_t L, uint32_t Rb, const uint8_t *vxl, const int32_t *mlid,
                      cudaStream_t s, uint32_t ld = 0, uint32_t col0 = 0) {
    const dim3 grid(div_up(N, TPB), L);
    if (aligned16(grad) && aligned16(gt) && (ld & 3u) == 0 && (col0 & 3u) == 0)
        grid_bwd_kernel<D, F, true><<<grid, TPB, 0, s>>>(grad, x, offsets, resolutions, gt, N, Rb, vxl, mlid, ld, col0);
    else
        grid_bwd_kernel<D, F, false><<<grid, TPB, 0, s>>>(grad, x, offsets, resolutions, gt, N, Rb, vxl, mlid, ld, col0);
    return check_launch("grid_encode_bwd");
}

template <int D>
static int dispatch_bwd(uint32_t F, const float *grad, const float *x, const int32_t *offsets,
                        const int32_t *resolutions, float *gt, uint32_t N, uint32_t L, uint32_t Rb,
                        const uint8_t *vxl, const int32_t *mlid, cudaStream_t s, uint32_t ld = 0, uint32_t col0 = 0) {
    switch (F) {
        case 1: return launch_bwd<D, 1>(grad, x, offsets, resolutions, gt, N, L, Rb, vxl, mlid, s, ld, col0);
        case 2: return launch_bwd<D, 2>(grad, x, offsets, resolutions, gt, N, L, Rb, vxl, mlid, s, ld, col0);
        case 4: return launch_bwd<D, 4>(grad, x, offsets, resolutions, gt, N, L, Rb, vxl, mlid, s, ld, col0);
        case 8: return launch_bwd<D, 8>(grad, x, offsets, resolutions, gt, N, L, Rb, vxl, mlid, s, ld, col0);
        case 16: return launch_bwd<D, 16>(grad, x, offsets, resolutions, gt, N, L, Rb, vxl, mlid, s, ld, col0);
        case 32: return launch_bwd<D, 32>(grad, x, offsets, resolutions, gt, N, L, Rb, vxl, mlid, s, ld, col0);
        default: set_error("GridEncoding: n_features must be 1, 2, 4, 8, 16 or 32."); return CNC_ENOTSUP;
    }
}

}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_grid_encode_fwd(const float *x, const float *table, const int32_t *offsets,
                        const int32_t *resolutions, float *out, uint32_t N, uint32_t D, uint32_t F,
                        uint32_t L_calc, uint32_t Rb, const uint8_t *binary_vxl,
                        const int32_t *min_level_id, cnc_stream_t stream) {
    return encode_fwd_any(false, x, table, offsets, resolutions, out, N, D, F, L_calc, Rb, binary_vxl, min_level_id, stream);
}

int cnc_grid_encode_fwd_bits(const float *x, const uint8_t *sign_bits, const int32_t *offsets,
                             const int32_t *resolutions, float *out, uint32_t N, uint32_t D,
                             uint32_t F, uint32_t L_calc, uint32_t Rb, const uint8_t *binary_vxl,
                             const int32_t *min_level_id, cnc_stream_t stream) {
    return encode_fwd_any(true, x, sign_bits, offsets, resolutions, out, N, D, F, L_calc, Rb, binary_vxl, min_level_id, stream);
}

int cnc_grid_encode_bwd(const float *grad, const float *x, const int32_t *offsets,
                        const int32_t *resolutions, float *grad_table, uint32_t N, uint32_t D,
                        uint32_t F, uint32_t L_calc, uint32_t Rb, const uint8_t *binary_vxl,
                        const int32_t *min_level_id, cnc_stream_t stream) {
    if (N == 0 || L_calc == 0) return CNC_OK;
    if (!grad || !x || !offsets || !resolutions || !grad_table) { set_error("grid_encode_bwd: null pointer"); return CNC_EINVAL; }
    if (L_calc > 65535) { set_error("grid_encode_bwd: too many levels"); return CNC_EINVAL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (D) {
        case 1: return dispatch_bwd<1>(F, grad, x, offsets, resolutions, grad_table, N, L_calc, Rb, binary_vxl, min_level_id, s);
        case 2: return dispatch_bwd<2>(F, grad, x, offsets, resolutions, grad_table, N, L_calc, Rb, binary_vxl, min_level_id, s);
        case 3: return dispatch_bwd<3>(F, grad, x, offsets, resolutions, grad_table, N, L_calc, Rb, binary_vxl, min_level_id, s);
        default: set_error("GridEncoding: num_dim must be 1, 2, 3."); return CNC_ENOTSUP;
    }
}

int cnc_grid_encode_bwd_rows(const float *grad_rows, uint32_t ld, uint32_t col0, const float *x, const int32_t *offsets,
                             const int32_t *resolutions, float *grad_table, uint32_t N, uint32_t D, uint32_t F, uint32_t L_calc,
                             uint32_t Rb, const uint8_t *binary_vxl, const int32_t *min_level_id, cnc_stream_t stream) {
    if (N == 0 || L_calc == 0) return CNC_OK;
    if (!grad_rows || !x || !offsets || !resolutions || !grad_table) { set_error("grid_encode_bwd_rows: null pointer"); return CNC_EINVAL; }
    if (L_calc > 65535 || ld == 0 || col0 + L_calc * F > ld) { set_error("grid_encode_bwd_rows: column block outside the row"); return CNC_EINVAL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (D) {
        case 1: return dispatch_bwd<1>(F, grad_rows, x, offsets, resolutions, grad_table, N, L_calc, Rb, binary_vxl, min_level_id, s, ld, col0);
        case 2: return dispatch_bwd<2>(F, grad_rows, x, offsets, resolutions, grad_table, N, L_calc, Rb, binary_vxl, min_level_id, s, ld, col0);
        case 3: return dispatch_bwd<3>(F, grad_rows, x, offsets, resolutions, grad_table, N, L_calc, Rb, binary_vxl, min_level_id, s, ld, col0);
        default: set_error("GridEncoding: num_dim must be 1, 2, 3."); return CNC_ENOTSUP;
    }
}

static inline uint32_t ew_blocks(uint64_t work_items) {
    // elementwise streaming kernels: enough CTAs for 148 SMs x 8 resident, grid-stride beyond
    const uint64_t b = (work_items + 255) / 256;
    return (uint32_t)(b < 148ull * 16 ? (b ? b : 1) : 148ull * 16);
}

int cnc_ste_binary_fwd(const float *params, float *out, uint64_t n, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!params || !out) { set_error("ste_binary_fwd: null pointer"); return CNC_EINVAL; }
    if (!aligned16(params) || !aligned16(out)) { set_error("ste_binary_fwd: pointers must be 16-byte aligned"); return CNC_EINVAL; }
    ste_fwd_kernel<<<ew_blocks((n + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, out, n);
    return check_launch("ste_binary_fwd");
}

int cnc_ste_binary_bwd(const float *params, const float *gout, float *gin, uint64_t n, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!params || !gout || !gin) { set_error("ste_binary_bwd: null pointer"); return CNC_EINVAL; }
    ste_bwd_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, gout, gin, n);
    return check_launch("ste_binary_bwd");
}

int cnc_sign_pack(const float *params, uint8_t *bits, uint64_t n, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!params || !bits) { set_error("sign_pack: null pointer"); return CNC_EINVAL; }
    if (n % 8) { set_error("sign_pack: n must be a multiple of 8"); return CNC_EINVAL; }
    if (!aligned16(params) || (reinterpret_cast<uintptr_t>(bits) & 3u)) { set_error("sign_pack: misaligned pointer"); return CNC_EINVAL; }
    sign_pack_kernel<<<ew_blocks(n / 32 + 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, bits, n);
    return check_launch("sign_pack");
}

int cnc_sign_unpack(const uint8_t *bits, float *out, uint64_t n, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!bits || !out) { set_error("sign_unpack: null pointer"); return CNC_EINVAL; }
    sign_unpack_kernel<<<ew_blocks(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(bits, out, n);
    return check_launch("sign_unpack");
}

}  // extern "C"
