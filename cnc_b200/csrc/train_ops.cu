// train_ops.cu -- elementwise passes of the data-parallel training step over the latent hash tables (sm_100a).
//
// The forward of the CNC field reads a latent p only through sign(p) (STE_binary forward, ngp.py:26-31) and the backward
// only through the window |p| <= 1 (STE_binary backward, ngp.py:33-39): two bits per parameter carry everything a replica
// needs from a table row it does not own.  Under data parallelism (cnc_b200/dp.py) every rank runs Adam on 1/N of the rows
// (`cnc_adam_planes`: one pass that also emits the two bit planes of the updated rows), the planes are all-gathered
// (2 x 5 MB instead of the 161 MB an all-reduce moves twice) and `cnc_surrogate_fill` writes a stand-in latent with the
// same sign and window bit into the rows a rank does not own, so that everything downstream keeps reading `.params`.
//
// All three are HBM-streaming kernels: one pass, 16-byte accesses, grid = a multiple of the SM count.
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"

namespace cnc {

static inline int stream_blocks(uint64_t n_words) {
    // 148 SMs x 8 resident CTAs of 256 threads (8 warps, one 32-word chunk per warp and pass); never more CTAs than work
    const uint64_t want = ((n_words + 31) / 32 + 7) / 8;
    const uint64_t cap = 148ull * 8ull;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// Work layout shared by the three kernels: a warp owns 1024 consecutive parameters (32 plane words) per outer iteration;
// in inner iteration `it` lane l touches the float4 at chunk*1024 + it*128 + l*4, so every warp instruction is one fully
// coalesced 512-byte access.  The 4 bits a lane produces belong to plane word chunk*32 + it*4 + l/8 at bit (l%8)*4; an
// OR-reduction over the 8-lane group (3 shuffles) leaves the word in the group's first lane.
__device__ __forceinline__ uint32_t group8_or(uint32_t x) {
    x |= __shfl_xor_sync(0xffffffffu, x, 1);
    x |= __shfl_xor_sync(0xffffffffu, x, 2);
    x |= __shfl_xor_sync(0xffffffffu, x, 4);
    return x;
}

__global__ void planes_pack_kernel(const float *__restrict__ p, uint32_t *__restrict__ sign, uint32_t *__restrict__ mask,
                                   uint64_t n_words) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_chunks = (n_words + 31) / 32;
    for (uint64_t c = warp; c < n_chunks; c += n_warps) {
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const uint64_t w = c * 32 + it * 4 + (lane >> 3);
            uint32_t s = 0, m = 0;
            if (w < n_words) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(p + c * 1024 + it * 128 + lane * 4));
                const float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    s |= (uint32_t)(a[j] >= 0.f) << ((lane & 7u) * 4 + j);
                    m |= (uint32_t)(a[j] >= -1.f && a[j] <= 1.f) << ((lane & 7u) * 4 + j);
                }
            }
            s = group8_or(s);
            m = group8_or(m);
            if ((lane & 7u) == 0 && w < n_words) {
                sign[w] = s;
                if (mask) mask[w] = m;
            }
        }
    }
}

// parameters outside [keep_lo, keep_hi) (in plane words) get +-0.5 (inside the STE window) or +-1.5 (outside)
__global__ void surrogate_fill_kernel(float *__restrict__ p, const uint32_t *__restrict__ sign, const uint32_t *__restrict__ mask,
                                      uint64_t n_words, uint64_t keep_lo, uint64_t keep_hi) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_chunks = (n_words + 31) / 32;
    for (uint64_t c = warp; c < n_chunks; c += n_warps) {
        if (c * 32 >= keep_lo && c * 32 + 32 <= keep_hi) continue;   // the whole chunk is owned
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const uint64_t w = c * 32 + it * 4 + (lane >> 3);
            if (w >= n_words || (w >= keep_lo && w < keep_hi)) continue;
            const uint32_t s = __ldg(sign + w) >> ((lane & 7u) * 4), m = __ldg(mask + w) >> ((lane & 7u) * 4);
            float a[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float mag = ((m >> j) & 1u) ? 0.5f : 1.5f;
                a[j] = ((s >> j) & 1u) ? mag : -mag;
            }
            *reinterpret_cast<float4 *>(p + c * 1024 + it * 128 + lane * 4) = make_float4(a[0], a[1], a[2], a[3]);
        }
    }
}

// torch.optim.Adam (L2 weight decay folded into the gradient, no amsgrad, not maximize), fp32 state, same operation
// order as torch's fused kernel: g += wd*p; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g;
// p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).  grad_scale: the gradient arrives multiplied by it (loss scaling /
// sum-instead-of-mean reductions) and is divided out first.
__global__ void adam_planes_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m1, float *__restrict__ v2,
                                   uint32_t *__restrict__ sign, uint32_t *__restrict__ mask, uint64_t n_words, float lr, float b1,
                                   float b2, float eps, float wd, float bc1, float bc2_sqrt, float inv_scale, int ste_window) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_chunks = (n_words + 31) / 32;
    const float step = lr / bc1;
    for (uint64_t c = warp; c < n_chunks; c += n_warps) {
#pragma unroll 2
        for (int it = 0; it < 8; it++) {
            const uint64_t w = c * 32 + it * 4 + (lane >> 3);
            uint32_t s = 0, mk = 0;
            if (w < n_words) {
                const uint64_t i = c * 1024 + it * 128 + lane * 4;
                const float4 pv = *reinterpret_cast<const float4 *>(p + i);
                const float4 gv = __ldg(reinterpret_cast<const float4 *>(g + i));
                const float4 mv = *reinterpret_cast<const float4 *>(m1 + i);
                const float4 vv = *reinterpret_cast<const float4 *>(v2 + i);
                float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
                float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float gr = __fmul_rn(ga[j], inv_scale);
                    if (ste_window && !(pa[j] >= -1.f && pa[j] <= 1.f)) gr = 0.f;   // STE_binary backward (ngp.py:33-39), folded in
                    gr = __fmaf_rn(wd, pa[j], gr);
                    ma[j] = __fmaf_rn(1.f - b1, __fsub_rn(gr, ma[j]), ma[j]);
                    va[j] = __fmaf_rn(b2, va[j], __fmul_rn(__fmul_rn(1.f - b2, gr), gr));
                    const float den = __fadd_rn(__fdiv_rn(sqrtf(va[j]), bc2_sqrt), eps);
                    pa[j] = __fsub_rn(pa[j], __fmul_rn(step, __fdiv_rn(ma[j], den)));
                    s |= (uint32_t)(pa[j] >= 0.f) << ((lane & 7u) * 4 + j);
                    mk |= (uint32_t)(pa[j] >= -1.f && pa[j] <= 1.f) << ((lane & 7u) * 4 + j);
                }
                *reinterpret_cast<float4 *>(p + i) = make_float4(pa[0], pa[1], pa[2], pa[3]);
                *reinterpret_cast<float4 *>(m1 + i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
                *reinterpret_cast<float4 *>(v2 + i) = make_float4(va[0], va[1], va[2], va[3]);
            }
            s = group8_or(s);
            mk = group8_or(mk);
            if ((lane & 7u) == 0 && w < n_words) {
                if (sign) sign[w] = s;
                if (mask) mask[w] = mk;
            }
        }
    }
}

static bool ok16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool ok4(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// set bits per level of a sign plane: block (c, l) covers a chunk of level l's bytes (levels are byte aligned: row counts are
// multiples of 8).  The zeroth-order statistics of the rate term (+1 frequency per level, utils_bpp_acc.py:472-486) from
// 5 MB of bits in one launch instead of one fp32 reduction over the level's slice per level.
__global__ void __launch_bounds__(256) level_popcount_kernel(const uint8_t *__restrict__ bits, const int64_t *__restrict__ byte_offs,
                                                             unsigned long long *__restrict__ out) {
    const int l = blockIdx.y;
    const int64_t b0 = byte_offs[l], b1 = byte_offs[l + 1];
    unsigned long long acc = 0;
    for (int64_t i = b0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b1; i += (int64_t)gridDim.x * blockDim.x)
        acc += (unsigned)__popc((unsigned)__ldg(bits + i));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + l, acc);
}

}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_ste_planes_pack(const float *params, uint8_t *sign_bits, uint8_t *mask_bits, uint64_t n, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!params || !sign_bits) { set_error("ste_planes_pack: null pointer"); return CNC_EINVAL; }
    if (n % 32) { set_error("ste_planes_pack: n must be a multiple of 32 (whole uint32 words of the planes)"); return CNC_EINVAL; }
    if (!ok16(params) || !ok4(sign_bits) || (mask_bits && !ok4(mask_bits))) { set_error("ste_planes_pack: misaligned pointer"); return CNC_EINVAL; }
    planes_pack_kernel<<<stream_blocks(n / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        params, reinterpret_cast<uint32_t *>(sign_bits), reinterpret_cast<uint32_t *>(mask_bits), n / 32);
    return check_launch("ste_planes_pack");
}

int cnc_surrogate_fill(float *params, const uint8_t *sign_bits, const uint8_t *mask_bits, uint64_t n, uint64_t keep_lo,
                       uint64_t keep_hi, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!params || !sign_bits || !mask_bits) { set_error("surrogate_fill: null pointer"); return CNC_EINVAL; }
    if (n % 32 || keep_lo % 32 || keep_hi % 32 || keep_lo > keep_hi || keep_hi > n) {
        set_error("surrogate_fill: n, keep_lo, keep_hi must be multiples of 32 with keep_lo <= keep_hi <= n");
        return CNC_EINVAL;
    }
    if (!ok16(params) || !ok4(sign_bits) || !ok4(mask_bits)) { set_error("surrogate_fill: misaligned pointer"); return CNC_EINVAL; }
    surrogate_fill_kernel<<<stream_blocks(n / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        params, reinterpret_cast<const uint32_t *>(sign_bits), reinterpret_cast<const uint32_t *>(mask_bits), n / 32, keep_lo / 32,
        keep_hi / 32);
    return check_launch("surrogate_fill");
}

int cnc_level_popcount(const uint8_t *sign_bits, const int64_t *byte_offsets, int32_t n_levels, int64_t *out, cnc_stream_t stream) {
    if (n_levels <= 0) return CNC_OK;
    if (!sign_bits || !byte_offsets || !out) { set_error("level_popcount: null pointer"); return CNC_EINVAL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(out, 0, sizeof(int64_t) * (size_t)n_levels, s) != cudaSuccess) { set_error("level_popcount: memset failed"); return CNC_ECUDA; }
    level_popcount_kernel<<<dim3(64, (unsigned)n_levels), 256, 0, s>>>(sign_bits, byte_offsets, reinterpret_cast<unsigned long long *>(out));
    return check_launch("level_popcount");
}

int cnc_adam_planes(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, uint8_t *sign_bits, uint8_t *mask_bits,
                    uint64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    int32_t ste_window, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!params || !grad || !exp_avg || !exp_avg_sq) { set_error("adam_planes: null pointer"); return CNC_EINVAL; }
    if (n % 32 || step < 1 || grad_scale == 0.f) { set_error("adam_planes: n must be a multiple of 32, step >= 1, grad_scale != 0"); return CNC_EINVAL; }
    if (!ok16(params) || !ok16(grad) || !ok16(exp_avg) || !ok16(exp_avg_sq) || (sign_bits && !ok4(sign_bits)) || (mask_bits && !ok4(mask_bits))) {
        set_error("adam_planes: misaligned pointer");
        return CNC_EINVAL;
    }
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_planes_kernel<<<stream_blocks(n / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        params, grad, exp_avg, exp_avg_sq, reinterpret_cast<uint32_t *>(sign_bits), reinterpret_cast<uint32_t *>(mask_bits), n / 32, lr,
        beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2), 1.f / grad_scale, ste_window);
    return check_launch("adam_planes");
}

}  // extern "C"
