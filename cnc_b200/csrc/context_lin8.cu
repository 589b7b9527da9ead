// context_lin8.cu -- the plane context models of the rate term: y = x W^T + b with 8 outputs, forward and backward.
//
// Reference: context_model_2D[n-1] = nn.Linear(8 * (min(n,3) + 1) + 1, 8) over every plane vertex of the occupied region
// (utils_bpp_acc.py:386-393, :558-561): [N, 17 / 25 / 33] x [., 8] with N ~ 10^5..10^6 -- a tall-skinny GEMM that cuBLAS
// serves with SIMT sgemm kernels at a few per cent of anything (forward 0.6 ms, backward 1.3 ms per step for nine calls).
// The arithmetic is 8 K FMAs per row; the kernels below are one pass over x (forward) or x and gy (backward):
//
//   block = 256 rows.  x is staged through shared memory (coalesced 16-byte global loads -> [256][K] tile, K odd -> the
//   per-thread row reads are conflict-free); thread = row computes the 8 dot products (forward) or gx = gy W (backward,
//   written back through the same tile, coalesced).  Weight gradient: thread (o, k) walks the tile's 256 rows
//   (gy[n][o] * x[n][k], both from shared memory) -> one partial [8, K] + bias partial [8] per block, summed in index order
//   by the caller (deterministic, no atomics).
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {
namespace l8 {

constexpr int ROWS = 256;

template <int K>
__device__ __forceinline__ void load_tile(const float *__restrict__ src, int64_t row0, int64_t N, float *tile) {
    // [ROWS, K] floats, contiguous in global memory from row0: element e -> tile[e] (same linear layout)
    const int64_t total = (N - row0 < ROWS ? N - row0 : ROWS) * K;
    const float *g = src + row0 * K;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) tile[e] = __ldg(g + e);
}

template <int K>
__global__ void __launch_bounds__(ROWS) lin8_fwd_kernel(const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ b,
                                                        float *__restrict__ y, int64_t N) {
    __shared__ float tile[ROWS * K];
    __shared__ float w[8 * K + 8];
    for (int e = threadIdx.x; e < 8 * K; e += blockDim.x) w[e] = __ldg(W + e);
    if (threadIdx.x < 8) w[8 * K + threadIdx.x] = __ldg(b + threadIdx.x);
    const int64_t row0 = (int64_t)blockIdx.x * ROWS;
    load_tile<K>(x, row0, N, tile);
    __syncthreads();
    const int64_t n = row0 + threadIdx.x;
    if (n >= N) return;
    const float *xr = tile + threadIdx.x * K;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; o++) acc[o] = w[8 * K + o];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float v = xr[k];
#pragma unroll
        for (int o = 0; o < 8; o++) acc[o] = __fmaf_rn(v, w[o * K + k], acc[o]);
    }
    float4 *yo = reinterpret_cast<float4 *>(y + n * 8);
    yo[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    yo[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

template <int K>
__global__ void __launch_bounds__(ROWS) lin8_bwd_kernel(const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ gy,
                                                        float *__restrict__ gx, float *__restrict__ parts, int64_t N) {
    __shared__ float tile[ROWS * K];
    __shared__ float gys[ROWS * 9];     // [256][8] padded to 9: column reads by the (o, k) threads are conflict-free
    __shared__ float w[8 * K];
    for (int e = threadIdx.x; e < 8 * K; e += blockDim.x) w[e] = __ldg(W + e);
    const int64_t row0 = (int64_t)blockIdx.x * ROWS;
    const int rows = (int)(N - row0 < ROWS ? N - row0 : ROWS);
    load_tile<K>(x, row0, N, tile);
    const int64_t n = row0 + threadIdx.x;
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (n < N) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(gy + n * 8)), c = __ldg(reinterpret_cast<const float4 *>(gy + n * 8) + 1);
        g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = c.x; g[5] = c.y; g[6] = c.z; g[7] = c.w;
    }
#pragma unroll
    for (int o = 0; o < 8; o++) gys[threadIdx.x * 9 + o] = g[o];
    __syncthreads();
    // weight / bias gradient partial of this block: thread e = (o, k), k == K -> bias
    float *part = parts + (size_t)blockIdx.x * (8 * K + 8);
    for (int e = threadIdx.x; e < 8 * K + 8; e += blockDim.x) {
        const int o = e < 8 * K ? e / K : e - 8 * K, k = e < 8 * K ? e % K : -1;
        float acc = 0.f;
        for (int r = 0; r < rows; r++) acc = __fmaf_rn(gys[r * 9 + o], k >= 0 ? tile[r * K + k] : 1.0f, acc);
        part[e] = acc;
    }
    __syncthreads();
    // input gradient of this thread's row, through the x tile (no longer needed) for a coalesced store
    if (gx != nullptr) {
        float *xr = tile + threadIdx.x * K;
#pragma unroll
        for (int k = 0; k < K; k++) {
            float acc = 0.f;
#pragma unroll
            for (int o = 0; o < 8; o++) acc = __fmaf_rn(g[o], w[o * K + k], acc);
            xr[k] = acc;
        }
        __syncthreads();
        float *dst = gx + row0 * K;
        for (int64_t e = threadIdx.x; e < (int64_t)rows * K; e += blockDim.x) dst[e] = tile[e];
    }
}

template <int K>
static int run_fwd(const float *x, const float *W, const float *b, float *y, int64_t N, cudaStream_t s) {
    lin8_fwd_kernel<K><<<(unsigned)((N + ROWS - 1) / ROWS), ROWS, 0, s>>>(x, W, b, y, N);
    return check_launch("lin8_fwd");
}
template <int K>
static int run_bwd(const float *x, const float *W, const float *gy, float *gx, float *parts, int64_t N, cudaStream_t s) {
    lin8_bwd_kernel<K><<<(unsigned)((N + ROWS - 1) / ROWS), ROWS, 0, s>>>(x, W, gy, gx, parts, N);
    return check_launch("lin8_bwd");
}

// Bernoulli_entropy (utils_bpp_acc.py:1002-1013) summed: bits = sum_i -log2(p_i) (1 + x_i)/2 - log2(1 - p_i) (1 - x_i)/2 with p
// clamped to [1e-6, 1 - 1e-6]; block partials (summed by the caller in index order).  The backward writes d/dp (zero where the
// clamp is active, like torch.clamp's) and d/dx = (log2(1 - p) - log2(p)) / 2, both times the upstream scalar.
constexpr int EB = 256;
__global__ void __launch_bounds__(EB) bern_fwd_kernel(const float *__restrict__ x, const float *__restrict__ p, int64_t n,
                                                      float *__restrict__ parts) {
    __shared__ float red[EB / 32];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * EB + threadIdx.x; i < n; i += (int64_t)gridDim.x * EB) {
        const float pc = fminf(fmaxf(__ldg(p + i), 1e-6f), 1.0f - 1e-6f), xv = __ldg(x + i);
        const float a = __fmul_rn(-log2f(pc), __fdiv_rn(__fadd_rn(1.0f, xv), 2.0f));
        const float b = __fmul_rn(-log2f(__fsub_rn(1.0f, pc)), __fdiv_rn(__fsub_rn(1.0f, xv), 2.0f));
        acc = __fadd_rn(acc, __fadd_rn(a, b));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, d));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < EB / 32; w++) t = __fadd_rn(t, red[w]);
        parts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(EB) bern_bwd_kernel(const float *__restrict__ x, const float *__restrict__ p, const float *__restrict__ g,
                                                      int64_t n, float *__restrict__ gx, float *__restrict__ gp) {
    const float gs = __ldg(g);
    for (int64_t i = (int64_t)blockIdx.x * EB + threadIdx.x; i < n; i += (int64_t)gridDim.x * EB) {
        const float pv = __ldg(p + i), xv = __ldg(x + i);
        const float pc = fminf(fmaxf(pv, 1e-6f), 1.0f - 1e-6f), qc = __fsub_rn(1.0f, pc);
        const float wa = __fdiv_rn(__fadd_rn(1.0f, xv), 2.0f), wb = __fdiv_rn(__fsub_rn(1.0f, xv), 2.0f);
        if (gp) {
            // d/dpc [-log2(pc) wa - log2(1 - pc) wb] = (-wa / pc + wb / (1 - pc)) / ln 2; clamp passes the gradient inside [lo, hi]
            const bool inside = pv >= 1e-6f && pv <= 1.0f - 1e-6f;
            const float d = __fmul_rn(__fadd_rn(__fdiv_rn(-wa, pc), __fdiv_rn(wb, qc)), 1.4426950408889634f);
            gp[i] = inside ? __fmul_rn(gs, d) : 0.f;
        }
        if (gx) gx[i] = __fmul_rn(gs, __fmul_rn(0.5f, __fsub_rn(log2f(qc), log2f(pc))));
    }
}

}  // namespace l8
}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_lin8_rows_per_block(void) { return l8::ROWS; }

int cnc_lin8_fwd(const float *x, const float *W, const float *b, float *y, int64_t N, int32_t K, cnc_stream_t stream) {
    if (N == 0) return CNC_OK;
    if (!x || !W || !b || !y) { set_error("lin8_fwd: null pointer"); return CNC_EINVAL; }
    if (reinterpret_cast<uintptr_t>(y) & 15u) { set_error("lin8_fwd: y must be 16-byte aligned"); return CNC_EINVAL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (K) {
        case 9: return l8::run_fwd<9>(x, W, b, y, N, s);
        case 17: return l8::run_fwd<17>(x, W, b, y, N, s);
        case 25: return l8::run_fwd<25>(x, W, b, y, N, s);
        case 33: return l8::run_fwd<33>(x, W, b, y, N, s);
        default: set_error("lin8_fwd: K must be 9, 17, 25 or 33 (8 * context levels + 1)"); return CNC_ENOTSUP;
    }
}

int cnc_bernoulli_bits_blocks(int64_t n) {
    const int64_t b = (n + l8::EB - 1) / l8::EB;
    return (int)(b < 1 ? 1 : (b > 1184 ? 1184 : b));     // 148 SMs x 8 resident CTAs
}

int cnc_bernoulli_bits_fwd(const float *x, const float *p, int64_t n, float *parts, cnc_stream_t stream) {
    if (!x || !p || !parts || n < 0) { set_error("bernoulli_bits_fwd: bad argument"); return CNC_EINVAL; }
    l8::bern_fwd_kernel<<<cnc_bernoulli_bits_blocks(n), l8::EB, 0, static_cast<cudaStream_t>(stream)>>>(x, p, n, parts);
    return check_launch("bernoulli_bits_fwd");
}

int cnc_bernoulli_bits_bwd(const float *x, const float *p, const float *g, int64_t n, float *gx, float *gp, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!x || !p || !g || !(gx || gp) || n < 0) { set_error("bernoulli_bits_bwd: bad argument"); return CNC_EINVAL; }
    l8::bern_bwd_kernel<<<cnc_bernoulli_bits_blocks(n), l8::EB, 0, static_cast<cudaStream_t>(stream)>>>(x, p, g, n, gx, gp);
    return check_launch("bernoulli_bits_bwd");
}

int cnc_lin8_bwd(const float *x, const float *W, const float *gy, float *gx, float *parts, int64_t N, int32_t K, cnc_stream_t stream) {
    if (N == 0) return CNC_OK;
    if (!x || !W || !gy || !parts) { set_error("lin8_bwd: null pointer"); return CNC_EINVAL; }
    if (reinterpret_cast<uintptr_t>(gy) & 15u) { set_error("lin8_bwd: gy must be 16-byte aligned"); return CNC_EINVAL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (K) {
        case 9: return l8::run_bwd<9>(x, W, gy, gx, parts, N, s);
        case 17: return l8::run_bwd<17>(x, W, gy, gx, parts, N, s);
        case 25: return l8::run_bwd<25>(x, W, gy, gx, parts, N, s);
        case 33: return l8::run_bwd<33>(x, W, gy, gx, parts, N, s);
        default: set_error("lin8_bwd: K must be 9, 17, 25 or 33 (8 * context levels + 1)"); return CNC_ENOTSUP;
    }
}

}  // extern "C"
