// context_mlp.cu -- context_model_3D (Linear(25,32) LeakyReLU Linear(32,32) LeakyReLU Linear(32,8), utils_bpp_acc.py:
// 247-251) as ONE forward and ONE backward kernel for the rate term of the training loss (utils_bpp_acc.py:533-706).
//
// Under autograd the reference (and the op-by-op path here) runs six cuBLAS SIMT GEMMs plus elementwise passes over
// [voxels, 32] activations that only exist to be read back once.  Here a thread owns a voxel:
//   forward : x [25] -> h1 -> h2 -> y [8] in registers (weights broadcast from shared memory), nothing else stored;
//   backward: h1, h2 recomputed, dz3 -> dz2 -> dz1 -> dx back-propagated in registers, and the weight gradients of a
//             256-voxel batch formed from [feature][voxel] tiles in shared memory (row stride 260 floats: the 16-byte
//             reads of 32 different rows are bank-conflict free): warp = 4 output rows (their tiles are broadcast
//             reads), lane = input column.  Every CTA keeps its sums in registers over all its batches and writes one
//             partial [2152] vector; the caller adds the partials in index order (deterministic, like cnc_wgrad).
// Packed weight layout = the one of context_fused.cu: W1T [25][32], b1 [32], W2T [32][32], b2 [32], W3T [32][8], b3 [8].
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {
namespace cmlp {

constexpr int NIN = 25, NH = 32, NO = 8, CT = 256;
constexpr int O_W1 = 0, O_B1 = O_W1 + NIN * NH, O_W2 = O_B1 + NH, O_B2 = O_W2 + NH * NH, O_W3 = O_B2 + NH,
              O_B3 = O_W3 + NH * NO, MLP_FLOATS = O_B3 + NO;  // 2152
constexpr int RS = 260;                                        // row stride of the [feature][voxel] tiles (floats)
constexpr int T_ROWS = 64;                                     // rows 0..31: dz, rows 32..63: layer inputs
constexpr uint32_t SM_W = 0, SM_X = SM_W + 2176 * 4, SM_T = SM_X + CT * NIN * 4, SM_BWD = SM_T + T_ROWS * RS * 4;  // 100864 B

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : __fmul_rn(x, 0.01f); }
__device__ __forceinline__ float slope(float h) { return h > 0.f ? 1.f : 0.01f; }   // leaky(z) > 0 <=> z > 0

// x -> h1, h2 (after LeakyReLU), y; bias first, inputs in ascending order (the FMA order of voxel_probs, context_fused.cu)
__device__ __forceinline__ void mlp_forward(const float *__restrict__ w, const float (&in)[NIN], float (&h1)[NH], float (&h2)[NH],
                                            float (&o)[NO]) {
#pragma unroll
    for (int j = 0; j < NH; j++) h1[j] = w[O_B1 + j];
#pragma unroll
    for (int i = 0; i < NIN; i++) {
#pragma unroll
        for (int j4 = 0; j4 < NH / 4; j4++) {
            const float4 t = *reinterpret_cast<const float4 *>(w + O_W1 + i * NH + 4 * j4);
            h1[4 * j4 + 0] = __fmaf_rn(in[i], t.x, h1[4 * j4 + 0]);
            h1[4 * j4 + 1] = __fmaf_rn(in[i], t.y, h1[4 * j4 + 1]);
            h1[4 * j4 + 2] = __fmaf_rn(in[i], t.z, h1[4 * j4 + 2]);
            h1[4 * j4 + 3] = __fmaf_rn(in[i], t.w, h1[4 * j4 + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < NH; j++) { h1[j] = leaky(h1[j]); h2[j] = w[O_B2 + j]; }
#pragma unroll
    for (int i = 0; i < NH; i++) {
#pragma unroll
        for (int j4 = 0; j4 < NH / 4; j4++) {
            const float4 t = *reinterpret_cast<const float4 *>(w + O_W2 + i * NH + 4 * j4);
            h2[4 * j4 + 0] = __fmaf_rn(h1[i], t.x, h2[4 * j4 + 0]);
            h2[4 * j4 + 1] = __fmaf_rn(h1[i], t.y, h2[4 * j4 + 1]);
            h2[4 * j4 + 2] = __fmaf_rn(h1[i], t.z, h2[4 * j4 + 2]);
            h2[4 * j4 + 3] = __fmaf_rn(h1[i], t.w, h2[4 * j4 + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < NH; j++) h2[j] = leaky(h2[j]);
#pragma unroll
    for (int k = 0; k < NO; k++) o[k] = w[O_B3 + k];
#pragma unroll
    for (int i = 0; i < NH; i++) {
        const float4 t0 = *reinterpret_cast<const float4 *>(w + O_W3 + i * NO), t1 = *reinterpret_cast<const float4 *>(w + O_W3 + i * NO + 4);
        o[0] = __fmaf_rn(h2[i], t0.x, o[0]); o[1] = __fmaf_rn(h2[i], t0.y, o[1]);
        o[2] = __fmaf_rn(h2[i], t0.z, o[2]); o[3] = __fmaf_rn(h2[i], t0.w, o[3]);
        o[4] = __fmaf_rn(h2[i], t1.x, o[4]); o[5] = __fmaf_rn(h2[i], t1.y, o[5]);
        o[6] = __fmaf_rn(h2[i], t1.z, o[6]); o[7] = __fmaf_rn(h2[i], t1.w, o[7]);
    }
}

__global__ void __launch_bounds__(CT) ctx_mlp_fwd_kernel(const float *__restrict__ X, const float *__restrict__ mlp,
                                                          float *__restrict__ Y, int64_t M) {
    __shared__ __align__(16) float w[MLP_FLOATS];
    __shared__ float xs[CT * NIN];
    const int tid = threadIdx.x;
    for (int i = tid; i < MLP_FLOATS; i += CT) w[i] = __ldg(mlp + i);
    for (int64_t base = (int64_t)blockIdx.x * CT; base < M; base += (int64_t)gridDim.x * CT) {
        const int n = (int)(M - base < CT ? M - base : CT);
        __syncthreads();   // weights visible / the previous tile is consumed
        for (int i = tid; i < n * NIN; i += CT) xs[i] = __ldg(X + base * NIN + i);   // coalesced
        __syncthreads();
        if (tid < n) {
            float in[NIN], h1[NH], h2[NH], o[NO];
#pragma unroll
            for (int i = 0; i < NIN; i++) in[i] = xs[tid * NIN + i];   // stride 25: conflict free
            mlp_forward(w, in, h1, h2, o);
            float4 *y = reinterpret_cast<float4 *>(Y + (base + tid) * NO);
            y[0] = make_float4(o[0], o[1], o[2], o[3]);
            y[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

// sum over the 256 voxels of a tile of T[ra][v] * T[rb][v], added to acc (4 rows ra0..ra0+3 at once: broadcast reads)
__device__ __forceinline__ void dot4(const float *__restrict__ T, int ra0, int rb, float (&acc)[4]) {
    const float4 *b = reinterpret_cast<const float4 *>(T + rb * RS);
    const float4 *a0 = reinterpret_cast<const float4 *>(T + (ra0 + 0) * RS), *a1 = reinterpret_cast<const float4 *>(T + (ra0 + 1) * RS),
                 *a2 = reinterpret_cast<const float4 *>(T + (ra0 + 2) * RS), *a3 = reinterpret_cast<const float4 *>(T + (ra0 + 3) * RS);
#pragma unroll 4
    for (int v = 0; v < CT / 4; v++) {
        const float4 x = b[v], p = a0[v], q = a1[v], r = a2[v], s = a3[v];
        acc[0] = __fmaf_rn(p.x, x.x, acc[0]); acc[0] = __fmaf_rn(p.y, x.y, acc[0]); acc[0] = __fmaf_rn(p.z, x.z, acc[0]); acc[0] = __fmaf_rn(p.w, x.w, acc[0]);
        acc[1] = __fmaf_rn(q.x, x.x, acc[1]); acc[1] = __fmaf_rn(q.y, x.y, acc[1]); acc[1] = __fmaf_rn(q.z, x.z, acc[1]); acc[1] = __fmaf_rn(q.w, x.w, acc[1]);
        acc[2] = __fmaf_rn(r.x, x.x, acc[2]); acc[2] = __fmaf_rn(r.y, x.y, acc[2]); acc[2] = __fmaf_rn(r.z, x.z, acc[2]); acc[2] = __fmaf_rn(r.w, x.w, acc[2]);
        acc[3] = __fmaf_rn(s.x, x.x, acc[3]); acc[3] = __fmaf_rn(s.y, x.y, acc[3]); acc[3] = __fmaf_rn(s.z, x.z, acc[3]); acc[3] = __fmaf_rn(s.w, x.w, acc[3]);
    }
}
__device__ __forceinline__ float dot1(const float *__restrict__ T, int ra, int rb, float acc) {
    const float4 *a = reinterpret_cast<const float4 *>(T + ra * RS), *b = reinterpret_cast<const float4 *>(T + rb * RS);
#pragma unroll 4
    for (int v = 0; v < CT / 4; v++) {
        const float4 p = a[v], x = b[v];
        acc = __fmaf_rn(p.x, x.x, acc); acc = __fmaf_rn(p.y, x.y, acc); acc = __fmaf_rn(p.z, x.z, acc); acc = __fmaf_rn(p.w, x.w, acc);
    }
    return acc;
}
__device__ __forceinline__ float rowsum(const float *__restrict__ T, int r, float acc) {
    const float4 *a = reinterpret_cast<const float4 *>(T + r * RS);
#pragma unroll 4
    for (int v = 0; v < CT / 4; v++) {
        const float4 p = a[v];
        acc = __fadd_rn(acc, p.x); acc = __fadd_rn(acc, p.y); acc = __fadd_rn(acc, p.z); acc = __fadd_rn(acc, p.w);
    }
    return acc;
}

__global__ void __launch_bounds__(CT) ctx_mlp_bwd_kernel(const float *__restrict__ X, const float *__restrict__ mlp,
                                                          const float *__restrict__ gY, float *__restrict__ gX,
                                                          float *__restrict__ partials, int64_t M) {
    extern __shared__ __align__(16) uint8_t smem[];
    float *w = reinterpret_cast<float *>(smem + SM_W), *xs = reinterpret_cast<float *>(smem + SM_X), *T = reinterpret_cast<float *>(smem + SM_T);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < MLP_FLOATS; i += CT) w[i] = __ldg(mlp + i);
    // this thread's share of the weight gradients, summed over all batches of the CTA
    float a1[4] = {0.f, 0.f, 0.f, 0.f};   // dW1T[i = lane][j = 4 warp + a]   (lane < 25)
    float a2[4] = {0.f, 0.f, 0.f, 0.f};   // dW2T[i = lane][j = 4 warp + a]
    float a3 = 0.f;                       // dW3T[j = lane][k = warp]
    float d1 = 0.f, d2 = 0.f, d3 = 0.f;   // warp 0: db1[lane], db2[lane], db3[lane < 8]
    for (int64_t base = (int64_t)blockIdx.x * CT; base < M; base += (int64_t)gridDim.x * CT) {
        const int n = (int)(M - base < CT ? M - base : CT);
        const bool live = tid < n;
        __syncthreads();
        for (int i = tid; i < n * NIN; i += CT) xs[i] = __ldg(X + base * NIN + i);
        __syncthreads();
        float h1[NH], h2[NH], dz3[NO];
        {
            float in[NIN], o[NO];
#pragma unroll
            for (int i = 0; i < NIN; i++) in[i] = live ? xs[tid * NIN + i] : 0.f;
            mlp_forward(w, in, h1, h2, o);
        }
        {
            float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
            if (live) {
                g0 = __ldg(reinterpret_cast<const float4 *>(gY + (base + tid) * NO));
                g1 = __ldg(reinterpret_cast<const float4 *>(gY + (base + tid) * NO) + 1);
            }
            dz3[0] = g0.x; dz3[1] = g0.y; dz3[2] = g0.z; dz3[3] = g0.w; dz3[4] = g1.x; dz3[5] = g1.y; dz3[6] = g1.z; dz3[7] = g1.w;
        }
        // ---- layer 3: dW3 = h2^T dz3, db3 = sum dz3          (rows of dead threads are zero: dz3 = 0)
#pragma unroll
        for (int k = 0; k < NO; k++) T[k * RS + tid] = dz3[k];
#pragma unroll
        for (int j = 0; j < NH; j++) T[(32 + j) * RS + tid] = live ? h2[j] : 0.f;
        __syncthreads();
        a3 = dot1(T, warp, 32 + lane, a3);
        if (warp == 0 && lane < NO) d3 = rowsum(T, lane, d3);
        __syncthreads();
        // dz2 = (dz3 W3) * leaky'(z2), in place of h2
#pragma unroll
        for (int j = 0; j < NH; j++) {
            const float4 t0 = *reinterpret_cast<const float4 *>(w + O_W3 + j * NO), t1 = *reinterpret_cast<const float4 *>(w + O_W3 + j * NO + 4);
            float s = __fmul_rn(dz3[0], t0.x);
            s = __fmaf_rn(dz3[1], t0.y, s); s = __fmaf_rn(dz3[2], t0.z, s); s = __fmaf_rn(dz3[3], t0.w, s);
            s = __fmaf_rn(dz3[4], t1.x, s); s = __fmaf_rn(dz3[5], t1.y, s); s = __fmaf_rn(dz3[6], t1.z, s); s = __fmaf_rn(dz3[7], t1.w, s);
            h2[j] = __fmul_rn(s, slope(h2[j]));
        }
        // ---- layer 2: dW2 = h1^T dz2, db2 = sum dz2
#pragma unroll
        for (int j = 0; j < NH; j++) { T[j * RS + tid] = live ? h2[j] : 0.f; T[(32 + j) * RS + tid] = live ? h1[j] : 0.f; }
        __syncthreads();
        dot4(T, 4 * warp, 32 + lane, a2);
        if (warp == 0) d2 = rowsum(T, lane, d2);
        __syncthreads();
        // dz1 = (dz2 W2) * leaky'(z1), in place of h1:  dh1[i] = sum_j W2T[i][j] dz2[j]
#pragma unroll
        for (int i = 0; i < NH; i++) {
            float s = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < NH / 4; j4++) {
                const float4 t = *reinterpret_cast<const float4 *>(w + O_W2 + i * NH + 4 * j4);
                s = __fmaf_rn(h2[4 * j4 + 0], t.x, s); s = __fmaf_rn(h2[4 * j4 + 1], t.y, s);
                s = __fmaf_rn(h2[4 * j4 + 2], t.z, s); s = __fmaf_rn(h2[4 * j4 + 3], t.w, s);
            }
            h1[i] = __fmul_rn(s, slope(h1[i]));
        }
        // ---- layer 1: dW1 = x^T dz1, db1 = sum dz1
#pragma unroll
        for (int j = 0; j < NH; j++) T[j * RS + tid] = live ? h1[j] : 0.f;
#pragma unroll
        for (int i = 0; i < NIN; i++) T[(32 + i) * RS + tid] = live ? xs[tid * NIN + i] : 0.f;
        __syncthreads();
        if (lane < NIN) dot4(T, 4 * warp, 32 + lane, a1);
        if (warp == 0) d1 = rowsum(T, lane, d1);
        // dx[i] = sum_j W1T[i][j] dz1[j]  -> own row of the x tile -> coalesced store
        if (live) {
#pragma unroll
            for (int i = 0; i < NIN; i++) {
                float s = 0.f;
#pragma unroll
                for (int j4 = 0; j4 < NH / 4; j4++) {
                    const float4 t = *reinterpret_cast<const float4 *>(w + O_W1 + i * NH + 4 * j4);
                    s = __fmaf_rn(h1[4 * j4 + 0], t.x, s); s = __fmaf_rn(h1[4 * j4 + 1], t.y, s);
                    s = __fmaf_rn(h1[4 * j4 + 2], t.z, s); s = __fmaf_rn(h1[4 * j4 + 3], t.w, s);
                }
                xs[tid * NIN + i] = s;
            }
        }
        __syncthreads();
        for (int i = tid; i < n * NIN; i += CT) gX[base * NIN + i] = xs[i];
    }
    float *P = partials + (size_t)blockIdx.x * MLP_FLOATS;
    if (lane < NIN) {
#pragma unroll
        for (int a = 0; a < 4; a++) P[O_W1 + lane * NH + 4 * warp + a] = a1[a];
    }
#pragma unroll
    for (int a = 0; a < 4; a++) P[O_W2 + lane * NH + 4 * warp + a] = a2[a];
    P[O_W3 + lane * NO + warp] = a3;
    if (warp == 0) {
        P[O_B1 + lane] = d1;
        P[O_B2 + lane] = d2;
        if (lane < NO) P[O_B3 + lane] = d3;
    }
}

}  // namespace cmlp
}  // namespace cnc

using namespace cnc;

extern "C" {

uint32_t cnc_ctx_mlp_floats(void) { return cmlp::MLP_FLOATS; }

int cnc_ctx_mlp_max_partials(void) {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    return 2 * n_sm;
}

int cnc_ctx_mlp_fwd(const float *X, const float *mlp, float *Y, int64_t M, cnc_stream_t stream) {
    if (M == 0) return CNC_OK;
    if (!X || !mlp || !Y) { set_error("ctx_mlp_fwd: null pointer"); return CNC_EINVAL; }
    if (reinterpret_cast<uintptr_t>(Y) & 15u) { set_error("ctx_mlp_fwd: Y must be 16-byte aligned"); return CNC_EINVAL; }
    const int64_t tiles = (M + cmlp::CT - 1) / cmlp::CT;
    const int64_t cap = 148ll * 8;
    cmlp::ctx_mlp_fwd_kernel<<<(unsigned)(tiles < cap ? tiles : cap), cmlp::CT, 0, static_cast<cudaStream_t>(stream)>>>(X, mlp, Y, M);
    return check_launch("ctx_mlp_fwd");
}

int cnc_ctx_mlp_bwd(const float *X, const float *mlp, const float *gY, float *gX, float *partials, uint32_t n_partials,
                    int64_t M, cnc_stream_t stream) {
    if (!X || !mlp || !gY || !gX || !partials || n_partials == 0) { set_error("ctx_mlp_bwd: null pointer"); return CNC_EINVAL; }
    if (reinterpret_cast<uintptr_t>(gY) & 15u) { set_error("ctx_mlp_bwd: gY must be 16-byte aligned"); return CNC_EINVAL; }
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(cmlp::ctx_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cmlp::SM_BWD) != cudaSuccess) {
            set_error("ctx_mlp_bwd: cannot reserve %u bytes of shared memory", cmlp::SM_BWD);
            return CNC_ECUDA;
        }
        attr = true;
    }
    // every CTA writes its partial vector, also the ones without a batch (zeros)
    cmlp::ctx_mlp_bwd_kernel<<<n_partials, cmlp::CT, cmlp::SM_BWD, static_cast<cudaStream_t>(stream)>>>(X, mlp, gY, gX, partials, M);
    return check_launch("ctx_mlp_bwd");
}

}  // extern "C"
