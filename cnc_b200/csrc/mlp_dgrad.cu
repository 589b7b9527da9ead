// mlp_dgrad.cu -- input gradients of the field MLPs on the tensor cores, fp32-equivalent, ReLU mask fused.
//
// Reference behaviour restated: the backward of nn.Linear + ReLU in examples/radiance_fields/ngp.py:428-505
// (torch autograd: dX = dY W, an fp32 cuBLAS GEMM, then threshold_backward as a separate elementwise pass).
//
//   C[s, i] = [H[s, i] > 0] * sum_o Z[s, o] * W[o, i]        Z = dL/d(layer output) [Ns, ldz], W = nn.Linear weight,
//                                                             H = the ReLU output that fed this layer (nullable)
//
// Same machinery as the forward kernel, one layer at a time: the weight is packed once per step into K chunks of 32
// (tf32 "hi" plane + bf16 (lo, hi) pair plane, K-major, 128-byte swizzle) and streamed from L2 by bulk copies; four
// warps stage the 128-sample tile of Z chunk by chunk (hi plane + (hi, lo) pair plane); one warp issues, per k-step,
// one kind::tf32 MMA (hi*hi) and one kind::f16 MMA on the bf16 pairs (both correction terms) into two TMEM
// accumulators; four warps run the epilogue (sum of the accumulators, mask, 32-byte stores).  HBM-bound by design:
// 4 (No + 2 Ni) bytes per sample.
#include <cuda_runtime.h>

#include "common.cuh"
#include "tc05.cuh"

namespace cnc {
namespace dg {

using namespace tc;

constexpr int MAX_N = 192;
constexpr uint32_t A_HALF = 128u * 128u;                 // [128 x 32] fp32 chunk of Z (hi or pair plane): 16 KB
// a stage = Z chunk (hi + pair) + weight chunk [N x 32] (hi + pair): 72 KB at N = 160 (3 stages), 80 KB at N = 192 (2)
template <int N> struct Cfg {
    static constexpr uint32_t B_HALF = N * 128u;
    static constexpr uint32_t STAGE_BYTES = 2 * A_HALF + 2 * B_HALF;
    static constexpr uint32_t NSTAGE = N > 160 ? 2u : 3u;
    static constexpr uint32_t SMEM_BAR = NSTAGE * STAGE_BYTES;
    static constexpr uint32_t SMEM_EPI = SMEM_BAR + 128;            // 4 warps x [32 rows][20 floats]: store transposition
    static constexpr uint32_t SMEM_DYN = SMEM_EPI + 4 * 32 * 80;
};
constexpr int NTHREADS = 320;                            // warps 0-3 epilogue, 4-7 staging, 8 MMA, 9 weight stream

struct Args {
    const float *Z;      // [Ns, ldz]
    const float *blob;   // packed weight: KC chunks of (hi [N x 32], pair [N x 32]) swizzled
    const float *H;      // [Ns, ldh] nullable
    float *C;            // [Ns, ldc]
    uint32_t ldz, ldh, ldc, No, Ns;
};

// ---- weight packing --------------------------------------------------------------------------------------------------
// blob element (chunk kc, plane p, row n, column kk): o = 32 kc + kk; value = W[o, n + col_off] for n_first <= n < n_valid
// and o < No, else 0.  K-major SW128: row n of a chunk is 128 B, rows in groups of 8 (1 KB), 16-byte chunk index ^= n & 7.
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const float *__restrict__ W, uint32_t ldw, uint32_t No, int col_off,
                                                         uint32_t n_first, uint32_t n_valid, uint32_t N, uint32_t KC,
                                                         float *__restrict__ blob) {
    const uint32_t total = KC * 2u * N * 32u;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint32_t per_chunk = 2u * N * 32u;
    const uint32_t kc = i / per_chunk, rem = i - kc * per_chunk;
    const uint32_t plane = rem / (N * 32u), e = rem - plane * N * 32u;
    // physical position e inside the plane -> (row n, logical column kk)
    const uint32_t n = (e >> 8) * 8u + ((e >> 5) & 7u);          // 256 floats per 8-row group, 32 per row
    const uint32_t phys16 = (e & 31u) >> 2, kk = ((phys16 ^ (n & 7u)) << 2) | (e & 3u);
    const uint32_t o = kc * 32u + kk;
    float v = 0.f;
    if (o < No && n >= n_first && n < n_valid) v = W[(size_t)o * ldw + (uint32_t)((int)n + col_off)];
    const uint32_t hi = rna_tf32(v);
    const float h = __uint_as_float(hi);
    blob[i] = __uint_as_float(plane ? pack_bf16(__fsub_rn(v, h), h) : hi);   // B side pair (lo, hi)
}

template <int N>
__global__ void __launch_bounds__(NTHREADS, 1) dgrad_kernel(const Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t B_HALF = Cfg<N>::B_HALF, STAGE_BYTES = Cfg<N>::STAGE_BYTES, NSTAGE = Cfg<N>::NSTAGE, SMEM_BAR = Cfg<N>::SMEM_BAR;
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t KC = (a.No + 31u) / 32u;
    auto full = [&](uint32_t s) { return sbase + SMEM_BAR + 8u * s; };        // A staged (4 warps) + B landed (tx)
    auto empty = [&](uint32_t s) { return sbase + SMEM_BAR + 32u + 8u * s; }; // MMAs that read the stage retired
    const uint32_t acc_full = sbase + SMEM_BAR + 64u, acc_free = sbase + SMEM_BAR + 72u;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NSTAGE; s++) { mbar_init(full(s), 5); mbar_init(empty(s), 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_free, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SMEM_BAR + 112u), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *reinterpret_cast<volatile uint32_t *>(smem + SMEM_BAR + 112u);   // written by tcgen05.alloc
    const uint32_t ntiles = (a.Ns + 127u) / 128u;
    const uint32_t my = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

    if (warp < 4) {
        // =============================== epilogue ===============================
        // Global traffic of the epilogue is laid out so that a warp instruction touches few 128-byte lines: with one row
        // per lane a 16-byte access per lane hits 32 different lines (measured: the mask loads and the result stores
        // written that way cost 2/3 of the kernel).  Loads: 8 lanes per row.  Stores: through a [32][16] transposition
        // buffer, 4 lanes per row.
        const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16);
        float *scratch = reinterpret_cast<float *>(smem + Cfg<N>::SMEM_EPI) + warp * 32 * 20;
        constexpr int NJ = (N + 31) / 32;
        for (uint32_t it = 0; it < my; it++) {
            const uint32_t row0 = (blockIdx.x + it * gridDim.x) * 128u + (uint32_t)warp * 32u;   // first row of this warp
            // ReLU mask of this lane's row (row0 + lane), fetched while the MMAs of the tile run.
            // word j: bit 8e + m <=> column 32j + 4m + e
            uint32_t mbits[NJ];
#pragma unroll
            for (int j = 0; j < NJ; j++) mbits[j] = 0xFFFFFFFFu;
            if (a.H != nullptr) {
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    float4 hv[8];   // all eight loads of the column group in flight before the first ballot consumes one
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t r = row0 + 4u * i + ((uint32_t)lane >> 3), col = 32u * j + 4u * ((uint32_t)lane & 7u);
                        hv[i] = (r < a.Ns && col < (uint32_t)N) ? __ldg(reinterpret_cast<const float4 *>(a.H + (size_t)r * a.ldh + col))
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const uint32_t bx = __ballot_sync(0xFFFFFFFFu, hv[i].x > 0.f), by = __ballot_sync(0xFFFFFFFFu, hv[i].y > 0.f),
                                       bz = __ballot_sync(0xFFFFFFFFu, hv[i].z > 0.f), bw = __ballot_sync(0xFFFFFFFFu, hv[i].w > 0.f);
                        if (i == (lane >> 2)) {   // the iteration that carried this lane's row: its 8 lanes are 8 (lane & 3) ..
                            const uint32_t sh = 8u * ((uint32_t)lane & 3u);
                            mbits[j] = ((bx >> sh) & 0xFFu) | (((by >> sh) & 0xFFu) << 8) | (((bz >> sh) & 0xFFu) << 16) | (((bw >> sh) & 0xFFu) << 24);
                        }
                    }
                }
            }
            mbar_wait(acc_full, it & 1u);
            tc_fence_after();
#pragma unroll
            for (uint32_t c = 0; c < (uint32_t)N; c += 16) {
                uint32_t m[16], sm[16];
                tmem_ld16(tl + c, m);
                tmem_ld16(tl + (uint32_t)N + c, sm);
                tc_wait_ld();
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t cc = c + k;   // column; mask bit 8 (cc & 3) + ((cc & 31) >> 2) of word cc >> 5
                    const bool on = ((mbits[cc >> 5] >> (8u * (cc & 3u) + ((cc & 31u) >> 2))) & 1u) != 0u;
                    v[k] = on ? __fadd_rn(__uint_as_float(m[k]), __uint_as_float(sm[k])) : 0.f;
                }
                __syncwarp();   // the previous block has left the buffer
#pragma unroll
                for (int q = 0; q < 4; q++)
                    *reinterpret_cast<float4 *>(scratch + lane * 20 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t rl = 8u * i + ((uint32_t)lane >> 2), c4 = (uint32_t)lane & 3u, r = row0 + rl;
                    const float4 o = *reinterpret_cast<const float4 *>(scratch + rl * 20 + 4 * c4);
                    if (r < a.Ns) *reinterpret_cast<float4 *>(a.C + (size_t)r * a.ldc + c + 4u * c4) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);
        }
    } else if (warp < 8) {
        // =============================== Z staging ===============================
        const uint32_t t = threadIdx.x - 128u;        // 0..127: row of the tile
        const uint32_t total = my * KC;               // chunks this CTA stages, flat over (tile, k chunk)
        // the 8 x 16-byte loads of a chunk are issued two chunks ahead of their use: with four staging warps one chunk
        // in flight is 16 KB per SM, a third of what the DRAM latency needs at this SM's share of the bandwidth
        // lanes 0-7 read the eight float4 of one row, the next eight lanes the next row: a warp instruction covers four
        // whole 128-byte lines (row-per-thread would touch 32 lines per instruction)
        const uint32_t wq = (t >> 5) * 32u + ((uint32_t)lane >> 3), jq = (uint32_t)lane & 7u;   // first row of this lane, its float4
        auto load = [&](uint32_t g, float4 (&v)[8]) {
            const uint32_t it = g / KC, kc = g - it * KC;
            const uint32_t row0 = (blockIdx.x + it * gridDim.x) * 128u + wq;
            const uint32_t col = kc * 32u + 4u * jq;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t row = row0 + 4u * i;
                v[i] = (row < a.Ns && col < a.No) ? __ldg(reinterpret_cast<const float4 *>(a.Z + (size_t)row * a.ldz + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto process = [&](uint32_t g, const float4 (&v)[8]) {
            const uint32_t s = g % NSTAGE, use = g / NSTAGE;
            if (use > 0) mbar_wait(empty(s), (use - 1) & 1u);
            uint8_t *st = smem + s * STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t r = wq + 4u * i;       // row inside the tile
                uint4 hi, pr;
                split_tf32(v[i].x, hi.x, pr.x); split_tf32(v[i].y, hi.y, pr.y);
                split_tf32(v[i].z, hi.z, pr.z); split_tf32(v[i].w, hi.w, pr.w);
                uint8_t *p = st + (r >> 3) * 1024u + (r & 7u) * 128u + ((jq ^ (r & 7u)) << 4);
                *reinterpret_cast<uint4 *>(p) = hi;
                *reinterpret_cast<uint4 *>(p + A_HALF) = pr;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full(s));
        };
        float4 v0[8], v1[8], v2[8];
        if (0 < total) load(0, v0);
        if (1 < total) load(1, v1);
        for (uint32_t g = 0; g < total; g += 3) {
            if (g + 2 < total) load(g + 2, v2);
            process(g, v0);
            if (g + 1 < total) {
                if (g + 3 < total) load(g + 3, v0);
                process(g + 1, v1);
            }
            if (g + 2 < total) {
                if (g + 4 < total) load(g + 4, v1);
                process(g + 2, v2);
            }
        }
    } else if (warp == 8) {
        // =============================== MMA issue ===============================
        constexpr uint32_t idt = idesc_tf32<N>(), idb = idesc_bf16<N>();
        uint32_t g = 0;
        for (uint32_t it = 0; it < my; it++) {
            if (it > 0) { mbar_wait_spin(acc_free, (it - 1) & 1u); tc_fence_after(); }
            for (uint32_t kc = 0; kc < KC; kc++, g++) {
                const uint32_t s = g % NSTAGE;
                mbar_wait_spin(full(s), (g / NSTAGE) & 1u);
                tc_fence_after();
                const uint32_t st = sbase + s * STAGE_BYTES;
                const uint64_t ah = smem_desc(st), ap = smem_desc(st + A_HALF), bh = smem_desc(st + 2 * A_HALF),
                               bp = smem_desc(st + 2 * A_HALF + B_HALF);
#pragma unroll
                for (uint32_t k4 = 0; k4 < 4; k4++) {
                    const uint32_t acc = (kc == 0 && k4 == 0) ? 0u : 1u;
                    mma_ss_bf16(tbase + (uint32_t)N, ap + 2 * k4, bp + 2 * k4, idb, acc);   // Zhi*Wlo + Zlo*Whi
                    mma_ss(tbase, ah + 2 * k4, bh + 2 * k4, idt, acc);                      // Zhi*Whi
                }
                tc_commit_elect(empty(s));
            }
            tc_commit_elect(acc_full);
        }
    } else {
        // =============================== weight stream ===============================
        const uint8_t *blob = reinterpret_cast<const uint8_t *>(a.blob);
        uint32_t g = 0;
        for (uint32_t it = 0; it < my; it++) {
            for (uint32_t kc = 0; kc < KC; kc++, g++) {
                const uint32_t s = g % NSTAGE, use = g / NSTAGE;
                if (use > 0) mbar_wait_spin(empty(s), (use - 1) & 1u);
                bulk_g2s_elect(sbase + s * STAGE_BYTES + 2 * A_HALF, blob + (size_t)kc * 2u * B_HALF, 2u * B_HALF, full(s));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

}  // namespace dg
}  // namespace cnc

using namespace cnc;

template <int N>
static int launch_dgrad(const dg::Args &a, cudaStream_t s) {
    static bool attr_set = false;
    static int n_sm = 0;
    if (!attr_set) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(dg::dgrad_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dg::Cfg<N>::SMEM_DYN) != cudaSuccess) {
            set_error("dgrad: cannot reserve %u bytes of shared memory", dg::Cfg<N>::SMEM_DYN);
            return CNC_ECUDA;
        }
        attr_set = true;
    }
    const uint32_t ntiles = (a.Ns + 127u) / 128u;
    const uint32_t grid = ntiles < (uint32_t)n_sm ? ntiles : (uint32_t)n_sm;
    dg::dgrad_kernel<N><<<grid, dg::NTHREADS, dg::Cfg<N>::SMEM_DYN, s>>>(a);
    return check_launch("dgrad");
}

extern "C" {

uint32_t cnc_dgrad_blob_floats(uint32_t No, uint32_t N) { return ((No + 31u) / 32u) * 2u * N * 32u; }

int cnc_dgrad_pack(const float *W, uint32_t ldw, uint32_t No, int32_t col_off, uint32_t n_first, uint32_t n_valid, uint32_t N,
                   float *blob, cnc_stream_t stream) {
    if (!W || !blob) { set_error("dgrad_pack: null pointer"); return CNC_EINVAL; }
    if (N == 0 || N > dg::MAX_N || (N & 15u) || No == 0 || No > 160 || n_valid > N || n_first > n_valid) {
        set_error("dgrad_pack: unsupported shape");
        return CNC_ENOTSUP;
    }
    const uint32_t KC = (No + 31u) / 32u, total = KC * 2u * N * 32u;
    dg::pack_dgrad_kernel<<<div_up(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(W, ldw, No, col_off, n_first, n_valid, N, KC, blob);
    return check_launch("dgrad_pack");
}

int cnc_dgrad(const float *Z, uint32_t ldz, uint32_t No, const float *blob, uint32_t N, const float *H, uint32_t ldh, float *C,
              uint32_t ldc, uint32_t Ns, cnc_stream_t stream) {
    if (Ns == 0) return CNC_OK;
    if (!Z || !blob || !C) { set_error("dgrad: null pointer"); return CNC_EINVAL; }
    if ((ldz & 3u) || (ldc & 3u) || (H && (ldh & 3u)) || ldz < No || ldc < N || (H && ldh < N) || No == 0 || No > 160 || (No & 3u)) {
        set_error("dgrad: leading dimensions must be multiples of 4 and cover the columns used (No <= 160, multiple of 4)");
        return CNC_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(blob) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(H)) & 15u) {
        set_error("dgrad: pointers must be 16-byte aligned");
        return CNC_EINVAL;
    }
    dg::Args a{Z, blob, H, C, ldz, ldh, ldc, No, Ns};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (N) {
        case 80: return launch_dgrad<80>(a, s);
        case 160: return launch_dgrad<160>(a, s);
        case 192: return launch_dgrad<192>(a, s);
        case 32: return launch_dgrad<32>(a, s);
        default: set_error("dgrad: output width %u is not instantiated (32, 80, 160, 192)", N); return CNC_ENOTSUP;
    }
}

}  // extern "C"
