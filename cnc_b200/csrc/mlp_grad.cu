// mlp_grad.cu -- weight gradients of the field MLPs on the tensor cores, fp32-equivalent (3xTF32).
//
// Reference behaviour restated: the backward of the five nn.Linear layers of examples/radiance_fields/ngp.py:428-505
// (torch autograd: dW = dY^T X, an fp32 cuBLAS GEMM whose contraction runs over the N_s samples of the batch).
//
//   C[i, o] = sum_s X[s, i] * Z[s, o]          X = layer input [N_s, ldx] (first Mi columns), Z = dL/d(layer output)
//
// Both operands live in HBM sample-major, i.e. the contraction index is the slow one: exactly the "MN-major" operand
// form of tcgen05 (shared-memory atoms of 4 samples x 32 features, 128-byte rows swizzled in 32-byte units), so no transpose pass exists
// anywhere.  One persistent CTA per SM walks slabs of 32 samples: eight warps stage the slab (coalesced 16-byte loads,
// error-compensated hi/lo split, swizzled st.shared) into a two-stage ring while one warp issues
// D += Xhi^T Zhi + Xhi^T Zlo + Xlo^T Zhi (kind::tf32, M = 128 per block of input features, N = padded output width,
// K = 8 samples) into TMEM.  Every CTA writes its partial [Mi, No] tile; the caller sums the <= 148 partials (fixed
// order -> deterministic gradients, unlike an atomic reduction).  The kernel is HBM-bound by design: 4 (Mi + No)
// bytes per sample against 6 Mi No tensor FLOPs.
#include <cuda_runtime.h>

#include "common.cuh"
#include "tc05.cuh"

namespace cnc {
namespace mg {

using namespace tc;

constexpr int KS = 32;                 // samples per slab (4 k-steps of 8)
constexpr int NSTAGE = 2;
constexpr int MAX_MI = 256, MAX_NO = 160;
constexpr uint32_t X_HALF = KS * MAX_MI * 4;     // 32 KB: hi (or lo) of a slab of X, [k-step][atom][1 KB]
constexpr uint32_t Z_HALF = KS * MAX_NO * 4;     // 20 KB
constexpr uint32_t STAGE_BYTES = 2 * X_HALF + 2 * Z_HALF;   // 104 KB
constexpr uint32_t SMEM_BAR = NSTAGE * STAGE_BYTES;
constexpr uint32_t SMEM_DYN = SMEM_BAR + 64;
constexpr int NSTAGER = 256;           // staging threads (8 warps); warp 8 issues the MMAs
constexpr int NTHREADS = NSTAGER + 32;

struct Args {
    const float *X;
    const float *Z;
    float *P;          // [gridDim.x, Mi, No] partial sums
    uint32_t ldx, ldz, Mi, No, Ns;
};

__device__ __forceinline__ void split3(float v, uint32_t &hi, uint32_t &lo) {
    hi = rna_tf32(v);
    lo = __float_as_uint(__fsub_rn(v, __uint_as_float(hi)));   // the MMA reads its top 19 bits: 2^-23 |v|
}

// byte offset of the float4 m4 (= feature / 4) of sample row kr (0..7) inside a k-step tile: consecutive 1 KB blocks of
// 32 features, each two 4-row atoms of the SW128_32B layout (tc05.cuh)
__device__ __forceinline__ uint32_t mn_off(uint32_t kr, uint32_t m4) {
    return (m4 >> 3) * 1024u + kr * 128u + ((((m4 >> 1) & 3u) ^ (kr & 3u)) << 5) + ((m4 & 1u) << 4);
}

// MI: X columns (multiple of 32), NO: Z columns (multiple of 16); ONES: row MI of the result = column sums of Z
// (a virtual all-ones X column, synthesised in shared memory: the bias gradient for free)
template <int MI, int NO, bool ONES>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_kernel(const Args a) {
    constexpr int NO_PAD = NO <= 96 ? 96 : 160;   // MMA N
    constexpr uint32_t MI_EFF = MI + (ONES ? 1 : 0);
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5;
    constexpr uint32_t nblk = (MI_EFF + 127u) / 128u;                        // M blocks of 128 input features
    constexpr uint32_t natx_pad = nblk * 4u, natz = (NO_PAD + 31) / 32;
    constexpr uint32_t xstep = natx_pad * 1024u, zstep = natz * 1024u;       // bytes per k-step tile
    auto full = [&](uint32_t s) { return sbase + SMEM_BAR + 8u * s; };
    auto empty = [&](uint32_t s) { return sbase + SMEM_BAR + 16u + 8u * s; };
    const uint32_t done = sbase + SMEM_BAR + 32u;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NSTAGE; s++) { mbar_init(full(s), NSTAGER / 32); mbar_init(empty(s), 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // padding atoms / columns are never written by the staging loop: clear both stages once
    for (uint32_t i = threadIdx.x; i < NSTAGE * STAGE_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;

    const uint32_t nslab = (a.Ns + KS - 1) / KS;
    const uint32_t my = blockIdx.x < nslab ? (nslab - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

    if (warp < 8) {
        // =============================== staging ===============================
        constexpr uint32_t x4 = MI / 4, z4 = NO / 4;               // float4 per row
        constexpr uint32_t nx = KS * x4, nz = KS * z4;
        constexpr int JX = (nx + NSTAGER - 1) / NSTAGER, JZ = (nz + NSTAGER - 1) / NSTAGER;
        // The loads of slab it + 1 are issued BEFORE slab it is converted and stored (a second register set): the global loads
        // of a slab and the split / swizzled stores of the previous one overlap instead of alternating, which is what kept the
        // kernel at 60 % of the HBM peak (and made it the most sensitive kernel of the step to anything else using memory).
        auto load_slab = [&](uint32_t it, float4 (&vx)[JX], float4 (&vz)[JZ]) {
            const uint32_t row0 = (blockIdx.x + it * gridDim.x) * KS;
#pragma unroll
            for (int j = 0; j < JX; j++) {
                const uint32_t i = threadIdx.x + j * NSTAGER, r = i / x4, c4 = i % x4, row = row0 + r;
                vx[j] = (i < nx && row < a.Ns) ? __ldg(reinterpret_cast<const float4 *>(a.X + (size_t)row * a.ldx) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < JZ; j++) {
                const uint32_t i = threadIdx.x + j * NSTAGER, r = i / z4, c4 = i % z4, row = row0 + r;
                vz[j] = (i < nz && row < a.Ns) ? __ldg(reinterpret_cast<const float4 *>(a.Z + (size_t)row * a.ldz) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        constexpr bool AHEAD = MI <= 160;      // (the 256-wide layer holds 13 float4 per slab and thread: a second set costs more than it hides)
        float4 vx[JX], vz[JZ];
        if (AHEAD && my) load_slab(0, vx, vz);
        for (uint32_t it = 0; it < my; it++) {
            const uint32_t slab = blockIdx.x + it * gridDim.x, s = it % NSTAGE, use = it / NSTAGE;
            const uint32_t row0 = slab * KS;
            float4 wx[AHEAD ? JX : 1], wz[AHEAD ? JZ : 1];
            if (AHEAD) {
                if (it + 1 < my) load_slab(it + 1, reinterpret_cast<float4 (&)[JX]>(wx), reinterpret_cast<float4 (&)[JZ]>(wz));
            } else {
                load_slab(it, vx, vz);
            }
            if (use > 0) mbar_wait(empty(s), (use - 1) & 1u);
            uint8_t *st = smem + s * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < JX; j++) {
                const uint32_t i = threadIdx.x + j * NSTAGER, r = i / x4, c4 = i % x4;
                if (i < nx) {
                    uint4 hi, lo;
                    split3(vx[j].x, hi.x, lo.x); split3(vx[j].y, hi.y, lo.y); split3(vx[j].z, hi.z, lo.z); split3(vx[j].w, hi.w, lo.w);
                    const uint32_t off = (r >> 3) * xstep + mn_off(r & 7u, c4);
                    *reinterpret_cast<uint4 *>(st + off) = hi;
                    *reinterpret_cast<uint4 *>(st + X_HALF + off) = lo;
                }
            }
#pragma unroll
            for (int j = 0; j < JZ; j++) {
                const uint32_t i = threadIdx.x + j * NSTAGER, r = i / z4, c4 = i % z4;
                if (i < nz) {
                    uint4 hi, lo;
                    split3(vz[j].x, hi.x, lo.x); split3(vz[j].y, hi.y, lo.y); split3(vz[j].z, hi.z, lo.z); split3(vz[j].w, hi.w, lo.w);
                    const uint32_t off = (r >> 3) * zstep + mn_off(r & 7u, c4);
                    *reinterpret_cast<uint4 *>(st + 2 * X_HALF + off) = hi;
                    *reinterpret_cast<uint4 *>(st + 2 * X_HALF + Z_HALF + off) = lo;
                }
            }
            if (ONES && threadIdx.x < KS) {   // the virtual X column MI: 1 for live samples
                const uint32_t r = threadIdx.x;
                const uint32_t off = (r >> 3) * xstep + mn_off(r & 7u, MI / 4);
                *reinterpret_cast<uint32_t *>(st + off) = (row0 + r < a.Ns) ? 0x3F800000u : 0u;
            }
            fence_async_smem();
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(full(s));
            if (AHEAD && it + 1 < my) {
#pragma unroll
                for (int j = 0; j < (AHEAD ? JX : 0); j++) vx[j] = wx[j];
#pragma unroll
                for (int j = 0; j < (AHEAD ? JZ : 0); j++) vz[j] = wz[j];
            }
        }
        // =============================== epilogue: TMEM -> partial tile ===============================
        if (warp < 4) {
            mbar_wait(done, 0);
            tc_fence_after();
            float *P = a.P + (size_t)blockIdx.x * MI_EFF * NO;
            for (uint32_t blk = 0; blk < nblk; blk++) {
                const uint32_t m = blk * 128u + (uint32_t)warp * 32u + (threadIdx.x & 31u);
                const uint32_t tl = tbase + ((uint32_t)(warp * 32) << 16) + blk * (uint32_t)NO_PAD;
                for (uint32_t c = 0; c < (uint32_t)NO_PAD; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(tl + c, v);
                    tc_wait_ld();
                    if (m < MI_EFF) {
#pragma unroll
                        for (int k = 0; k < 16; k += 4)
                            if (c + k < (uint32_t)NO)
                                *reinterpret_cast<float4 *>(P + (size_t)m * NO + c + k) =
                                    make_float4(my ? __uint_as_float(v[k]) : 0.f, my ? __uint_as_float(v[k + 1]) : 0.f,
                                                my ? __uint_as_float(v[k + 2]) : 0.f, my ? __uint_as_float(v[k + 3]) : 0.f);
                    }
                }
            }
        }
    } else {
        // =============================== MMA issue ===============================
        constexpr uint32_t id = idesc_tf32_mn<NO_PAD>();
        for (uint32_t it = 0; it < my; it++) {
            const uint32_t s = it % NSTAGE;
            mbar_wait_spin(full(s), (it / NSTAGE) & 1u);
            tc_fence_after();
            const uint32_t st = sbase + s * STAGE_BYTES;
#pragma unroll 1
            for (uint32_t ks = 0; ks < KS / 8; ks++) {
                const uint64_t zh = smem_desc_mn(st + 2 * X_HALF + ks * zstep, 1024u, 512u),
                               zl = smem_desc_mn(st + 2 * X_HALF + Z_HALF + ks * zstep, 1024u, 512u);
                for (uint32_t blk = 0; blk < nblk; blk++) {
                    const uint64_t xh = smem_desc_mn(st + ks * xstep + blk * 4096u, 1024u, 512u),
                                   xl = smem_desc_mn(st + X_HALF + ks * xstep + blk * 4096u, 1024u, 512u);
                    const uint32_t d = tbase + blk * (uint32_t)NO_PAD;
                    mma_ss(d, xh, zl, id, (it == 0 && ks == 0) ? 0u : 1u);   // the two small products first
                    mma_ss(d, xl, zh, id, 1u);
                    mma_ss(d, xh, zh, id, 1u);
                }
            }
            tc_commit_elect(empty(s));
        }
        tc_commit_elect(done);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

}  // namespace mg
}  // namespace cnc

using namespace cnc;

template <int MI, int NO, bool ONES>
static int launch_wgrad(const mg::Args &a, uint32_t n_partials, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(mg::wgrad_kernel<MI, NO, ONES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mg::SMEM_DYN) != cudaSuccess) {
            set_error("wgrad: cannot reserve %u bytes of shared memory", mg::SMEM_DYN);
            return CNC_ECUDA;
        }
        attr_set = true;
    }
    mg::wgrad_kernel<MI, NO, ONES><<<n_partials, mg::NTHREADS, mg::SMEM_DYN, s>>>(a);
    return check_launch("wgrad");
}

extern "C" {

int cnc_wgrad_max_partials(void) {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    return n_sm;
}

int cnc_wgrad(const float *X, uint32_t ldx, uint32_t Mi, const float *Z, uint32_t ldz, uint32_t No, int with_ones,
              float *partials, uint32_t n_partials, uint32_t Ns, cnc_stream_t stream) {
    if (!X || !Z || !partials) { set_error("wgrad: null pointer"); return CNC_EINVAL; }
    if ((ldx & 3u) || (ldz & 3u) || ldx < Mi || ldz < No || n_partials == 0) {
        set_error("wgrad: leading dimensions must be multiples of 4 and cover the columns used");
        return CNC_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(partials)) & 15u) {
        set_error("wgrad: pointers must be 16-byte aligned");
        return CNC_EINVAL;
    }
    if (with_ones && Mi + 1 > 2 * 128) { set_error("wgrad: with_ones needs Mi < 256 (two M blocks of shared memory)"); return CNC_ENOTSUP; }
    mg::Args a{X, Z, partials, ldx, ldz, Mi, No, Ns};
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool o = with_ones != 0;
#define CNC_WG(MI, NO) if (Mi == MI && No == NO) return o ? launch_wgrad<MI, NO, true>(a, n_partials, s) : launch_wgrad<MI, NO, false>(a, n_partials, s);
    if (Mi == 256 && No == 160) return launch_wgrad<256, 160, false>(a, n_partials, s);
    CNC_WG(160, 160) CNC_WG(96, 160) CNC_WG(160, 80) CNC_WG(160, 16) CNC_WG(32, 16) CNC_WG(64, 32)
#undef CNC_WG
    set_error("wgrad: shape (Mi=%u, No=%u) is not instantiated: (256,160) (160,160) (96,160) (160,80) (160,16) (32,16) (64,32)", Mi, No);
    return CNC_ENOTSUP;
}

}  // extern "C"
