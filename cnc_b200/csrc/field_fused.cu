// field_fused.cu -- the radiance field forward as ONE persistent kernel per batch of samples:
//   hash-grid encode (3D 12 levels + 3 planes x 4 levels, 1-bit sign tables) + 63-d frequency
//   embedding -> Linear(255,160)+ReLU -> Linear(160,80) -> sigma = exp(h0-1)*selector,
//   [SH16(dir) | geo79] -> Linear(95,160)+ReLU -> Linear(160,160)+ReLU -> Linear(160,3) -> sigmoid.
//
// Reference behaviour restated: examples/radiance_fields/ngp.py:514-566 (query_density, _query_rgb,
// forward), :620-645 (compose_3D_2D_embed), :569-617 (Embedder), :412-425 (tcnn SH, Appendix C of
// SURVEY.md).  The reference runs this as 4 gather kernels + ~60 elementwise kernels + 5 cuBLAS
// SGEMMs with a [N,255] activation round trip through HBM; here the 255-wide activation never
// leaves the SM.
//
// B200 mapping (as built; DESIGN.md 3.1 has the measurements behind each choice)
//   * persistent grid, one CTA per SM (576 threads), one tile = 128 samples = one UMMA M tile.
//   * warps 0-15, the gather / epilogue warps: thread (r, q) = sample row r (= TMEM lane) x octet q of a 32-column K chunk.
//     Per chunk it gathers one level of one encoder (8 features from 1-bit sign planes; the corner bytes are loaded one
//     chunk ahead), splits every value into a tf32 hi word and a bf16 (hi, lo) pair word and stores both into a
//     128B-swizzled, double-buffered smem operand slot; chunks 6-7 are the 63 sin/cos columns.  The same warps run the
//     layer epilogues (tcgen05.ld -> bias / ReLU -> split -> tcgen05.st) that turn an accumulator into the next layer's A
//     operand *inside TMEM*, and pre-gather chunks 0-1 of the CTA's next tile while layers 2 and 4 run.
//   * warp 16, the MMA warp (warp-uniform control flow, one elected lane issues): per k-step ONE kind::tf32 MMA
//     (Ahi*Bhi, K = 8) and ONE kind::f16 MMA on the bf16 pairs (Ahi*Blo + Alo*Bhi, K = 16) into different TMEM
//     accumulators -- fp32 parity from two half-rate MMAs instead of three.  Layer 1 reads A from smem, layers 2-4 from TMEM.
//   * warp 17, the weight-stream warp: pre-split, pre-swizzled weight chunks by cp.async.bulk (mbarrier complete_tx)
//     through a 2-stage smem ring (147 KB of shared memory in total, requested explicitly so that ~90 KB of the SM's
//     array stay L1 for the byte gathers).
//   * TMEM (all 512 columns) is recycled layer to layer: L1 main [0,160) / [160,320) (even / odd chunks), small [320,480);
//     h1 hi in place [0,160), pairs [160,320); L2 acc [320,400) + [400,480); head input hi [0,96), pairs [96,192);
//     L3 acc [192,352) + [352,512) -> h3 in place; L4 acc [0,160) (ep3 -> L4 wavefront per 32-column group);
//     Linear(160,3) runs as FFMA in the last epilogue.
//   * variants: DENSITY_ONLY (sigma pass of the sampler), SAVE (training forward: x0 / h1 / h3 / h4 / geo to HBM), POLL
//     (host-buffer pipeline, cnc_field_fwd_host); the sample count may live in device memory (cnc_field_fwd_n).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "tc05.cuh"

namespace cnc {
namespace ff {

#ifndef CNC_FF_NSTAGE
#define CNC_FF_NSTAGE 2
#endif
#ifndef CNC_FF_NA
#define CNC_FF_NA 2
#endif
constexpr int NSTAGE = CNC_FF_NSTAGE;   // weight ring depth
constexpr int NA = CNC_FF_NA;           // layer-1 operand slots (feature chunks in flight between gather warps and MMA warp)
constexpr uint32_t STAGE_BYTES = 2u * 160u * 128u;  // hi + lo of a [160 x 32] fp32 chunk (or two [80 x 32] chunks)
constexpr uint32_t A_HALF = 128u * 128u;            // one [128 x 32] fp32 chunk
constexpr uint32_t A_SLOT_BYTES = 2u * A_HALF;
constexpr uint32_t SMEM_B = 0;
constexpr uint32_t SMEM_A = NSTAGE * STAGE_BYTES;            // 163840
constexpr uint32_t SMEM_BAR = SMEM_A + NA * A_SLOT_BYTES;
constexpr uint32_t SMEM_LVL = SMEM_BAR + 256;                // 16 x LevelTab (32 B)
constexpr uint32_t SMEM_W5 = SMEM_LVL + 16 * 32;             // W5 [3][160] + b5 [3] (+pad) fp32
constexpr uint32_t SMEM_DYN = SMEM_W5 + 484 * 4;             // 231968 <= 232448

// weight stream of one tile (stage loads): L1 8 x [160x32], L2 [80x64] [80x64] [80x32], L3 3 x [160x32], L4 5 x [160x32]
constexpr int CPT_FULL = 19, CPT_DENSITY = 11;
constexpr uint32_t BLOB_W_FLOATS = 189440;
constexpr uint32_t BLOB_W5 = BLOB_W_FLOATS;  // raw fp32 W5 [3][160]: the last layer runs as FFMA in the L4 epilogue
constexpr uint32_t BIAS1 = BLOB_W5 + 480, BIAS2 = BIAS1 + 160, BIAS3 = BIAS2 + 80, BIAS4 = BIAS3 + 160,
                   BIAS5 = BIAS4 + 160;
constexpr uint32_t BLOB_FLOATS = BIAS5 + 16;  // 190496

__host__ __device__ __forceinline__ void chunk_meta(int g, uint32_t &ofs, uint32_t &bytes) {
    if (g < 8) { bytes = 40960; ofs = (uint32_t)g * 40960u; }                               // L1
    else if (g < 11) { bytes = g < 10 ? 40960u : 20480u; ofs = 327680u + (uint32_t)(g - 8) * 40960u; }  // L2
    else if (g < 14) { bytes = 40960; ofs = 430080u + (uint32_t)(g - 11) * 40960u; }        // L3
    else { bytes = 40960; ofs = 552960u + (uint32_t)(g - 14) * 40960u; }                    // L4
}

using namespace tc;

// ------------------------------------------------------------------------------------------
// weight packing: nn.Linear weights -> blob of pre-split, pre-swizzled K chunks + biases
// ------------------------------------------------------------------------------------------
struct PackArgs {
    const float *W[5];
    const float *b[5];
};

// head input K order: k=0 <- SH0, k=1..79 <- geo0..78, k=80..94 <- SH1..15, k=95 <- pad
__device__ __forceinline__ int head_src_col(int k) {
    if (k == 0) return 0;
    if (k < 80) return 15 + k;  // geo j = k-1 sits at column 16 + j of cat[SH16, geo]
    if (k < 95) return k - 79;  // SH 1..15
    return -1;
}

__global__ void __launch_bounds__(256) pack_weights_kernel(PackArgs a, float *__restrict__ blob) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BLOB_FLOATS) return;
    if (i >= BLOB_W_FLOATS) {
        const uint32_t j = i - BLOB_W_FLOATS;
        float v;
        if (j < 480) v = a.W[4][j];  // Linear(160,3) weight, [3][160] row-major, unsplit
        else if (j < 640) v = a.b[0][j - 480];
        else if (j < 720) v = a.b[1][j - 640];
        else if (j < 880) v = a.b[2][j - 720];
        else if (j < 1040) v = a.b[3][j - 880];
        else v = (j - 1040 < 3) ? a.b[4][j - 1040] : 0.f;
        blob[i] = v;
        return;
    }
    // which layer / chunk
    const uint32_t byte = i * 4u;
    int layer, kc, N;
    uint32_t in_chunk;
    if (byte < 327680u) { layer = 0; N = 160; kc = byte / 40960u; in_chunk = byte % 40960u; }
    else if (byte < 430080u) { layer = 1; N = 80; kc = (byte - 327680u) / 20480u; in_chunk = (byte - 327680u) % 20480u; }
    else if (byte < 552960u) { layer = 2; N = 160; kc = (byte - 430080u) / 40960u; in_chunk = (byte - 430080u) % 40960u; }
    else { layer = 3; N = 160; kc = (byte - 552960u) / 40960u; in_chunk = (byte - 552960u) % 40960u; }
    const uint32_t half_bytes = (uint32_t)N * 128u;
    const bool is_lo = in_chunk >= half_bytes;
    const uint32_t p = is_lo ? in_chunk - half_bytes : in_chunk;
    const uint32_t r = (p % 1024u) / 128u, n = (p / 1024u) * 8u + r;
    const uint32_t c16 = ((p % 128u) / 16u) ^ r;  // undo the 128B swizzle
    const int k = kc * 32 + (int)(c16 * 4u + (p % 16u) / 4u);
    float v = 0.f;
    switch (layer) {
        case 0: if (k < 255) v = a.W[0][n * 255 + k]; break;                       // Linear(255,160)
        case 1: v = a.W[1][n * 160 + k]; break;                                    // Linear(160,80)
        case 2: { const int c = head_src_col(k); if (c >= 0) v = a.W[2][n * 95 + c]; } break;  // Linear(95,160)
        default: v = a.W[3][n * 160 + k]; break;                                   // Linear(160,160)
    }
    const uint32_t hi = rna_tf32(v);
    const float h = __uint_as_float(hi);
    blob[i] = __uint_as_float(is_lo ? pack_bf16(__fsub_rn(v, h), h) : hi);   // B side pair (Blo, Bhi)
}

// ------------------------------------------------------------------------------------------
// fused forward
// ------------------------------------------------------------------------------------------
struct FieldArgs {
    const float *pos;   // [N,3] world
    const float *dirs;  // [N,3] unit view directions (nullptr -> density only)
    float aabb[6];
    const uint8_t *bits3, *bits_xy, *bits_xz, *bits_yz;
    const int32_t *offs3, *res3, *offs2, *res2;  // 12+1 / 12 and 4+1 / 4 entries
    const float *blob;
    float *sigma;  // [N]
    float *rgb;    // [N,3]   (nullptr in density-only mode)
    float *geo;    // [N,79]  nullable
    // training forward (all nullable, together): what the backward pass needs, written as the values go by
    // single-launch host pipeline (cnc_field_fwd_host, all nullable): wave w of tiles may be read once ready[w] == epoch
    // (written behind the H2D copy of its samples); every CTA bumps done[w] when its tile of the wave is in HBM
    const uint32_t *ready;
    uint32_t *done;
    uint32_t epoch;
    float *sv_x0;  // [N,256] layer-1 input (192 grid features | x, sin/cos 63 | 1)
    float *sv_h1;  // [N,160] relu(L1)
    float *sv_h3;  // [N,160] relu(L3)
    float *sv_h4;  // [N,160] relu(L4)
    uint32_t N;
    const uint32_t *n_dev;    // nullable: the sample count lives in device memory (<= N; wavefront renderer, no host sync)
    unsigned long long *dbg;  // optional timeline buffer (cnc_field_set_timeline_buffer), see DESIGN.md
};

// Per-level constants, built once per CTA in shared memory (12 levels of the 3D grid, then 4 of the planes).
struct LevelTab {
    uint32_t base_row, res, mask, hashed;  // hashed: 1 -> row = (c0 ^ c1*P1 ^ c2*P2) & mask, 0 -> dense c0 + c1*res + c2*res^2
    float scale, scale_re;                 // float(res-2), 1/float(res-2)
    uint32_t res2, T;
};

__device__ __forceinline__ void make_level_tab(const int32_t *__restrict__ offs, const int32_t *__restrict__ res, int level,
                                               int D, LevelTab &L) {
    const LevelConst lc = load_level(offs, res, (uint32_t)level);
    L.base_row = lc.base_row; L.res = lc.res; L.T = lc.T; L.scale = lc.scale; L.scale_re = lc.scale_re;
    L.res2 = lc.res * lc.res;
    // gridencoder.cu:63-87: the index is dense while the running stride stays <= T, hashed otherwise
    uint64_t stride = 1;
    for (int d = 0; d < D; d++) stride = stride <= lc.T ? stride * lc.res : (uint64_t)1 << 40;
    L.hashed = stride > lc.T ? 1u : 0u;
    L.mask = ((lc.T & (lc.T - 1u)) == 0u) ? lc.T - 1u : 0u;  // 0: T is not a power of two -> generic modulo
}

// One level of one encoder for one sample, split in two halves so that the table loads of chunk c+1 are
// in flight while chunk c is accumulated and stored.  Float arithmetic is the one of make_corners
// (common.cuh); the integer side factors the index into per-dimension terms, which is what makes the
// gather cheap enough to hide behind the MMAs (the generic kernel spends ~400 instructions per level on it).
struct Pend {
    float ww[8];      // renormalised corner weights w_i / sum(w); 0 for corners that do not contribute (+-0 adds are exact)
    uint32_t sb[8];   // sign bytes of the corner rows (bit k = feature k >= 0)
};

template <int D>
__device__ __forceinline__ void gather_issue(const float (&x)[D], const uint8_t *__restrict__ bits, const LevelTab &L,
                                             Pend &p) {
    constexpr uint32_t P1 = 2654435761u, P2 = 805459861u;
    bool inb = true;
    uint32_t t[D][2];   // per-dimension index term of the low / high cell
    float wd[D][2];     // per-dimension weight of the low / high cell
    bool z[D][2];       // low / high cell lies on the zero border (gridencoder.cu:212-219)
#pragma unroll
    for (int d = 0; d < D; d++) {
        inb = inb && !(x[d] < 0.f) && !(x[d] > 1.f);
        const float pos = __fmaf_rn(x[d], L.scale, 0.5f);
        const uint32_t g = (uint32_t)floorf(pos);
        const float f = __fsub_rn(pos, (float)g);
        const uint32_t c0 = g, c1 = min(g + 1u, L.res - 1u);
        wd[d][0] = __fsub_rn(1.f, f);
        wd[d][1] = f;
        z[d][0] = (c0 == 0u) || (c0 == L.res - 1u);
        z[d][1] = (c1 == 0u) || (c1 == L.res - 1u);
        const uint32_t m = d == 0 ? 1u : (L.hashed ? (d == 1 ? P1 : P2) : (d == 1 ? L.res : L.res2));
        t[d][0] = c0 * m;
        t[d][1] = c1 * m;
    }
    float w[1 << D];
    uint32_t valid = 0;
    float wn = 0.f;
#pragma unroll
    for (int i = 0; i < (1 << D); i++) {
        float wi = wd[0][i & 1];
        bool zi = z[0][i & 1];
        uint32_t row = t[0][i & 1];
#pragma unroll
        for (int d = 1; d < D; d++) {
            wi = __fmul_rn(wi, wd[d][(i >> d) & 1]);
            zi = zi || z[d][(i >> d) & 1];
            row = L.hashed ? (row ^ t[d][(i >> d) & 1]) : (row + t[d][(i >> d) & 1]);
        }
        if (L.hashed) row = L.mask ? (row & L.mask) : (row % L.T);
        w[i] = wi;
        const bool on = inb && !zi;
        if (on) {
            wn = __fadd_rn(wn, wi);
            valid |= 1u << i;
        }
        p.sb[i] = on ? (uint32_t)__ldg(bits + L.base_row + row) : 0u;
    }
    if (wn == 0.f) wn = 1e-9f;
    const float wn_re = __frcp_rn(wn);
#pragma unroll
    for (int i = 0; i < (1 << D); i++) p.ww[i] = ((valid >> i) & 1u) ? __fmul_rn(w[i], wn_re) : 0.f;
}

template <int NC>
__device__ __forceinline__ void gather_finish(const Pend &p, float (&f)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k++) f[k] = 0.f;
#pragma unroll
    for (int i = 0; i < NC; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = __fadd_rn(f[k], ((p.sb[i] >> k) & 1u) ? p.ww[i] : -p.ww[i]);
    }
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// one 256-bit global store (STG.E.256, sm_100): a full 32-byte sector per thread and instruction.  The training forward writes
// 3.3 KB per sample row by row (thread = row), i.e. every warp store touches 32 different lines; as pairs of 16-byte stores
// that was 7.5e7 half-sector requests per 380 k samples and the store path, not HBM, bounded the kernel (1.4 TB/s).
__device__ __forceinline__ void st_global_v8(float *p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// store 8 feature values (columns 8q..8q+7 of a 32-column chunk) of row r into the swizzled A slot
__device__ __forceinline__ void store_oct(uint8_t *slot, int r, int q, const float (&f)[8]) {
    uint8_t *rowp = slot + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        uint4 hi, lo;
        split_tf32(f[4 * j + 0], hi.x, lo.x);
        split_tf32(f[4 * j + 1], hi.y, lo.y);
        split_tf32(f[4 * j + 2], hi.z, lo.z);
        split_tf32(f[4 * j + 3], hi.w, lo.w);
        const int phys = ((2 * q + j) ^ (r & 7)) << 4;
        *reinterpret_cast<uint4 *>(rowp + phys) = hi;
        *reinterpret_cast<uint4 *>(rowp + A_HALF + phys) = lo;
    }
}

__device__ __forceinline__ void sh16_eval_h(float x, float y, float z, float (&o)[16]) {
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
#pragma unroll
    for (int k = 0; k < 16; k++) o[k] = __half2float(__float2half_rn(o[k]));  // tcnn emits fp16
}

// timeline probe: CTA 0, its second tile; slots 0..31 = compute thread 0, 32..63 = MMA lane
#define CNC_TL(slot) do { if (a.dbg != nullptr && blockIdx.x == 0 && it == 1) a.dbg[slot] = clock64(); } while (0)

constexpr uint32_t NONE = 0xFFFFu;

// 8 accumulator columns (block b8) of this thread's TMEM lane: main (+ main2) (+ small), summed in fp32 (RN)
template <uint32_t MAIN2, uint32_t SMALL>
__device__ __forceinline__ void acc_block(uint32_t tl, uint32_t c_main, int b8, float (&v)[8]) {
    uint32_t m[8], m2[8], sm[8];
    tmem_ld8(tl + c_main + 8 * b8, m);
    if (MAIN2 != NONE) tmem_ld8(tl + MAIN2 + 8 * b8, m2);
    if (SMALL != NONE) tmem_ld8(tl + SMALL + 8 * b8, sm);
    tc_wait_ld();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float x = __uint_as_float(m[k]);
        if (MAIN2 != NONE) x = __fadd_rn(x, __uint_as_float(m2[k]));
        if (SMALL != NONE) x = __fadd_rn(x, __uint_as_float(sm[k]));
        v[k] = x;
    }
}

// hidden-layer epilogue (160 columns = 20 blocks of 8, this thread's column group takes every 4th):
// (+bias, ReLU, split) -> hi at dst_hi, lo at dst_lo
template <uint32_t MAIN2, uint32_t SMALL>
__device__ __forceinline__ void ep_hidden(uint32_t tl, int cg, uint32_t c_main, uint32_t dst_hi, uint32_t dst_lo,
                                          const float *__restrict__ bias, float *__restrict__ save_row, uint32_t part_bar = 0u) {
#pragma unroll 1
    for (int b = cg; b < 20; b += 4) {
        float v[8];
        uint32_t hi[8], lo[8];
        acc_block<MAIN2, SMALL>(tl, c_main, b, v);
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + 8 * b)), b1 = __ldg(reinterpret_cast<const float4 *>(bias + 8 * b) + 1);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[k] = fmaxf(__fadd_rn(v[k], bb[k]), 0.f);
            split_tf32(v[k], hi[k], lo[k]);
        }
        if (save_row != nullptr) st_global_v8(save_row + 8 * b, v);
        tmem_st8(tl + dst_hi + 8 * b, hi);
        tmem_st8(tl + dst_lo + 8 * b, lo);
        if (part_bar != 0u && b < 16) {   // this 32-column group is complete once all four column groups have stored
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(part_bar + 8u * (uint32_t)(b >> 2));
        }
    }
    tc_wait_st();
}

constexpr int NCOMPUTE = 512;              // 16 gather/epilogue warps
constexpr int NTHREADS = NCOMPUTE + 64;    // + the MMA warp + the weight-stream (bulk copy) warp

template <bool DENSITY_ONLY, bool SAVE, bool POLL = false>
__global__ void __launch_bounds__(NTHREADS, 1) field_fwd_kernel(const FieldArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5;
    constexpr int CPT = DENSITY_ONLY ? CPT_DENSITY : CPT_FULL;
    constexpr int MMA_WARP = NCOMPUTE / 32;

    // barriers
    const uint32_t bar0 = sbase + SMEM_BAR;
    auto b_full = [&](uint32_t s) { return bar0 + 8u * s; };
    auto b_empty = [&](uint32_t s) { return bar0 + 8u * (NSTAGE + s); };
    auto a_full = [&](uint32_t s) { return bar0 + 8u * (2 * NSTAGE + s); };
    auto a_empty = [&](uint32_t s) { return bar0 + 8u * (2 * NSTAGE + NA + s); };
    const uint32_t layer_done = bar0 + 8u * (2 * NSTAGE + 2 * NA);
    const uint32_t act_ready = bar0 + 8u * (2 * NSTAGE + 2 * NA + 1);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SMEM_BAR + 8 * (2 * NSTAGE + 2 * NA + 2));
    // act_part(i): columns [32 i, 32 i + 32) of h3 are in TMEM (ep3 -> L4 wavefront), one completion per tile each
    auto act_part = [&](uint32_t i) { return bar0 + 8u * (2 * NSTAGE + 2 * NA + 3 + i); };
    static_assert(2 * NSTAGE + 2 * NA + 7 <= 32, "barrier block is 256 bytes");
    LevelTab *lvl = reinterpret_cast<LevelTab *>(smem + SMEM_LVL);
    float *w5s = reinterpret_cast<float *>(smem + SMEM_W5);

    if (threadIdx.x == 0) {
        if (sbase & 1023u) __trap();  // the 128B-swizzled operand tiles need a 1024-byte aligned base
        for (int s = 0; s < NSTAGE; s++) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int s = 0; s < NA; s++) { mbar_init(a_full(s), NCOMPUTE); mbar_init(a_empty(s), 1); }
        mbar_init(layer_done, 1);
        mbar_init(act_ready, NCOMPUTE);
        for (int i = 0; i < 4; i++) mbar_init(act_part(i), NCOMPUTE);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 16) {
        if (threadIdx.x < 12) make_level_tab(a.offs3, a.res3, threadIdx.x, 3, lvl[threadIdx.x]);
        else make_level_tab(a.offs2, a.res2, threadIdx.x - 12, 2, lvl[threadIdx.x]);
    }
    for (int i = threadIdx.x; i < 484; i += NTHREADS) w5s[i] = i < 480 ? __ldg(a.blob + BLOB_W5 + i) : (i < 483 ? __ldg(a.blob + BIAS5 + (i - 480)) : 0.f);
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // This CTA owns the whole SM (227 KB of shared memory) and asks for all 512 columns, so the allocation
    // starts at lane 0 / column 0.  Using the constant keeps every MMA operand in uniform registers.
    if (*tmem_slot != 0u) __trap();
    constexpr uint32_t tbase = 0u;

    const uint32_t Nrt = a.n_dev ? min(__ldg(a.n_dev), a.N) : a.N;
    const uint32_t ntiles = (Nrt + TILE_M - 1) / TILE_M;
    const uint32_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == MMA_WARP) {
        // =============================== weight stream + MMA issue ===============================
        // The whole warp runs this control flow (all values warp-uniform -> uniform datapath); one elected
        // lane issues the copies, MMAs and commits.
        uint32_t Gm = 0, a_cons = 0, act_cnt = 0;

        // One stage load = NATOM K-atoms (32 columns = 4 k-steps of 8) of a layer with N output columns.
        // The two small products (Ahi*Blo, Alo*Bhi) go to their own accumulator d_small when the layer has
        // one, so the round-toward-zero of the tensor-core accumulate acts on them at their own 2^-11 scale.
        // bar1 / bar2: extra mbarriers (0 = none) that the completion of these MMAs arrives on.
        auto chunk_mma = [&](auto n_tag, auto smem_tag, int natom, uint32_t a_hi, uint32_t a_lo, uint32_t d_main,
                             bool first_main, uint32_t d_small, bool first_small, uint32_t bar1, uint32_t bar2,
                             uint32_t it = 0, int tl_slot = -1) {
            constexpr int N = decltype(n_tag)::value;
            constexpr bool A_IN_SMEM = decltype(smem_tag)::value;
            const uint32_t s = Gm % NSTAGE;
            mbar_wait_spin(b_full(s), (Gm / NSTAGE) & 1u);
            tc_fence_after();
            if (tl_slot >= 0 && elect_one()) CNC_TL(tl_slot);
            const uint32_t bst = sbase + SMEM_B + s * STAGE_BYTES;
            constexpr uint32_t id = idesc_tf32<N>(), idb = idesc_bf16<N>();
            const bool shared_acc = (d_small == d_main);
            for (int at = 0; at < natom; at++) {
                const uint64_t dbh0 = smem_desc(bst + (uint32_t)at * (2u * N * 128u)),
                               dbl0 = smem_desc(bst + (uint32_t)at * (2u * N * 128u) + (uint32_t)N * 128u);
                const uint64_t dah0 = A_IN_SMEM ? smem_desc(a_hi) : 0ull, dal0 = A_IN_SMEM ? smem_desc(a_lo) : 0ull;
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) {
                    // +32 bytes along K inside the 128-byte swizzle row == +2 in the descriptor's address field.
                    // The two MMAs of a k-step go to different accumulators, so their accumulate latencies overlap.
                    const uint64_t dbh = dbh0 + (uint64_t)(2 * k4), dbl = dbl0 + (uint64_t)(2 * k4);
                    const bool head = (at == 0 && k4 == 0);
                    const uint32_t acc_s = (first_small && head) ? 0u : 1u;
                    const uint32_t acc_m = shared_acc ? 1u : ((first_main && head) ? 0u : 1u);
                    if (A_IN_SMEM) {
                        const uint64_t dah = dah0 + (uint64_t)(2 * k4), dal = dal0 + (uint64_t)(2 * k4);
                        mma_ss_bf16(tbase + d_small, dal, dbl, idb, acc_s);   // Ahi*Blo + Alo*Bhi
                        mma_ss(tbase + d_main, dah, dbh, id, acc_m);          // Ahi*Bhi
                    } else {
                        const uint32_t ah = tbase + a_hi + 32u * at + 8u * k4, al = tbase + a_lo + 32u * at + 8u * k4;
                        mma_ts_bf16(tbase + d_small, al, dbl, idb, acc_s);
                        mma_ts(tbase + d_main, ah, dbh, id, acc_m);
                    }
                }
            }
            tc_commit_elect(b_empty(s));
            if (bar1) tc_commit_elect(bar1);
            if (bar2) tc_commit_elect(bar2);
            Gm++;
        };
        using N160 = std::integral_constant<int, 160>;
        using N80 = std::integral_constant<int, 80>;
        using FromSmem = std::true_type;
        using FromTmem = std::false_type;

        for (uint32_t it = 0; it < my_tiles; it++) {
            // L1: A from smem slots; main [0,160) (even chunks) / [160,320) (odd chunks: halves the round-toward-zero
            // bias of the 32-step accumulation), small [320,480)
#pragma unroll 1
            for (int kc = 0; kc < 8; kc++) {
                const uint32_t sl = a_cons % NA;
                mbar_wait_spin(a_full(sl), (a_cons / NA) & 1u);
                const uint32_t ab = sbase + SMEM_A + sl * A_SLOT_BYTES;
                if (elect_one()) CNC_TL(32 + kc);
                chunk_mma(N160{}, FromSmem{}, 1, ab, ab + A_HALF, (kc & 1) ? 160u : 0u, kc < 2, 320u, kc == 0, a_empty(sl),
                          kc == 7 ? layer_done : 0u);
                a_cons++;
            }
            if (elect_one()) CNC_TL(40);
            // L2: A = h1 hi [0,160) lo [160,320) -> main [320,400), small [400,480); K = 64 + 64 + 32
            mbar_wait_spin(act_ready, act_cnt & 1u); act_cnt++;
            tc_fence_after();
            if (elect_one()) CNC_TL(41);
#pragma unroll 1
            for (int j = 0; j < 3; j++)
                chunk_mma(N80{}, FromTmem{}, j < 2 ? 2 : 1, 64u * j, 160u + 64u * j, 320u, j == 0, 400u, j == 0,
                          j == 2 ? layer_done : 0u, 0u, it, 48 + j);
            if (elect_one()) CNC_TL(42);
            if (!DENSITY_ONLY) {
                // L3: A = head input hi [0,96) lo [96,192) -> main [192,352), small [352,512)
                mbar_wait_spin(act_ready, act_cnt & 1u); act_cnt++;
                tc_fence_after();
                if (elect_one()) CNC_TL(43);
#pragma unroll 1
                for (int kc = 0; kc < 3; kc++)
                    chunk_mma(N160{}, FromTmem{}, 1, 32u * kc, 96u + 32u * kc, 192u, kc == 0, 352u, kc == 0,
                              kc == 2 ? layer_done : 0u, 0u);
                // L4: A hi [192,352) lo [352,512) -> [0,160) (no room for a second accumulator).  Wavefront behind ep3:
                // K chunk kc needs only columns [32 kc, 32 kc + 32) of h3, and [0,160) is dead once L3 has completed.
#pragma unroll 1
                for (int kc = 0; kc < 5; kc++) {
                    if (kc < 4) mbar_wait_spin(act_part(kc), it & 1u);
                    else { mbar_wait_spin(act_ready, act_cnt & 1u); act_cnt++; }
                    tc_fence_after();
                    if (kc == 0 && elect_one()) CNC_TL(44);
                    chunk_mma(N160{}, FromTmem{}, 1, 192u + 32u * kc, 352u + 32u * kc, 0u, false, 0u, kc == 0,
                              kc == 4 ? layer_done : 0u, 0u, it, 53 + kc);
                }
                if (elect_one()) CNC_TL(45);
            }
            // The next tile's first feature chunks are usually already waiting in smem, and its L1 overwrites
            // the accumulators the last epilogue of this tile is still reading: wait until it has read them.
            mbar_wait_spin(act_ready, act_cnt & 1u); act_cnt++;
            tc_fence_after();
            if (elect_one()) CNC_TL(46);
        }
    } else if (warp == MMA_WARP + 1) {
        // =============================== weight stream ===============================
        // Keeps the 4-stage ring full: a stage is refilled as soon as the MMAs that read it have retired
        // (tcgen05.commit -> b_empty), independently of where the MMA warp is -> up to 3 loads (120 KB) in flight.
        const uint8_t *blob = reinterpret_cast<const uint8_t *>(a.blob);
        const uint32_t total = my_tiles * CPT;
        for (uint32_t G = 0; G < total; G++) {
            const uint32_t s = G % NSTAGE, use = G / NSTAGE;
            if (use > 0) mbar_wait(b_empty(s), (use - 1) & 1u);
            uint32_t ofs, bytes;
            chunk_meta((int)(G % CPT), ofs, bytes);
            if (a.dbg != nullptr && blockIdx.x == 0 && G / CPT == 1 && elect_one()) a.dbg[64 + G % CPT] = clock64();
            bulk_g2s_elect(sbase + SMEM_B + s * STAGE_BYTES, blob + ofs, bytes, b_full(s));
        }
    } else {
        // =============================== gather / epilogue warps ===============================
        const int r = threadIdx.x & 127;   // row in tile == TMEM lane; (warp & 3) is the lane quarter this warp may touch
        const int q = threadIdx.x >> 7;    // feature stage: which 8 of a chunk's 32 columns; epilogues: column group
        const uint32_t tl = tbase + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t a_prod = 0, done_cnt = 0;
        const float *bias = a.blob;
        const float3 amin = make_float3(a.aabb[0], a.aabb[1], a.aabb[2]);
        const float3 ainv = make_float3(__fsub_rn(a.aabb[3], a.aabb[0]), __fsub_rn(a.aabb[4], a.aabb[1]), __fsub_rn(a.aabb[5], a.aabb[2]));

        // normalised position of this thread's row in tile `tile` (ngp.py:517-518)
        auto load_pos = [&](uint32_t tile, float (&p)[3]) {
            const uint32_t row = tile * TILE_M + r;
            const bool live = row < Nrt;
            if (POLL) {   // the samples of this wave are being uploaded while earlier waves are evaluated
                const uint32_t *flag = a.ready + tile / gridDim.x;
                uint32_t v, spins = 0;
                do {   // acquire: the positions written before the stamp are visible after it; bounded (~2 s): never hang the GPU
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                    if ((int32_t)(v - a.epoch) >= 0) break;
                    __nanosleep(64);
                } while (++spins < (1u << 23));
            }
            // L2 loads (each value is read once; with POLL the buffer is rewritten by the copy engine between launches)
            p[0] = live ? __ldcg(a.pos + (size_t)row * 3 + 0) : __fadd_rn(amin.x, 0.5f * ainv.x);
            p[1] = live ? __ldcg(a.pos + (size_t)row * 3 + 1) : __fadd_rn(amin.y, 0.5f * ainv.y);
            p[2] = live ? __ldcg(a.pos + (size_t)row * 3 + 2) : __fadd_rn(amin.z, 0.5f * ainv.z);
        };
        auto normalise = [&](const float (&p)[3], float (&x)[3]) {
            x[0] = __fdiv_rn(__fsub_rn(p[0], amin.x), ainv.x);
            x[1] = __fdiv_rn(__fsub_rn(p[1], amin.y), ainv.y);
            x[2] = __fdiv_rn(__fsub_rn(p[2], amin.z), ainv.z);
        };
        auto load_x = [&](uint32_t tile, float (&x)[3]) {
            float p[3];
            load_pos(tile, p);
            normalise(p, x);
        };
        // chunk c of the layer-1 input: [xyz 96 | xy 32 | xz 32 | yz 32 | x, sin/cos 63 | pad]; this thread: columns 8q..8q+7
        auto issue = [&](int c, const float (&x)[3], Pend &p) {   // table loads of chunk c (c < 6)
            if (c < 3) {
                gather_issue<3>(x, a.bits3, lvl[4 * c + q], p);
            } else {
                const float x2[2] = {c == 5 ? x[1] : x[0], c == 3 ? x[1] : x[2]};
                const uint8_t *bt = c == 3 ? a.bits_xy : (c == 4 ? a.bits_xz : a.bits_yz);
                gather_issue<2>(x2, bt, lvl[12 + q], p);
            }
        };
        auto publish = [&](int c, uint32_t prow, const float (&f)[8]) {  // 8 feature values -> smem slot -> MMA warp
            if (SAVE && prow < Nrt) {   // column 255 (K padding, weight 0) is saved as 1: its weight-gradient row is the bias gradient
                float *d = a.sv_x0 + (size_t)prow * 256 + 32 * c + 8 * q;
                const float fs[8] = {f[0], f[1], f[2], f[3], f[4], f[5], f[6], (c == 7 && q == 3) ? 1.0f : f[7]};
                st_global_v8(d, fs);
            }
            const uint32_t sl = a_prod % NA, use = a_prod / NA;
            if (use > 0) mbar_wait(a_empty(sl), (use - 1) & 1u);
            store_oct(smem + SMEM_A + sl * A_SLOT_BYTES, r, q, f);
            fence_async_smem();
            mbar_arrive(a_full(sl));
            a_prod++;
        };
        auto finish = [&](int c, uint32_t prow, const Pend &p) {  // gathered chunk c (c < 6) of global row prow
            float f[8];
            if (c < 3) gather_finish<8>(p, f);
            else gather_finish<4>(p, f);
            publish(c, prow, f);
        };
        auto embed = [&](int c, uint32_t prow, const float (&x)[3]) {  // chunks 6, 7
            // embed column j (0..62): j<3 -> x[j]; else g=(j-3)/3, d=(j-3)%3: even g -> sin(2^(g/2) x_d), odd -> cos.
            // The cosine of an angle sits three columns after its sine: where both fall into this thread's octet one
            // sincosf (one range reduction) serves both.
            float f[8];
            bool done[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) { f[jj] = 0.f; done[jj] = false; }
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const int j = (c - 6) * 32 + 8 * q + jj;
                if (j < 3) f[jj] = x[j];
                else if (j < 63 && !done[jj]) {
                    const int g = (j - 3) / 3, d = (j - 3) - 3 * g;
                    const float xd = d == 0 ? x[0] : (d == 1 ? x[1] : x[2]);
                    const float ang = __fmul_rn(xd, (float)(1 << (g >> 1)));
                    if (g & 1) f[jj] = cosf(ang);
                    else if (jj + 3 < 8) {
                        float sv, cv;
                        sincosf(ang, &sv, &cv);
                        f[jj] = sv;
                        f[jj + 3 < 8 ? jj + 3 : 7] = cv;
                        done[jj + 3 < 8 ? jj + 3 : 7] = true;
                    } else f[jj] = sinf(ang);
                }
            }
            publish(c, prow, f);
        };

        float x[3], xn[3];
        Pend pa, pb;   // gathers in flight: even chunks use pa, odd chunks pb
        int c0 = 0;  // chunks of the current tile that were already produced during the previous tile's layer chain
        // chunks of the NEXT tile produced while this tile's L2 / L4 (density-only: the tile's tail) run on the tensor
        // pipe.  All NA operand slots are free by then (L1 is complete), and no more than NA chunks may be produced
        // before this tile's last epilogue has arrived on act_ready (the MMA warp frees slots only after that).
        constexpr int PRE2 = NA >= 4 ? 2 : 1, PRE4 = NA >= 3 ? 2 : 1;
        static_assert(PRE2 + PRE4 <= NA, "pre-gathered chunks must fit the operand slots");
        auto pre = [&](auto ctag, uint32_t prow) {   // finish chunk c of the next tile (gather already in flight), start c + 1
            constexpr int c = decltype(ctag)::value;
            if (c + 1 < 6) issue(c + 1, xn, (c & 1) ? pa : pb);
            finish(c, prow, (c & 1) ? pb : pa);
        };
        using C0 = std::integral_constant<int, 0>;
        using C1 = std::integral_constant<int, 1>;
        using CA = std::integral_constant<int, PRE2>;
        using CB = std::integral_constant<int, PRE2 + 1>;
        if (my_tiles > 0) {
            load_x(blockIdx.x, x);
            issue(0, x, pa);
        }
        for (uint32_t it = 0; it < my_tiles; it++) {
            const uint32_t tile = blockIdx.x + it * gridDim.x;
            const uint32_t row = tile * TILE_M + r;
            const bool live = row < Nrt;
            const bool sel = (x[0] > 0.f) && (x[0] < 1.f) && (x[1] > 0.f) && (x[1] < 1.f) && (x[2] > 0.f) && (x[2] < 1.f);
            const bool has_next = it + 1 < my_tiles;
            if (threadIdx.x == 0) CNC_TL(0);
            // ---- L1 feature chunks c0..7; pa holds the gather of chunk c
            if (c0 & 1) {   // pb holds chunk c0
                if (c0 + 1 < 6) issue(c0 + 1, x, pa);
                finish(c0, row, pb);
                c0++;
            }
#pragma unroll 1
            for (int c = c0; c < 6; c += 2) {
                issue(c + 1, x, pb);
                finish(c, row, pa);
                if (threadIdx.x == 0) CNC_TL(1 + c);
                if (c + 2 < 6) issue(c + 2, x, pa);
                finish(c + 1, row, pb);
                if (threadIdx.x == 0) CNC_TL(2 + c);
            }
            if (has_next) load_pos(tile + gridDim.x, xn);   // raw positions of the next tile: in flight behind the sin/cos chunks
#pragma unroll 1
            for (int c = 6; c < 8; c++) {
                embed(c, row, x);
                if (threadIdx.x == 0) CNC_TL(1 + c);
            }
            c0 = 0;
            if (has_next) {
                normalise(xn, xn);
                issue(0, xn, pa);  // in flight during the L1 tail and ep1
            }
            // ---- ep1: h1 = relu(acc1 + b1) -> hi in place [0,160), (hi, lo) bf16 pairs over the second accumulator [160,320)
            if (threadIdx.x == 0) CNC_TL(9);
            mbar_wait(layer_done, done_cnt & 1u); done_cnt++;
            tc_fence_after();
            if (threadIdx.x == 0) CNC_TL(10);
            ep_hidden<160u, 320u>(tl, q, 0u, 0u, 160u, bias + BIAS1, (SAVE && live) ? a.sv_h1 + (size_t)row * 160 : nullptr);
            tc_fence_before();
            mbar_arrive(act_ready);
            if (threadIdx.x == 0) CNC_TL(11);
            float dv[3] = {0.f, 0.f, 0.f};   // view direction of the row (column groups 2, 3 build the SH block in ep2)
            if (!DENSITY_ONLY && q >= 2 && live) {
#pragma unroll
                for (int d = 0; d < 3; d++) dv[d] = __ldcg(a.dirs + (size_t)row * 3 + d);
            }
            if (has_next) {  // while L2 runs: the first chunk(s) of the next tile
                pre(C0{}, row + gridDim.x * TILE_M);
                if (PRE2 >= 2) pre(C1{}, row + gridDim.x * TILE_M);
            }
            // ---- ep2: acc2 [320,400)+[400,480): col 0 -> sigma, cols 1..79 -> geo -> head input (+ SH16 block)
            mbar_wait(layer_done, done_cnt & 1u); done_cnt++;
            tc_fence_after();
            if (threadIdx.x == 0) CNC_TL(12);
#pragma unroll 1
            for (int b = q; b < 12; b += 4) {
                uint32_t hi[8], lo[8];
                if (b < 10) {
                    float v[8], hs[8];
                    acc_block<NONE, 400u>(tl, 320u, b, v);
                    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + BIAS2 + 8 * b)),
                                 b1 = __ldg(reinterpret_cast<const float4 *>(bias + BIAS2 + 8 * b) + 1);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        float h = __fadd_rn(v[k], bb[k]);
                        if (k == 0 && b == 0) {
                            // density = trunc_exp(h0 - 1) * selector   (ngp.py:528-532, :318-334)
                            const float sg = __fmul_rn(expf(__fsub_rn(h, 1.0f)), sel ? 1.f : 0.f);
                            if (live) a.sigma[row] = sg;
                            h = 0.28198242f;  // SH band 0 (0.28209479 rounded to fp16) takes head-input column 0
                        } else if (!SAVE && a.geo != nullptr && live) {
                            a.geo[(size_t)row * 79 + (8 * b + k - 1)] = h;
                        }
                        hs[k] = h;
                        split_tf32(h, hi[k], lo[k]);
                    }
                    // training forward: the head input goes out as rows of 96 in the kernel's own column order (SH0 | geo 0..78 |
                    // SH1..15 | 0, see head_src_col) -- what the weight gradient of the first head layer contracts with -- one
                    // 32-byte store per block of 8 columns
                    if (SAVE && a.geo != nullptr && live) st_global_v8(a.geo + (size_t)row * 96 + 8 * b, hs);
                    if (!DENSITY_ONLY) {
                        tmem_st8(tl + 0u + 8 * b, hi);
                        tmem_st8(tl + 96u + 8 * b, lo);
                    }
                } else if (!DENSITY_ONLY) {
                    // head-input columns 80..87 (b == 10) / 88..95 (b == 11): SH bands 1..8 / 9..15, pad
                    float d3[3], sh[16];
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const float d01 = __fdiv_rn(__fadd_rn(dv[d], 1.0f), 2.0f);  // ngp.py:540
                        d3[d] = __fsub_rn(__fmul_rn(d01, 2.f), 1.f);             // tcnn maps back to [-1,1]
                    }
                    sh16_eval_h(d3[0], d3[1], d3[2], sh);
                    float vs[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float v = (b == 10) ? sh[1 + k] : (k < 7 ? sh[9 + k] : 0.f);
                        vs[k] = v;
                        split_tf32(v, hi[k], lo[k]);
                    }
                    if (SAVE && a.geo != nullptr && live) st_global_v8(a.geo + (size_t)row * 96 + 8 * b, vs);
                    tmem_st8(tl + 8u * b, hi);
                    tmem_st8(tl + 96u + 8u * b, lo);
                }
            }
            if (!DENSITY_ONLY) {
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(act_ready);
                if (threadIdx.x == 0) CNC_TL(13);
                // ---- ep3: [192,352)+[352,512) -> hi in place, lo over the small accumulator
                mbar_wait(layer_done, done_cnt & 1u); done_cnt++;
                tc_fence_after();
                if (threadIdx.x == 0) CNC_TL(14);
                ep_hidden<NONE, 352u>(tl, q, 192u, 192u, 352u, bias + BIAS3, (SAVE && live) ? a.sv_h3 + (size_t)row * 160 : nullptr,
                                      act_part(0));
                tc_fence_before();
                mbar_arrive(act_ready);
                if (threadIdx.x == 0) CNC_TL(15);
                if (has_next) {  // while L4 (the longest of the small layers) runs: the next chunk(s) of the next tile
                    pre(CA{}, row + gridDim.x * TILE_M);
                    if (PRE4 >= 2) pre(CB{}, row + gridDim.x * TILE_M);
                    c0 = PRE2 + PRE4;
                }
                // ---- ep4 + the last layer: h4 = relu(acc4 + b4) stays in registers; Linear(160,3) is 3 x 40 FFMA per
                // thread on its own columns (W5 broadcast from smem), the four column groups of a row meet in TMEM
                mbar_wait(layer_done, done_cnt & 1u); done_cnt++;
                tc_fence_after();
                if (threadIdx.x == 0) CNC_TL(16);
                float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll 1
                for (int b = q; b < 20; b += 4) {
                    float v[8];
                    acc_block<NONE, NONE>(tl, 0u, b, v);
                    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + BIAS4 + 8 * b)),
                                 b1 = __ldg(reinterpret_cast<const float4 *>(bias + BIAS4 + 8 * b) + 1);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    const float4 *w0 = reinterpret_cast<const float4 *>(w5s + 8 * b), *w1 = reinterpret_cast<const float4 *>(w5s + 160 + 8 * b),
                                 *w2 = reinterpret_cast<const float4 *>(w5s + 320 + 8 * b);
                    const float4 wa0 = w0[0], wb0 = w0[1], wa1 = w1[0], wb1 = w1[1], wa2 = w2[0], wb2 = w2[1];
                    const float W0[8] = {wa0.x, wa0.y, wa0.z, wa0.w, wb0.x, wb0.y, wb0.z, wb0.w};
                    const float W1[8] = {wa1.x, wa1.y, wa1.z, wa1.w, wb1.x, wb1.y, wb1.z, wb1.w};
                    const float W2[8] = {wa2.x, wa2.y, wa2.z, wa2.w, wb2.x, wb2.y, wb2.z, wb2.w};
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float h = fmaxf(__fadd_rn(v[k], bb[k]), 0.f);
                        v[k] = h;
                        p0 = __fmaf_rn(h, W0[k], p0);
                        p1 = __fmaf_rn(h, W1[k], p1);
                        p2 = __fmaf_rn(h, W2[k], p2);
                    }
                    if (SAVE && live) st_global_v8(a.sv_h4 + (size_t)row * 160 + 8 * b, v);
                }
                tmem_st4(tl + 480u + 4u * q, __float_as_uint(p0), __float_as_uint(p1), __float_as_uint(p2), 0u);
                tc_wait_st();
                tc_fence_before();
                asm volatile("bar.sync 1, %0;" ::"n"(NCOMPUTE) : "memory");
                tc_fence_after();
                if (q == 0) {
                    uint32_t e[16];
                    tmem_ld16(tl + 480u, e);
                    tc_wait_ld();
                    if (live) {
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            float h = __fadd_rn(__uint_as_float(e[k]), __uint_as_float(e[4 + k]));
                            h = __fadd_rn(h, __uint_as_float(e[8 + k]));
                            h = __fadd_rn(h, __uint_as_float(e[12 + k]));
                            h = __fadd_rn(h, w5s[480 + k]);
                            a.rgb[(size_t)row * 3 + k] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-h)));  // torch.sigmoid
                        }
                    }
                }
            }
            if (threadIdx.x == 0) CNC_TL(19);
            if (POLL) {   // this tile's sigma / rgb are stored: publish for the download stream
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(NCOMPUTE) : "memory");
                if (threadIdx.x == 0) atomicAdd(a.done + it, 1u);
            }
            tc_fence_before();
            mbar_arrive(act_ready);  // this tile's accumulators are consumed: the next tile's L1 may overwrite them
            if (has_next) {
                if (DENSITY_ONLY) { pre(CA{}, row + gridDim.x * TILE_M); c0 = PRE2 + 1; }
                x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
    }
}

}  // namespace ff
}  // namespace cnc

using namespace cnc;

static unsigned long long *g_timeline = nullptr;

extern "C" {

/* debug facility: device buffer of 64 uint64 that CTA 0 fills with clock64() stamps of its second tile */
int cnc_field_set_timeline_buffer(uint64_t *device_buf) { g_timeline = reinterpret_cast<unsigned long long *>(device_buf); return CNC_OK; }

uint32_t cnc_field_blob_floats(void) { return ff::BLOB_FLOATS; }

int cnc_field_pack_weights(const float *W1, const float *b1, const float *W2, const float *b2, const float *W3,
                           const float *b3, const float *W4, const float *b4, const float *W5, const float *b5,
                           float *blob, cnc_stream_t stream) {
    if (!W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !W4 || !b4 || !W5 || !b5 || !blob) {
        set_error("field_pack_weights: null pointer");
        return CNC_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(blob) & 15u) { set_error("field_pack_weights: blob must be 16-byte aligned"); return CNC_EINVAL; }
    ff::PackArgs a{{W1, W2, W3, W4, W5}, {b1, b2, b3, b4, b5}};
    ff::pack_weights_kernel<<<div_up(ff::BLOB_FLOATS, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, blob);
    return check_launch("field_pack_weights");
}

static int field_fwd_impl(const float *pos, const float *dirs, const float *aabb6_host, const uint8_t *bits_xyz,
                          const uint8_t *bits_xy, const uint8_t *bits_xz, const uint8_t *bits_yz, const int32_t *offsets3,
                          const int32_t *resolutions3, const int32_t *offsets2, const int32_t *resolutions2,
                          const float *blob, float *sigma, float *rgb, float *geo, float *sv_x0, float *sv_h1, float *sv_h3,
                          float *sv_h4, uint32_t N, cnc_stream_t stream, const uint32_t *ready = nullptr, uint32_t *done = nullptr,
                          uint32_t epoch = 0, const uint32_t *n_dev = nullptr) {
    if (N == 0) return CNC_OK;
    if (!pos || !aabb6_host || !bits_xyz || !bits_xy || !bits_xz || !bits_yz || !offsets3 || !resolutions3 ||
        !offsets2 || !resolutions2 || !blob || !sigma || (dirs && !rgb)) {
        set_error("field_fwd: null pointer");
        return CNC_EINVAL;
    }
    if (reinterpret_cast<uintptr_t>(blob) & 15u) { set_error("field_fwd: blob must be 16-byte aligned"); return CNC_EINVAL; }
    static int n_sm = 0;
    static bool attr_set = false;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!attr_set) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e1 = cudaFuncSetAttribute(ff::field_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff::SMEM_DYN);
        cudaError_t e2 = cudaFuncSetAttribute(ff::field_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff::SMEM_DYN);
        cudaError_t e3 = cudaFuncSetAttribute(ff::field_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff::SMEM_DYN);
        if (e1 == cudaSuccess) e1 = e3;
        cudaError_t e4 = cudaFuncSetAttribute(ff::field_fwd_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff::SMEM_DYN);
        if (e1 == cudaSuccess) e1 = e4;
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("field_fwd: cannot reserve %u bytes of shared memory (%s)", ff::SMEM_DYN,
                      cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            return CNC_ECUDA;
        }
        // The gathers live on L1: ask for the smallest shared-memory carve-out that holds the CTA, so that the rest of the
        // 256 KB array serves as cache (measured: 0.405 ms with the 228 KB carve-out, 0.34 ms with <= 196 KB).
        // CNC_FF_CARVEOUT=<percent> overrides (profiling aid).
        int carve = (int)((ff::SMEM_DYN + 1024u) * 100u / (228u * 1024u)) + 1;
        if (const char *e = getenv("CNC_FF_CARVEOUT")) carve = atoi(e);
        cudaFuncSetAttribute(ff::field_fwd_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(ff::field_fwd_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(ff::field_fwd_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(ff::field_fwd_kernel<false, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        attr_set = true;
    }
    ff::FieldArgs a;
    a.pos = pos; a.dirs = dirs;
    for (int i = 0; i < 6; i++) a.aabb[i] = aabb6_host[i];
    a.bits3 = bits_xyz; a.bits_xy = bits_xy; a.bits_xz = bits_xz; a.bits_yz = bits_yz;
    a.offs3 = offsets3; a.res3 = resolutions3; a.offs2 = offsets2; a.res2 = resolutions2;
    a.blob = blob; a.sigma = sigma; a.rgb = rgb; a.geo = geo; a.N = N; a.dbg = g_timeline;
    a.sv_x0 = sv_x0; a.sv_h1 = sv_h1; a.sv_h3 = sv_h3; a.sv_h4 = sv_h4;
    a.ready = ready; a.done = done; a.epoch = epoch; a.n_dev = n_dev;
    const uint32_t ntiles = (N + ff::TILE_M - 1) / ff::TILE_M;
    uint32_t grid = ntiles < (uint32_t)n_sm ? ntiles : (uint32_t)n_sm;
    if (const char *g = getenv("CNC_FIELD_GRID")) { const uint32_t v = (uint32_t)atoi(g); if (v >= 1 && v < grid) grid = v; }  // profiling aid
    if (ready) ff::field_fwd_kernel<false, false, true><<<grid, ff::NTHREADS, ff::SMEM_DYN, s>>>(a);
    else if (sv_x0) ff::field_fwd_kernel<false, true><<<grid, ff::NTHREADS, ff::SMEM_DYN, s>>>(a);
    else if (dirs) ff::field_fwd_kernel<false, false><<<grid, ff::NTHREADS, ff::SMEM_DYN, s>>>(a);
    else ff::field_fwd_kernel<true, false><<<grid, ff::NTHREADS, ff::SMEM_DYN, s>>>(a);
    return check_launch("field_fwd");
}

int cnc_field_fwd(const float *pos, const float *dirs, const float *aabb6_host, const uint8_t *bits_xyz,
                  const uint8_t *bits_xy, const uint8_t *bits_xz, const uint8_t *bits_yz, const int32_t *offsets3,
                  const int32_t *resolutions3, const int32_t *offsets2, const int32_t *resolutions2,
                  const float *blob, float *sigma, float *rgb, float *geo, uint32_t N, cnc_stream_t stream) {
    return field_fwd_impl(pos, dirs, aabb6_host, bits_xyz, bits_xy, bits_xz, bits_yz, offsets3, resolutions3, offsets2,
                          resolutions2, blob, sigma, rgb, geo, nullptr, nullptr, nullptr, nullptr, N, stream);
}

/* The same forward with the sample count read from device memory at kernel start (*n_dev, clamped to n_max): the caller
 * of a wavefront loop whose per-round sample count is produced on the device launches it without knowing the count. */
int cnc_field_fwd_n(const float *pos, const float *dirs, const float *aabb6_host, const uint8_t *bits_xyz,
                    const uint8_t *bits_xy, const uint8_t *bits_xz, const uint8_t *bits_yz, const int32_t *offsets3,
                    const int32_t *resolutions3, const int32_t *offsets2, const int32_t *resolutions2,
                    const float *blob, float *sigma, float *rgb, const uint32_t *n_dev, uint32_t n_max, cnc_stream_t stream) {
    if (!n_dev) { set_error("field_fwd_n: null count pointer"); return CNC_EINVAL; }
    return field_fwd_impl(pos, dirs, aabb6_host, bits_xyz, bits_xy, bits_xz, bits_yz, offsets3, resolutions3, offsets2,
                          resolutions2, blob, sigma, rgb, nullptr, nullptr, nullptr, nullptr, nullptr, n_max, stream, nullptr,
                          nullptr, 0, n_dev);
}

/* Host-buffer entry point, ONE launch of the persistent kernel: the samples are uploaded in chunks of whole waves
 * (1, 2, 4, .. max .. 4, 2, 1) on s_in, each followed by a 4-byte-per-wave "ready" stamp; the kernel, started at once on
 * s_compute, waits for the stamp of a wave before it reads the wave's positions; CTAs count finished tiles per wave and
 * a one-thread kernel per chunk on s_out waits for those counts before the chunk's download is queued behind it.  Only
 * the first upload and the last download are exposed, and the kernel keeps its inter-tile overlap (six chunk launches
 * cost 0.08 ms of ramp-up and tails at 262 144 samples). */
int cnc_field_fwd_host(const float *pos_host, const float *dirs_host, const float *aabb6_host, const uint8_t *bits_xyz,
                       const uint8_t *bits_xy, const uint8_t *bits_xz, const uint8_t *bits_yz, const int32_t *offsets3,
                       const int32_t *resolutions3, const int32_t *offsets2, const int32_t *resolutions2, const float *blob,
                       float *sigma_host, float *rgb_host, uint32_t N, float *d_pos, float *d_dirs, float *d_sigma,
                       float *d_rgb, uint32_t wave_samples, uint32_t max_chunk_waves, cnc_stream_t s_compute, cnc_stream_t s_in,
                       cnc_stream_t s_out) {
    if (N == 0) return CNC_OK;
    if (!pos_host || !dirs_host || !sigma_host || !rgb_host || !d_pos || !d_dirs || !d_sigma || !d_rgb || wave_samples == 0 || max_chunk_waves == 0) {
        set_error("field_fwd_host: null pointer / zero chunk");
        return CNC_EINVAL;
    }
    constexpr uint32_t MAXW = 4096, NSLOT = 8, NPIPE = 4;
    static cudaEvent_t ev[4];
    static uint32_t *h_stamp = nullptr, *d_ready = nullptr, *d_done = nullptr, epoch = 0;
    static uint32_t *cum = nullptr;                       // tiles published per wave since start-up (the counters never reset)
    typedef CUresult (*wait32_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    static wait32_t wait32 = nullptr;                     // cuStreamWaitValue32: a stream-side wait that needs no SM
    static int n_sm = 0;
    static const float *pipe_key[NPIPE] = {nullptr, nullptr, nullptr, nullptr};   // staging buffer -> its own flags / counters
    if (!h_stamp) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        bool ok = cudaMallocHost(&h_stamp, NSLOT * MAXW * 4) == cudaSuccess && cudaMalloc(&d_ready, NPIPE * MAXW * 4) == cudaSuccess &&
                  cudaMalloc(&d_done, NPIPE * MAXW * 4) == cudaSuccess && cudaMemset(d_ready, 0, NPIPE * MAXW * 4) == cudaSuccess &&
                  cudaMemset(d_done, 0, NPIPE * MAXW * 4) == cudaSuccess;
        for (int i = 0; i < 4 && ok; i++) ok = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) == cudaSuccess;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        ok = ok && cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess && fn != nullptr;
        wait32 = reinterpret_cast<wait32_t>(fn);
        cum = static_cast<uint32_t *>(calloc(NPIPE * MAXW, 4));
        ok = ok && cum != nullptr;
        if (!ok) { h_stamp = nullptr; set_error("field_fwd_host: cannot allocate the pipeline state"); return CNC_ECUDA; }
    }
    if (wave_samples != (uint32_t)n_sm * ff::TILE_M) { set_error("field_fwd_host: wave_samples must be SMs x 128 = %d", n_sm * ff::TILE_M); return CNC_EINVAL; }
    const uint32_t ntiles = (N + ff::TILE_M - 1) / ff::TILE_M, grid = ntiles < (uint32_t)n_sm ? ntiles : (uint32_t)n_sm;
    const uint32_t nwaves = (ntiles + grid - 1) / grid;
    if (nwaves > MAXW) { set_error("field_fwd_host: more than %u waves", MAXW); return CNC_EINVAL; }
    cudaStream_t sc = static_cast<cudaStream_t>(s_compute), si = static_cast<cudaStream_t>(s_in), so = static_cast<cudaStream_t>(s_out);
    // Calls on different staging buffers issued from different streams overlap (the upload of one runs beside the kernel of
    // the other): each staging buffer gets its own ready flags and tile counters.
    uint32_t pipe = 0;
    while (pipe < NPIPE && pipe_key[pipe] != d_pos && pipe_key[pipe] != nullptr) pipe++;
    if (pipe == NPIPE) { pipe = 0; cudaDeviceSynchronize(); for (uint32_t i = 1; i < NPIPE; i++) pipe_key[i] = nullptr; }   // table full: start over
    pipe_key[pipe] = d_pos;
    uint32_t *const ready_p = d_ready + pipe * MAXW, *const done_p = d_done + pipe * MAXW, *const cum_p = cum + pipe * MAXW;
    epoch++;
    uint32_t *stamp = h_stamp + (epoch % NSLOT) * MAXW;   // a slot per call: earlier calls may still be copying from theirs
    for (uint32_t w = 0; w < nwaves; w++) stamp[w] = epoch;
    // chunk schedule in waves: 1, 2, 4, .. max .. 4, 2, 1
    uint32_t front[64], back[64];
    int nf = 0, nb = 0;
    uint32_t left = nwaves;
    for (uint32_t sz = 1; left > 0 && nf < 63 && nb < 63; sz = sz * 2 < max_chunk_waves ? sz * 2 : max_chunk_waves) {
        uint32_t t = sz < left ? sz : left;
        front[nf++] = t; left -= t;
        if (left == 0) break;
        t = sz < left ? sz : left;
        back[nb++] = t; left -= t;
    }
    if (left > 0) front[nf - 1] += left;   // (only for tiny max_chunk_waves on huge batches)
    // the staging buffers and the stamps may still be in use by earlier work on the caller's stream
    cudaEventRecord(ev[0], sc); cudaStreamWaitEvent(si, ev[0], 0); cudaStreamWaitEvent(so, ev[0], 0);
    const int rc = field_fwd_impl(d_pos, d_dirs, aabb6_host, bits_xyz, bits_xy, bits_xz, bits_yz, offsets3, resolutions3, offsets2,
                                  resolutions2, blob, d_sigma, d_rgb, nullptr, nullptr, nullptr, nullptr, nullptr, N, s_compute, ready_p,
                                  done_p, epoch);
    if (rc != CNC_OK) return rc;
    const uint32_t last_ctas = ntiles - (nwaves - 1) * grid;
    uint32_t w0 = 0;
    for (int c = 0; c < nf + nb; c++) {
        const uint32_t waves = c < nf ? front[c] : back[nb - 1 - (c - nf)], w1 = w0 + waves;
        const size_t lo = (size_t)w0 * grid * ff::TILE_M;
        size_t hi = (size_t)w1 * grid * ff::TILE_M;
        if (hi > N) hi = N;
        const size_t n = hi - lo;
        cudaMemcpyAsync(d_pos + lo * 3, pos_host + lo * 3, n * 12, cudaMemcpyHostToDevice, si);
        cudaMemcpyAsync(d_dirs + lo * 3, dirs_host + lo * 3, n * 12, cudaMemcpyHostToDevice, si);
        cudaMemcpyAsync(ready_p + w0, stamp + w0, (size_t)waves * 4, cudaMemcpyHostToDevice, si);   // behind the data, same stream
        for (uint32_t w = w0; w < w1; w++) {   // the download waits (on the stream, no kernel) for every tile of the chunk
            cum_p[w] += w == nwaves - 1 ? last_ctas : grid;
            if (wait32((CUstream)so, (CUdeviceptr)(uintptr_t)(done_p + w), cum_p[w], CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
                set_error("field_fwd_host: cuStreamWaitValue32 failed");
                return CNC_ECUDA;
            }
        }
        cudaMemcpyAsync(rgb_host + lo * 3, d_rgb + lo * 3, n * 12, cudaMemcpyDeviceToHost, so);
        cudaMemcpyAsync(sigma_host + lo, d_sigma + lo, n * 4, cudaMemcpyDeviceToHost, so);
        w0 = w1;
    }
    cudaEventRecord(ev[1], so); cudaStreamWaitEvent(sc, ev[1], 0);   // the caller's stream ends after the last download
    cudaEventRecord(ev[2], si); cudaStreamWaitEvent(sc, ev[2], 0);
    return check_launch("field_fwd_host");
}

int cnc_field_fwd_train(const float *pos, const float *dirs, const float *aabb6_host, const uint8_t *bits_xyz,
                        const uint8_t *bits_xy, const uint8_t *bits_xz, const uint8_t *bits_yz, const int32_t *offsets3,
                        const int32_t *resolutions3, const int32_t *offsets2, const int32_t *resolutions2,
                        const float *blob, float *sigma, float *rgb, float *geo, float *x0, float *h1, float *h3, float *h4,
                        uint32_t N, cnc_stream_t stream) {
    if (!dirs || !geo || !x0 || !h1 || !h3 || !h4) { set_error("field_fwd_train: null pointer"); return CNC_EINVAL; }
    if ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(h1) | reinterpret_cast<uintptr_t>(h3) |
         reinterpret_cast<uintptr_t>(h4) | reinterpret_cast<uintptr_t>(geo)) & 31u) {
        set_error("field_fwd_train: activation buffers must be 32-byte aligned (rows leave as 256-bit stores); head_in is [N,96]");
        return CNC_EINVAL;
    }
    return field_fwd_impl(pos, dirs, aabb6_host, bits_xyz, bits_xy, bits_xz, bits_yz, offsets3, resolutions3, offsets2,
                          resolutions2, blob, sigma, rgb, geo, x0, h1, h3, h4, N, stream);
}

}  // extern "C"
