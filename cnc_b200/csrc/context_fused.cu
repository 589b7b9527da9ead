// context_fused.cu -- level-wise context model of the 3D hash grid as ONE kernel per coded chunk:
//   for every hash entry of level n: walk the voxels that map to it (inverse hash table), keep those whose
//   neighbourhood touches the occupancy grid (K6), interpolate the three coarser, already coded levels at
//   the voxel centre (K1 with occupancy mask), run context_model_3D (25 -> 32 -> 32 -> 8, LeakyReLU) and
//   average the predictions with the overlap volumes as weights -> P(+1) of the entry's 8 features.
//
// Reference behaviour restated: examples/utils_bpp_acc.py:798-852 (encode) == :929-968 (decode):
//   query_mask_3D -> boolean compaction -> align_and_pack(mask) -> Encoding_xyz(points, n-3, n, binary_vxl)
//   -> cat Pg -> context_model_3D -> align_and_pack -> * overlap/sum(overlap) -> sum(dim=1) -> clamp.
// The reference materialises up to 2e7 voxels x 25 floats plus two padded [entries, max_count, F] tensors
// per chunk; here nothing but the [entries, 8] probabilities reaches HBM.
//
// Mapping: see context3d_kernel below (CTA per batch of 64 entries, passing voxels compacted into a queue so that
// every lane of the evaluation carries a voxel).  The MLP weights sit in shared memory, transposed so that one
// 128-bit broadcast load feeds four FMAs.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {
namespace cf {

constexpr int F = 8, NIN = 25, NH = 32;
// packed MLP (floats): W1T [25][32], b1 [32], W2T [32][32], b2 [32], W3T [32][8], b3 [8]
constexpr int O_W1 = 0, O_B1 = O_W1 + NIN * NH, O_W2 = O_B1 + NH, O_B2 = O_W2 + NH * NH, O_W3 = O_B2 + NH,
              O_B3 = O_W3 + NH * F, MLP_FLOATS = O_B3 + F;  // 2152

struct Args {
    const int16_t *pts;   // [Nv,3] voxel coordinates of level n, grouped by hash entry
    const int64_t *seg;   // [Ne+1] running voxel count per entry (absolute), seg[0] == seg_base
    int64_t seg_base;
    const uint8_t *vxl;   // [Rb]^3 occupancy
    int32_t Rb;
    const uint8_t *bits;  // 1-bit sign table of the whole 3D encoder (cnc_sign_pack)
    const uint32_t *vbits;     // vertex validity bitmaps of all levels (cnc_vertex_valid_bits), nullable
    const int64_t *vbit_off;   // [L+1] bit offset of each level's bitmap (multiples of 32)
    const int32_t *offs, *res;  // level arrays of the encoder
    int32_t level;        // n >= 3; context = levels n-3, n-2, n-1
    float Pg;             // level-wide frequency of +1 (utils_bpp_acc.py:472-486)
    const float *mlp;     // MLP_FLOATS packed weights
    float *prob;          // [Ne,8] clamp(mean, 1e-6, 1-1e-6); 0 where the entry is not coded
    float *mean;          // [Ne,8] unclamped (nullable)
    uint8_t *exist;       // [Ne] entry has at least one voxel touching the occupancy
    int64_t Ne;
    int64_t entry_base;   // absolute index (inside the level) of the chunk's first entry: batches are aligned to it
};

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : __fmul_rn(x, 0.01f); }

// context of one voxel (three masked coarser levels at the voxel centre) -> context_model_3D -> o[8]
__device__ __forceinline__ void voxel_probs(const Args &a, const float *__restrict__ w, const LevelConst (&lc)[3],
                                            const uint32_t *const (&vb)[3], const int (&c)[3], float scale_n, float (&o)[F]) {
    // voxel centre in [0,1]: (c - 0.5) / (res - 2)   (utils_bpp_acc.py:810)
    float x[3];
#pragma unroll
    for (int d = 0; d < 3; d++) x[d] = __fdiv_rn(__fsub_rn((float)c[d], 0.5f), scale_n);
    float in[NIN];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        Corners<3> cs;
        float f[F];
#pragma unroll
        for (int k = 0; k < F; k++) f[k] = 0.f;
        // "vertex touches the occupancy" (gridencoder.cu:221-276): read from the per-level bitmap when the
        // caller built one (same predicate, evaluated once per vertex instead of once per use)
        const uint32_t *vbl = vb[l];
        const LevelConst &lcl = lc[l];
        const bool inside = vbl
            ? make_corners_fn<3>(x, lcl, [&](const uint32_t (&cc)[3]) {
                  const uint32_t v = (cc[0] * lcl.res + cc[1]) * lcl.res + cc[2];
                  return ((__ldg(vbl + (v >> 5)) >> (v & 31u)) & 1u) != 0u; }, cs)
            : make_corners<3>(x, lcl, (uint32_t)a.Rb, a.vxl, cs);
        if (inside) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if ((cs.valid >> i) & 1u) {
                    const uint32_t sb = __ldg(a.bits + lc[l].base_row + cs.row[i]);
                    const float ww = __fmul_rn(cs.w[i], cs.wn_re);
#pragma unroll
                    for (int k = 0; k < F; k++) f[k] = __fadd_rn(f[k], ((sb >> k) & 1u) ? ww : -ww);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < F; k++) in[l * F + k] = f[k];
    }
    in[NIN - 1] = a.Pg;
    // context_model_3D: Linear(25,32) LeakyReLU Linear(32,32) LeakyReLU Linear(32,8); fixed fma order
    float h1[NH], h2[NH];
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
    for (int k = 0; k < F; k++) o[k] = w[O_B3 + k];
#pragma unroll
    for (int i = 0; i < NH; i++) {
        const float hv = leaky(h2[i]);
        const float4 t0 = *reinterpret_cast<const float4 *>(w + O_W3 + i * F), t1 = *reinterpret_cast<const float4 *>(w + O_W3 + i * F + 4);
        o[0] = __fmaf_rn(hv, t0.x, o[0]); o[1] = __fmaf_rn(hv, t0.y, o[1]);
        o[2] = __fmaf_rn(hv, t0.z, o[2]); o[3] = __fmaf_rn(hv, t0.w, o[3]);
        o[4] = __fmaf_rn(hv, t1.x, o[4]); o[5] = __fmaf_rn(hv, t1.y, o[5]);
        o[6] = __fmaf_rn(hv, t1.z, o[6]); o[7] = __fmaf_rn(hv, t1.w, o[7]);
    }
}

constexpr int EB = 64;      // entries per batch (aligned to the absolute entry index -> partition independent)
constexpr int CT = 256;     // threads per CTA = voxels per fill / evaluation round
constexpr int QCAP = 2 * CT;

// One CTA per batch of EB consecutive hash entries (their voxel lists are contiguous in the inverse table):
//   fill:     256 voxels at a time, one per thread, K6 mask + overlap (cheap) -> the passing ones are appended, in order,
//             to a queue in shared memory;
//   evaluate: as soon as 256 are queued (or the batch ends) every thread takes ONE passing voxel through the context
//             gather and the MLP -- no lane idles behind a masked-out neighbour, which is what held the
//             warp-per-entry mapping at 2-12 active lanes of 32;
//   reduce:   overlap-weighted sums per entry with a segmented warp scan (queue order = voxel order, entries are
//             runs), run tails add into per-entry accumulators warp after warp.
// The summation order depends only on the voxel list and on the chunk / (absolute) batch boundaries, never on the
// grid size or on which GPU takes a chunk, so encoder and decoder -- which cut a level into the same chunks,
// utils_bpp_acc.py:798-802 == :929-933 -- agree bit for bit.
__global__ void __launch_bounds__(CT, 2) context3d_kernel(const Args a) {
    __shared__ __align__(16) float w[MLP_FLOATS];
    __shared__ int32_t segs[EB + 1];
    __shared__ int32_t qv[QCAP];
    __shared__ float qw[QCAP];
    __shared__ float acc[EB][F + 1];
    __shared__ int32_t wcnt[CT / 32];
    for (int i = threadIdx.x; i < MLP_FLOATS; i += CT) w[i] = __ldg(a.mlp + i);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float res_n = (float)__ldg(a.res + a.level);
    const float scale_n = __fsub_rn(res_n, 2.0f);
    LevelConst lc[3];
    const uint32_t *vb[3];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        lc[l] = load_level(a.offs, a.res, (uint32_t)(a.level - 3 + l));
        vb[l] = a.vbits ? a.vbits + (__ldg(a.vbit_off + a.level - 3 + l) >> 5) : nullptr;
    }
    const int64_t b_first = a.entry_base / EB, b_last = (a.entry_base + a.Ne + EB - 1) / EB;

    for (int64_t b = b_first + blockIdx.x; b < b_last; b += gridDim.x) {
        const int64_t abs0 = b * EB;
        const int64_t e0 = (abs0 > a.entry_base ? abs0 : a.entry_base) - a.entry_base;
        const int64_t e1x = abs0 + EB < a.entry_base + a.Ne ? abs0 + EB : a.entry_base + a.Ne;
        const int nE = (int)(e1x - a.entry_base - e0);
        __syncthreads();
        if (tid <= nE) segs[tid] = (int32_t)(__ldg(a.seg + e0 + tid) - a.seg_base);
        for (int i = tid; i < EB * (F + 1); i += CT) (&acc[0][0])[i] = 0.f;
        __syncthreads();
        const int32_t v0 = segs[0], v1 = segs[nE];
        int32_t tile = v0;
        int qn = 0;   // queued voxels (kept identically by every thread)
        while (true) {
            // ---- fill
            while (qn < CT && tile < v1) {
                const int32_t v = tile + tid;
                bool pass = false;
                float wv = 0.f;
                if (v < v1) {
                    const int c[3] = {(int)__ldg(a.pts + (int64_t)v * 3), (int)__ldg(a.pts + (int64_t)v * 3 + 1), (int)__ldg(a.pts + (int64_t)v * 3 + 2)};
                    int32_t ov;
                    pass = voxel_mask_overlap<3>(c, res_n, a.Rb, a.vxl, ov);   // utils_bpp_acc.py:811-814
                    wv = (float)(ov < 1 ? 1 : ov);                             // clamp(min=1), :826
                }
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
                if (lane == 0) wcnt[warp] = __popc(bal);
                __syncthreads();
                int base = qn, total = 0;
#pragma unroll
                for (int k = 0; k < CT / 32; k++) {
                    const int ck = wcnt[k];
                    if (k < warp) base += ck;
                    total += ck;
                }
                if (pass) {
                    const int p = base + __popc(bal & ((1u << lane) - 1u));
                    qv[p] = v;
                    qw[p] = wv;
                }
                __syncthreads();
                qn += total;
                tile += CT;
            }
            if (qn == 0) break;
            // ---- evaluate one queued voxel per thread
            const int n = qn < CT ? qn : CT;
            int e = -1 - tid;   // entry of this thread's voxel (negative and distinct = not part of any run)
            float p[F + 1];
#pragma unroll
            for (int k = 0; k <= F; k++) p[k] = 0.f;
            if (tid < n) {
                const int32_t v = qv[tid];
                const float wv = qw[tid];
                int lo = 0, hi = nE;           // largest e with segs[e] <= v
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (segs[mid] <= v) lo = mid; else hi = mid;
                }
                e = lo;
                const int c[3] = {(int)__ldg(a.pts + (int64_t)v * 3), (int)__ldg(a.pts + (int64_t)v * 3 + 1), (int)__ldg(a.pts + (int64_t)v * 3 + 2)};
                float o[F];
                voxel_probs(a, w, lc, vb, c, scale_n, o);
#pragma unroll
                for (int k = 0; k < F; k++) p[k] = __fmul_rn(o[k], wv);   // mean * overlap, summed per entry (:843-848)
                p[F] = wv;
            }
            // ---- segmented inclusive scan over the lanes of a warp (key = entry), fixed shape
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int pe = __shfl_up_sync(0xFFFFFFFFu, e, d);
                const bool take = lane >= d && pe == e;
#pragma unroll
                for (int k = 0; k <= F; k++) {
                    const float t = __shfl_up_sync(0xFFFFFFFFu, p[k], d);
                    if (take) p[k] = __fadd_rn(p[k], t);
                }
            }
            const int en = __shfl_down_sync(0xFFFFFFFFu, e, 1);
            const bool tail = e >= 0 && (lane == 31 || en != e);
            for (int wq = 0; wq < CT / 32; wq++) {      // warp after warp: a run that spans warps is added in order
                if (warp == wq && tail) {
#pragma unroll
                    for (int k = 0; k <= F; k++) acc[e][k] = __fadd_rn(acc[e][k], p[k]);
                }
                __syncthreads();
            }
            // ---- drop the evaluated voxels from the queue
            const int rest = qn - n;
            int32_t tv = 0;
            float tw = 0.f;
            if (tid < rest) { tv = qv[n + tid]; tw = qw[n + tid]; }
            __syncthreads();
            if (tid < rest) { qv[tid] = tv; qw[tid] = tw; }
            __syncthreads();
            qn = rest;
        }
        // ---- per-entry mean, clamp
        for (int i = tid; i < nE * F; i += CT) {
            const int el = i / F, k = i - el * F;
            const float osum = acc[el][F];
            const bool ex = osum > 0.f;
            const float m = ex ? __fdiv_rn(acc[el][k], osum) : 0.f;
            const int64_t eg = e0 + el;
            if (a.mean) a.mean[eg * F + k] = m;
            a.prob[eg * F + k] = ex ? fminf(fmaxf(m, 1e-6f), 1.0f - 1e-6f) : 0.f;   // :852, :1006
            if (k == 0) a.exist[eg] = ex ? 1 : 0;
        }
    }
}

// One bit per grid vertex of every level: does the +-1-voxel box around the vertex touch an occupied cell
// (the per-corner test of the reference's masked gather, gridencoder.cu:221-276).  Bit v = (c0*res + c1)*res + c2
// of level l lives at bit_off[l] + v.  A warp covers 32 consecutive vertices -> one ballot, one store.
__global__ void __launch_bounds__(256) vertex_valid_kernel(const uint8_t *__restrict__ vxl, int32_t Rb, const int32_t *__restrict__ res_list,
                                                           int32_t n_levels, const int64_t *__restrict__ bit_off, uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t total_words = __ldg(bit_off + n_levels) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int level = 0;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total_words; w += nwarp) {
        while (level + 1 < n_levels && (w << 5) >= __ldg(bit_off + level + 1)) level++;
        while (level > 0 && (w << 5) < __ldg(bit_off + level)) level--;
        const uint32_t res = (uint32_t)__ldg(res_list + level);
        const int64_t v = (w << 5) - __ldg(bit_off + level) + lane;
        bool ok = false;
        if (v < (int64_t)res * res * res) {
            const uint32_t c[3] = {(uint32_t)(v / ((int64_t)res * res)), (uint32_t)((v / res) % res), (uint32_t)(v % res)};
            const float scale_re = __frcp_rn((float)(res - 2u));
            ok = occ_box_any<3>(c, scale_re, (uint32_t)Rb, vxl);
        }
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, ok);
        if (lane == 0) out[w] = word;
    }
}

}  // namespace cf
}  // namespace cnc

using namespace cnc;

extern "C" {

uint32_t cnc_context3d_mlp_floats(void) { return cf::MLP_FLOATS; }

int cnc_vertex_valid_bits(const uint8_t *binary_vxl, int32_t Rb, const int32_t *resolutions, int32_t n_levels,
                           const int64_t *bit_offsets, int64_t total_bits, uint32_t *out_words, cnc_stream_t stream) {
    if (n_levels <= 0 || total_bits <= 0) return CNC_OK;
    if (!binary_vxl || !resolutions || !bit_offsets || !out_words || Rb <= 0 || (total_bits & 31)) {
        set_error("vertex_valid_bits: bad argument (offsets must be multiples of 32 bits)");
        return CNC_EINVAL;
    }
    const int64_t words = total_bits >> 5;
    int64_t blocks = (words + 7) / 8;
    if (blocks > 148 * 8 * 8) blocks = 148 * 8 * 8;
    cf::vertex_valid_kernel<<<(uint32_t)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(binary_vxl, Rb, resolutions, n_levels,
                                                                                          bit_offsets, out_words);
    return check_launch("vertex_valid_bits");
}

int cnc_context3d_probs(const int16_t *pts, const int64_t *seg, int64_t n_entries, const uint8_t *binary_vxl, int32_t Rb,
                        const uint8_t *sign_bits, const int32_t *offsets, const int32_t *resolutions, int32_t level,
                        float Pg, const float *mlp_packed, float *prob, float *mean, uint8_t *exist, int64_t seg_base,
                        int64_t entry_base, const uint32_t *vertex_bits, const int64_t *vertex_bit_offsets, cnc_stream_t stream) {
    if (n_entries == 0) return CNC_OK;
    if (!pts || !seg || !binary_vxl || !sign_bits || !offsets || !resolutions || !mlp_packed || !prob || !exist || Rb <= 0) {
        set_error("context3d_probs: bad argument");
        return CNC_EINVAL;
    }
    if ((vertex_bits == nullptr) != (vertex_bit_offsets == nullptr)) { set_error("context3d_probs: vertex_bits and vertex_bit_offsets go together"); return CNC_EINVAL; }
    if (level < 3) { set_error("context3d_probs: needs three coarser context levels (level >= 3)"); return CNC_ENOTSUP; }
    cf::Args a{pts, seg, seg_base, binary_vxl, Rb, sign_bits, vertex_bits, vertex_bit_offsets, offsets, resolutions, level, Pg, mlp_packed, prob, mean, exist, n_entries, entry_base};
    if (entry_base < 0) { set_error("context3d_probs: entry_base must be >= 0"); return CNC_EINVAL; }
    const int64_t nbatch = (entry_base + n_entries + cf::EB - 1) / cf::EB - entry_base / cf::EB;
    int64_t blocks = nbatch;
    const int64_t cap = 148 * 2 * 8;   // two resident CTAs per SM; further batches are taken grid-stride
    if (blocks > cap) blocks = cap;
    cf::context3d_kernel<<<(uint32_t)blocks, cf::CT, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("context3d_probs");
}

}  // extern "C"
