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
// Mapping: one warp per hash entry (its voxel list is contiguous in the inverse table), lanes stride over
// the voxels, per-lane partial sums are combined with a fixed xor-shuffle tree -> the result depends only on
// the voxel list, never on the grid size or GPU count, so encoder and decoder agree bit for bit.  The MLP
// weights sit in shared memory, transposed so that one 128-bit broadcast load feeds four FMAs.
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
};

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : __fmul_rn(x, 0.01f); }

__global__ void __launch_bounds__(256) context3d_kernel(const Args a) {
    __shared__ __align__(16) float w[MLP_FLOATS];
    for (int i = threadIdx.x; i < MLP_FLOATS; i += blockDim.x) w[i] = __ldg(a.mlp + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float res_n = (float)__ldg(a.res + a.level);
    const float scale_n = __fsub_rn(res_n, 2.0f);
    LevelConst lc[3];
    const uint32_t *vb[3];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        lc[l] = load_level(a.offs, a.res, (uint32_t)(a.level - 3 + l));
        vb[l] = a.vbits ? a.vbits + (__ldg(a.vbit_off + a.level - 3 + l) >> 5) : nullptr;
    }

    for (int64_t e = warp0; e < a.Ne; e += nwarp) {
        const int64_t v0 = __ldg(a.seg + e) - a.seg_base, v1 = __ldg(a.seg + e + 1) - a.seg_base;
        float acc[F];
#pragma unroll
        for (int k = 0; k < F; k++) acc[k] = 0.f;
        float osum = 0.f;
        for (int64_t v = v0 + lane; v < v1; v += 32) {
            const int c[3] = {(int)__ldg(a.pts + v * 3), (int)__ldg(a.pts + v * 3 + 1), (int)__ldg(a.pts + v * 3 + 2)};
            int32_t ov;
            if (!voxel_mask_overlap<3>(c, res_n, a.Rb, a.vxl, ov)) continue;   // utils_bpp_acc.py:811-814
            const float wv = (float)(ov < 1 ? 1 : ov);                           // clamp(min=1), :826
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
            float h1[NH], h2[NH], o[F];
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
#pragma unroll
            for (int k = 0; k < F; k++) acc[k] = __fmaf_rn(o[k], wv, acc[k]);   // mean * overlap, summed per entry (:843-848)
            osum = __fadd_rn(osum, wv);
        }
        // fixed-shape butterfly: deterministic for a given voxel list
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            osum = __fadd_rn(osum, __shfl_xor_sync(0xFFFFFFFFu, osum, s));
#pragma unroll
            for (int k = 0; k < F; k++) acc[k] = __fadd_rn(acc[k], __shfl_xor_sync(0xFFFFFFFFu, acc[k], s));
        }
        if (lane < F) {
            float m = 0.f;
#pragma unroll
            for (int k = 0; k < F; k++) if (k == lane) m = acc[k];
            const bool ex = osum > 0.f;
            m = ex ? __fdiv_rn(m, osum) : 0.f;
            if (a.mean) a.mean[e * F + lane] = m;
            a.prob[e * F + lane] = ex ? fminf(fmaxf(m, 1e-6f), 1.0f - 1e-6f) : 0.f;   // :852, :1006
            if (lane == 0) a.exist[e] = ex ? 1 : 0;
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
                        const uint32_t *vertex_bits, const int64_t *vertex_bit_offsets, cnc_stream_t stream) {
    if (n_entries == 0) return CNC_OK;
    if (!pts || !seg || !binary_vxl || !sign_bits || !offsets || !resolutions || !mlp_packed || !prob || !exist || Rb <= 0) {
        set_error("context3d_probs: bad argument");
        return CNC_EINVAL;
    }
    if ((vertex_bits == nullptr) != (vertex_bit_offsets == nullptr)) { set_error("context3d_probs: vertex_bits and vertex_bit_offsets go together"); return CNC_EINVAL; }
    if (level < 3) { set_error("context3d_probs: needs three coarser context levels (level >= 3)"); return CNC_ENOTSUP; }
    cf::Args a{pts, seg, seg_base, binary_vxl, Rb, sign_bits, vertex_bits, vertex_bit_offsets, offsets, resolutions, level, Pg, mlp_packed, prob, mean, exist, n_entries};
    const int64_t warps = n_entries;
    int64_t blocks = (warps + 7) / 8;
    const int64_t cap = 148 * 8 * 4;  // persistent-ish: a few waves of 8 resident CTAs per SM, warps stride over entries
    if (blocks > cap) blocks = cap;
    cf::context3d_kernel<<<(uint32_t)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("context3d_probs");
}

}  // extern "C"
