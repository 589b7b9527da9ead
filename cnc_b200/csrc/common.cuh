// common.cuh -- shared device helpers of libcnc_b200 (sm_100a).
//
// The corner logic below is the single source of truth for the CNC grid semantics
// (SURVEY Appendix A): 1-cell zero border, (res-2) scaling + 0.5 offset, optional
// occupancy test per corner, weight renormalisation over the valid corners.  It is written
// with explicit-rounding intrinsics (__fmul_rn / __fadd_rn / __fmaf_rn) so that the sequence
// of fp32 roundings is fixed by this source and not by the compiler's contraction choices;
// it reproduces the roundings of the reference kernels (gridencoder.cu:160-291, whose
// double-literal intermediates all collapse to correctly rounded fp32 operations).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cnc_b200.h"

namespace cnc {

void set_error(const char *fmt, ...);
int check_launch(const char *what);  // cudaGetLastError -> CNC_ECUDA

__host__ __device__ inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// ---- index: gridencoder.cu:45-87 -----------------------------------------------------------
template <int D>
__device__ __forceinline__ uint32_t grid_row(const uint32_t (&c)[D], uint32_t T, uint32_t res) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        if (stride <= T) {
            index += c[d] * stride;
            stride *= res;
        }
    }
    if (stride > T) {
        constexpr uint32_t P[3] = {1u, 2654435761u, 805459861u};
        index = 0;
#pragma unroll
        for (int d = 0; d < D; d++) index ^= c[d] * P[d];
    }
    // T is a power of two for every hashed level of the shipped layouts, but dense levels
    // (and user layouts) are not: keep the generic modulo unless it is a mask.
    return ((T & (T - 1)) == 0) ? (index & (T - 1)) : (index % T);
}

// per-level constants, computed once per thread (cheap) ------------------------------------
struct LevelConst {
    uint32_t T, res, base_row;
    float scale;     // float(res-2)
    float scale_re;  // fl(1/(res-2))  == float(1.0/(double(float(res))-2.0)), gridencoder.cu:224
};

__device__ __forceinline__ LevelConst load_level(const int32_t *__restrict__ offsets,
                                                 const int32_t *__restrict__ resolutions,
                                                 uint32_t level) {
    LevelConst lc;
    const uint32_t o0 = (uint32_t)__ldg(offsets + level), o1 = (uint32_t)__ldg(offsets + level + 1);
    lc.base_row = o0;
    lc.T = o1 - o0;
    lc.res = (uint32_t)__ldg(resolutions + level);
    lc.scale = (float)(lc.res - 2u);
    lc.scale_re = __frcp_rn(lc.scale);
    return lc;
}

// "does the +-1-voxel box around grid vertex c touch an occupied cell" (gridencoder.cu:221-276)
template <int D>
__device__ __forceinline__ bool occ_box_any(const uint32_t (&c)[D], float scale_re, uint32_t Rb,
                                            const uint8_t *__restrict__ vxl) {
    int lo[D], hi[D];
    const float fRb = (float)Rb, fRb1 = (float)(Rb - 1u);
    const float mhalf = __fmul_rn(-0.5f, scale_re);
#pragma unroll
    for (int d = 0; d < D; d++) {
        // float((double(c) - 0.5) * double(scale_re)): exact product, one rounding == fma
        const float pn = __fmaf_rn((float)c[d], scale_re, mhalf);
        float g1 = __fmul_rn(__fsub_rn(pn, scale_re), fRb);
        g1 = g1 < 0.f ? 0.f : g1;
        g1 = g1 > fRb1 ? fRb1 : g1;
        lo[d] = (int)g1;
        float g2 = __fmul_rn(__fadd_rn(pn, scale_re), fRb);
        g2 = g2 < 0.f ? 0.f : g2;
        g2 = g2 > fRb1 ? fRb1 : g2;
        hi[d] = (int)g2;
    }
    if (D == 1) {
        for (int a = lo[0]; a <= hi[0]; a++)
            if (vxl[a]) return true;
    } else if (D == 2) {
        for (int a = lo[0]; a <= hi[0]; a++)
            for (int b = lo[1]; b <= hi[1]; b++)
                if (vxl[(size_t)a * Rb + b]) return true;
    } else {
        for (int a = lo[0]; a <= hi[0]; a++)
            for (int b = lo[1]; b <= hi[1]; b++) {
                const uint8_t *row = vxl + ((size_t)a * Rb + b) * Rb;
                for (int cc = lo[D - 1]; cc <= hi[D - 1]; cc++)
                    if (row[cc]) return true;
            }
    }
    return false;
}

template <int D>
struct Corners {
    float w[1 << D];       // raw D-linear weights (not yet renormalised)
    uint32_t row[1 << D];  // row inside the level slab
    uint32_t valid;        // bit i set <=> corner i contributes
    float wn_re;           // 1 / sum of valid weights
};

// returns false if the point lies outside [0,1]^D (forward writes zeros, backward skips)
// `occ(c)` decides whether grid vertex c may contribute (occupancy test of the reference, or a precomputed bitmap)
template <int D, class OccFn>
__device__ __forceinline__ bool make_corners_fn(const float (&x)[D], const LevelConst &lc, OccFn occ, Corners<D> &cs) {
    bool oob = false;
#pragma unroll
    for (int d = 0; d < D; d++) oob |= (x[d] < 0.f) | (x[d] > 1.f);  // NaN -> in range, like the ref
    if (oob) return false;
    float f[D];
    uint32_t g[D];
#pragma unroll
    for (int d = 0; d < D; d++) {
        // gridencoder.cu:173 `x * float(res-2) + 0.5`: nvcc contracts the widened multiply-add of the
        // reference into a single DFMA, i.e. one rounding of x*s + 0.5 == fmaf (see oracle/cnc_oracle.c)
        const float p = __fmaf_rn(x[d], lc.scale, 0.5f);
        const float fl = floorf(p);
        g[d] = (uint32_t)fl;
        f[d] = __fsub_rn(p, (float)g[d]);
    }
    float wn = 0.f;
    cs.valid = 0;
#pragma unroll
    for (int i = 0; i < (1 << D); i++) {
        float w = 1.f;
        uint32_t c[D];
        bool zero = false;
#pragma unroll
        for (int d = 0; d < D; d++) {
            if ((i >> d) & 1) {
                w = __fmul_rn(w, f[d]);
                c[d] = min(g[d] + 1u, lc.res - 1u);
            } else {
                w = __fmul_rn(w, __fsub_rn(1.f, f[d]));
                c[d] = g[d];
            }
            zero |= (c[d] == 0u) | (c[d] == lc.res - 1u);  // gridencoder.cu:212-219
        }
        bool ok = !zero;
        if (ok) ok = occ(c);
        cs.w[i] = w;
        cs.row[i] = 0;
        if (ok) {
            cs.row[i] = grid_row<D>(c, lc.T, lc.res);
            wn = __fadd_rn(wn, w);
            cs.valid |= 1u << i;
        }
    }
    if (wn == 0.f) wn = 1e-9f;  // float(0.0 + 1e-9), gridencoder.cu:288-290
    cs.wn_re = __frcp_rn(wn);   // float(1.0 / double(wn)): correctly rounded reciprocal
    return true;
}

template <int D>
__device__ __forceinline__ bool make_corners(const float (&x)[D], const LevelConst &lc, uint32_t Rb,
                                             const uint8_t *__restrict__ vxl, Corners<D> &cs) {
    return make_corners_fn<D>(x, lc, [&](const uint32_t (&c)[D]) { return vxl ? occ_box_any<D>(c, lc.scale_re, Rb, vxl) : true; }, cs);
}


// "does the +-1-voxel box around voxel `c` (int coords at a level of resolution r) touch an occupied cell, and
// with how much overlap": aligner_kernel.cu:161-242 (3D) / :4-80 (2D).  Returns the mask; `overlap` receives
// int(overlap_volume * Rb^D * 1000) (truncation), the reference's overlap_area_pool entry.
template <int D>
__device__ __forceinline__ bool voxel_mask_overlap(const int (&c)[D], float r, int32_t Rb, const uint8_t *__restrict__ vxl,
                                                   int32_t &overlap) {
    const float fRb = (float)Rb, fRb1 = (float)(Rb - 1);
    const float Rb_re = __frcp_rn(fRb);
    const float scale_re = __frcp_rn(__fsub_rn(r, 2.0f));  // float(1.0/(double(r)-2.0)), exact sub
    const float mhalf = __fmul_rn(-0.5f, scale_re);
    float pn[D];
    int lo[D], hi[D];
#pragma unroll
    for (int d = 0; d < D; d++) {
        pn[d] = __fmaf_rn((float)c[d], scale_re, mhalf);
        float g1 = __fmul_rn(__fsub_rn(pn[d], scale_re), fRb);
        g1 = g1 < 0.f ? 0.f : g1;
        g1 = g1 > fRb1 ? fRb1 : g1;
        lo[d] = (int)g1;
        float g2 = __fmul_rn(__fadd_rn(pn[d], scale_re), fRb);
        g2 = g2 < 0.f ? 0.f : g2;
        g2 = g2 > fRb1 ? fRb1 : g2;
        hi[d] = (int)g2;
    }
    bool m = false;
    float area = 0.f;
    for (int a = lo[0]; a <= hi[0]; a++) {
        const float ra = fminf(__fmaf_rn((float)a, Rb_re, Rb_re), __fadd_rn(pn[0], scale_re));
        const float la = fmaxf(__fmul_rn((float)a, Rb_re), __fsub_rn(pn[0], scale_re));
        const float oa = __fsub_rn(ra, la);
        for (int b = lo[1]; b <= hi[1]; b++) {
            const float rb = fminf(__fmaf_rn((float)b, Rb_re, Rb_re), __fadd_rn(pn[1], scale_re));
            const float lb = fmaxf(__fmul_rn((float)b, Rb_re), __fsub_rn(pn[1], scale_re));
            const float ob = __fsub_rn(rb, lb);
            if constexpr (D == 2) {
                if (vxl[(size_t)a * Rb + b]) {
                    m = true;
                    area = __fmaf_rn(oa, ob, area);
                }
            } else {
                const uint8_t *row = vxl + ((size_t)a * Rb + b) * Rb;
                const float oab = __fmul_rn(oa, ob);
                for (int cc = lo[D - 1]; cc <= hi[D - 1]; cc++) {
                    if (row[cc]) {
                        const float rc = fminf(__fmaf_rn((float)cc, Rb_re, Rb_re), __fadd_rn(pn[D - 1], scale_re));
                        const float lc = fmaxf(__fmul_rn((float)cc, Rb_re), __fsub_rn(pn[D - 1], scale_re));
                        m = true;
                        area = __fmaf_rn(oab, __fsub_rn(rc, lc), area);
                    }
                }
            }
        }
    }
    area = __fmul_rn(area, fRb);
    area = __fmul_rn(area, fRb);
    if constexpr (D == 3) area = __fmul_rn(area, fRb);
    overlap = (int32_t)__fmul_rn(area, 1000.0f);
    return m;
}

// Candidate occupancy cells of finest-level coordinate c in the reference's vote list (get_idx_coords2,
// utils_bpp_acc.py:498-512: c = occ * t + k + 1, k in [-1, t]): o in [ceil((c - t - 1) / t), floor(c / t)].
__device__ __forceinline__ void vote_cand(uint32_t c, uint32_t t, uint32_t Rb, uint32_t &o_lo, uint32_t &o_hi) {
    o_hi = c / t;
    o_lo = c > t + 1u ? (c - t - 2u) / t + 1u : 0u;     // ceil((c - t - 1) / t)
    if (o_hi > Rb - 1u) o_hi = Rb - 1u;                  // (o_lo > o_hi -> no candidate)
}

// is finest-level vertex c in that list (and inside the border the vote kernels skip, gridencoder.cu:895-898)?
__device__ __forceinline__ bool vote_member(const uint32_t (&c)[3], uint32_t res, uint32_t Rb, const uint8_t *__restrict__ vxl) {
    const uint32_t t = (res - 2u) / Rb;
    uint32_t lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (c[d] == 0u || c[d] >= res - 1u) return false;
        vote_cand(c[d], t, Rb, lo[d], hi[d]);
    }
    for (uint32_t o0 = lo[0]; o0 <= hi[0]; o0++)
        for (uint32_t o1 = lo[1]; o1 <= hi[1]; o1++)
            for (uint32_t o2 = lo[2]; o2 <= hi[2]; o2++)
                if (vxl[((size_t)o0 * Rb + o1) * Rb + o2]) return true;
    return false;
}

}  // namespace cnc
