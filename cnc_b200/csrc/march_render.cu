// march_render.cu -- occupancy-grid ray marching, packed scans and the volume-rendering tail.
//
// Reference behaviour restated: nerfacc/cuda/csrc/grid.cu:68-318 (traverse_grids_kernel), :320-349
// (ray_aabb_intersect_kernel), include/utils_grid.cuh:11-149 (slab test, DDA set-up and step),
// include/utils_scan.cuh + scan.cu (segmented inclusive/exclusive sum/prod), nerfacc/volrend.py:211-266,
// :314-364, :485-549 (transmittance / weights / accumulate_along_rays).  SURVEY Appendix E.
//
// B200 notes: the march of a ray is a sequential walk over a 2 MiB occupancy grid (L2 resident) -- latency-bound, one
// ray per warp for training-size batches (see traverse_kernel).  The scans and the rendering tail are one warp per ray,
// 32 samples per pass with shuffle scans: where the reference launches exclusive_sum + exp + mul + 3 x index_add_ (plus
// pack_info), `cnc_render_from_density` does it in one pass per ray with no atomics and a fixed summation tree.
// Multiply-adds that nvcc contracts in the reference build are written as explicit __fmaf_rn (this library is
// compiled with --fmad=false), matching oracle/cnc_oracle_march.c.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {
namespace mr {

// utils_grid.cuh:11-57
__device__ __forceinline__ bool aabb_hit(const float *o, const float *d, float near, float far, const float *bb,
                                         float &tmin_o, float &tmax_o) {
    const float inv[3] = {__fdiv_rn(1.0f, d[0]), __fdiv_rn(1.0f, d[1]), __fdiv_rn(1.0f, d[2])};
    float tmin, tmax;
    if (inv[0] >= 0) { tmin = __fmul_rn(__fsub_rn(bb[0], o[0]), inv[0]); tmax = __fmul_rn(__fsub_rn(bb[3], o[0]), inv[0]); }
    else             { tmin = __fmul_rn(__fsub_rn(bb[3], o[0]), inv[0]); tmax = __fmul_rn(__fsub_rn(bb[0], o[0]), inv[0]); }
#pragma unroll
    for (int k = 1; k < 3; k++) {
        float a, b;
        if (inv[k] >= 0) { a = __fmul_rn(__fsub_rn(bb[k], o[k]), inv[k]); b = __fmul_rn(__fsub_rn(bb[3 + k], o[k]), inv[k]); }
        else             { a = __fmul_rn(__fsub_rn(bb[3 + k], o[k]), inv[k]); b = __fmul_rn(__fsub_rn(bb[k], o[k]), inv[k]); }
        if (tmin > b || a > tmax) return false;
        if (a > tmin) tmin = a;
        if (b < tmax) tmax = b;
    }
    if (tmax <= 0) return false;
    tmin_o = fmaxf(tmin, near);
    tmax_o = fminf(tmax, far);
    return true;
}

__global__ void __launch_bounds__(256)
ray_aabb_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d, int64_t n_rays, float near, float far,
                const float *__restrict__ aabbs, int32_t n_aabbs, float miss, float *__restrict__ t_mins,
                float *__restrict__ t_maxs, uint8_t *__restrict__ hits) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rays * n_aabbs) return;
    const int64_t r = t / n_aabbs, a = t % n_aabbs;
    const float o[3] = {rays_o[r * 3], rays_o[r * 3 + 1], rays_o[r * 3 + 2]};
    const float d[3] = {rays_d[r * 3], rays_d[r * 3 + 1], rays_d[r * 3 + 2]};
    const float bb[6] = {aabbs[a * 6], aabbs[a * 6 + 1], aabbs[a * 6 + 2], aabbs[a * 6 + 3], aabbs[a * 6 + 4], aabbs[a * 6 + 5]};
    float lo, hi;
    const bool h = aabb_hit(o, d, near, far, bb, lo, hi);
    t_mins[t] = h ? lo : miss;
    t_maxs[t] = h ? hi : miss;
    hits[t] = h ? 1 : 0;
}

__device__ __forceinline__ float calc_dt(float t, float cone, float dmin, float dmax) {
    return fmaxf(dmin, fminf(__fmul_rn(t, cone), dmax));  // grid.cu:23-28
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct MarchArgs {
    const float *rays_o, *rays_d;
    const uint8_t *rays_mask;  // nullable
    int64_t n_rays;
    int32_t n_grids, rx, ry, rz;
    const uint8_t *binaries;   // [n_grids, rx, ry, rz]
    const float *aabbs;        // [n_grids, 6]
    const uint8_t *hits;       // [n_rays, n_grids]
    const float *t_sorted;     // [n_rays, 2*n_grids]
    const int64_t *t_indices;  // [n_rays, 2*n_grids]
    const float *near_planes, *far_planes;
    float step_size, cone_angle;
    int32_t steps_limit;
    const int64_t *chunk_starts;  // fill pass: output offset per ray; nullptr in the count pass
    int64_t *cnt;                 // samples per ray (count pass)
    float *t_starts, *t_ends;     // fill pass
    int64_t *ray_idx;
    float *terminate;             // nullable: where the march of the ray stopped
};

// The walk of one ray: grid.cu:68-318 with the interval edges folded into (t_start, t_end) per sample.  `emit(j, t0, t1)`
// receives sample j; returns the number of samples and leaves the stopping point of the march in t_end.
//
// WARP = true: the whole warp walks ONE ray, every lane holding the same state.  The DDA (which cells, where the ray leaves
// them) does not depend on the sampling state, so the warp first runs the DDA 32 cells ahead on a copy, lane j keeps cell j
// and ALL 32 occupancy bytes are loaded at once (one memory latency per 32 cells instead of one per cell: the walk is a
// latency chain, 390 cycles per cell when measured one load at a time); a ballot hands every lane the 32 bits, and the walk
// proper replays the same cells with the same arithmetic in the same order -- sample placement stays bit for bit.
template <bool WARP = false, bool CDT = false, class Emit>
__device__ __forceinline__ int64_t march_ray(const MarchArgs &a, int64_t tid, float near, float far, int32_t steps_limit, Emit emit,
                                             float &t_end) {
    const float eps = 1e-6f;
    const float dtc = a.step_size, half = __fmul_rn(a.step_size, 0.5f);
    const float o[3] = {a.rays_o[tid * 3], a.rays_o[tid * 3 + 1], a.rays_o[tid * 3 + 2]};
    const float d[3] = {a.rays_d[tid * 3], a.rays_d[tid * 3 + 1], a.rays_d[tid * 3 + 2]};
    const float inv[3] = {__fdiv_rn(1.0f, d[0]), __fdiv_rn(1.0f, d[1]), __fdiv_rn(1.0f, d[2])};
    int64_t n_samples = 0;
    float t_last = near;
    bool continuous = false;
    const int64_t bh = tid * a.n_grids, bt = tid * a.n_grids * 2;
    const int ires[3] = {a.rx, a.ry, a.rz};
    for (int64_t i = bt; i < bt + a.n_grids * 2 - 1; i++) {
        const bool entering = a.t_indices[i] < a.n_grids;
        int64_t level = a.t_indices[i] % a.n_grids;
        if (!a.hits[bh + level]) continue;
        if (!entering) {
            if (a.t_indices[i + 1] < a.n_grids) continue;
            level = a.t_indices[i + 1] % a.n_grids;
            if (!a.hits[bh + level]) continue;
        }
        const float this_tmin = fmaxf(a.t_sorted[i], near);
        const float this_tmax = fminf(a.t_sorted[i + 1], far);
        if (this_tmin >= this_tmax) continue;
        const bool first_skip = !continuous;   // (the skip itself follows the traversal setup: it only touches t_last)
        const float *bb = a.aabbs + level * 6;
        // setup_traversal, utils_grid.cuh:59-118
        float tdist[3], delta[3];
        int step[3], cur[3], over[3];
        const float ts = __fadd_rn(this_tmin, eps), te = __fsub_rn(this_tmax, eps);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float res = (float)ires[k];
            const float ext = __fsub_rn(bb[3 + k], bb[k]);
            const float vox = __fdiv_rn(ext, res);
            const float rs = __fmaf_rn(d[k], ts, o[k]), re = __fmaf_rn(d[k], te, o[k]);
            cur[k] = clampi((int)__fmul_rn(__fdiv_rn(__fsub_rn(rs, bb[k]), ext), res), 0, ires[k] - 1);
            const int fin = clampi((int)__fmul_rn(__fdiv_rn(__fsub_rn(re, bb[k]), ext), res), 0, ires[k] - 1);
            const int si = cur[k] + (d[k] > 0 ? 1 : 0);
            const float txyz = __fmaf_rn(__fadd_rn(bb[k], __fmaf_rn((float)si, vox, -rs)), inv[k], this_tmin);
            tdist[k] = (d[k] == 0.0f) ? this_tmax : txyz;
            const float sf = (d[k] == 0.0f) ? 0.0f : (d[k] > 0.0f ? 1.0f : -1.0f);
            step[k] = (int)sf;
            delta[k] = (d[k] == 0.0f) ? this_tmax : __fmul_rn(__fmul_rn(vox, inv[k]), sf);
            over[k] = fin + step[k];
        }
        // the two ways t_last moves (grid.cu:225-290): through an empty cell / up to a segment start without samples, and
        // through an occupied cell with one sample per step.  CDT (cone_angle == 0): dt is the step size, and
        // fma(dt, 0.5, t) == t + dt/2 (the product is exact), so a step is two additions.
        auto skip_to = [&](float target) {
            if (a.step_size <= 0.0f) { t_last = target; return; }
            if (CDT) {
                for (;;) {   // four steps per round trip: the additions are the chain, the comparisons hang off it
                    const float t1 = __fadd_rn(t_last, dtc), t2 = __fadd_rn(t1, dtc), t3 = __fadd_rn(t2, dtc), t4 = __fadd_rn(t3, dtc);
                    if (__fadd_rn(t_last, half) >= target) break;
                    if (__fadd_rn(t1, half) >= target) { t_last = t1; break; }
                    if (__fadd_rn(t2, half) >= target) { t_last = t2; break; }
                    if (__fadd_rn(t3, half) >= target) { t_last = t3; break; }
                    t_last = t4;
                }
            } else {
                const float dt = calc_dt(t_last, a.cone_angle, a.step_size, 1e10f);
                while (!(__fmaf_rn(dt, 0.5f, t_last) >= target)) t_last = __fadd_rn(t_last, dt);
            }
        };
        auto sample_to = [&](float t_trav) {
            while (steps_limit <= 0 || n_samples < steps_limit) {
                float t_next;
                if (a.step_size <= 0.0f) t_next = t_trav;
                else if (CDT) {
                    if (__fadd_rn(t_last, half) >= t_trav) break;
                    t_next = __fadd_rn(t_last, dtc);
                } else {
                    const float dt = calc_dt(t_last, a.cone_angle, a.step_size, 1e10f);
                    if (__fmaf_rn(dt, 0.5f, t_last) >= t_trav) break;
                    t_next = __fadd_rn(t_last, dt);
                }
                emit(n_samples, t_last, t_next);
                n_samples++;
                continuous = true;
                t_last = t_next;
                if (t_next >= t_trav) break;
            }
        };
        if (first_skip) skip_to(this_tmin);
        if (WARP) {
            // The DDA does not depend on the sampling state: the warp runs it 32 cells ahead (every lane the same arithmetic,
            // lane j keeps cell j and the time the ray leaves it), loads the 32 occupancy bytes at once, and then walks the
            // 32 cells against a ballot and shuffles -- same cells, same exit times, same order as one cell at a time.
            const uint32_t lane = threadIdx.x & 31u;
            const int64_t sx = (int64_t)a.ry * a.rz, sy = a.rz;
            const int64_t cstep[3] = {step[0] * sx, step[1] * sy, (int64_t)step[2]};
            int64_t cell = cur[0] * sx + cur[1] * sy + cur[2] + level * (int64_t)a.rx * sx;
            bool seg_done = false, stop = false;
            while (!seg_done && !stop) {
                float my_tt = 0.f;
                int64_t my_cell = cell;
                int n_ahead = 0;
                for (int j = 0; j < 32; j++) {
                    const float tt = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), this_tmax);
                    if (lane == (uint32_t)j) { my_tt = tt; my_cell = cell; }
                    n_ahead = j + 1;
                    const int ax = ((tdist[0] < tdist[1]) && (tdist[0] < tdist[2])) ? 0 : ((tdist[1] < tdist[2]) ? 1 : 2);
                    bool fin = false;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (k == ax) {
                            cur[k] += step[k];
                            tdist[k] = __fadd_rn(tdist[k], delta[k]);
                            cell += cstep[k];
                            fin = cur[k] == over[k];
                        }
                    }
                    if (fin) { seg_done = true; break; }
                }
                const uint32_t occ = __ballot_sync(0xffffffffu, lane < (uint32_t)n_ahead && a.binaries[my_cell] != 0);
                for (int j = 0; j < n_ahead; j++) {
                    if (!(steps_limit <= 0 || n_samples < steps_limit)) { stop = true; break; }
                    const float t_trav = __shfl_sync(0xffffffffu, my_tt, j);
                    if ((occ >> j) & 1u) sample_to(t_trav);
                    else { skip_to(t_trav); continuous = false; }
                }
            }
        } else {
            while (steps_limit <= 0 || n_samples < steps_limit) {
                const float t_trav = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), this_tmax);
                const int64_t cell = (int64_t)cur[0] * a.ry * a.rz + (int64_t)cur[1] * a.rz + cur[2] + level * (int64_t)a.rx * a.ry * a.rz;
                if (a.binaries[cell]) sample_to(t_trav);
                else { skip_to(t_trav); continuous = false; }
                // single_traversal, utils_grid.cuh:121-149
                const int ax = ((tdist[0] < tdist[1]) && (tdist[0] < tdist[2])) ? 0 : ((tdist[1] < tdist[2]) ? 1 : 2);
                bool done = false;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (k == ax) {
                        cur[k] += step[k];
                        tdist[k] = __fadd_rn(tdist[k], delta[k]);
                        done = cur[k] == over[k];
                    }
                }
                if (done) break;
            }
        }
    }
    t_end = t_last;
    return n_samples;
}

// RAY_PER_WARP: lane 0 of every warp walks one ray and the other lanes retire at once -- the walk is one sequential chain
// (t += dt in fp32, DDA boundaries tdist += delta: both defined by their rounding sequence) with nested data-dependent
// loops, and with 32 rays per warp the warp executes the UNION of 32 different loop nests (measured: 292 us for 1100
// rays, ten times one ray's chain); a training batch has a few thousand rays, far fewer than the machine has warp slots.
// Large batches (test-time wavefronts of 10^5..10^6 rays with a step limit) keep one thread per ray.
template <bool RAY_PER_WARP, bool CDT = false>
__global__ void __launch_bounds__(128) traverse_kernel(const MarchArgs a) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    if (RAY_PER_WARP) tid >>= 5;
    if (tid >= a.n_rays) return;
    if (a.rays_mask && !a.rays_mask[tid]) {
        if (a.cnt && (!RAY_PER_WARP || lane == 0)) a.cnt[tid] = 0;
        return;
    }
    const bool fill = a.t_starts != nullptr;
    const int64_t start = a.chunk_starts ? a.chunk_starts[tid] : 0;
    float t_last;
    int64_t n_samples;
    if (RAY_PER_WARP) {
        // every lane sees every sample; lane (j % 32) keeps sample j, and 32 of them leave as one coalesced store
        float k0 = 0.f, k1 = 0.f;
        n_samples = march_ray<true, CDT>(a, tid, a.near_planes[tid], a.far_planes[tid], a.steps_limit,
                                    [&](int64_t j, float t0, float t1) {
                                        if (fill) {
                                            if ((uint32_t)(j & 31) == lane) { k0 = t0; k1 = t1; }
                                            if ((j & 31) == 31) {
                                                const int64_t at = start + j - 31 + lane;
                                                a.t_starts[at] = k0;
                                                a.t_ends[at] = k1;
                                                if (a.ray_idx) a.ray_idx[at] = tid;
                                            }
                                        }
                                    }, t_last);
        if (fill && lane < (uint32_t)(n_samples & 31)) {
            const int64_t at = start + (n_samples & ~(int64_t)31) + lane;
            a.t_starts[at] = k0;
            a.t_ends[at] = k1;
            if (a.ray_idx) a.ray_idx[at] = tid;
        }
        if (lane) return;
    } else {
        n_samples = march_ray(a, tid, a.near_planes[tid], a.far_planes[tid], a.steps_limit,
                              [&](int64_t j, float t0, float t1) {
                                  if (fill) {
                                      a.t_starts[start + j] = t0;
                                      a.t_ends[start + j] = t1;
                                      if (a.ray_idx) a.ray_idx[start + j] = tid;
                                  }
                              }, t_last);
    }
    if (a.terminate) a.terminate[tid] = t_last;
    if (a.cnt) a.cnt[tid] = n_samples;
}

__device__ __forceinline__ float warp_incl_scan(float v, int op, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v = op ? __fmul_rn(o, v) : __fadd_rn(o, v);
    }
    return v;
}

// op: 0 sum, 1 prod.  One warp per ray: 32 consecutive samples per pass (coalesced), a shuffle scan inside the pass and a
// running carry between passes -- a fixed summation tree per ray (deterministic; the reference's smem tree, 32 elements per
// tile as well, differs in the last bits: utils_scan.cuh:153-245).
__global__ void __launch_bounds__(128)
packed_scan_kernel(const float *__restrict__ in, const int64_t *__restrict__ packed, int64_t n_rays, float *__restrict__ out,
                   int op, int inclusive, int reverse) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (r >= n_rays) return;
    const int64_t s = packed[r * 2], c = packed[r * 2 + 1];
    float carry = op ? 1.0f : 0.0f;
    for (int64_t j0 = 0; j0 < c; j0 += 32) {
        const int64_t j = j0 + lane;
        const bool ok = j < c;
        const int64_t i = reverse ? s + c - 1 - j : s + j;
        const float v = ok ? in[i] : (op ? 1.0f : 0.0f);
        float inc = warp_incl_scan(v, op, lane);
        inc = op ? __fmul_rn(carry, inc) : __fadd_rn(carry, inc);
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = carry;
        if (ok) out[i] = inclusive ? inc : exc;
        carry = __shfl_sync(0xffffffffu, inc, 31);
    }
}

// Backward of render_density_kernel (+ the three accumulations) in one pass per ray, walking the ray from its far end:
//   gw_i  = gC_r . c_i + gO_r + gD_r * mid_i (+ the gradient that arrives at the weights themselves)
//   gc_i  = w_i * gC_r
//   gs_i  = (gw_i * T_i * (1 - a_i) - sum_{j > i} gw_j * a_j * T_j) * (t1_i - t0_i)
// (w = T a, T_i = exp(-sum_{j<i} s_j d_j), a_i = 1 - exp(-s_i d_i): d w_j / d s_i = -w_j d_i for j > i, T_i (1 - a_i) d_i for j = i).
__global__ void __launch_bounds__(128)
render_bwd_kernel(const float *__restrict__ t0, const float *__restrict__ t1, const float *__restrict__ trans,
                  const float *__restrict__ alphas, const float *__restrict__ weights, const float *__restrict__ rgb,
                  const int64_t *__restrict__ packed, int64_t n_rays, const float *__restrict__ g_colors,
                  const float *__restrict__ g_opac, const float *__restrict__ g_depth, const float *__restrict__ g_weights,
                  float *__restrict__ g_sigma, float *__restrict__ g_rgb) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (r >= n_rays) return;
    const int64_t s = packed[r * 2], c = packed[r * 2 + 1];
    const float gr = g_colors ? g_colors[r * 3] : 0.f, gg = g_colors ? g_colors[r * 3 + 1] : 0.f, gb = g_colors ? g_colors[r * 3 + 2] : 0.f;
    const float go = g_opac ? g_opac[r] : 0.f, gd = g_depth ? g_depth[r] : 0.f;
    float carry = 0.f;
    for (int64_t j0 = 0; j0 < c; j0 += 32) {
        const int64_t j = j0 + lane;
        const bool ok = j < c;
        const int64_t i = s + c - 1 - j;
        float a0 = 0.f, a1 = 0.f, T = 0.f, al = 0.f, gw = 0.f;
        if (ok) {
            a0 = t0[i];
            a1 = t1[i];
            T = trans[i];
            al = alphas[i];
            const float mid = __fmul_rn(__fadd_rn(a0, a1), 0.5f);
            const float cr = rgb[i * 3], cg = rgb[i * 3 + 1], cb = rgb[i * 3 + 2];
            gw = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(gr, cr), __fmul_rn(gg, cg)), __fmul_rn(gb, cb)), go), __fmul_rn(gd, mid));
            if (g_weights) gw = __fadd_rn(gw, g_weights[i]);
            const float w = weights[i];
            g_rgb[i * 3] = __fmul_rn(w, gr);
            g_rgb[i * 3 + 1] = __fmul_rn(w, gg);
            g_rgb[i * 3 + 2] = __fmul_rn(w, gb);
        }
        const float gt = ok ? __fmul_rn(__fmul_rn(gw, al), T) : 0.f;
        const float inc = __fadd_rn(carry, warp_incl_scan(gt, 0, lane));
        float exc = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exc = carry;
        carry = __shfl_sync(0xffffffffu, inc, 31);
        if (ok) g_sigma[i] = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn(gw, T), __fsub_rn(1.0f, al)), exc), __fsub_rn(a1, a0));
    }
}

// ray samples -> query points: positions = o[r] + d[r] * (t0 + t1) / 2 and the direction of the ray, per sample
// (examples/utils.py:250-262: the rgb_sigma_fn / sigma_fn closures of the training and test renderers)
__global__ void __launch_bounds__(256)
sample_points_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const int64_t *__restrict__ ray_idx,
                     const float *__restrict__ t0, const float *__restrict__ t1, int64_t n, float *__restrict__ pos,
                     float *__restrict__ dirs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = ray_idx[i];
    const float h = __fadd_rn(t0[i], t1[i]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float d = rays_d[r * 3 + k];
        pos[i * 3 + k] = __fadd_rn(rays_o[r * 3 + k], __fdiv_rn(__fmul_rn(d, h), 2.0f));
        if (dirs) dirs[i * 3 + k] = d;
    }
}

// one walk instead of two: the march wrote ray r's samples at scratch[r * cap ..] and its count; this moves them to their
// packed place (start_r = exclusive prefix sum of the counts) and writes the ray index.  One warp per ray, coalesced.
__global__ void __launch_bounds__(128)
pack_ray_chunks_kernel(const float *__restrict__ s0, const float *__restrict__ s1, int64_t cap, const int64_t *__restrict__ packed,
                       int64_t n_rays, float *__restrict__ o0, float *__restrict__ o1, int64_t *__restrict__ oi) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (r >= n_rays) return;
    const int64_t start = packed[r * 2], c = packed[r * 2 + 1];
    for (int64_t j = lane; j < c; j += 32) {
        o0[start + j] = s0[r * cap + j];
        o1[start + j] = s1[r * cap + j];
        oi[start + j] = r;
    }
}

// keep[i] != 0 samples move to the front, order kept: out index = exclusive count of kept samples before i (`rank`, an
// int64 inclusive prefix sum of keep computed by the caller, so rank[i] - 1 is the slot of a kept sample)
__global__ void __launch_bounds__(256)
compact_samples_kernel(const uint8_t *__restrict__ keep, const int64_t *__restrict__ rank, int64_t n, const float *__restrict__ t0,
                       const float *__restrict__ t1, const int64_t *__restrict__ ray_idx, float *__restrict__ o0,
                       float *__restrict__ o1, int64_t *__restrict__ oi, const int64_t *__restrict__ packed, int64_t n_rays,
                       int64_t *__restrict__ packed_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (packed_out && i < n_rays) {   // (start, count) of the ray's kept samples: pack_info of the compacted ray_indices
        const int64_t s = packed[i * 2], c = packed[i * 2 + 1];
        const int64_t before = s > 0 ? rank[(s < n ? s : n) - 1] : 0;
        packed_out[i * 2] = before;
        packed_out[i * 2 + 1] = (c > 0 ? rank[s + c - 1] : before) - before;
    }
    if (i >= n || !keep[i]) return;
    const int64_t k = rank[i] - 1;
    o0[k] = t0[i];
    o1[k] = t1[i];
    oi[k] = ray_idx[i];
}

// volrend.py:211-266 + :314-364 + :485-549 in one pass per ray, one warp per ray (32 samples per pass)
__global__ void __launch_bounds__(128)
render_density_kernel(const float *__restrict__ t0, const float *__restrict__ t1, const float *__restrict__ sigma,
                      const float *__restrict__ rgb, const int64_t *__restrict__ packed, int64_t n_rays,
                      const float *__restrict__ prefix_trans, float *__restrict__ weights, float *__restrict__ trans,
                      float *__restrict__ alphas, float *__restrict__ colors, float *__restrict__ opac,
                      float *__restrict__ depth) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    if (r >= n_rays) return;
    const int64_t s = packed[r * 2], c = packed[r * 2 + 1];
    float carry = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, op = 0.f, dp = 0.f;
    for (int64_t j0 = 0; j0 < c; j0 += 32) {
        const int64_t i = s + j0 + lane;
        const bool ok = j0 + lane < c;
        float a0 = 0.f, a1 = 0.f, sd = 0.f;
        if (ok) {
            a0 = t0[i];
            a1 = t1[i];
            sd = __fmul_rn(sigma[i], __fsub_rn(a1, a0));
        }
        const float inc = __fadd_rn(carry, warp_incl_scan(sd, 0, lane));
        float acc = __shfl_up_sync(0xffffffffu, inc, 1);     // exclusive sum of sigma * delta along the ray
        if (lane == 0) acc = carry;
        carry = __shfl_sync(0xffffffffu, inc, 31);
        if (ok) {
            const float al = __fsub_rn(1.0f, expf(-sd));
            float T = expf(-acc);
            if (prefix_trans) T = __fmul_rn(T, prefix_trans[i]);
            const float w = __fmul_rn(T, al);
            if (weights) weights[i] = w;
            if (trans) trans[i] = T;
            if (alphas) alphas[i] = al;
            if (rgb) {
                cr = __fadd_rn(cr, __fmul_rn(w, rgb[i * 3 + 0]));   // accumulate_along_rays: src = weights * values, then add
                cg = __fadd_rn(cg, __fmul_rn(w, rgb[i * 3 + 1]));
                cb = __fadd_rn(cb, __fmul_rn(w, rgb[i * 3 + 2]));
            }
            op = __fadd_rn(op, w);
            dp = __fadd_rn(dp, __fmul_rn(w, __fmul_rn(__fadd_rn(a0, a1), 0.5f)));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cr = __fadd_rn(cr, __shfl_xor_sync(0xffffffffu, cr, d));
        cg = __fadd_rn(cg, __shfl_xor_sync(0xffffffffu, cg, d));
        cb = __fadd_rn(cb, __shfl_xor_sync(0xffffffffu, cb, d));
        op = __fadd_rn(op, __shfl_xor_sync(0xffffffffu, op, d));
        dp = __fadd_rn(dp, __shfl_xor_sync(0xffffffffu, dp, d));
    }
    if (lane == 0) {
        if (colors) { colors[r * 3 + 0] = cr; colors[r * 3 + 1] = cg; colors[r * 3 + 2] = cb; }
        if (opac) opac[r] = op;
        if (depth) depth[r] = dp;
    }
}

// ------------------------------------------------------------------------------------------
// Test-time wavefront renderer without host round trips (examples/utils.py:395-479).  One round = wf_begin -> wf_march ->
// field forward (cnc_field_fwd_n, sample count read from `st`) -> wf_composite; everything the reference's python loop
// decides on the host per round (number of live rays -> samples per ray of the round, the stop conditions) lives in `st`.
// st[0] live rays of this round   st[1] samples per ray of this round   st[2] iter_samples   st[3] done
// st[4] samples marched this round (atomic cursor)   st[5] live rays after this round   st[6] rounds   st[8..9] total samples (u64)
// ------------------------------------------------------------------------------------------
__global__ void wf_begin_kernel(uint32_t *__restrict__ st, uint32_t n_rays, uint32_t min_samples, uint32_t max_samples) {
    st[4] = 0;
    if (st[3]) return;
    const uint32_t alive = st[5];
    if (alive == 0 || st[2] >= max_samples) {   // utils.py:395-399
        st[3] = 1;
        return;
    }
    uint32_t n = n_rays / alive;
    n = n < 64u ? n : 64u;
    n = n > min_samples ? n : min_samples;       // utils.py:402
    st[0] = alive;
    st[1] = n;
    st[2] += n;
    st[5] = 0;
    st[6] += 1;
}

struct WaveArgs {
    MarchArgs m;               // rays, grids, precomputed intersections; near_planes = the running planes (updated in place)
    uint32_t *st;
    uint8_t *ray_mask;         // [n_rays] live flags (updated by wf_composite)
    float *near_planes;        // [n_rays]
    uint32_t capacity;         // sample slots of the round buffers
    // per-ray outputs of the march
    uint32_t *ray_base, *ray_cnt;
    float *ray_term;
    // per-sample outputs of the march = inputs of the field kernel and of the compositing
    float *t0, *t1, *pos, *dir;
};

// one thread per ray: count the samples of the round, reserve their slots (warp-aggregated atomic), march again writing
// them.  The walk itself is march_ray, i.e. the sample placement of traverse_grids bit for bit.
template <bool CDT>
__global__ void __launch_bounds__(128) wf_march_kernel(const WaveArgs w) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    if (w.st[3]) return;
    const int32_t limit = (int32_t)w.st[1];
    const bool live = tid < w.m.n_rays && w.ray_mask[tid];
    float near = 0.f, far = 0.f, t_end = 0.f;
    uint32_t cnt = 0;
    // the round's samples of this ray (the schedule hands out at most 64 per ray and round): kept from the one walk, so that
    // the slot assignment below does not have to be followed by a second walk
    constexpr int KEEP = 64;
    float k0[KEEP], k1[KEEP];
    const bool keep = limit > 0 && limit <= KEEP;
    if (live) {
        near = w.near_planes[tid];
        far = w.m.far_planes[tid];
        cnt = (uint32_t)march_ray<false, CDT>(w.m, tid, near, far, limit, [&](int64_t j, float ta, float tb) {
            if (keep) { k0[j] = ta; k1[j] = tb; }
        }, t_end);
    }
    // slots: exclusive prefix inside the warp + one atomic per warp
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(w.st + 4, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
    if (!live) return;
    w.ray_base[tid] = base;
    w.ray_cnt[tid] = cnt;
    w.ray_term[tid] = t_end;
    if (cnt == 0 || base + cnt > w.capacity) return;   // (capacity is sized for the schedule: n_live * n <= n_rays * min_samples)
    const float o[3] = {w.m.rays_o[tid * 3], w.m.rays_o[tid * 3 + 1], w.m.rays_o[tid * 3 + 2]};
    const float d[3] = {w.m.rays_d[tid * 3], w.m.rays_d[tid * 3 + 1], w.m.rays_d[tid * 3 + 2]};
    auto put = [&](int64_t j, float ta, float tb) {
        const uint32_t k = base + (uint32_t)j;
        w.t0[k] = ta;
        w.t1[k] = tb;
        // positions = origins + dirs * (t_starts + t_ends) / 2 (utils.py:350-352): add, multiply by the direction, divide, add
        const float ts = __fadd_rn(ta, tb);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            w.pos[k * 3 + c] = __fadd_rn(o[c], __fdiv_rn(__fmul_rn(d[c], ts), 2.0f));
            w.dir[k * 3 + c] = d[c];
        }
    };
    if (keep) {
        for (uint32_t j = 0; j < cnt; j++) put(j, k0[j], k1[j]);
    } else {
        float dummy;
        march_ray<false, CDT>(w.m, tid, near, far, limit, put, dummy);
    }
}

struct CompArgs {
    uint32_t *st;
    uint8_t *ray_mask;
    float *near_planes;
    const uint32_t *ray_base, *ray_cnt;
    const float *ray_term, *t0, *t1, *sigma, *rgbs;
    float *rgb, *opacity, *depth;   // [n_rays,3], [n_rays], [n_rays] accumulators
    int64_t n_rays;
    uint32_t capacity;
    float alpha_thre, opc_thre;
};

// one thread per ray: weights of the round's samples with the ray's running transmittance as prefix
// (render_weight_from_density(prefix_trans=1-opacity), utils.py:436-443), accumulation into the image buffers
// (accumulate_along_rays_ x3, :454-471), new near plane, new live flag (:473-478)
__global__ void __launch_bounds__(128) wf_composite_kernel(const CompArgs a) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a.st[3]) return;
    const bool live = tid < a.n_rays && a.ray_mask[tid];
    bool stay = false;
    uint32_t cnt = 0, n_vis = 0;
    if (live) {
        cnt = a.ray_cnt[tid];
        const uint32_t base = a.ray_base[tid];
        const float op0 = a.opacity[tid];
        const float prefix = __fsub_rn(1.0f, op0);
        float acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, op = 0.f, dp = 0.f;
        if (base + cnt <= a.capacity) {
            for (uint32_t j = 0; j < cnt; j++) {
                const uint32_t i = base + j;
                const float ta = a.t0[i], tb = a.t1[i];
                const float sd = __fmul_rn(a.sigma[i], __fsub_rn(tb, ta));
                const float al = __fsub_rn(1.0f, expf(-sd));
                const float w = __fmul_rn(__fmul_rn(expf(-acc), prefix), al);
                acc = __fadd_rn(acc, sd);
                if (a.alpha_thre > 0.f && !(al >= a.alpha_thre)) continue;   // vis_mask (utils.py:444-452)
                n_vis++;
                cr = __fadd_rn(cr, __fmul_rn(w, a.rgbs[i * 3 + 0]));
                cg = __fadd_rn(cg, __fmul_rn(w, a.rgbs[i * 3 + 1]));
                cb = __fadd_rn(cb, __fmul_rn(w, a.rgbs[i * 3 + 2]));
                op = __fadd_rn(op, w);
                dp = __fadd_rn(dp, __fmul_rn(w, __fdiv_rn(__fadd_rn(ta, tb), 2.0f)));
            }
        }
        a.rgb[tid * 3 + 0] = __fadd_rn(a.rgb[tid * 3 + 0], cr);
        a.rgb[tid * 3 + 1] = __fadd_rn(a.rgb[tid * 3 + 1], cg);
        a.rgb[tid * 3 + 2] = __fadd_rn(a.rgb[tid * 3 + 2], cb);
        const float op1 = __fadd_rn(op0, op);
        a.opacity[tid] = op1;
        a.depth[tid] = __fadd_rn(a.depth[tid], dp);
        a.near_planes[tid] = a.ray_term[tid];
        stay = (op1 <= a.opc_thre) && (cnt == a.st[1]);
        a.ray_mask[tid] = stay ? 1 : 0;
    }
    const uint32_t n_stay = __popc(__ballot_sync(0xffffffffu, stay));
    uint32_t c = n_vis;                      // total_samples counts what is accumulated (utils.py:479)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31u) == 0) {
        if (n_stay) atomicAdd(a.st + 5, n_stay);
        if (c) atomicAdd(reinterpret_cast<unsigned long long *>(a.st + 8), (unsigned long long)c);
    }
}

}  // namespace mr
}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_ray_aabb_intersect(const float *rays_o, const float *rays_d, int64_t n_rays, float near_plane, float far_plane,
                           const float *aabbs, int32_t n_aabbs, float miss_value, float *t_mins, float *t_maxs,
                           uint8_t *hits, cnc_stream_t stream) {
    if (n_rays * n_aabbs == 0) return CNC_OK;
    if (!rays_o || !rays_d || !aabbs || !t_mins || !t_maxs || !hits) { set_error("ray_aabb_intersect: null pointer"); return CNC_EINVAL; }
    mr::ray_aabb_kernel<<<div_up((uint64_t)(n_rays * n_aabbs), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rays_o, rays_d, n_rays, near_plane, far_plane, aabbs, n_aabbs, miss_value, t_mins, t_maxs, hits);
    return check_launch("ray_aabb_intersect");
}

int cnc_traverse_grids(const float *rays_o, const float *rays_d, const uint8_t *rays_mask, int64_t n_rays, int32_t n_grids,
                       int32_t rx, int32_t ry, int32_t rz, const uint8_t *binaries, const float *aabbs, const uint8_t *hits,
                       const float *t_sorted, const int64_t *t_indices, const float *near_planes, const float *far_planes,
                       float step_size, float cone_angle, int32_t steps_limit, const int64_t *chunk_starts, int64_t *cnt,
                       float *t_starts, float *t_ends, int64_t *ray_indices, float *terminate_planes, cnc_stream_t stream) {
    if (n_rays == 0) return CNC_OK;
    if (!rays_o || !rays_d || !binaries || !aabbs || !hits || !t_sorted || !t_indices || !near_planes || !far_planes) {
        set_error("traverse_grids: null pointer");
        return CNC_EINVAL;
    }
    if ((t_starts != nullptr) != (chunk_starts != nullptr) || (t_starts && !t_ends) || (!t_starts && !cnt)) {
        set_error("traverse_grids: count pass needs cnt, fill pass needs chunk_starts + t_starts + t_ends (+ ray_indices)");
        return CNC_EINVAL;
    }
    mr::MarchArgs a{rays_o, rays_d, rays_mask, n_rays, n_grids, rx, ry, rz, binaries, aabbs, hits, t_sorted, t_indices,
                    near_planes, far_planes, step_size, cone_angle, steps_limit, chunk_starts, cnt, t_starts, t_ends,
                    ray_indices, terminate_planes};
    if (n_rays <= 148 * 64 * 2) {   // at most two waves of one-ray warps
        if (cone_angle == 0.0f && step_size > 0.0f)
            mr::traverse_kernel<true, true><<<div_up((uint64_t)n_rays * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
        else
            mr::traverse_kernel<true, false><<<div_up((uint64_t)n_rays * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    } else
        mr::traverse_kernel<false><<<div_up((uint64_t)n_rays, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("traverse_grids");
}

int cnc_wavefront_begin(uint32_t *state, uint32_t n_rays, uint32_t min_samples, uint32_t max_samples, cnc_stream_t stream) {
    if (!state || n_rays == 0) { set_error("wavefront_begin: bad argument"); return CNC_EINVAL; }
    mr::wf_begin_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(state, n_rays, min_samples < 1 ? 1 : min_samples, max_samples);
    return check_launch("wavefront_begin");
}

int cnc_wavefront_march(const float *rays_o, const float *rays_d, int64_t n_rays, int32_t n_grids, int32_t rx, int32_t ry, int32_t rz,
                        const uint8_t *binaries, const float *aabbs, const uint8_t *hits, const float *t_sorted,
                        const int64_t *t_indices, const float *far_planes, float step_size, float cone_angle, uint32_t *state,
                        uint8_t *ray_mask, float *near_planes, uint32_t capacity, uint32_t *ray_base, uint32_t *ray_cnt,
                        float *ray_term, float *t0, float *t1, float *pos, float *dirs, cnc_stream_t stream) {
    if (n_rays == 0) return CNC_OK;
    if (!rays_o || !rays_d || !binaries || !aabbs || !hits || !t_sorted || !t_indices || !far_planes || !state || !ray_mask ||
        !near_planes || !ray_base || !ray_cnt || !ray_term || !t0 || !t1 || !pos || !dirs) {
        set_error("wavefront_march: null pointer");
        return CNC_EINVAL;
    }
    mr::WaveArgs w;
    w.m = mr::MarchArgs{rays_o, rays_d, nullptr, n_rays, n_grids, rx, ry, rz, binaries, aabbs, hits, t_sorted, t_indices,
                        near_planes, far_planes, step_size, cone_angle, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    w.st = state; w.ray_mask = ray_mask; w.near_planes = near_planes; w.capacity = capacity;
    w.ray_base = ray_base; w.ray_cnt = ray_cnt; w.ray_term = ray_term; w.t0 = t0; w.t1 = t1; w.pos = pos; w.dir = dirs;
    if (cone_angle == 0.0f && step_size > 0.0f)
        mr::wf_march_kernel<true><<<div_up((uint64_t)n_rays, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(w);
    else
        mr::wf_march_kernel<false><<<div_up((uint64_t)n_rays, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(w);
    return check_launch("wavefront_march");
}

int cnc_wavefront_composite(uint32_t *state, uint8_t *ray_mask, float *near_planes, const uint32_t *ray_base, const uint32_t *ray_cnt,
                            const float *ray_term, const float *t0, const float *t1, const float *sigma, const float *rgbs,
                            float *rgb, float *opacity, float *depth, int64_t n_rays, uint32_t capacity, float alpha_thre,
                            float opc_thre, cnc_stream_t stream) {
    if (n_rays == 0) return CNC_OK;
    if (!state || !ray_mask || !near_planes || !ray_base || !ray_cnt || !ray_term || !t0 || !t1 || !sigma || !rgbs || !rgb ||
        !opacity || !depth) {
        set_error("wavefront_composite: null pointer");
        return CNC_EINVAL;
    }
    mr::CompArgs a{state, ray_mask, near_planes, ray_base, ray_cnt, ray_term, t0, t1, sigma, rgbs, rgb, opacity, depth, n_rays,
                   capacity, alpha_thre, opc_thre};
    mr::wf_composite_kernel<<<div_up((uint64_t)n_rays, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("wavefront_composite");
}

int cnc_packed_scan(const float *in, const int64_t *packed_info, int64_t n_rays, float *out, int32_t op, int32_t inclusive,
                    int32_t reverse, cnc_stream_t stream) {
    if (n_rays == 0) return CNC_OK;
    if (!in || !packed_info || !out) { set_error("packed_scan: null pointer"); return CNC_EINVAL; }
    mr::packed_scan_kernel<<<div_up((uint64_t)n_rays * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(in, packed_info, n_rays, out, op, inclusive, reverse);
    return check_launch("packed_scan");
}

int cnc_render_bwd(const float *t_starts, const float *t_ends, const float *trans, const float *alphas, const float *weights,
                   const float *rgbs, const int64_t *packed_info, int64_t n_rays, const float *g_colors, const float *g_opacities,
                   const float *g_depths, const float *g_weights, float *g_sigmas, float *g_rgbs, cnc_stream_t stream) {
    if (n_rays == 0) return CNC_OK;
    if (!t_starts || !t_ends || !trans || !alphas || !weights || !rgbs || !packed_info || !g_sigmas || !g_rgbs) {
        set_error("render_bwd: null pointer");
        return CNC_EINVAL;
    }
    mr::render_bwd_kernel<<<div_up((uint64_t)n_rays * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        t_starts, t_ends, trans, alphas, weights, rgbs, packed_info, n_rays, g_colors, g_opacities, g_depths, g_weights, g_sigmas, g_rgbs);
    return check_launch("render_bwd");
}

int cnc_sample_points(const float *rays_o, const float *rays_d, const int64_t *ray_indices, const float *t_starts, const float *t_ends,
                      int64_t n, float *positions, float *dirs, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!rays_o || !rays_d || !ray_indices || !t_starts || !t_ends || !positions) { set_error("sample_points: null pointer"); return CNC_EINVAL; }
    mr::sample_points_kernel<<<div_up((uint64_t)n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rays_o, rays_d, ray_indices, t_starts,
                                                                                                  t_ends, n, positions, dirs);
    return check_launch("sample_points");
}

int cnc_pack_ray_chunks(const float *scratch_t_starts, const float *scratch_t_ends, int64_t cap, const int64_t *packed_info,
                        int64_t n_rays, float *t_starts, float *t_ends, int64_t *ray_indices, cnc_stream_t stream) {
    if (n_rays == 0) return CNC_OK;
    if (!scratch_t_starts || !scratch_t_ends || !packed_info || !t_starts || !t_ends || !ray_indices || cap <= 0) {
        set_error("pack_ray_chunks: bad argument");
        return CNC_EINVAL;
    }
    mr::pack_ray_chunks_kernel<<<div_up((uint64_t)n_rays * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        scratch_t_starts, scratch_t_ends, cap, packed_info, n_rays, t_starts, t_ends, ray_indices);
    return check_launch("pack_ray_chunks");
}

int cnc_compact_samples(const uint8_t *keep, const int64_t *rank, int64_t n, const float *t_starts, const float *t_ends,
                        const int64_t *ray_indices, float *out_t_starts, float *out_t_ends, int64_t *out_ray_indices,
                        const int64_t *packed_info, int64_t n_rays, int64_t *out_packed_info, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!keep || !rank || !t_starts || !t_ends || !ray_indices || !out_t_starts || !out_t_ends || !out_ray_indices ||
        (out_packed_info && !packed_info)) {
        set_error("compact_samples: null pointer");
        return CNC_EINVAL;
    }
    const uint64_t threads = (uint64_t)(out_packed_info && n_rays > n ? n_rays : n);
    mr::compact_samples_kernel<<<div_up(threads, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        keep, rank, n, t_starts, t_ends, ray_indices, out_t_starts, out_t_ends, out_ray_indices, packed_info, n_rays, out_packed_info);
    return check_launch("compact_samples");
}

int cnc_render_from_density(const float *t_starts, const float *t_ends, const float *sigmas, const float *rgbs,
                            const int64_t *packed_info, int64_t n_rays, const float *prefix_trans, float *weights,
                            float *trans, float *alphas, float *colors, float *opacities, float *depths,
                            cnc_stream_t stream) {
    if (n_rays == 0) return CNC_OK;
    if (!t_starts || !t_ends || !sigmas || !packed_info) { set_error("render_from_density: null pointer"); return CNC_EINVAL; }
    mr::render_density_kernel<<<div_up((uint64_t)n_rays * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        t_starts, t_ends, sigmas, rgbs, packed_info, n_rays, prefix_trans, weights, trans, alphas, colors, opacities, depths);
    return check_launch("render_from_density");
}

}  // extern "C"
