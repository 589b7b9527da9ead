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

// grid.cu:68-318 with the interval edges folded into (t_start, t_end) per sample.  The walk of a ray is one sequential
// chain (t += dt in fp32, DDA boundaries tdist += delta: both defined by their rounding sequence) with nested
// data-dependent loops.  RAY_PER_WARP: lane 0 of every warp walks one ray and the other lanes retire at once -- with
// 32 rays per warp the warp executes the UNION of 32 different loop nests (measured: 292 us for 1100 rays, ten times one
// ray's chain); a training batch has a few thousand rays, far fewer than the machine has warp slots.  Large batches
// (test-time wavefronts of 10^5..10^6 rays with a step limit) keep one thread per ray.
template <bool RAY_PER_WARP>
__global__ void __launch_bounds__(128) traverse_kernel(const MarchArgs a) {
    int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (RAY_PER_WARP) {
        if (threadIdx.x & 31) return;
        tid >>= 5;
    }
    if (tid >= a.n_rays) return;
    if (a.rays_mask && !a.rays_mask[tid]) {
        if (a.cnt) a.cnt[tid] = 0;
        return;
    }
    const float eps = 1e-6f;
    const bool fill = a.t_starts != nullptr;
    const float o[3] = {a.rays_o[tid * 3], a.rays_o[tid * 3 + 1], a.rays_o[tid * 3 + 2]};
    const float d[3] = {a.rays_d[tid * 3], a.rays_d[tid * 3 + 1], a.rays_d[tid * 3 + 2]};
    const float inv[3] = {__fdiv_rn(1.0f, d[0]), __fdiv_rn(1.0f, d[1]), __fdiv_rn(1.0f, d[2])};
    const float near = a.near_planes[tid], far = a.far_planes[tid];
    int64_t n_samples = 0;
    const int64_t start = a.chunk_starts ? a.chunk_starts[tid] : 0;
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
        if (!continuous) {
            if (a.step_size <= 0.0f) t_last = this_tmin;
            else {
                const float dt = calc_dt(t_last, a.cone_angle, a.step_size, 1e10f);
                while (!(__fmaf_rn(dt, 0.5f, t_last) >= this_tmin)) t_last = __fadd_rn(t_last, dt);
            }
        }
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
        while (a.steps_limit <= 0 || n_samples < a.steps_limit) {
            const float t_trav = fminf(fminf(tdist[0], fminf(tdist[1], tdist[2])), this_tmax);
            const int64_t cell = (int64_t)cur[0] * a.ry * a.rz + (int64_t)cur[1] * a.rz + cur[2] + level * (int64_t)a.rx * a.ry * a.rz;
            if (!a.binaries[cell]) {
                if (a.step_size <= 0.0f) t_last = t_trav;
                else {
                    const float dt = calc_dt(t_last, a.cone_angle, a.step_size, 1e10f);
                    while (!(__fmaf_rn(dt, 0.5f, t_last) >= t_trav)) t_last = __fadd_rn(t_last, dt);
                }
                continuous = false;
            } else {
                while (a.steps_limit <= 0 || n_samples < a.steps_limit) {
                    float t_next;
                    if (a.step_size <= 0.0f) t_next = t_trav;
                    else {
                        const float dt = calc_dt(t_last, a.cone_angle, a.step_size, 1e10f);
                        if (__fmaf_rn(dt, 0.5f, t_last) >= t_trav) break;
                        t_next = __fadd_rn(t_last, dt);
                    }
                    if (fill) {
                        a.t_starts[start + n_samples] = t_last;
                        a.t_ends[start + n_samples] = t_next;
                        a.ray_idx[start + n_samples] = tid;
                    }
                    n_samples++;
                    continuous = true;
                    t_last = t_next;
                    if (t_next >= t_trav) break;
                }
            }
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
    if ((t_starts != nullptr) != (chunk_starts != nullptr) || (t_starts && (!t_ends || !ray_indices)) || (!t_starts && !cnt)) {
        set_error("traverse_grids: count pass needs cnt, fill pass needs chunk_starts + t_starts + t_ends + ray_indices");
        return CNC_EINVAL;
    }
    mr::MarchArgs a{rays_o, rays_d, rays_mask, n_rays, n_grids, rx, ry, rz, binaries, aabbs, hits, t_sorted, t_indices,
                    near_planes, far_planes, step_size, cone_angle, steps_limit, chunk_starts, cnt, t_starts, t_ends,
                    ray_indices, terminate_planes};
    if (n_rays <= 148 * 64 * 2)   // at most two waves of one-ray warps
        mr::traverse_kernel<true><<<div_up((uint64_t)n_rays * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    else
        mr::traverse_kernel<false><<<div_up((uint64_t)n_rays, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("traverse_grids");
}

int cnc_packed_scan(const float *in, const int64_t *packed_info, int64_t n_rays, float *out, int32_t op, int32_t inclusive,
                    int32_t reverse, cnc_stream_t stream) {
    if (n_rays == 0) return CNC_OK;
    if (!in || !packed_info || !out) { set_error("packed_scan: null pointer"); return CNC_EINVAL; }
    mr::packed_scan_kernel<<<div_up((uint64_t)n_rays * 32, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(in, packed_info, n_rays, out, op, inclusive, reverse);
    return check_launch("packed_scan");
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
