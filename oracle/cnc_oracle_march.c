/*
 * cnc_oracle_march.c -- CPU restatement of nerfacc's occupancy-grid ray marching as
 * vendored by the reference.  TEST INFRASTRUCTURE ONLY (see cnc_oracle.c header).
 *
 * Follows nerfacc/cuda/csrc/grid.cu:68-318 (traverse_grids_kernel), :320-349
 * (ray_aabb_intersect_kernel) and nerfacc/cuda/csrc/include/utils_grid.cuh:11-149.
 * fmaf() marks the multiply-adds nvcc contracts in the reference build.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { float x, y, z; } f3;

/* utils_grid.cuh:11-57.  near/far are the kernel-wide planes. */
static int aabb_hit(const float *o, const float *d, float near, float far, const float *bb,
                    float *tmin_o, float *tmax_o) {
    float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    float tmin, tmax, a, b;
    if (inv[0] >= 0) { tmin = (bb[0] - o[0]) * inv[0]; tmax = (bb[3] - o[0]) * inv[0]; }
    else             { tmin = (bb[3] - o[0]) * inv[0]; tmax = (bb[0] - o[0]) * inv[0]; }
    for (int k = 1; k < 3; k++) {
        if (inv[k] >= 0) { a = (bb[k] - o[k]) * inv[k]; b = (bb[3 + k] - o[k]) * inv[k]; }
        else             { a = (bb[3 + k] - o[k]) * inv[k]; b = (bb[k] - o[k]) * inv[k]; }
        if (tmin > b || a > tmax) return 0;
        if (a > tmin) tmin = a;
        if (b < tmax) tmax = b;
    }
    if (tmax <= 0) return 0;
    *tmin_o = fmaxf(tmin, near);
    *tmax_o = fminf(tmax, far);
    return 1;
}

/* grid.cu:320-349 */
void cnc_o_ray_aabb_intersect(const float *rays_o, const float *rays_d, int32_t n_rays, float near,
                              float far, const float *aabbs, int32_t n_aabbs, float miss,
                              float *t_mins, float *t_maxs, uint8_t *hits) {
    for (int32_t t = 0; t < n_rays * n_aabbs; t++) {
        int32_t r = t / n_aabbs, a = t % n_aabbs;
        float lo, hi;
        int h = aabb_hit(rays_o + r * 3, rays_d + r * 3, near, far, aabbs + a * 6, &lo, &hi);
        t_mins[t] = h ? lo : miss;
        t_maxs[t] = h ? hi : miss;
        hits[t] = (uint8_t)h;
    }
}

static float calc_dt(float t, float cone, float dmin, float dmax) {
    float v = t * cone; /* grid.cu:23-28 */
    return fmaxf(dmin, fminf(v, dmax));
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/*
 * One pass of traverse_grids_kernel for every ray.  If sample_t == NULL only the counts are
 * produced (the reference's first pass).  Intervals are not materialised: for a fixed step
 * the reference derives t_starts/t_ends from the interval edges (occ_grid.py:188-190), which
 * are exactly (t_last, t_next) of every emitted sample; both are returned directly.
 *   chunk_starts[n_rays] : output offsets (input when filling)
 *   cnt[n_rays]          : samples per ray (output)
 */
void cnc_o_traverse_grids(const float *rays_o, const float *rays_d, const uint8_t *rays_mask,
                          int32_t n_rays, int32_t n_grids, int32_t rx, int32_t ry, int32_t rz,
                          const uint8_t *binaries, const float *aabbs, const uint8_t *hits,
                          const float *t_sorted, const int64_t *t_indices, const float *near_planes,
                          const float *far_planes, float step_size, float cone_angle,
                          int32_t steps_limit, const int64_t *chunk_starts, int64_t *cnt,
                          float *t_starts, float *t_ends, int64_t *ray_idx, float *terminate) {
    const float eps = 1e-6f;
    for (int32_t tid = 0; tid < n_rays; tid++) {
        if (rays_mask && !rays_mask[tid]) { if (cnt) cnt[tid] = 0; continue; }
        const float *o = rays_o + tid * 3, *d = rays_d + tid * 3;
        float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        float near = near_planes[tid], far = far_planes[tid];
        int64_t n_samples = 0;
        int64_t start = chunk_starts ? chunk_starts[tid] : 0;
        float t_last = near;
        int continuous = 0;
        int32_t bh = tid * n_grids, bt = tid * n_grids * 2;
        for (int32_t i = bt; i < bt + n_grids * 2 - 1; i++) {
            int entering = t_indices[i] < n_grids;
            int64_t level = t_indices[i] % n_grids;
            if (!hits[bh + level]) continue;
            if (!entering) {
                int next_entering = t_indices[i + 1] < n_grids;
                if (next_entering) continue;
                level = t_indices[i + 1] % n_grids;
                if (!hits[bh + level]) continue;
            }
            float this_tmin = fmaxf(t_sorted[i], near);
            float this_tmax = fminf(t_sorted[i + 1], far);
            if (this_tmin >= this_tmax) continue;
            if (!continuous) {
                if (step_size <= 0.0f) t_last = this_tmin;
                else {
                    float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
                    while (!(fmaf(dt, 0.5f, t_last) >= this_tmin)) t_last += dt;
                }
            }
            const float *bb = aabbs + level * 6;
            /* setup_traversal, utils_grid.cuh:59-118 */
            float res[3] = {(float)rx, (float)ry, (float)rz};
            int ires[3] = {rx, ry, rz};
            float vox[3], rs[3], re[3], tdist[3], delta[3];
            int step[3], cur[3], fin[3], over[3];
            float ts = this_tmin + eps, te = this_tmax - eps;
            for (int k = 0; k < 3; k++) {
                vox[k] = (bb[3 + k] - bb[k]) / res[k];
                rs[k] = fmaf(d[k], ts, o[k]);
                re[k] = fmaf(d[k], te, o[k]);
                cur[k] = clampi((int)(((rs[k] - bb[k]) / (bb[3 + k] - bb[k])) * res[k]), 0, ires[k] - 1);
                fin[k] = clampi((int)(((re[k] - bb[k]) / (bb[3 + k] - bb[k])) * res[k]), 0, ires[k] - 1);
                int si = cur[k] + (d[k] > 0 ? 1 : 0);
                float txyz = fmaf(bb[k] + fmaf((float)si, vox[k], -rs[k]), inv[k], this_tmin);
                tdist[k] = (d[k] == 0.0f) ? this_tmax : txyz;
                float sf = (d[k] == 0.0f) ? 0.0f : (d[k] > 0.0f ? 1.0f : -1.0f);
                step[k] = (int)sf;
                float dtmp = vox[k] * inv[k] * sf;
                delta[k] = (d[k] == 0.0f) ? this_tmax : dtmp;
                over[k] = fin[k] + step[k];
            }
            while (steps_limit <= 0 || n_samples < steps_limit) {
                float t_trav = fminf(tdist[0], fminf(tdist[1], tdist[2]));
                t_trav = fminf(t_trav, this_tmax);
                int64_t cell = (int64_t)cur[0] * ry * rz + (int64_t)cur[1] * rz + cur[2] +
                               level * (int64_t)rx * ry * rz;
                if (!binaries[cell]) {
                    if (step_size <= 0.0f) t_last = t_trav;
                    else {
                        float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
                        while (!(fmaf(dt, 0.5f, t_last) >= t_trav)) t_last += dt;
                    }
                    continuous = 0;
                } else {
                    while (steps_limit <= 0 || n_samples < steps_limit) {
                        float t_next;
                        if (step_size <= 0.0f) t_next = t_trav;
                        else {
                            float dt = calc_dt(t_last, cone_angle, step_size, 1e10f);
                            if (fmaf(dt, 0.5f, t_last) >= t_trav) break;
                            t_next = t_last + dt;
                        }
                        if (t_starts) {
                            t_starts[start + n_samples] = t_last;
                            t_ends[start + n_samples] = t_next;
                            ray_idx[start + n_samples] = tid;
                        }
                        n_samples++;
                        continuous = 1;
                        t_last = t_next;
                        if (t_next >= t_trav) break;
                    }
                }
                /* single_traversal, utils_grid.cuh:121-149 */
                int ax = ((tdist[0] < tdist[1]) && (tdist[0] < tdist[2])) ? 0 : ((tdist[1] < tdist[2]) ? 1 : 2);
                cur[ax] += step[ax];
                tdist[ax] += delta[ax];
                if (cur[ax] == over[ax]) break;
            }
        }
        if (terminate) terminate[tid] = t_last;
        if (cnt) cnt[tid] = n_samples;
    }
}
