// table_build.cu -- inverse hash tables of the context model, built on the GPU and pruned by the occupancy grid.
//
// Reference: CNC_context_models.__init__ (examples/utils_bpp_acc.py:294-335) enumerates EVERY lattice vertex of every
// level (up to 514^3 = 135.8 M), hashes it with the python twin of the CUDA hash, sorts 135.8 M int64 keys and keeps the
// int16 coordinates grouped by table row: 815 MB for the finest level alone, 1.34 GB for the product layout.  The codec
// then throws most of it away again on every call: a vertex takes part only if its +-1-cell box touches an occupied
// occupancy cell (query_mask_3D, aligner_kernel.cu:161-242; `mask` / `mask_exist` of utils_bpp_acc.py:811-833).
//
// Here the row statistics the driver needs from ALL vertices (which rows of a level are hit at all: that fixes the entry
// numbering, the stream chunking and the sample windows) come from one histogram pass without storing anything per vertex,
// and the per-vertex list is built for the occupied neighbourhood only:
//
//   cnc_level_row_hist    : cnt[row] += 1 for every vertex of a level                       (no output per vertex)
//   cnc_level_pruned_keys : for every vertex that passes the reference's occupancy test, emit the 64-bit key
//                           (entry index << 28) | lattice index, lattice index = (x*res + y)*res + z.  Sorting the keys
//                           (one radix sort of the SURVIVORS, ~15 % of the vertices) reproduces the reference's order:
//                           entries ascending, vertices of an entry in lattice order (its stable sort of a lattice-ordered
//                           list).  keys == NULL counts only (the caller sizes the buffer exactly).
//   cnc_keys_to_points    : sorted keys -> int16 coordinates [n,3] + int32 entry index [n]
//
// All integer work; the mask test is `voxel_mask_overlap<3>` of common.cuh, the same code query_mask runs (bit-exact
// against the reference binary).  One thread per vertex, consecutive threads = consecutive z: the occupancy probes of a
// warp fall into a few rows of the 2 MiB grid (L1/L2 resident); the only HBM traffic is the survivors' 8-byte keys.
#include <cuda_runtime.h>

#include "common.cuh"

namespace cnc {

__global__ void __launch_bounds__(256) level_row_hist_kernel(uint32_t res, uint32_t T, uint32_t *__restrict__ cnt, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t lin = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; lin < n; lin += stride) {
        const uint32_t z = (uint32_t)(lin % res), xy = (uint32_t)(lin / res);
        const uint32_t c[3] = {xy / res, xy % res, z};
        atomicAdd(cnt + grid_row<3>(c, T, res), 1u);
    }
}

__global__ void __launch_bounds__(256)
level_pruned_keys_kernel(uint32_t res, uint32_t T, const uint8_t *__restrict__ vxl, int32_t Rb, const int32_t *__restrict__ entry_of_row,
                         unsigned long long *__restrict__ keys, unsigned long long *__restrict__ counter, uint64_t n, int mode) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u;
    // (the loop bound is rounded up to whole warps so that the ballot below is executed by all 32 lanes)
    const uint64_t n_up = (n + 31) / 32 * 32;
    for (uint64_t lin = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; lin < n_up; lin += stride) {
        bool pass = false;
        uint32_t row = 0;
        if (lin < n) {
            const uint32_t z = (uint32_t)(lin % res), xy = (uint32_t)(lin / res);
            const uint32_t cu[3] = {xy / res, xy % res, z};
            const int ci[3] = {(int)cu[0], (int)cu[1], (int)cu[2]};
            int32_t ov;
            pass = mode == 0 ? voxel_mask_overlap<3>(ci, (float)res, Rb, vxl, ov) : vote_member(cu, res, (uint32_t)Rb, vxl);
            if (pass) row = grid_row<3>(cu, T, res);
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, pass);
        if (ballot == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass && keys) {
            const uint64_t entry = entry_of_row ? (uint64_t)(uint32_t)__ldg(entry_of_row + row) : (uint64_t)row;
            keys[base + __popc(ballot & ((1u << lane) - 1u))] = (entry << 28) | lin;
        }
    }
}

__global__ void __launch_bounds__(256)
keys_to_points_kernel(const unsigned long long *__restrict__ keys, uint64_t n, uint32_t res, int16_t *__restrict__ pts,
                      int32_t *__restrict__ entry) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const uint32_t lin = (uint32_t)(k & ((1ull << 28) - 1ull));
    const uint32_t z = lin % res, xy = lin / res;
    pts[i * 3 + 0] = (int16_t)(xy / res);
    pts[i * 3 + 1] = (int16_t)(xy % res);
    pts[i * 3 + 2] = (int16_t)z;
    entry[i] = (int32_t)(k >> 28);
}

static inline int grid_for(uint64_t n) {
    const uint64_t want = (n + 255) / 256, cap = 148ull * 16ull;   // 148 SMs x 8 CTAs of 256 threads, two waves
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace cnc

using namespace cnc;

extern "C" {

int cnc_level_row_hist(uint32_t resolution, uint32_t hashmap_size, uint32_t *counts, cnc_stream_t stream) {
    if (!counts || resolution < 3 || hashmap_size == 0) { set_error("level_row_hist: bad argument"); return CNC_EINVAL; }
    const uint64_t n = (uint64_t)resolution * resolution * resolution;
    if (n >= (1ull << 28)) { set_error("level_row_hist: resolution^3 must stay below 2^28"); return CNC_EINVAL; }
    level_row_hist_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(resolution, hashmap_size, counts, n);
    return check_launch("level_row_hist");
}

int cnc_level_pruned_keys(uint32_t resolution, uint32_t hashmap_size, const uint8_t *binary_vxl, int32_t Rb, const int32_t *entry_of_row,
                          uint64_t *keys, uint64_t *counter, int32_t mode, cnc_stream_t stream) {
    if (!binary_vxl || !counter || resolution < 3 || hashmap_size == 0 || Rb < 1 || mode < 0 || mode > 1) { set_error("level_pruned_keys: bad argument"); return CNC_EINVAL; }
    if (mode == 1 && ((resolution - 2) % (uint32_t)Rb)) { set_error("level_pruned_keys: vote membership needs (resolution - 2) % Rb == 0"); return CNC_EINVAL; }
    const uint64_t n = (uint64_t)resolution * resolution * resolution;
    if (n >= (1ull << 28)) { set_error("level_pruned_keys: resolution^3 must stay below 2^28"); return CNC_EINVAL; }
    level_pruned_keys_kernel<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        resolution, hashmap_size, binary_vxl, Rb, entry_of_row, reinterpret_cast<unsigned long long *>(keys),
        reinterpret_cast<unsigned long long *>(counter), n, mode);
    return check_launch("level_pruned_keys");
}

int cnc_keys_to_points(const uint64_t *keys, uint64_t n, uint32_t resolution, int16_t *pts, int32_t *entry, cnc_stream_t stream) {
    if (n == 0) return CNC_OK;
    if (!keys || !pts || !entry || resolution < 3) { set_error("keys_to_points: bad argument"); return CNC_EINVAL; }
    keys_to_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long *>(keys), n, resolution, pts, entry);
    return check_launch("keys_to_points");
}

}  // extern "C"
