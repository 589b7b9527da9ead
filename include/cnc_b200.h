/*
 * cnc_b200.h -- C-ABI of libcnc_b200.so: the B200 (sm_100a) implementation of CNC's
 * data-parallel hot path (hash-grid encode, field MLP, context model, range coder).
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers, sizes and a
 * CUDA stream (passed as void* == cudaStream_t; NULL = legacy default stream), allocates
 * nothing the caller can see, keeps no state between calls, and returns an int status:
 *      0  CNC_OK
 *     -1  CNC_EINVAL      bad argument (null pointer, size overflow)
 *     -2  CNC_ENOTSUP     unsupported D / F combination (the reference throws
 *                         std::runtime_error for the same cases, gridencoder.cu:641,669)
 *     -3  CNC_ECUDA       a CUDA runtime error; cnc_last_error() has the text
 * The Python shims in cnc_b200/ map non-zero codes to RuntimeError, like the reference's
 * TORCH_CHECK / std::runtime_error paths (gridencoder.cu:15-18,764-783).
 *
 * Each prototype cites the reference interface it replaces (paths relative to the
 * reference repo YihangChen-ee/CNC); INTEGRATION.md shows the binding a maintainer would
 * add on the reference side.
 */
#ifndef CNC_B200_H
#define CNC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNC_OK 0
#define CNC_EINVAL (-1)
#define CNC_ENOTSUP (-2)
#define CNC_ECUDA (-3)

typedef void *cnc_stream_t; /* cudaStream_t */

int cnc_version(void);
const char *cnc_last_error(void); /* thread-local, valid until the next failing call */

/* ------------------------------------------------------------------------------------------
 * Hash-grid encode.
 * replaces: _gridencoder.grid_encode_forward   gridencoder/src/gridencoder.h:12-22,
 *           gridencoder.cu:752-806 (host), :99-316 (kernel_grid)
 *   x            [N,D] f32 in [0,1]            table [rows,F] f32
 *   offsets      [>=L_calc+1] i32 (already sliced by the caller when min_level_id is NULL,
 *                ngp.py:86-97)                 resolutions [>=L_calc] i32
 *   out          [L_calc,N,F] f32 (level-major, gridencoder.cu:131)
 *   binary_vxl   nullable, bool/u8 [Rb]^D occupancy; min_level_id nullable i32 [N]
 * D in {1,2,3}; F in {1,2,4,8,16,32}.
 * ---------------------------------------------------------------------------------------- */
int cnc_grid_encode_fwd(const float *x, const float *table, const int32_t *offsets,
                        const int32_t *resolutions, float *out, uint32_t N, uint32_t D, uint32_t F,
                        uint32_t L_calc, uint32_t Rb, const uint8_t *binary_vxl,
                        const int32_t *min_level_id, cnc_stream_t stream);

/* replaces: _gridencoder.grid_encode_backward  gridencoder.h:24-36, gridencoder.cu:808-866,
 *           :400-585 (kernel_grid_backward).  grad [L_calc,N,F]; grad_table [rows,F] must arrive
 *           zeroed (ngp.py:129) and is accumulated into with float atomics. */
int cnc_grid_encode_bwd(const float *grad, const float *x, const int32_t *offsets,
                        const int32_t *resolutions, float *grad_table, uint32_t N, uint32_t D,
                        uint32_t F, uint32_t L_calc, uint32_t Rb, const uint8_t *binary_vxl,
                        const int32_t *min_level_id, cnc_stream_t stream);
/* Same scatter-add, gradient read in place from a column block of a row-major [N, ld] matrix (level l at columns
 * col0 + l*F ..): the layout a GEMM that produced the feature gradients leaves behind; saves the permute + copy into
 * the reference's [L, N, F] layout (ngp.py:126). */
int cnc_grid_encode_bwd_rows(const float *grad_rows, uint32_t ld, uint32_t col0, const float *x,
                             const int32_t *offsets, const int32_t *resolutions, float *grad_table,
                             uint32_t N, uint32_t D, uint32_t F, uint32_t L_calc, uint32_t Rb,
                             const uint8_t *binary_vxl, const int32_t *min_level_id, cnc_stream_t stream);

/* Same gather, reading a 1-bit/parameter sign table (bit ch of byte-group row = param >= 0)
 * produced by cnc_sign_pack; the encoded features are bit-identical to cnc_grid_encode_fwd on
 * STE_binary(params) because every table value is exactly +-1.  (ngp.py:244-245 + K1) */
int cnc_grid_encode_fwd_bits(const float *x, const uint8_t *sign_bits, const int32_t *offsets,
                             const int32_t *resolutions, float *out, uint32_t N, uint32_t D,
                             uint32_t F, uint32_t L_calc, uint32_t Rb, const uint8_t *binary_vxl,
                             const int32_t *min_level_id, cnc_stream_t stream);

/* STE_binary (ngp.py:22-39 == utils_bpp_acc.py:164-181).
 * fwd: out = (clamp(p,-1,1) >= 0) ? +1 : -1.   bwd: gin = gout * (|p| <= 1). */
int cnc_ste_binary_fwd(const float *params, float *out, uint64_t n, cnc_stream_t stream);
int cnc_ste_binary_bwd(const float *params, const float *gout, float *gin, uint64_t n,
                       cnc_stream_t stream);
/* sign bit-planes: bits[(row*F + ch) / 8] bit ((row*F+ch) % 8) = params[row,ch] >= 0.
 * n = rows*F must be a multiple of 8 (row counts are, ngp.py:204). */
int cnc_sign_pack(const float *params, uint8_t *bits, uint64_t n, cnc_stream_t stream);
int cnc_sign_unpack(const uint8_t *bits, float *out, uint64_t n, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Plane context models: y [N,8] = x [N,K] W^T [8,K] + b, K = 8 * context levels + 1 in {9, 17, 25, 33}.
 * replaces: context_model_2D[n-1] (nn.Linear -> cuBLAS sgemm) and its autograd backward, utils_bpp_acc.py:386-393, :558-561.
 *   bwd: gx [N,K] (nullable) = gy W; parts [ceil(N / cnc_lin8_rows_per_block()), 8K + 8] = per-block partials of
 *   (dW [8,K] row-major | db [8]); the caller sums them over the blocks in index order.
 * ---------------------------------------------------------------------------------------- */
/* sum of Bernoulli_entropy(x, p) (utils_bpp_acc.py:1002-1013) and its gradient, as the rate term of the loss uses it:
 *   fwd: parts [cnc_bernoulli_bits_blocks(n)] = per-block partial sums of -log2(pc)(1+x)/2 - log2(1-pc)(1-x)/2, pc = clamp(p);
 *   bwd: gx / gp (nullable) = *g (upstream scalar, device) times d/dx, d/dp (zero where the clamp is active). */
int cnc_bernoulli_bits_blocks(int64_t n);
/* rows of 8 floats: out[i] = table[rows[i]] / out[rows[i]] = grad[i] (distinct rows; the caller zeroes `out` of the scatter):
 * `params_q[unique_value_list + offset]` of utils_bpp_acc.py:560,690 and its backward */
int cnc_rows8_gather(const float *table, const int64_t *rows, int64_t M, float *out, cnc_stream_t stream);
int cnc_rows8_scatter(const float *grad, const int64_t *rows, int64_t M, float *out, cnc_stream_t stream);
int cnc_bernoulli_bits_fwd(const float *x, const float *p, int64_t n, float *parts, cnc_stream_t stream);
int cnc_bernoulli_bits_bwd(const float *x, const float *p, const float *g, int64_t n, float *gx, float *gp, cnc_stream_t stream);
int cnc_lin8_rows_per_block(void);
int cnc_lin8_fwd(const float *x, const float *W, const float *b, float *y, int64_t N, int32_t K, cnc_stream_t stream);
int cnc_lin8_bwd(const float *x, const float *W, const float *gy, float *gx, float *parts, int64_t N, int32_t K,
                 cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 3D context gather of the rate term (training loss), forward and backward.
 * replaces: Encoding_xyz.forward_diff_levels(points, n_list - 3, 3, binary_vxl, PV=1001) + torch.cat([context, Pg]) and its
 *           autograd backward, examples/utils_bpp_acc.py:644-687 (K1 / K2 with per-point start level, gridencoder.cu:118-126).
 *   pts [M,3] i16 voxel coordinates at their own level `level[i]` (>= 3), sign_bits = cnc_sign_pack of the 3D table,
 *   vertex_bits / vertex_bit_offsets from cnc_vertex_valid_bits (the per-corner occupancy predicate as one bit per vertex),
 *   Pg [L] level frequencies.  fwd: x [M,25] = 3 x 8 interpolated context features of levels n-3..n-1 | Pg[n].
 *   bwd: grad_table [rows,8] += scatter of gx[:, 0:24], grad_pg [L] += gx[:,24] summed per level (caller zeroes both).
 * ---------------------------------------------------------------------------------------- */
int cnc_ctx3d_gather_fwd(const int16_t *pts, const int64_t *level, int64_t M, const uint8_t *sign_bits, const int32_t *offsets,
                         const int32_t *resolutions, const uint32_t *vertex_bits, const int64_t *vertex_bit_offsets,
                         const float *Pg, float *x, cnc_stream_t stream);
int cnc_ctx3d_gather_bwd(const int16_t *pts, const int64_t *level, int64_t M, const int32_t *offsets, const int32_t *resolutions,
                         const uint32_t *vertex_bits, const int64_t *vertex_bit_offsets, const float *gx, float *grad_table,
                         float *grad_pg, cnc_stream_t stream);

/* Plane context (utils_bpp_acc.py:551-558, :731-738) in one kernel each way:
 * replaces: Encoding_2D(points, n - c, n, binary_vxl_2D) + Encoding_2D.forward_given_params(points, ..., pn_embed_frac,
 *           binary_vxl_2D) + expand + torch.cat, and their autograd backward (2 x K1 / K2 with D = 2, ngp.py:228-315).
 *   pts [N,2] normalised plane vertices; sign_bits / offsets / resolutions of the plane encoder; level n, n_ctx_levels c (levels
 *   n-c..n-1); binary_vxl_2D [Rb,Rb]; frac [res_frac^2, 8] f32 (nullable: no dimension-wise context), Pg 1 float (device).
 *   fwd: x [N, 8c (+8) + 1].  bwd: grad_table [rows,8] +=, grad_frac [res_frac^2,8] += (nullable), grad_pg [1] += (caller zeroes). */
int cnc_ctx2d_gather_fwd(const float *pts, int64_t N, const uint8_t *sign_bits, const int32_t *offsets, const int32_t *resolutions,
                         int32_t level, int32_t n_ctx_levels, const uint8_t *binary_vxl_2D, int32_t Rb, const float *frac,
                         int32_t res_frac, const float *Pg, float *x, cnc_stream_t stream);
int cnc_ctx2d_gather_bwd(const float *pts, int64_t N, const int32_t *offsets, const int32_t *resolutions, int32_t level,
                         int32_t n_ctx_levels, const uint8_t *binary_vxl_2D, int32_t Rb, int32_t res_frac, const float *gx,
                         float *grad_table, float *grad_frac, float *grad_pg, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Test-time wavefront renderer without host round trips.
 * replaces: the python loop of render_image_with_occgrid_test, examples/utils.py:395-479 (per round: ray_mask.sum().item(),
 *           traverse_grids(over_allocate), boolean compactions, rgb_sigma_fn, render_weight_from_density, 3 x
 *           accumulate_along_rays_, ray-mask update).
 * `state` = 16 x u32 in device memory: [0] live rays of the round, [1] samples per ray of the round, [2] iter_samples,
 * [3] done, [4] samples of the round (the count cnc_field_fwd_n reads), [5] live rays after the round (initialise to
 * n_rays), [6] rounds, [8..9] total accumulated samples (u64).  One round = cnc_wavefront_begin -> cnc_wavefront_march
 * (sample placement == traverse_grids, bit for bit; writes t0/t1/pos/dirs of the round's samples into `capacity` slots)
 * -> cnc_field_fwd_n(pos, dirs, ..., n_dev = state + 4, n_max = capacity) -> cnc_wavefront_composite.  After `done` every
 * kernel returns at once, so a caller may queue rounds in batches and look at state[3] between batches.
 * ---------------------------------------------------------------------------------------- */
int cnc_wavefront_begin(uint32_t *state, uint32_t n_rays, uint32_t min_samples, uint32_t max_samples, cnc_stream_t stream);
int cnc_wavefront_march(const float *rays_o, const float *rays_d, int64_t n_rays, int32_t n_grids, int32_t rx, int32_t ry, int32_t rz,
                        const uint8_t *binaries, const float *aabbs, const uint8_t *hits, const float *t_sorted,
                        const int64_t *t_indices, const float *far_planes, float step_size, float cone_angle, uint32_t *state,
                        uint8_t *ray_mask, float *near_planes, uint32_t capacity, uint32_t *ray_base, uint32_t *ray_cnt,
                        float *ray_term, float *t0, float *t1, float *pos, float *dirs, cnc_stream_t stream);
int cnc_wavefront_composite(uint32_t *state, uint8_t *ray_mask, float *near_planes, const uint32_t *ray_base, const uint32_t *ray_cnt,
                            const float *ray_term, const float *t0, const float *t1, const float *sigma, const float *rgbs,
                            float *rgb, float *opacity, float *depth, int64_t n_rays, uint32_t capacity, float alpha_thre,
                            float opc_thre, cnc_stream_t stream);
int cnc_field_fwd_n(const float *pos, const float *dirs, const float *aabb6_host, const uint8_t *bits_xyz,
                    const uint8_t *bits_xy, const uint8_t *bits_xz, const uint8_t *bits_yz, const int32_t *offsets3,
                    const int32_t *resolutions3, const int32_t *offsets2, const int32_t *resolutions2,
                    const float *blob, float *sigma, float *rgb, const uint32_t *n_dev, uint32_t n_max, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Inverse hash tables of the context model, pruned by the occupancy grid.
 * replaces: CNC_context_models.__init__ table construction, examples/utils_bpp_acc.py:294-335 (meshgrid of every level ->
 *           get_grid_index -> torch.sort of up to 135.8 M keys -> unique), followed at every encode/decode by
 *           query_mask_3D + boolean compaction of the same lists (:811-833).
 *   cnc_level_row_hist    : counts [hashmap_size] u32 += number of lattice vertices of the level hashing to each row
 *                           (caller zeroes).  Which rows are hit fixes entry numbering / chunking (:337-352, :798-802).
 *   cnc_level_pruned_keys : every vertex whose +-1-cell box touches an occupied cell (the K6 test, aligner_kernel.cu:
 *                           161-242) -> key (entry << 28) | ((x*res + y)*res + z), entry = entry_of_row[row] (NULL:
 *                           entry = row), appended at *counter (u64, caller zeroes).  keys == NULL: count only.
 *                           Sorting the keys gives the reference's order (entries ascending, lattice order inside).
 *                           mode 1: the predicate is membership in the vote list of the dimension-wise context instead
 *                           (get_idx_coords2, utils_bpp_acc.py:498-512, minus the border the vote kernels skip).
 *   cnc_keys_to_points    : sorted keys -> pts [n,3] i16, entry [n] i32.
 * resolution^3 < 2^28.  binary_vxl [Rb,Rb,Rb] u8.
 * ---------------------------------------------------------------------------------------- */
int cnc_level_row_hist(uint32_t resolution, uint32_t hashmap_size, uint32_t *counts, cnc_stream_t stream);
int cnc_level_pruned_keys(uint32_t resolution, uint32_t hashmap_size, const uint8_t *binary_vxl, int32_t Rb,
                          const int32_t *entry_of_row, uint64_t *keys, uint64_t *counter, int32_t mode, cnc_stream_t stream);
int cnc_keys_to_points(const uint64_t *keys, uint64_t n, uint32_t resolution, int16_t *pts, int32_t *entry,
                       cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training step over the latent tables (caller side: the optimizer step of train_CNC_nerf_synthetic.py:254-266,
 * :362-364 and, under data parallelism, the exchange SURVEY 8(e) adds; the reference has neither kernel).
 * A replica needs from a table row only sign(p) (STE forward, ngp.py:26-31) and [|p| <= 1] (STE backward, ngp.py:33-39):
 *   cnc_ste_planes_pack : params [n] -> sign plane (cnc_sign_pack layout) and window plane (bit = -1 <= p <= 1), n % 32 == 0;
 *                         mask_bits may be NULL.
 *   cnc_surrogate_fill  : params[i] = +-0.5 (window bit set) / +-1.5 (clear), sign from the sign plane, for every i
 *                         outside [keep_lo, keep_hi); n, keep_lo, keep_hi multiples of 32.
 *   cnc_adam_planes     : torch.optim.Adam (L2 weight decay, fp32 state, bias correction with `step` >= 1) over
 *                         params/grad/exp_avg/exp_avg_sq [n], emitting the two planes of the updated values in the same
 *                         pass (sign_bits / mask_bits nullable); grad is divided by grad_scale first (GradScaler, :211,362);
 *                         ste_window != 0: the gradient of a latent outside [-1, 1] is dropped first (the STE_binary
 *                         backward, ngp.py:33-39, for callers that hand over the unmasked table gradient).
 * ---------------------------------------------------------------------------------------- */
/* set bits per level of a sign plane: out[l] = popcount of bytes [byte_offsets[l], byte_offsets[l+1]) -- the +1 count of a
 * level of a binarised table (get_BiRF_wentropy_leveln, utils_bpp_acc.py:472-486) without an fp32 pass over the level */
int cnc_level_popcount(const uint8_t *sign_bits, const int64_t *byte_offsets, int32_t n_levels, int64_t *out, cnc_stream_t stream);
int cnc_ste_planes_pack(const float *params, uint8_t *sign_bits, uint8_t *mask_bits, uint64_t n, cnc_stream_t stream);
int cnc_surrogate_fill(float *params, const uint8_t *sign_bits, const uint8_t *mask_bits, uint64_t n, uint64_t keep_lo,
                       uint64_t keep_hi, cnc_stream_t stream);
int cnc_adam_planes(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, uint8_t *sign_bits, uint8_t *mask_bits,
                    uint64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    int32_t ste_window, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Gradient exchange of the data-parallel step over NVLink peer memory (csrc/peer.cu; SURVEY 8(e): the reference has no
 * distributed code).  Buffers live in plain device allocations that every rank of the box maps (CUDA IPC); the 64-byte
 * handles travel through the caller's control plane (torch.distributed).
 *   cnc_peer_alloc / _free      : zero-filled device allocation that can be exported
 *   cnc_peer_export / _import   : cudaIpcGetMemHandle / cudaIpcOpenMemHandle (`handle`: cnc_peer_handle_bytes() bytes);
 *                                 _unmap closes an imported mapping
 *   cnc_peer_barrier            : all ranks meet: rank r stores `epoch` into word [slot][r] of every rank's signal pad
 *                                 (pads[k] = rank k's pad as mapped here, cnc_peer_pad_bytes() bytes, zero at start) and waits
 *                                 until its own pad holds >= epoch from everybody; epochs of a slot must increase.  A wait
 *                                 longer than timeout_ms traps the kernel (the next CUDA call reports it).
 *   cnc_peer_min                : every rank contributes `value` and receives the minimum over the ranks in *out (device
 *                                 word); same meeting discipline as the barrier, on pad slots slot .. slot + 2
 *   cnc_peer_reduce             : out[i] = scale * sum_{k < world} srcs[k][lo + i], i < count, summed in rank order (lo,
 *                                 count multiples of 4); `blocks` caps the grid (0 = one CTA per SM)
 *   cnc_peer_push               : words [off, off + words) of rank `rank`'s arena are stored at the same offsets of every
 *                                 other rank's arena (arenas[k] as mapped here; uint32 words; at most 8 segments)
 * ---------------------------------------------------------------------------------------- */
int cnc_peer_alloc(uint64_t bytes, void **out);
int cnc_peer_free(void *p);
int cnc_peer_handle_bytes(void);
int cnc_peer_pad_bytes(void);
int cnc_peer_export(void *p, uint8_t *handle);
int cnc_peer_import(const uint8_t *handle, void **out);
int cnc_peer_unmap(void *p);
int cnc_peer_barrier(void *const *pads, int32_t rank, int32_t world, int32_t slot, uint32_t epoch, uint32_t timeout_ms,
                     cnc_stream_t stream);
int cnc_peer_min(void *const *pads, int32_t rank, int32_t world, int32_t slot, uint32_t epoch, uint32_t value, uint32_t *out,
                 uint32_t timeout_ms, cnc_stream_t stream);
int cnc_peer_reduce(const void *const *srcs, int32_t world, int64_t lo, int64_t count, float scale, float *out, int32_t blocks,
                    cnc_stream_t stream);
int cnc_peer_push(void *const *arenas, int32_t rank, int32_t world, const int64_t *seg_off_words, const int64_t *seg_words,
                  int32_t n_seg, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dimension-wise context: 3D -> 2D vote planes.
 * replaces: _gridencoder.cnt_np_embed / cnt_np_embed_backward   gridencoder.h:39-53,
 *           gridencoder.cu:873-915, :972-1020
 *   pts [N,3] i16 voxel coords at `resolution`; table = finest 3D level [T,F] f32;
 *   out [res-2,res-2,F,2] f32 (+1 votes, -1 votes), accumulated into (caller zeroes).
 * ---------------------------------------------------------------------------------------- */
int cnc_vote_planes_fwd(const int16_t *pts, const float *table, float *out, uint32_t N,
                        uint32_t resolution, uint32_t F, uint32_t hashmap_size, uint32_t axis,
                        cnc_stream_t stream);
int cnc_vote_planes_bwd(const int16_t *pts, const float *table, const float *out_sum,
                        const float *grad, float *grad_table, uint32_t N, uint32_t resolution,
                        uint32_t F, uint32_t hashmap_size, uint32_t axis, cnc_stream_t stream);

/* The same vote planes for all three axes at once, from the occupancy grid instead of the enumerated voxel list and
 * without atomics (what get_idx_coords2 + 3 x cnt_np_embed compute, utils_bpp_acc.py:498-530): a finest-level voxel c is
 * in the reference's list iff one of the <= 8 occupancy cells o with o*t <= c <= o*t + t + 1 (t = (res-2)/Rb) is
 * occupied.  sign_bits: cnc_sign_pack of the level's rows (vote +1 <=> value > 0.9 <=> sign bit, the table is +-1).
 * out_* [(res-2),(res-2),8,2] are overwritten with exact integer counts.  cnc_vote3_bwd is the matching backward
 * (gridencoder.cu:1047-1087 summed over the three planes): pts_by_row / seg = the level's inverse hash table (voxel
 * coords grouped by table row, [T+1] running counts); grad_* [(res-2),(res-2),8,2] = d loss / d fraction already
 * divided by the cell's vote sum (the 1/sum of :1012, folded by the caller); grad_table [T,8] is overwritten.
 * F == 8, Rb <= 128.  *   A NULL output plane (forward) / NULL gradient plane (backward) skips that axis: a data-parallel rank builds only the
 *   planes its share of the plane terms reads.
  *   members_only != 0: pts_by_row lists vote-list members only (cnc_level_pruned_keys mode 1): the per-voxel membership
 *   test is skipped.
 */
int cnc_vote3_fwd(const uint8_t *binary_vxl, uint32_t Rb, const uint8_t *sign_bits, uint32_t resolution,
                  uint32_t F, uint32_t hashmap_size, float *out_xy, float *out_xz, float *out_yz,
                  cnc_stream_t stream);
int cnc_vote3_bwd(const int16_t *pts_by_row, const int64_t *seg, const uint8_t *binary_vxl, uint32_t Rb,
                  const uint8_t *sign_bits, uint32_t resolution, uint32_t F, uint32_t hashmap_size,
                  const float *grad_xy, const float *grad_xz, const float *grad_yz, float *grad_table,
                  int32_t members_only, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Occupancy query of voxels.
 * replaces: pack_and_align.query_mask_3D / query_mask_3D_qlist   my_cuda_backen/aligner.cpp:37-70,
 *           aligner_kernel.cu:4-80,161-242 (scalar resolution), :82-158,244-326 (per point)
 *   pts [N,D] i16; binary_vxl bool [Rb]^D; mask [N] i16; overlap [N] i32;
 *   res_per_point nullable i64 [N] (then `resolution` is ignored); D in {2,3}.
 * ---------------------------------------------------------------------------------------- */
int cnc_query_mask(const int16_t *pts, const uint8_t *binary_vxl, int32_t Rb, int16_t *mask,
                   int32_t *overlap, const int64_t *res_per_point, int32_t resolution, int64_t N,
                   int32_t D, cnc_stream_t stream);

/* replaces: pack_and_align.align_and_pack_forward / _backward   aligner.cpp:4-35,
 *           aligner_kernel.cu:413-495, :498-565.  packed [N,M,F]; feat [T,F]; cnt [N] i64;
 *           cumsum [N+1] i64.  bwd writes every row of dfeat it owns (caller zeroes). */
int cnc_align_pack_fwd(const float *feat, const int64_t *cnt, const int64_t *cumsum, float *packed,
                       int64_t N, int64_t M, int64_t F, float V, cnc_stream_t stream);
int cnc_align_pack_bwd(const float *dpacked, const int64_t *cnt, const int64_t *cumsum,
                       float *dfeat, int64_t N, int64_t M, int64_t F, cnc_stream_t stream);

/* Segment reduce that makes the padded [N,M,F] tensor unnecessary:
 * out[i,k] = sum_j w[cumsum[i]+j] * feat[cumsum[i]+j, k]  (w nullable -> plain sum), fixed
 * left-to-right order per segment (deterministic, GPU-count independent).
 * replaces the pattern align_and_pack -> mul -> sum(dim=1), utils_bpp_acc.py:842-848,563-566. */
int cnc_segment_wsum(const float *feat, const float *w, const int64_t *cumsum, float *out,
                     int64_t N, int64_t F, cnc_stream_t stream);
/* The same reduction with the preceding row gather folded in (out[i] = sum_j w[j] * feat[idx[j]], idx nullable) and its
 * backward w.r.t. feat (grad_feat[idx[j]] = w[j] * grad_out[i]; idx must be a permutation, rows outside every segment
 * are left untouched): the differentiable form used by the rate term (utils_bpp_acc.py:563-566,741-745). */
int cnc_segment_wsum_idx(const float *feat, const int64_t *idx, const float *w, const int64_t *cumsum,
                         float *out, int64_t N, int64_t F, cnc_stream_t stream);
int cnc_segment_wsum_idx_bwd(const float *grad_out, const int64_t *idx, const float *w,
                             const int64_t *cumsum, float *grad_feat, int64_t N, int64_t F,
                             cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Entropy coder (torchac-compatible 32-bit binary range coder).
 * replaces: encoder()/decoder() -> torchac.encode_float_cdf / decode_float_cdf,
 *           examples/utils_bpp_acc.py:77-110 (torchac==0.9.3, requirements.txt:32)
 * cnc_cdf_from_p: c1 = uint16(round((1-p)*65534) + 1)  (the int16 CDF entry torchac builds).
 * Streams are independent; stream k codes symbols [sym_off[k], sym_off[k+1]) and writes at
 * out + out_off[k] (capacity out_off[k+1]-out_off[k]); out_len[k] receives the byte count
 * (if it exceeds the capacity the stream is truncated and the call returns CNC_EINVAL after
 * completion -- capacity n/8*2+64 bytes is always enough for probabilities in [1e-6,1-1e-6]).
 * ---------------------------------------------------------------------------------------- */
int cnc_cdf_from_p(const float *p, uint16_t *c1, uint64_t n, cnc_stream_t stream);
int cnc_ac_encode(const uint16_t *c1, const uint8_t *sym, const int64_t *sym_off,
                  uint8_t *out, const int64_t *out_off, int64_t *out_len, int32_t n_streams,
                  cnc_stream_t stream);
int cnc_ac_decode(const uint16_t *c1, const int64_t *sym_off, const uint8_t *in,
                  const int64_t *in_off, const int64_t *in_len, uint8_t *sym, int32_t n_streams,
                  cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Per-sample field ops.
 * cnc_sh16 replaces tcnn.Encoding(SphericalHarmonics, degree 4) at ngp.py:412-425,541:
 *   d01 [n,3] = (dir+1)/2 -> out [n,16] f32; fp16_round != 0 rounds through fp16 like tcnn's output.
 * cnc_freq_embed replaces Embedder.embed (ngp.py:569-617): out [n, 3+6*n_freq].
 * ---------------------------------------------------------------------------------------- */
int cnc_sh16(const float *d01, float *out, uint64_t n, int fp16_round, cnc_stream_t stream);
int cnc_freq_embed(const float *x, float *out, uint64_t n, int n_freq, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused radiance-field forward (product layout: F=8, 3D 12 levels + 3 planes x 4 levels,
 * n_neurons=160, geo_feat_dim=79).
 * replaces: NGPRadianceField_mygrid_2D3D.query_density / _query_rgb / forward
 *           examples/radiance_fields/ngp.py:514-566, compose_3D_2D_embed.forward ngp.py:629-645
 *           (4 x _grid_encode + Embedder + torch.cat + 5 nn.Linear + trunc_exp + sigmoid).
 * cnc_field_pack_weights: nn.Linear weights ([out,in] row-major) and biases of
 *   mlp_base.network[0] (160x255), [2] (80x160), mlp_head[0] (160x95), [2] (160x160), [4] (3x160)
 *   -> blob of cnc_field_blob_floats() floats (tf32 hi/lo split, 128B-swizzled K chunks).
 * cnc_field_fwd: pos [N,3] world coords, dirs [N,3] (NULL = density only), aabb6 = 6 HOST floats
 *   (min xyz, max xyz), bits_* = cnc_sign_pack tables of the four encoders, offsets/resolutions
 *   of the 3D (12+1 / 12 entries) and 2D (4+1 / 4) encoders -> sigma [N], rgb [N,3],
 *   geo [N,79] (nullable).  fp32 results via 3xTF32 tcgen05 MMA with fp32 accumulation.
 * ---------------------------------------------------------------------------------------- */
uint32_t cnc_field_blob_floats(void);
/* profiling aid: 128 x uint64 device buffer that receives clock64() stamps of one tile (NULL = off) */
int cnc_field_set_timeline_buffer(uint64_t *device_buf);
int cnc_field_pack_weights(const float *W1, const float *b1, const float *W2, const float *b2,
                           const float *W3, const float *b3, const float *W4, const float *b4,
                           const float *W5, const float *b5, float *blob, cnc_stream_t stream);
int cnc_field_fwd(const float *pos, const float *dirs, const float *aabb6_host,
                  const uint8_t *bits_xyz, const uint8_t *bits_xy, const uint8_t *bits_xz,
                  const uint8_t *bits_yz, const int32_t *offsets3, const int32_t *resolutions3,
                  const int32_t *offsets2, const int32_t *resolutions2, const float *blob,
                  float *sigma, float *rgb, float *geo, uint32_t N, cnc_stream_t stream);
/* Host-buffer forward (the call a host-side caller of the reference makes: positions and directions in, colours and
 * densities out, all in host memory -- pinned, for the copies to be asynchronous).  ONE launch of the persistent kernel
 * on s_compute; the samples are uploaded on s_in in chunks of 1, 2, 4, .. max_chunk_waves .. 4, 2, 1 waves
 * (wave_samples = SMs x 128), each followed by a per-wave "ready" stamp the kernel waits for before it reads the wave;
 * CTAs count finished tiles per wave and the downloads on s_out wait for those counts with cuStreamWaitValue32, so only
 * a short first upload and last download are exposed.  d_pos / d_dirs / d_rgb [N,3], d_sigma [N] are device staging
 * buffers owned by the caller.  Returns after everything is enqueued; s_compute completes after the last download.
 * All device-side waits are bounded (2 s), a missing stamp cannot hang the GPU.
 * Ready flags and tile counters are kept per staging buffer (keyed by d_pos, up to 4): calls that use different staging
 * buffers and different s_compute streams (same s_in / s_out) overlap -- the first upload of one call runs beside the
 * kernel of the other -- which is how a stream of independent batches reaches the device-resident rate. */
int cnc_field_fwd_host(const float *pos_host, const float *dirs_host, const float *aabb6_host,
                       const uint8_t *bits_xyz, const uint8_t *bits_xy, const uint8_t *bits_xz,
                       const uint8_t *bits_yz, const int32_t *offsets3, const int32_t *resolutions3,
                       const int32_t *offsets2, const int32_t *resolutions2, const float *blob,
                       float *sigma_host, float *rgb_host, uint32_t N, float *d_pos, float *d_dirs,
                       float *d_sigma, float *d_rgb, uint32_t wave_samples, uint32_t max_chunk_waves,
                       cnc_stream_t s_compute, cnc_stream_t s_in, cnc_stream_t s_out);
/* Training forward: the same kernel, additionally leaving what the backward pass of ngp.py:514-566 needs in HBM as
 * the values go by (no recomputation, no extra pass): x0 [N,256] = input of Linear(255,160) (192 grid features |
 * x, sin/cos | a constant 1 in the pad column), h1 / h3 / h4 [N,160] = the ReLU outputs, head_in [N,96] = the input of the
 * first head layer in the kernel's column order (SH band 0 | 79 geo features | SH bands 1..15 | 0: the rows of that layer's
 * weight gradient come out in the same order); all 32-byte aligned (the rows leave as 256-bit stores), all required. */
int cnc_field_fwd_train(const float *pos, const float *dirs, const float *aabb6_host,
                        const uint8_t *bits_xyz, const uint8_t *bits_xy, const uint8_t *bits_xz,
                        const uint8_t *bits_yz, const int32_t *offsets3, const int32_t *resolutions3,
                        const int32_t *offsets2, const int32_t *resolutions2, const float *blob,
                        float *sigma, float *rgb, float *head_in, float *x0, float *h1, float *h3, float *h4,
                        uint32_t N, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Weight gradients of the field MLPs (backward of the nn.Linear layers of ngp.py:428-505; the reference leaves it to
 * torch autograd = an fp32 cuBLAS GEMM):  C[i, o] = sum_s X[s, i] * Z[s, o],  X = layer input [Ns, ldx] (first Mi
 * columns, Mi a multiple of 32, <= 256), Z = gradient of the layer output [Ns, ldz] (first No columns, No a multiple
 * of 16, <= 160), fp32 row-major, 16-byte aligned.  3xTF32 tcgen05 MMAs with fp32 accumulation (fp32-equivalent).
 * with_ones != 0 appends a virtual all-ones column to X (built in shared memory, no HBM traffic): row Mi of the result
 * is the column sum of Z, i.e. the bias gradient.  Instantiated (Mi, No): (256,160) (160,160) (96,160) (160,80) (160,16)
 * -- the five layers of the field -- plus (32,16) (64,32) for tests.
 * The kernel writes n_partials partial sums [n_partials, Mi (+1), No] (one per CTA, n_partials <=
 * cnc_wgrad_max_partials()); the caller adds them up in index order, which makes the result deterministic.
 * ---------------------------------------------------------------------------------------- */
int cnc_wgrad_max_partials(void);
int cnc_wgrad(const float *X, uint32_t ldx, uint32_t Mi, const float *Z, uint32_t ldz, uint32_t No,
              int with_ones, float *partials, uint32_t n_partials, uint32_t Ns, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Input gradients of the field MLPs with the ReLU mask fused (backward of nn.Linear + ReLU, ngp.py:428-505; the
 * reference leaves it to torch autograd = fp32 cuBLAS GEMM + threshold_backward):
 *   C[s, n] = [H[s, n] > 0] * sum_o Z[s, o] * W[o, n + col_off]      for n_first <= n < n_valid, 0 elsewhere (n < N)
 * cnc_dgrad_pack turns the nn.Linear weight W [No, ldw] (row-major, [out, in]) into the streamed operand
 * (cnc_dgrad_blob_floats(No, N) floats, once per step); cnc_dgrad runs one layer: Z [Ns, ldz] (first No columns,
 * No <= 160, multiple of 4), H [Ns, ldh] nullable (the ReLU output that fed the layer), C [Ns, ldc] (first N columns,
 * N in {32, 80, 160, 192}).  Error-compensated tf32 tcgen05 MMAs with fp32 accumulation (fp32-equivalent).
 * ---------------------------------------------------------------------------------------- */
uint32_t cnc_dgrad_blob_floats(uint32_t No, uint32_t N);
int cnc_dgrad_pack(const float *W, uint32_t ldw, uint32_t No, int32_t col_off, uint32_t n_first,
                   uint32_t n_valid, uint32_t N, float *blob, cnc_stream_t stream);
int cnc_dgrad(const float *Z, uint32_t ldz, uint32_t No, const float *blob, uint32_t N, const float *H,
              uint32_t ldh, float *C, uint32_t ldc, uint32_t Ns, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Level-wise context model of the 3D grid, fused (one kernel per coded chunk).
 * replaces the chunk body of encode_/decode_binary_vxl_mixPg_3D2D, examples/utils_bpp_acc.py:798-852
 *   == :929-968: query_mask_3D -> compaction -> align_and_pack -> Encoding_xyz(points, n-3, n,
 *   binary_vxl) -> context_model_3D -> align_and_pack -> overlap-weighted sum -> clamp.
 *   pts [Nv,3] i16 voxel coords of level `level`, grouped by hash entry; seg [n_entries+1] i64 running
 *   voxel counts with seg[0] == seg_base (a slice of unique_count_cumsum, utils_bpp_acc.py:803-808);
 *   sign_bits = cnc_sign_pack of the whole 3D table (encoder: STE(params); decoder: the partially
 *   reconstructed table); mlp_packed = cnc_context3d_mlp_floats() floats:
 *   W1^T [25][32], b1 [32], W2^T [32][32], b2 [32], W3^T [32][8], b3 [8] of context_model_3D.
 *   -> prob [n_entries,8] (clamped to [1e-6, 1-1e-6]; 0 for entries that are not coded),
 *      mean [n_entries,8] unclamped (nullable), exist [n_entries] u8 (mask_exist, :823).
 *   entry_base = index inside the level of the chunk's first entry (`lo` of utils_bpp_acc.py:798-802): the kernel
 *   works in batches of 64 entries aligned to the absolute index; the summation order per entry -- and with it
 *   every probability bit -- is a function of the voxel list and of the chunk boundaries only (never of the grid
 *   size or the GPU count), and cuts that fall on the batch grid do not change a bit.
 *   vertex_bits / vertex_bit_offsets (nullable, together): the per-vertex occupancy predicate of the masked
 *   gather (gridencoder.cu:221-276) precomputed by cnc_vertex_valid_bits for all levels of the encoder: one bit
 *   per grid vertex, bit (c0*res + c1)*res + c2 of level l at vertex_bit_offsets[l] (multiples of 32);
 *   the occupancy does not change during an encode/decode, so the box test runs once per vertex instead of
 *   once per (voxel, corner, context level).  Results are identical with and without it.
 * ---------------------------------------------------------------------------------------- */
uint32_t cnc_context3d_mlp_floats(void);
int cnc_vertex_valid_bits(const uint8_t *binary_vxl, int32_t Rb, const int32_t *resolutions,
                          int32_t n_levels, const int64_t *bit_offsets, int64_t total_bits,
                          uint32_t *out_words, cnc_stream_t stream);
int cnc_context3d_probs(const int16_t *pts, const int64_t *seg, int64_t n_entries,
                        const uint8_t *binary_vxl, int32_t Rb, const uint8_t *sign_bits,
                        const int32_t *offsets, const int32_t *resolutions, int32_t level, float Pg,
                        const float *mlp_packed, float *prob, float *mean, uint8_t *exist,
                        int64_t seg_base, int64_t entry_base, const uint32_t *vertex_bits,
                        const int64_t *vertex_bit_offsets, cnc_stream_t stream);

/* context_model_3D (Linear(25,32) LeakyReLU Linear(32,32) LeakyReLU Linear(32,8), utils_bpp_acc.py:247-251) for the rate
 * term of the training loss (utils_bpp_acc.py:533-706), where the reference runs it under autograd (six cuBLAS GEMMs +
 * elementwise passes over [voxels, 32] activations).  One forward and one backward kernel, thread = voxel:
 *   cnc_ctx_mlp_fwd: X [M,25] row-major, mlp_packed (layout of cnc_context3d_probs, cnc_ctx_mlp_floats() floats) -> Y [M,8]
 *   cnc_ctx_mlp_bwd: recomputes the hidden activations, gY [M,8] -> gX [M,25] and n_partials partial gradient vectors
 *                    [n_partials, cnc_ctx_mlp_floats()] in the packed layout (every CTA writes one; the caller adds them
 *                    in index order: deterministic).  n_partials <= cnc_ctx_mlp_max_partials() is a sensible grid. */
uint32_t cnc_ctx_mlp_floats(void);
int cnc_ctx_mlp_max_partials(void);
int cnc_ctx_mlp_fwd(const float *X, const float *mlp_packed, float *Y, int64_t M, cnc_stream_t stream);
int cnc_ctx_mlp_bwd(const float *X, const float *mlp_packed, const float *gY, float *gX, float *partials,
                    uint32_t n_partials, int64_t M, cnc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Occupancy-grid ray marching, packed scans, volume rendering (vendored nerfacc 0.5.3).
 * replaces: nerfacc.csrc ray_aabb_intersect / traverse_grids  nerfacc/cuda/csrc/grid.cu:320-349, :68-318
 *           (python: nerfacc/grid.py:20-91, :94-194); exclusive/inclusive_sum/prod scan.cu:9-304
 *           (nerfacc/scan.py); render_weight_from_density + accumulate_along_rays volrend.py:314-364,485-549.
 * cnc_traverse_grids is called twice like the reference host code (grid.cu:441-507): a count pass
 *   (t_starts == chunk_starts == NULL, cnt filled) and, after an exclusive scan of cnt, a fill pass.
 *   hits [n_rays,n_grids] u8, t_sorted / t_indices [n_rays, 2*n_grids] = sorted (t_min | t_max) of
 *   cnc_ray_aabb_intersect; binaries [n_grids,rx,ry,rz] bool.  Samples come out as (t_start, t_end, ray).
 * cnc_packed_scan: op 0 sum / 1 prod over the chunks packed_info [n_rays,2] (start, count).
 * cnc_render_from_density: per ray, weights = exp(-exclusive_sum(sigma*dt)) * (1 - exp(-sigma*dt)) and the
 *   accumulations sum(w*rgb), sum(w), sum(w*t_mid); every output pointer is nullable.
 * ---------------------------------------------------------------------------------------- */
int cnc_ray_aabb_intersect(const float *rays_o, const float *rays_d, int64_t n_rays, float near_plane,
                           float far_plane, const float *aabbs, int32_t n_aabbs, float miss_value,
                           float *t_mins, float *t_maxs, uint8_t *hits, cnc_stream_t stream);
int cnc_traverse_grids(const float *rays_o, const float *rays_d, const uint8_t *rays_mask, int64_t n_rays,
                       int32_t n_grids, int32_t rx, int32_t ry, int32_t rz, const uint8_t *binaries,
                       const float *aabbs, const uint8_t *hits, const float *t_sorted,
                       const int64_t *t_indices, const float *near_planes, const float *far_planes,
                       float step_size, float cone_angle, int32_t steps_limit,
                       const int64_t *chunk_starts, int64_t *cnt, float *t_starts, float *t_ends,
                       int64_t *ray_indices, float *terminate_planes, cnc_stream_t stream);
int cnc_packed_scan(const float *in, const int64_t *packed_info, int64_t n_rays, float *out, int32_t op,
                    int32_t inclusive, int32_t reverse, cnc_stream_t stream);
int cnc_render_from_density(const float *t_starts, const float *t_ends, const float *sigmas,
                            const float *rgbs, const int64_t *packed_info, int64_t n_rays,
                            const float *prefix_trans, float *weights, float *trans, float *alphas,
                            float *colors, float *opacities, float *depths, cnc_stream_t stream);
/* backward of cnc_render_from_density (weights, colours, opacities, weighted depths) in one pass per ray: the analytic
 * gradient nerfacc's autograd assembles from render_weight_from_density + accumulate_along_rays (volrend.py:14-160);
 * g_colors [R,3] / g_opacities [R] / g_depths [R] / g_weights [n] nullable (= zero) */
int cnc_render_bwd(const float *t_starts, const float *t_ends, const float *trans, const float *alphas, const float *weights,
                   const float *rgbs, const int64_t *packed_info, int64_t n_rays, const float *g_colors, const float *g_opacities,
                   const float *g_depths, const float *g_weights, float *g_sigmas, float *g_rgbs, cnc_stream_t stream);
/* query points of ray samples: positions[i] = o[r] + d[r] * (t0 + t1) / 2, dirs[i] = d[r], r = ray_indices[i]
 * (examples/utils.py:250-262, the sigma_fn / rgb_sigma_fn closures); dirs nullable */
int cnc_sample_points(const float *rays_o, const float *rays_d, const int64_t *ray_indices, const float *t_starts, const float *t_ends,
                      int64_t n, float *positions, float *dirs, cnc_stream_t stream);
/* One-walk marching: cnc_traverse_grids called with chunk_starts[r] = r * cap, steps_limit = cap, cnt AND the fill outputs
 * (ray_indices may be NULL) leaves ray r's samples at scratch[r * cap ..] and its count; this moves them to the packed layout
 * of packed_info [n_rays,2] = (exclusive prefix sum of the counts, counts) and writes ray_indices.  The caller picks cap as
 * an upper bound of the samples of one ray (box diagonal / step) and falls back to count + fill when a count reaches it. */
int cnc_pack_ray_chunks(const float *scratch_t_starts, const float *scratch_t_ends, int64_t cap, const int64_t *packed_info,
                        int64_t n_rays, float *t_starts, float *t_ends, int64_t *ray_indices, cnc_stream_t stream);
/* samples with keep[i] != 0 move to slot rank[i] - 1 (rank = inclusive int64 prefix sum of keep), order kept: the
 * `t_starts[masks], t_ends[masks], ray_indices[masks]` of OccGridEstimator.sampling (occ_grid.py:192-197);
 * out_packed_info [n_rays,2] (nullable) = (start, count) per ray of what is kept, from packed_info [n_rays,2] of the input
 * (= pack_info(ray_indices[masks]), pack.py:9-37) */
int cnc_compact_samples(const uint8_t *keep, const int64_t *rank, int64_t n, const float *t_starts, const float *t_ends,
                        const int64_t *ray_indices, float *out_t_starts, float *out_t_ends, int64_t *out_ray_indices,
                        const int64_t *packed_info, int64_t n_rays, int64_t *out_packed_info, cnc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CNC_B200_H */
