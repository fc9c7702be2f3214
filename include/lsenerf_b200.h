/*
 * lsenerf_b200.h -- C ABI of the B200-native LSENeRF per-ray hot path (liblsenerf_b200.so).
 *
 * Plain pointers and sizes only; every pointer is DEVICE memory unless marked [host].  All entry points
 * enqueue work on `stream` (a cudaStream_t passed as void*) and return immediately: 0 on success, a
 * negative code on error with a message retrievable through lse_last_error().  Nothing throws across the
 * ABI; there is no CPU fallback.  Each function cites the reference interface it replaces (paths relative
 * to ubc-vision/LSENeRF; "3P" = third-party kernel the reference reaches at that call site).
 *
 * Ragged sizes: sample-level kernels take a capacity `S_cap` and an optional device pointer `n_dev` holding
 * the live sample count (written by lse_scan_counts); when n_dev == NULL all S_cap rows are live.  This
 * keeps the whole step free of host synchronisation (the reference syncs with .item(),
 * lse_grid_estimator.py:112, lse_cameras.py:360-362).
 */
#ifndef LSENERF_B200_H
#define LSENERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSE_ABI_VERSION 1

typedef void *lse_stream_t; /* cudaStream_t */

int lse_abi_version(void);
const char *lse_last_error(void);
/* number of kernel launches issued through this library since load (bench.py: gpu_launches) */
int64_t lse_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * (a) ray generation
 * ------------------------------------------------------------------------------------------------- */

/* SplineCameraOptimizer.get_rgb_cameras / get_evs_cameras / get_deblur_cameras
 * (lse_nerf/ns_camera_optimizer.py:130-197) = exp_map_to_quat_map + vectorized_generalized_interpolation +
 * slerp + quat_map_to_mtx (lse_nerf/interpolation_utils.py:56-246).
 *   ctrl_tangents [K,6] (t, rotvec), ctrl_ts [K], times [n]  ->  c2w [n*n_virtual,3,4]
 *   n_virtual = 1: pose at clip(times[i]); n_virtual = 4 (deblur): (times[i]-t_offset) + t_step*k, pixel-major,
 *   with t_offset = exp_t/2 and t_step = exp_t/(n_virtual-1) rounded to fp32 by the caller (as torch does).
 *   dM [4,4] or NULL: right-multiply the relative event-camera transform; scale [1] or NULL scales dM[:3,3].
 *   seg_out [n*n_virtual] int32 or NULL: the control segment chosen (bit-exact parity hook). */
int lse_pose_interp_fwd(const float *ctrl_tangents, const float *ctrl_ts, int K, const float *times, int64_t n,
                        int n_virtual, float t_offset, float t_step, const float *dM, const float *scale,
                        float *c2w, int32_t *seg_out, lse_stream_t stream);
/* backward: d_c2w [n*n_virtual,3,4] -> d_ctrl_tangents [K,6] (atomically accumulated), d_scale [1] (accumulated,
 * may be NULL). */
int lse_pose_interp_bwd(const float *ctrl_tangents, const float *ctrl_ts, int K, const float *times, int64_t n,
                        int n_virtual, float t_offset, float t_step, const float *dM, const float *scale,
                        const float *d_c2w, float *d_ctrl_tangents, float *d_scale, lse_stream_t stream);

/* (a1) EvPixelSampler.collate_image_dataset_batch (lse_nerf/lse_pixel_sampler.py:6-48) on image stacks resident in
 * HBM: rand3 [B,3] uniforms in [0,1) (drawn by the host generator, so the pick is bit-identical to the host sampler) ->
 * (c,y,x) = floor(rand * [N,H,W]); image_out [B,C] = images[c,y,x,:]; indices [B,3] = (image_idx[c], y, x);
 * appearance_out [B] = appearance_id[c]; e_thresh_out [B] = e_thresh[c] (both optional). */
int lse_pixel_gather(const float *rand3, int64_t B, int64_t N, int H, int W, const float *images, int C,
                     const int64_t *image_idx, const int64_t *appearance_id, const float *e_thresh, int64_t *indices,
                     float *image_out, int64_t *appearance_out, float *e_thresh_out, lse_stream_t stream);

/* EdCameras._generate_rays_from_coords, perspective branch (lse_nerf/lse_cameras.py:349-414,487-493,539-561).
 *   c2w [n,3,4], coords [n,2]=(y,x)  ->  origins [n,3], directions [n,3], pixel_area [n], dir_norm [n]
 *   distortion: NULL, or 6 floats (k1,k2,k3,k4,p1,p2) -- the OpenCV lens model the scene files carry (lse_parser.py:187-198): the
 *   pixel and its two +1 px neighbours are undistorted by nerfstudio's radial_and_tangential_undistort (10 Newton steps,
 *   lse_cameras.py:397-414) before they become directions. */
int lse_rays_from_c2w_fwd(const float *c2w, const float *coords, int64_t n, float fx, float fy, float cx, float cy,
                          const float *distortion, float *origins, float *directions, float *pixel_area, float *dir_norm,
                          lse_stream_t stream);
int lse_rays_from_c2w_bwd(const float *c2w, const float *coords, int64_t n, float fx, float fy, float cx, float cy,
                          const float *distortion, const float *d_origins, const float *d_directions, float *d_c2w,
                          lse_stream_t stream);

/* CameraOptimizer.forward + apply_to_raybundle, mode SO3xR3 (lse_nerf/ns_camera_optimizer.py:258-329):
 *   origins += t(pose_adjustment[cam]); directions = R(pose_adjustment[cam]) @ directions.  cam_idx int64 [n]. */
int lse_pose_delta_fwd(const float *pose_adjustment, int64_t n_cams, const int64_t *cam_idx, int64_t n,
                       const float *origins, const float *directions, float *origins_out, float *directions_out,
                       lse_stream_t stream);
int lse_pose_delta_bwd(const float *pose_adjustment, int64_t n_cams, const int64_t *cam_idx, int64_t n,
                       const float *directions, const float *d_origins_out, const float *d_directions_out,
                       float *d_pose_adjustment, float *d_origins, float *d_directions, lse_stream_t stream);

/* The same application with the [n_cams,3,4] correction matrices given explicitly (CameraOptimizer mode "SE3",
 * lse_nerf/ns_camera_optimizer.py:276-277: nerfstudio exp_map_SE3 of the pose table is evaluated per CAMERA by the host mirror, these
 * kernels apply it per RAY): origins += M[cam][:,3]; directions = M[cam][:,:3] @ directions.  Backward: d_mats [n_cams,3,4]
 * accumulated atomically, d_origins / d_directions written. */
int lse_apply_delta_fwd(const float *mats, int64_t n_cams, const int64_t *cam_idx, int64_t n, const float *origins,
                        const float *directions, float *origins_out, float *directions_out, lse_stream_t stream);
int lse_apply_delta_bwd(const float *mats, int64_t n_cams, const int64_t *cam_idx, int64_t n, const float *directions,
                        const float *d_origins_out, const float *d_directions_out, float *d_mats, float *d_origins,
                        float *d_directions, lse_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * (b) occupancy-grid marching  (3P nerfacc.grid.traverse_grids, lse_nerf/lse_grid_estimator.py:83-106)
 * ------------------------------------------------------------------------------------------------- */

/* Pack the byte occupancy tensor (nerfacc `binaries` [levels,res,res,res] bool) into 1 bit per cell (occ_bits,
 * levels*res^3/32 words; brick-major: one 64-bit word per 4x4x4 brick of cells, opaque to the caller) plus an any-occupied mip over 8^3 superblocks (occ_mip, ceil(levels*(res/8)^3/32) words) that
 * the marching kernels keep in shared memory.  res % 32 == 0. */
int lse_occ_pack(const uint8_t *binaries, int levels, int res, uint32_t *occ_bits, uint32_t *occ_mip, lse_stream_t stream);

/* near = near_plane (+ jitter[r]*step_size when jitter != NULL: the stratified draw of :91-92).
 * pass 1: counts[r] = number of samples of ray r. */
int lse_march_count(const float *rays_o, const float *rays_d, int64_t n_rays, const uint32_t *occ_bits,
                    const uint32_t *occ_mip, const float *aabbs, int levels, int res, float near_plane, float far_plane,
                    const float *jitter, float step_size, float cone_angle, int32_t *counts, lse_stream_t stream);
/* exclusive scan of counts -> starts [n] ; total -> total_dev [1] (int32).  Also used for compaction. */
int lse_scan_counts(const int32_t *counts, int64_t n, int32_t *starts, int32_t *total_dev, lse_stream_t stream);
/* The same scan against a sample capacity (the reference has none: nerfacc sizes its outputs from a host read of the count): starts are
 * clamped to cap, counts [n] are REWRITTEN so that starts[i] + counts[i] <= cap, *total_dev = min(sum, cap), and *overflow_dev
 * (int32, may be NULL) is set to 1 when samples were dropped.  After this call no sample-level kernel can index past `cap`. */
int lse_scan_counts_capped(int32_t *counts, int64_t n, int64_t cap, int32_t *starts, int32_t *total_dev,
                           int32_t *overflow_dev, lse_stream_t stream);
/* pass 2: writes ray_indices/t_starts/t_ends at starts[r]; rays whose samples would exceed S_cap are truncated
 * (overflow_dev [1] int32 set to 1). */
int lse_march_fill(const float *rays_o, const float *rays_d, int64_t n_rays, const uint32_t *occ_bits,
                   const uint32_t *occ_mip, const float *aabbs, int levels, int res, float near_plane, float far_plane,
                   const float *jitter, float step_size, float cone_angle, const int32_t *starts, int64_t S_cap,
                   int32_t *ray_indices, float *t_starts, float *t_ends, int32_t *overflow_dev, lse_stream_t stream);
/* Single-pass variant of count+fill (same samples, bit for bit).  Kernel 1 (thread = ray x level segment) runs the DDA and
 * collapses consecutive cells of equal occupancy into (bound, occupied) runs in `workspace`
 * (`lse_march_workspace(n_rays, levels, res)` bytes); kernel 2 (thread = ray) replays the sample-step recurrence once against the
 * runs and writes (t_start, t_end) pairs to per-ray slabs `slabs` [n_rays, max_per_ray, 2] fp32 (max_per_ray >= the longest
 * possible ray, 1024 for the default cone/step); counts[r] = samples of ray r (clamped to max_per_ray, *overflow_dev = 1 if
 * clamped).  lse_scan_counts + lse_march_pack then produce the packed ray-major (ray_indices, t_starts, t_ends) of
 * lse_march_fill. */
int64_t lse_march_workspace(int64_t n_rays, int levels, int res);
int lse_march_single(const float *rays_o, const float *rays_d, int64_t n_rays, const uint32_t *occ_bits,
                     const uint32_t *occ_mip, const float *aabbs, int levels, int res, float near_plane,
                     float far_plane, const float *jitter, float step_size, float cone_angle, float *slabs,
                     int max_per_ray, int32_t *counts, int32_t *overflow_dev, void *workspace, int64_t workspace_bytes,
                     lse_stream_t stream);
int lse_march_pack(const float *slabs, int max_per_ray, const int32_t *counts, const int32_t *starts, int64_t n_rays,
                   int64_t S_cap, int32_t *ray_indices, float *t_starts, float *t_ends, int32_t *overflow_dev,
                   lse_stream_t stream);

/* One generation of the early-terminating culling pass (exactly the kept set of render_visibility_from_density, lse_grid_estimator.py:
 * 109-127, without evaluating the field behind opaque surfaces): for every ray still visible after its first `lo` samples
 * (transmittance from depth_acc [n_rays] + sigmas over [prev_lo, lo) >= early_stop_eps) append the ids of its samples [lo, hi) to
 * sample_list and bump *n_list_dev.  Call with prev_lo == lo == 0 for the first generation: it also zero-fills the sigmas of every ray's samples
 * behind `hi` (samples never evaluated keep sigma = 0, which the visibility test rejects), so the buffer needs no memset. */
int lse_cull_generation(const int32_t *starts, const int32_t *counts, int64_t n_rays, const float *t_starts,
                        const float *t_ends, const float *sigmas, float *depth_acc, int prev_lo, int lo, int hi,
                        float early_stop_eps, int32_t *sample_list, int64_t list_cap, int32_t *n_list_dev, lse_stream_t stream);

/* 3P nerfacc.render_visibility_from_density (lse_nerf/lse_grid_estimator.py:109-127,139-143):
 * per ray, T = exp(-exclusive_sum(sigma*dt)), alpha = 1-exp(-sigma*dt), keep = T >= early_stop_eps &&
 * alpha >= min(alpha_thre, *occ_mean_dev).  Writes keep [S] (uint8) and keep_counts [n_rays]. */
int lse_visibility(const int32_t *starts, const int32_t *counts, int64_t n_rays, const float *t_starts,
                   const float *t_ends, const float *sigmas, float early_stop_eps, float alpha_thre,
                   const float *occ_mean_dev, uint8_t *keep, int32_t *keep_counts, lse_stream_t stream);
/* stable per-ray compaction of the kept samples (ray_indices, t_starts, t_ends, h16 [S,16], sigma [S], and -- optional -- the base-MLP
 * activation masks uint64 [S] of LseFieldParams.relu_mask_density). */
int lse_compact(const int32_t *starts, const int32_t *counts, const int32_t *new_starts, int64_t n_rays,
                const uint8_t *keep, const float *t_starts, const float *t_ends, const float *h16,
                const float *sigma, const unsigned long long *mask, int32_t *ray_indices_out, float *t_starts_out,
                float *t_ends_out, float *h16_out, float *sigma_out, unsigned long long *mask_out, lse_stream_t stream);

/* 3P nerfacc OccGridEstimator.update_every_n_steps (inherited NGP callback; grid built at lse_nerf/lsenerf.py:184-188), step 1:
 * the cells to re-evaluate and their jittered sample points, without a host read of the occupied-cell count.
 *   warmup_all != 0 (step < warmup_steps): every cell of every level, level-major (u_uniform / u_occupied / workspace unused);
 *   else per level (N = res^3 cells): N/4 uniform draws floor(u_uniform[l,i]*N), then the occupied cells of `binaries` in ascending
 *   order -- all of them if at most N/4, otherwise N/4 draws occupied[floor(u_occupied[l,i]*n_occupied)]
 *   (nerfacc _sample_uniform_and_occupied_cells with the uniforms supplied by the caller's generator: u_* are [levels, N/4] in [0,1)).
 * Outputs compacted level-major with the live count in *n_out_dev: cell_out [cap] int32 = level*N + cell, pos_out [cap,3] =
 * aabb_lo + (cell coords + u_jitter[j]) / res * (aabb_hi - aabb_lo); cap = levels*N (warm-up) or levels*2*(N/4).  res % 16 == 0.
 * workspace: lse_occ_select_workspace(levels, res) bytes. */
int64_t lse_occ_select_workspace(int levels, int res);
int lse_occ_select_cells(const uint8_t *binaries, int levels, int res, int warmup_all, const float *u_uniform,
                         const float *u_occupied, const float *u_jitter, const float *aabbs, void *workspace,
                         int64_t workspace_bytes, int32_t *cell_out, float *pos_out, int32_t *n_out_dev, lse_stream_t stream);
/* ... step 2: occs[cell] = max(occs[cell]*decay, occ[i]*occ_scale) for the evaluated cells (live prefix *n_dev of cap; n_dev NULL =
 * all), then binaries = occs > min(mean(occs), occ_thre).  Deterministic: a cell listed several times decays ONCE from its old value
 * and takes the largest of its new values (the reference's indexed assignment lets an arbitrary duplicate win), and mean(occs) is a
 * fixed-order sum -- replicas holding identical weights derive bit-identical grids.  occ_mean_dev [1] receives mean(occs) (the value
 * lse_grid_estimator.py:112 reads with .item()).  scratch_max: int32 [n_cells_total] that the caller fills with -1 ONCE (the call
 * leaves it at -1); scratch_partials: 1024 doubles. */
int lse_occ_update(float *occs, int64_t n_cells_total, const int32_t *cell_idx, const float *occ, int64_t cap,
                   const int32_t *n_dev, float occ_scale, float decay, float occ_thre, uint8_t *binaries,
                   float *occ_mean_dev, int32_t *scratch_max, double *scratch_partials, lse_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * (c) multiresolution hash grid (3P nerfstudio HashEncoding.pytorch_fwd semantics,
 *     lse_nerf/lse_field.py:43-91,208,279)
 * ------------------------------------------------------------------------------------------------- */
/* x [S,3] in [0,1), table [L*2^log2_T, 2], scalings [L] (taken from the module buffer, never recomputed)
 * -> out [S, 2L] level-major. */
int lse_hashgrid_fwd(const float *x, int64_t S, const float *table, const float *scalings, int L, int log2_T,
                     float *out, lse_stream_t stream);
/* d_out [S,2L] -> d_table (atomically accumulated, same shape as table), d_x [S,3] (written; may be NULL). */
int lse_hashgrid_bwd(const float *x, int64_t S, const float *table, const float *scalings, int L, int log2_T,
                     const float *d_out, float *d_table, float *d_x, lse_stream_t stream);
/* parity hook: the 8 table rows touched per level, idx [S,L,8] int64, corner order of pytorch_fwd. */
int lse_hashgrid_indices(const float *x, int64_t S, const float *scalings, int L, int log2_T, int64_t *idx,
                         lse_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * (c+d) field: LSEField.get_density / get_outputs (lse_nerf/lse_field.py:264-360)
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
    /* hash grid */
    const float *table; const float *scalings; int32_t levels; int32_t log2_T;
    /* base MLP 32->64->16 (nn.Linear layout [out,in] + bias), lse_field.py:199-207 */
    const float *W1, *b1, *W2, *b2;
    /* colour MLP 63->64->64->3, lse_field.py:254-262 */
    const float *W3, *b3, *W4, *b4, *W5, *b5;
    /* appearance embedding [emb_rows,32]; emb_mode 0: row app_id[ray]; 1: row 0 (GlobalEmbedding);
       2: zeros (eval "zero"); 3: explicit vector emb (1 row, eval "mean"/"param") */
    const float *emb; int32_t emb_rows; int32_t emb_mode;
    int32_t sh_shifted;        /* 1: SH evaluated on (d+1)/2 (nerfstudio torch path, lse_field.py:298) */
    int32_t use_tensor_cores;  /* 1: tcgen05/TMEM 3xTF32 MLP kernels, 0: fp32 SIMT kernels */
    /* Activation patterns of the tensor-core path (optional, may be NULL): the forward kernels WRITE the ReLU masks of the hidden layers
       (bit j = unit j active), the backward kernels READ them instead of re-deriving them from their own, differently rounded
       recomputation of the hidden layer -- a unit within rounding of zero would otherwise be differentiated as active by one pass and
       as inactive by the other.  relu_mask_density: uint64 [S] (base MLP, written at each sample's slot; lse_compact moves it with the
       sample); relu_mask_color: uint64 [S,2] (the two hidden layers of the colour MLP).  Ignored by the fp32 CUDA-core kernels, whose
       backward repeats the forward arithmetic exactly. */
    unsigned long long *relu_mask_density; unsigned long long *relu_mask_color;
    /* rgb [S,3] as lse_field_color_fwd wrote it (optional, may be NULL; read by the tensor-core lse_field_color_bwd only): the
       derivative of the output sigmoid is rgb (1 - rgb), so the backward pass need not recompute the output layer -- one MMA phase
       and one tensor-memory round trip less per tile.  NULL: recomputed. */
    const float *rgb_fwd;
    /* Hidden activations of the colour MLP (optional, may be NULL; tensor-core kernels only): lse_field_color_fwd WRITES, per 128-sample
       tile, the two hidden layers as the bf16 hi/lo operand panels of the backward kernel (64 KB per tile, lse_field_color_act_bytes);
       lse_field_color_bwd then loads them straight into shared memory (cp.async.bulk) instead of recomputing them: two MMA phases,
       two tensor-memory round trips and the operand staging of both hidden layers less per tile.  Requires rgb_fwd in the backward. */
    void *color_act;
} LseFieldParams;

typedef struct {
    float *d_table;                 /* [L*T,2] accumulated */
    float *dW1, *db1, *dW2, *db2;   /* accumulated */
    float *dW3, *db3, *dW4, *db4, *dW5, *db5;
    float *d_emb;                   /* [emb_rows,32] accumulated (NULL for emb_mode 2) */
} LseFieldGrads;

/* samples: position = origins[r] + directions[r] * (t0+t1)/2 with r = ray_idx ? ray_idx[i] : i; if directions ==
 * NULL the position is origins[r] (Field.density_fn).  -> sigma [S] (density after trunc_exp * selector) and
 * h16 [S,16] = base-MLP output (col 0 = density pre-activation, cols 1..15 = geo features); either output may be NULL. */
int lse_field_density_fwd(const LseFieldParams *p /*[host]*/, const float *origins, const float *directions,
                          const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap,
                          const int32_t *n_dev, float *sigma, float *h16, lse_stream_t stream);
/* The same forward over an explicit list of sample ids (int32 [list_cap], live prefix *n_list_dev): thread j evaluates sample
 * sample_list[j] and writes sigma / h16 at that sample's own slot.  Used by the early-terminating culling pass. */
int lse_field_density_fwd_list(const LseFieldParams *p /*[host]*/, const float *origins, const float *directions,
                               const int32_t *ray_idx, const float *t0, const float *t1, const int32_t *sample_list,
                               int64_t list_cap, const int32_t *n_list_dev, float *sigma, float *h16, lse_stream_t stream);
/* d_sigma [S], d_h16 [S,16] (col 0 ignored; geo grads in cols 1..15) -> parameter grads, and when not NULL
 * d_origins/d_directions [n_rays,3] (atomically accumulated per ray). */
int lse_field_density_bwd(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                          const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                          int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                          const float *d_h16, float *d_origins, float *d_directions, lse_stream_t stream);
/* The same backward with caller-provided scratch memory (device, `lse_field_density_bwd_workspace(S_cap)` bytes, 256-byte
 * aligned; may be NULL when p->use_tensor_cores == 0).  With tensor cores the work is split into three launches: re-encode
 * of the samples at full occupancy (feat [S,32] into the workspace), base-MLP backward on tcgen05, table scatter +
 * position gradients at full occupancy.  Replaces the same reference lines as lse_field_density_bwd (autograd of
 * lse_field.py:264-288). */
int64_t lse_field_density_bwd_workspace(int64_t S_cap);
int lse_field_density_bwd_ws(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                             const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                             int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                             const float *d_h16, float *d_origins, float *d_directions, void *workspace,
                             int64_t workspace_bytes, lse_stream_t stream);
/* The two halves of lse_field_density_bwd_ws.  `encode` re-gathers the features of the kept samples and needs nothing from the
 * backward pass (d_sigma is ignored and may be NULL: `finish` / `rays` apply it), so it may run on a side stream as soon as the
 * kept samples exist -- during the forward pass -- and be joined before `finish`.  with_jacobian != 0: encode also stores
 * d(features)/d(grid coordinate) (96 floats per sample) so that lse_field_density_bwd_rays(with_jacobian) computes the ray
 * gradients without re-gathering the table; lse_field_density_bwd_finish always re-gathers. */
int lse_field_density_bwd_encode(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                 const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                 int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                                 int with_jacobian, void *workspace, int64_t workspace_bytes, lse_stream_t stream);
/* lse_field_density_bwd_encode as a BACKGROUND launch for a side stream: at most ctas_per_sm (1..8) resident CTAs per SM, so the
 * kernels of the main stream keep their registers and issue slots (the gathers saturate the L1TEX sector rate with far fewer than
 * 64 warps per SM).  Writes no gradient and reads nothing of the backward pass. */
int lse_field_density_bwd_encode_bg(const LseFieldParams *p, const float *origins, const float *directions,
                                    const int32_t *ray_idx, const float *t0, const float *t1, int64_t S_cap,
                                    const int32_t *n_dev, const float *h16, int with_jacobian, void *workspace,
                                    int64_t workspace_bytes, int ctas_per_sm, lse_stream_t stream);
int lse_field_density_bwd_finish(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                 const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                 int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                                 const float *d_h16, float *d_origins, float *d_directions, void *workspace,
                                 int64_t workspace_bytes, lse_stream_t stream);
/* lse_field_density_bwd_finish as two calls (tensor-core path): `rays` = base-MLP backward + ray / position gradients, `table` =
 * the hash-table gradient alone, which only the optimizer needs and which may therefore run on the optimizer's stream.
 * With with_jacobian and the ray-major packing of the samples given (starts / counts [n_rays] int32, as lse_composite_fwd takes them)
 * the ray gradients are reduced by one warp per ray in a fixed order and WRITTEN to d_origins / d_directions [n_rays,3] (no
 * atomics, no zero-fill needed, bit-reproducible); d_dir_samples [S,3] (optional: the per-sample direction gradients of
 * lse_field_color_bwd(dir_per_sample = 1)) is folded into the same reduction.  starts == NULL: accumulated atomically as before. */
int lse_field_density_bwd_rays(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                               const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                               int64_t S_cap, const int32_t *n_dev, const float *h16, const float *d_sigma,
                               const float *d_h16, float *d_origins, float *d_directions, int with_jacobian,
                               const int32_t *starts, const int32_t *counts, int64_t n_rays, const float *d_dir_samples,
                               void *workspace, int64_t workspace_bytes, lse_stream_t stream);
int lse_field_density_bwd_table(const LseFieldParams *p, const LseFieldGrads *g, const float *origins,
                                const float *directions, const int32_t *ray_idx, const float *t0, const float *t1,
                                int64_t S_cap, const int32_t *n_dev, void *workspace, int64_t workspace_bytes,
                                lse_stream_t stream);
/* bytes of LseFieldParams.color_act for a capacity of S_cap samples */
int64_t lse_field_color_act_bytes(int64_t S_cap);
/* rgb [S,3] = sigmoid(head(cat[SH16(dir), geo15, emb32])).  app_id int32 [n_rays] (per RAY) or NULL. */
int lse_field_color_fwd(const LseFieldParams *p, const float *directions, const int32_t *ray_idx,
                        const int32_t *app_id, int64_t S_cap, const int32_t *n_dev, const float *h16, float *rgb,
                        lse_stream_t stream);
/* d_rgb [S,3] -> d_h16 [S,16] (cols 1..15 written, col 0 zeroed), head/embedding grads, and the direction gradient of the SH
 * encoding: dir_per_sample == 0: accumulated atomically into d_directions [n_rays,3]; != 0: written per sample into
 * d_directions [S,3] (to be reduced per ray by lse_field_density_bwd_rays). */
int lse_field_color_bwd(const LseFieldParams *p, const LseFieldGrads *g, const float *directions,
                        const int32_t *ray_idx, const int32_t *app_id, int64_t S_cap, const int32_t *n_dev,
                        const float *h16, const float *d_rgb, float *d_h16, float *d_directions, int dir_per_sample,
                        lse_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * (e) compositing (3P nerfacc.pack_info / render_weight_from_density, nerfstudio RGB/Depth/Accumulation
 *     renderers; lse_nerf/lsenerf.py:300-318, lse_nerf/lse_renderer.py:6-11)
 * ------------------------------------------------------------------------------------------------- */
/* per ray r with samples [starts[r], starts[r]+counts[r]):  w = T*alpha;  rgb_out [R,3] = sum w*rgb;
 * acc [R] = sum w; depth_num [R] = sum w*t_mid; tmid_minmax [R,2] = (min,max) sample mid-point of the ray
 * (+inf,-inf when empty) -- DepthRenderer("expected") divides by acc+1e-10 and clips to the global min/max on the
 * host side of the API; weights [S] stored.  acc/depth_num/tmid_minmax/weights may be NULL. */
int lse_composite_fwd(const int32_t *starts, const int32_t *counts, int64_t n_rays, const float *t_starts,
                      const float *t_ends, const float *sigma, const float *rgb, float *weights, float *rgb_out,
                      float *acc_out, float *depth_num_out, float *tmid_minmax, lse_stream_t stream);
/* d_rgb_out [R,3], d_acc [R] / d_depth_num [R] (may be NULL) -> d_sigma [S], d_rgb [S,3]. */
int lse_composite_bwd(const int32_t *starts, const int32_t *counts, int64_t n_rays, const float *t_starts,
                      const float *t_ends, const float *sigma, const float *rgb, const float *d_rgb_out,
                      const float *d_acc, const float *d_depth_num, float *d_sigma, float *d_rgb,
                      lse_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * (f) render epilogue and losses (lse_nerf/lsenerf.py:329-439, lse_nerf/intensity_mappers.py:64-87)
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t use_mapping;      /* LSENeRFModelConfig.use_mapping */
    int32_t map_mode;         /* 0 co_map, 1 evs_rgb, 2 rgb_evs */
    int32_t rgb_mapper;       /* 0 identity, 1 gt (x^(1/2.4)), 2 powpow */
    int32_t evs_mapper;       /* 0 none/identity, 2 powpow */
    int32_t ev_one_dim;       /* 0 off, 1 learned (softmax weights), 2 gt gray */
    int32_t training;
    int32_t deblur;           /* mean over groups of 4 rays */
    int32_t want_ev;          /* ev_out flag of get_outputs */
} LseEpilogueCfg;

/* rgb_lin [R,3] (composited) -> rgb_final [R or R/4,3], ev_linear [R,C], ev_out [R,C] (C = 1 if ev_one_dim else 3)
 * rgb_pow / evs_pow: the Powpow coefficients [1] (or NULL); rgb_to_one_w [3] raw (pre-softmax) weights or NULL. */
int lse_epilogue_fwd(const LseEpilogueCfg *cfg /*[host]*/, const float *rgb_lin, int64_t n_rays, const float *rgb_pow,
                     const float *evs_pow, const float *rgb_to_one_w, float *rgb_final, float *ev_linear,
                     float *ev_out, lse_stream_t stream);
int lse_epilogue_bwd(const LseEpilogueCfg *cfg, const float *rgb_lin, int64_t n_rays, const float *rgb_pow,
                     const float *evs_pow, const float *rgb_to_one_w, const float *d_rgb_final,
                     const float *d_ev_out, float *d_rgb_lin, float *d_rgb_pow, float *d_evs_pow,
                     float *d_rgb_to_one_w, lse_stream_t stream);

/* The learned intensity mappers MLP_Mapper (dim 1) / RGB_MLP_Mapper (dim 3), lse_nerf/intensity_mappers.py:27-62: nerfstudio
 * MLP(in_dim = out_dim = dim, num_layers = 4, layer_width = 16, ReLU, Sigmoid), torch implementation -- four nn.Linear ([out,in] + bias).
 * Weights are taken as given (the reference's identity_init pre-fit, :8-25, is a training recipe, not part of the path).
 * y [n,dim] = mapper(x [n,dim]); backward: dy -> dx (written, may be NULL) and the parameter gradients (accumulated; NULL = skipped). */
typedef struct {
    const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4;
    int32_t dim; int32_t pad;
} LseMlpMapper;
typedef struct {
    float *dW1, *db1, *dW2, *db2, *dW3, *db3, *dW4, *db4;
} LseMlpMapperGrads;
int lse_mlp_mapper_fwd(const LseMlpMapper *m /*[host]*/, const float *x, int64_t n, float *y, lse_stream_t stream);
int lse_mlp_mapper_bwd(const LseMlpMapper *m /*[host]*/, const LseMlpMapperGrads *g /*[host]*/, const float *x, const float *dy,
                       int64_t n, float *dx, lse_stream_t stream);

/* get_loss_dict (lsenerf.py:392-439): losses[0] = MSE(col_target, col_pred);
 * losses[1] = evs_weight * MSE(log(next+1e-6) - log(prev+1e-6), evs) with to_gray when C == 3.
 * Also writes the gradients of each loss w.r.t. its predictions (unit upstream gradient). */
int lse_loss_fwd(const float *col_pred, const float *col_target, int64_t n_col, const float *prev, const float *next,
                 const float *evs, int64_t n_evs, int C, float evs_weight, float *losses, float *d_col_pred,
                 float *d_prev, float *d_next, lse_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * next-row: fused Adam over a flat fp32 arena (lse_nerf/lse_config.py:29-38)
 * ------------------------------------------------------------------------------------------------- */
int lse_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int step, float grad_scale, lse_stream_t stream);

/* Same update with the step-dependent scalars read from device memory, hp_dev = {lr / (1 - beta1^t), sqrt(1 - beta2^t),
 * grad_scale, unused}: nothing that changes from step to step is baked into the launch, so the whole training step can
 * be replayed as one CUDA graph. */
int lse_adam_step_hp(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float beta1,
                     float beta2, float eps, const float *hp_dev, lse_stream_t stream);

/* Multi-GPU optimizer step with the gradient exchange folded in (replaces DDP's all-reduce, lse_pipeline.py:96-97, plus the
 * Adam step, lse_config.py:29-38).  The parameter and gradient arenas are symmetric-memory allocations of `n` floats mapped by
 * every rank: peer_params / peer_grads [host] are `world` device pointers (this rank's own at index `rank`);
 * param_multicast / grad_multicast are the NVSwitch multicast addresses of the two arenas (both NULL: plain peer loads /
 * stores).  Rank r reduces slice r of the gradients over all ranks, applies Adam to it (exp_avg / exp_avg_sq are LOCAL, only
 * the slice is touched; hp_dev as in lse_adam_step_hp with grad_scale = 1/world) and writes the new parameters into every
 * rank's arena.  The caller brackets the call with two cross-rank barriers on the same stream. */
int lse_adam_step_collective(void *const *peer_params, void *const *peer_grads, float *param_multicast,
                             const float *grad_multicast, float *exp_avg, float *exp_avg_sq, int64_t n, int rank, int world,
                             float beta1, float beta2, float eps, const float *hp_dev, lse_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * reproducible training (bench.py burn-in; tests): order-independent gradient accumulation
 * ------------------------------------------------------------------------------------------------- */

/* Deterministic mode: every atomic fp32 accumulation whose target lies inside one of the registered ranges (the flat gradient arenas
 * of engine.ParamArena; bases / nbytes / shadows are HOST arrays of n_ranges entries, n_ranges <= 4) goes to a 64-bit fixed-point
 * shadow of 2 x nbytes (zero-filled by the caller) instead, whose integer sums do not depend on the order of the additions;
 * lse_det_flush(grad, shadow, n) adds the shadow into the n fp32 gradients and clears it (run it before the optimizer).
 * n_ranges = 0 switches the mode off.  Synchronises the device; not to be called while a stream is capturing. */
int lse_set_deterministic(int n_ranges, const void *const *bases, const int64_t *nbytes, void *const *shadows);
int lse_det_flush(float *grad, void *shadow, int64_t n, lse_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * measurement (SURVEY 8d "L2 to be measured with a resident-set microbench")
 * ------------------------------------------------------------------------------------------------- */

/* Random gathers of vec_bytes-byte entries (4, 8 = one float2 corner, 16) from a resident set of table_bytes (power of two):
 * 148 * ctas_per_sm CTAs of 256 threads, 8 independent loads in flight per thread, `iters` rounds; coherent != 0 makes the lanes
 * of a warp read 32 consecutive entries (the coarse-level pattern).  loads_out (host pointer, may be NULL) receives the number of
 * loads issued; the caller times the launch with CUDA events: GB/s = loads * vec_bytes / t.  The denominator bench.py reports
 * the hash-gather kernels against, next to the HBM copy peak. */
int lse_gather_microbench(const void *table, int64_t table_bytes, int vec_bytes, int iters, int coherent, int ctas_per_sm,
                          float *sink, int64_t *loads_out, lse_stream_t stream);

/* Issue-to-completion clocks of `n_mma` tcgen05.mma instructions (M = 128, N columns) issued round-robin over `n_acc` TMEM
 * accumulators by one elected thread (n_issuers = 2: two warps, one accumulator set each) on ONE CTA; best of `reps`.
 * mode 0: bf16, A / B K-major in shared memory; 1: bf16, A / B MN-major (the weight-gradient form); 2: tf32, A in tensor memory;
 * 3: tf32, A / B in shared memory.  cycles_out (device, 2 int64 per issuer): {first issue -> completion observed, first -> last
 * issue}.  n_acc = 1 is one dependent chain (every instruction accumulates into the previous one's columns): the figure the
 * 64-wide MLP layers of field_tc*.cu are bounded by. */
int lse_mma_chain_microbench(int mode, int n_mma, int n_acc, int N, int n_issuers, int reps, int64_t *cycles_out,
                             lse_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LSENERF_B200_H */
