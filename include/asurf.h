/*
 * alphasurf_b200 -- C ABI of the B200-native replacement for the `svox2.csrc` hot path.
 *
 * Every entry point takes plain device pointers, sizes, POD option blocks and a CUDA stream (passed as
 * void*), and returns 0 on success or a non-zero code (cudaError_t value, or ASURF_E_* below);
 * asurf_last_error() gives the message.  No torch types cross this boundary.  The pybind11 functions of the
 * reference that each entry replaces are cited as /root/reference/svox2/csrc/<file>:<line>.
 *
 * Tensors keep the layouts the reference's Python API owns:
 *   links   int32 (X,Y,Z) contiguous, value = row in the data tensors or < 0 for an empty vertex
 *   density float32 (N,1);  surface float32 (N,1);  sh float32 (N,D) channel-major, D = 3*basis_dim
 * Gradients are ACCUMULATED (+=) into caller-owned buffers; mask bytes are set to 1 for touched rows.
 *
 * Threading / streams: like the reference extension (called from the Python main thread under the GIL, svox2.cpp), the
 * library is NOT re-entrant: its device workspaces and caches (pre-march lists, wavefront records, work-pyramid cache, list
 * verdicts) are per process.  Calls may be issued on any stream, but calls that share a workspace must be ordered with
 * respect to each other: render calls (surf_trav / cuvol / msi) form one family, the regularisers another, the optimizer
 * and exchange helpers use no workspace.  alphasurf_b200/dist.py relies on exactly this split (render lane and regulariser
 * lane on two streams).
 */
#ifndef ASURF_H
#define ASURF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASURF_ABI_VERSION 3

enum {
    ASURF_OK = 0,
    ASURF_E_INVALID = -1,      /* bad argument (the reference raises TORCH_CHECK errors here) */
    ASURF_E_UNSUPPORTED = -2,  /* feature outside the hot path (MSI background, non-SH basis, ...) */
    ASURF_E_NOMEM = -3
};

/* include/data_spec.hpp:39-56 (SparseGridSpec) / include/data_spec_packed.cuh:11-83 */
typedef struct {
    const int32_t *links;
    int32_t size[3];
    const float *density;
    const float *surface;        /* NULL when surface_type == SURFACE_TYPE_NONE */
    const float *sh;
    const float *level_set;      /* device (L,) */
    int32_t level_set_num;
    int32_t basis_dim;
    int32_t sh_dim;
    int64_t capacity;            /* N */
    float offset[3];             /* SparseGridSpec._offset  (host values) */
    float scaling[3];            /* SparseGridSpec._scaling (host values) */
    float fake_sample_std;
    float truncated_vol_render_a;
    /* Optional occupancy pyramid built by asurf_accel_build() for exactly these `links`
     * (NULL: the library builds a transient one for this call). */
    const uint64_t *accel;
    /* Optional work pyramid built by asurf_work_build() for exactly this grid content (links, density, surface,
     * level sets) and the render options of the call (NULL: the library builds a transient one per call). */
    const uint64_t *work;
    /* Multi-sphere-image background (SparseGridSpec.background_links / background_data, data_spec.hpp:47-48): links int32
     * (2 reso, reso), value = row of background_data or < 0; data float32 (n, nlayers, 4) = r, g, b, sigma per layer.
     * NULL / 0 layers: no background (the renders then end on opt.background_brightness). */
    const int32_t *background_links;
    const float *background_data;
    int32_t background_reso;
    int32_t background_nlayers;
} asurf_grid_t;

/* include/data_spec.hpp:168-201 (RenderOptions); bools widened to int32 */
typedef struct {
    float background_brightness;
    float step_size;
    float sigma_thresh;
    float stop_thresh;
    float near_clip;
    int32_t use_spheric_clip;
    int32_t last_sample_opaque;
    int32_t surf_fake_sample;
    float surf_fake_sample_min_vox_len;
    int32_t limited_fake_sample;
    int32_t no_surf_grad_from_sh;
    int32_t alpha_activation_type;
    int32_t fake_sample_l_dist;
    int32_t fake_sample_normalize_surf;
    int32_t only_outward_intersect;
    int32_t truncated_vol_render;
    float trunc_vol_weight_min;
} asurf_opt_t;

/* include/data_spec.hpp:139-147 (RaysSpec) */
typedef struct {
    const float *origins;        /* (Q,3) */
    const float *dirs;           /* (Q,3) */
    int64_t n_rays;              /* Q */
} asurf_rays_t;

/* include/data_spec.hpp:83-122 (GridOutputGrads) */
typedef struct {
    float *grad_density;         /* (N,1) */
    float *grad_surface;         /* (N,1) or NULL */
    float *grad_sh;              /* (N,D) */
    float *grad_fake_sample_std; /* (1,) or NULL */
    uint8_t *mask;               /* (N,) bool or NULL */
    float *grad_background;      /* (n, nlayers, 4) or NULL: GridOutputGrads.grad_background_out */
    uint8_t *mask_background;    /* (n, nlayers) bool or NULL: GridOutputGrads.mask_background_out */
} asurf_grads_t;

/* scalar arguments of volume_render_surf_trav_fused, render_lerp_kernel_surf_trav.cu:3802-3828 */
typedef struct {
    float beta_loss;
    float sparsity_loss;
    float fused_surf_norm_reg_scale;     /* must be 0: the reference asserts (:2872-2875) */
    float lambda_l2;
    float lambda_l1;
    float lambda_l_dist;
    float lambda_l_entropy;
    int32_t no_norm_weight_l_entropy;
    float lambda_l_dist_a;
    float lambda_l_entropy_a;
    float lambda_l_samp_dist;
    float lambda_l_di;
    float l_di_alpha_thresh;
    float surf_sparse_alpha_thresh;
    float lambda_inplace_surf_sparse;
    float lambda_inwards_norm_loss;
    float lambda_conv_mode_samp;
    int32_t l_dist_max_sample;
    /* Ray count used to normalise the losses (2/(3Q), lambda/Q).  0 means "this call's Q" (the reference
     * behaviour, :3308-3309, :3896-3908); the ray-sharded multi-GPU wrapper passes the global batch size. */
    int64_t norm_rays;
} asurf_fused_t;

/* per-call statistics (SURVEY.md 8d counters), filled when a non-NULL device pointer is passed */
typedef struct {
    unsigned long long n_steps;      /* DDA voxel visits actually executed */
    unsigned long long n_skips;      /* hierarchical empty-block skips */
    unsigned long long n_linked;     /* visited voxels with all 8 links >= 0 (Nl) */
    unsigned long long n_active;     /* ... that also pass the density gate and need work (Na) */
    unsigned long long n_samples;    /* composited samples (S) */
    unsigned long long n_ref_steps;  /* reserved */
} asurf_stats_t;

const char *asurf_last_error(void);
int asurf_abi_version(void);

/* ---- occupancy pyramid (ours; the reference marches voxel by voxel with USE_ACC_SKIP=false,
 *      render_lerp_kernel_surf_trav.cu:31) ---- */
int64_t asurf_accel_words(const int32_t size[3]);            /* number of uint64 words needed (3 levels + block list) */
int asurf_accel_build(const int32_t *links, const int32_t size[3], uint64_t *accel_out, void *stream);

/* ---- work pyramid (ours): same layout as the occupancy pyramid, bit set iff the voxel can contribute a sample
 *      under `opt`: all 8 links >= 0, the 8-corner density gate passes (render_lerp_kernel_surf_trav.cu:230-239)
 *      and a level set lies inside the corner range (:273-277) -- or fake samples are taken everywhere (:423).
 *      The marcher visits every voxel of the reference DDA but touches grid data only where the bit is set. ---- */
int asurf_work_build(const asurf_grid_t *grid, const asurf_opt_t *opt, uint64_t *work_out, void *stream);

/* ---- surf_trav renderer ---- */
/* volume_render_surf_trav, render_lerp_kernel_surf_trav.cu:3596-3654 (forward only; rgb_out (Q,3)) */
int asurf_surf_trav_forward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                            float *rgb_out, asurf_stats_t *stats_dev, void *stream);
/* Scalar renders of the surf_trav backend, one entry for the five reference functions (host side :3944-4050+, kernels
 * :3458-3560):  volume_render_expected_term_surf_trav (mode EXPECTED_TERM), volume_render_mode_term_surf_trav (MODE_TERM,
 * param = weight_thresh), volume_render_sigma_thresh_surf_trav (THRESH_DEPTH, param = sigma_thresh),
 * volume_render_alpha_surf_trav (THRESH_ALPHA, param = thresh), render_normal_surf_trav (NORMAL; out is (Q,3)),
 * extract_pts_surf_trav (:4052-4081; EXTRACT_PTS, param = alpha_thresh: out = depths and out2 = alphas of the first
 * max_sample samples with alpha > param, both (Q, max_sample), zero-filled here).
 * out is (Q,) floats otherwise; out2 / max_sample are ignored unless EXTRACT_PTS.  Rays that miss give 0. */
enum {
    ASURF_SCALAR_EXPECTED_TERM = 0,
    ASURF_SCALAR_MODE_TERM = 1,
    ASURF_SCALAR_THRESH_DEPTH = 2,
    ASURF_SCALAR_THRESH_ALPHA = 3,
    ASURF_SCALAR_NORMAL = 4,
    ASURF_SCALAR_EXTRACT_PTS = 5
};
int asurf_surf_trav_scalar(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, int32_t mode,
                           float param, int32_t max_sample, float *out, float *out2, void *stream);
/* volume_render_surf_trav_backward, :3708-3800 (grad_out = dL/dRGB (Q,3), color_cache = forward RGB) */
int asurf_surf_trav_backward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                             const float *grad_out, const float *color_cache, const asurf_grads_t *grads,
                             void *stream);
/* volume_render_surf_trav_fused, :3802-3942 (rgb_gt (Q,3) in, rgb_out (Q,3) out, grads +=) */
int asurf_surf_trav_fused(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                          const float *rgb_gt, const asurf_fused_t *fused, float *rgb_out,
                          const asurf_grads_t *grads, asurf_stats_t *stats_dev, void *stream);
/* test hooks: grid-space rays as the kernels see them, (Q,9) = origin3, dir3, tmin, tmax, world_step
 * (ray_find_bounds, include/render_util.cuh:651-701); composited-sample trace of the last march:
 * per ray up to max_hits entries of (cell, kind = st_id | fake<<2 | intersect_i<<3, t). */
int asurf_debug_ray_bounds(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                           float *xf_out, void *stream);
int asurf_debug_trace(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                      int32_t max_hits, int32_t *hit_count, int32_t *hit_cell, int32_t *hit_kind, float *hit_t,
                      void *stream);

/* The asurf_debug_set_* switches below select algorithm variants that must give the same results (the tests compare them).
 * They are INERT unless the process environment holds ASURF_DEBUG_HOOKS=1 when the library is first used. */
/* test hook: switch the hierarchical empty-block skipping of the marchers off (0) / on (non-zero, default).  Results
 * must be bit-identical either way; tests/ use it as a full-size property check. */
void asurf_debug_set_skip(int32_t enabled);

/* test hook: route every ray through the persistent shading kernels (0) instead of sending the rays whose march fits the
 * pre-march list through the wavefront kernels (non-zero, default).  Colours must be bit-identical either way, gradients
 * equal up to atomic order. */
void asurf_debug_set_wave(int32_t enabled);

/* test hook: two-level pre-march (block jumps per ray, then one thread per non-empty 16^3 block crossed; used for batches
 * >= 8192 rays) off (0) / on (non-zero, default).  The listed voxels must be identical either way. */
void asurf_debug_set_seg(int32_t enabled);

/* test hook (synchronises): the 8 device counters of the last render call -- [0],[1] ray fetch cursors of the persistent
 * kernels, [2] long rays, [3] short rays, [4] (ray, voxel) items, [5] samples, [6] fine items of the pre-march. */
int asurf_debug_counters(uint64_t *out8);

/* test hooks: the work pyramid the library keeps between render calls on the same grid (updated incrementally from the
 * vertices whose level-set side / density gate changed).  copy: the pyramid used by the last render call (device buffer
 * of asurf_accel_words-sized pyramid part, i.e. the first three levels); valid: whether a cache exists. */
int asurf_debug_work_cache_copy(uint64_t *out, int64_t words, void *stream);
int32_t asurf_debug_work_cache_valid(void);

/* test hook: tiled kernels of surf_tv_grad_sparse / surface_normal_grad_sparse for lists that are a window of the stored
 * vertices on (non-zero, default) / off (0: always the list kernels).  Same result either way up to summation order. */
void asurf_debug_set_normal_tile(int32_t enabled);
/* test hook (synchronises): verdict of the last device-side list check, 4 ints = {bad (0: the tiled kernel ran), lo, hi
 * (flat ids of the window), tiles fetched by the persistent CTAs} */
int asurf_debug_last_verdict(int32_t *out4);

/* ---- Plenoxels "cuvol" renderer, render_lerp_kernel_cuvol.cu:1120-1354 (grid->surface / level_set / accel / work unused;
 *      links may hold the negative skip codes written by accel_dist_prop) ---- */
/* volume_render_cuvol, :1120-1160 (rgb_out (Q,3); log_transmit_out (Q,) or NULL) */
int asurf_cuvol_forward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, float *rgb_out,
                        float *log_transmit_out, void *stream);
/* Depth renders of the cuvol backend, one entry for four reference functions (render_lerp_kernel_cuvol.cu:1356-1442):
 * volume_render_expected_term (mode EXPECTED_TERM, param = weight_thresh), volume_render_mode_term (MODE_TERM, param =
 * weight_thresh), volume_render_med_term (MED_TERM: out = depths, out2 = sigmas, both (Q, max_sample), zero-filled here),
 * volume_render_sigma_thresh (SIGMA_THRESH, param = sigma_thresh).  out is (Q,) floats unless MED_TERM. */
enum { ASURF_CUVOL_EXPECTED_TERM = 0, ASURF_CUVOL_MODE_TERM = 1, ASURF_CUVOL_MED_TERM = 2, ASURF_CUVOL_SIGMA_THRESH = 3 };
int asurf_cuvol_scalar(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, int32_t mode, float param,
                       int32_t max_sample, float *out, float *out2, void *stream);
/* march counters of one forward pass (the caller zero-fills stats_dev): n_steps = sample positions, n_skips =
 * positions that jump over an empty block, n_linked = n_active = samples whose 8 links + 8 densities are gathered,
 * n_samples = samples with sigma > sigma_thresh (SH gather and gradients) */
int asurf_cuvol_stats(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, asurf_stats_t *stats_dev,
                      void *stream);
/* volume_render_cuvol_image, :1162-1209: rays of a pinhole camera (c2w: 12 host floats, row-major 3x4), rgb_out (H,W,3) */
int asurf_cuvol_image(const asurf_grid_t *grid, const float *c2w_host, float fx, float fy, float cx, float cy, int32_t width,
                      int32_t height, const asurf_opt_t *opt, float *rgb_out, void *stream);
/* volume_render_cuvol_backward, :1211-1270 */
int asurf_cuvol_backward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, const float *grad_out,
                         const float *color_cache, const asurf_grads_t *grads, void *stream);
/* volume_render_cuvol_fused, :1272-1354 (norm_rays: 0 = this call's Q, see asurf_fused_t.norm_rays) */
int asurf_cuvol_fused(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, const float *rgb_gt,
                      float beta_loss, float sparsity_loss, int64_t norm_rays, float *rgb_out, const asurf_grads_t *grads,
                      void *stream);

/* ---- grid maintenance in front of the render path ----
 * accel_dist_prop, misc_kernel.cu:1022-1058: rewrites every negative entry of links (in place) with -(1 + number of empty
 * octree levels above the vertex), the skip codes the cuvol marcher reads. */
int asurf_accel_dist_prop(int32_t *links, const int32_t size[3], void *stream);
/* dilate, misc_kernel.cu:1005-1020: out = 3x3x3 OR of a bool (one byte per vertex) grid of the given size. */
int asurf_dilate(const uint8_t *grid, const int32_t size[3], uint8_t *out, void *stream);
/* Camera passes that decide which voxels to keep.  c2w_host: 12 floats, row-major 3x4, HOST memory (CameraSpec.c2w);
 * offset / scaling: 3 host floats each (the reference passes them as device tensors).
 * grid_weight_render, :1084-1111: data is a DENSE (X,Y,Z) sigma volume; grid_weight_out (X,Y,Z) takes, per vertex, the
 * max over pixels of the rendering weight of a sample in an adjacent voxel (atomic max; the caller zero-fills it). */
int asurf_grid_weight_render(const float *data, const int32_t size[3], const float offset[3], const float scaling[3],
                             const float *c2w_host, float fx, float fy, float cx, float cy, int32_t width, int32_t height,
                             float step_size, float stop_thresh, int32_t last_sample_opaque, float *grid_weight_out,
                             void *stream);
/* sparse_grid_weight_render, :1113-1138: same march through links / density (N,1); grid_weight_out (X,Y,Z) takes the max
 * TRANSMITTANCE in front of the sample. */
int asurf_sparse_grid_weight_render(const int32_t *links, const float *density, const int32_t size[3], const float offset[3],
                                    const float scaling[3], const float *c2w_host, float fx, float fy, float cx, float cy,
                                    int32_t width, int32_t height, float step_size, float stop_thresh, float *grid_weight_out,
                                    void *stream);
/* sparse_grid_mask_render, :1158-1175: grid_mask (N,) float rows of every stored vertex of a voxel some ray samples
 * (step 0.1 voxel from max(t_enter, near_clip)) are set to 1. */
int asurf_sparse_grid_mask_render(const int32_t *links, const int32_t size[3], const float offset[3], const float scaling[3],
                                  const float *origins, const float *dirs, int64_t n_rays, float near_clip, float *grid_mask,
                                  void *stream);
/* sparse_grid_visbility_render_surf, :1140-1156: visibility_out (N,) += 1 per pixel for the stored vertices of every voxel
 * the pixel's ray crosses up to and including the voxel of its first level-set intersection. */
int asurf_sparse_grid_visibility_render_surf(const int32_t *links, const float *surface, const float *level_set,
                                             int32_t level_set_num, const int32_t size[3], const float offset[3],
                                             const float scaling[3], const float *c2w_host, float fx, float fy, float cx,
                                             float cy, int32_t width, int32_t height, float *visibility_out, void *stream);

/* ---- point queries, svox2_kernel.cu:384-582 ----
 * Trilinear gather of one (N, n_cols) tensor at n_points world-space points (P,3); corners without a stored vertex read
 * `missing`.  Serves sample_grid (density with 0, SH with 0), sample_grid_sh_surf (surface with default_surf) and
 * sample_grid_raw_alpha (density with empty_raw); out is (P, n_cols). */
int asurf_sample_grid(const int32_t *links, const int32_t size[3], const float offset[3], const float scaling[3],
                      const float *data, int32_t n_cols, float missing, const float *points, int64_t n_points, float *out,
                      void *stream);
/* sample_grid_backward, :491-539: grad_data (N, n_cols) += transposed gather of grad_out (P, n_cols). */
int asurf_sample_grid_backward(const int32_t *links, const int32_t size[3], const float offset[3], const float scaling[3],
                               const float *points, int64_t n_points, const float *grad_out, int32_t n_cols,
                               float *grad_data, void *stream);
/* cubic_extract_iso_pts, :542-582: for each listed cell (flat vertex id, all 8 corners stored) and each of the 3 n_sample^2
 * axis-parallel lattice lines through it, the first zero of the trilinear level function in [0,1] whose interpolated
 * mask value is >= density_thresh, in grid coordinates; out is (n_cells, 3 n_sample^2, 3), zero where there is none. */
int asurf_cubic_extract_iso_pts(const int32_t *links, const int32_t size[3], const float *level_data, const float *mask_data,
                                const int32_t *cell_ids, int64_t n_cells, int32_t n_sample, float density_thresh, float *out,
                                void *stream);

/* ---- optimizer steps, optim_kernel.cu:154-267 ----
 * indexer_kind: 0 = all rows, 1 = bool mask (n rows), 2 = int64 row indices (n_index entries). */
int asurf_rmsprop_step(float *data, float *rms, float *grad, int64_t n_rows, int32_t n_cols, int32_t indexer_kind,
                       const void *indexer, int64_t n_index, float beta, float lr, float eps, float minval,
                       float lr_last, void *stream);
int asurf_sgd_step(float *data, float *grad, int64_t n_rows, int32_t n_cols, int32_t indexer_kind,
                   const void *indexer, int64_t n_index, float lr, float lr_last, void *stream);

/* ---- grid-side regularisers, loss_kernel.cu (gradients are ADDED into grad_*; mask_out may be NULL) ----
 * `size` is links' shape; data tensors are (N, n_cols) row-major; rand_cells holds flat x*Y*Z + y*Z + z cell ids. */
/* tv, loss_kernel.cu:1214-1247: mean over cells and channels [start_dim, end_dim) of sqrt(1e-5 + |grad|^2) -> *out_scalar */
int asurf_tv(const int32_t *links, const int32_t size[3], const float *data, int32_t n_cols, int32_t start_dim,
             int32_t end_dim, int32_t ignore_edge, float *out_scalar, void *stream);
/* tv_grad, :1249-1287 (dense) */
int asurf_tv_grad(const int32_t *links, const int32_t size[3], const float *data, int32_t n_cols, int32_t start_dim,
                  int32_t end_dim, float scale, int32_t ignore_edge, float *grad_data, void *stream);
/* tv_grad_sparse, :1327-1373 */
int asurf_tv_grad_sparse(const int32_t *links, const int32_t size[3], const float *data, int32_t n_cols,
                         const int32_t *rand_cells, int64_t n_cells, uint8_t *mask_out, int32_t start_dim,
                         int32_t end_dim, float scale, int32_t ignore_edge, int32_t ignore_last_z, float *grad_data,
                         void *stream);
/* surf_tv_grad_sparse, :1375-1427.  accel: optional occupancy buffer of `links` (asurf_accel_build); with it, a list that is
 * a contiguous window of the ascending list of all stored vertices (what svox2.py:6354-6372 produces) is recognised on the
 * device and processed tile by tile instead of cell by cell (same result up to summation order). */
int asurf_surf_tv_grad_sparse(const int32_t *links, const int32_t size[3], const float *surf, int32_t n_cols,
                              const float *density, int32_t density_cols, const int32_t *rand_cells, int64_t n_cells,
                              uint8_t *mask_out, int32_t start_dim, int32_t end_dim, float scale, int32_t ignore_edge,
                              float edge_value, int32_t ignore_last_z, int32_t alpha_dependency, float *grad_data,
                              const uint64_t *accel, void *stream);
/* surf_sign_change_grad_sparse, :1429-1466 (kernel :895-977, with its loop counter started at 0: the reference leaves it
 * uninitialised) */
int asurf_surf_sign_change_grad_sparse(const int32_t *links, const int32_t size[3], const float *data, int32_t n_cols,
                                       const int32_t *rand_cells, int64_t n_cells, uint8_t *mask_out, int32_t start_dim,
                                       int32_t end_dim, float scale, float *grad_data, void *stream);
/* surface_normal_grad (dense), loss_kernel.cu:1289-1325 (kernel :245-396): the normal-consistency loss over EVERY cell of the
 * (size - 1)^3 lattice on column(s) [start_dim, end_dim) of `data`, connectivity test always on, squared-difference form, no
 * touched mask.  (Unreachable from the reference's Python -- svox2.py:5725 raises first -- present for module completeness.) */
int asurf_surface_normal_grad(const int32_t *links, const int32_t size[3], const float *data, int32_t n_cols, float lv_set,
                              int32_t start_dim, int32_t end_dim, float scale, float *grad_data, void *stream);
/* lumisphere_tv_grad_sparse, :1661-1697 (kernel :1067-1177): TV of the radiance seen from one direction (basis values
 * basis_fn[basis_dim]) plus its change towards a perturbed direction (basis_fn_u[basis_dim], weight dir_factor), per colour
 * channel; cells are flat ids on the (size - 1) lattice.  One warp lane per SH coefficient: sh_data_dim <= 32. */
int asurf_lumisphere_tv_grad_sparse(const int32_t *links, const int32_t size[3], const float *sh_data, int32_t sh_data_dim,
                                    int32_t basis_dim, const int32_t *rand_cells, int64_t n_cells, const float *basis_fn,
                                    const float *basis_fn_u, float scale, float dir_factor, uint8_t *mask_out, float *grad_sh,
                                    void *stream);
/* alpha_surf_sparsify_grad_sparse, :1512-1570 */
int asurf_alpha_surf_sparsify_grad_sparse(const int32_t *links, const int32_t size[3], const float *alpha,
                                          int32_t alpha_cols, const float *surf, int32_t surf_cols,
                                          const int32_t *rand_cells, int64_t n_cells, uint8_t *mask_out, float scale_alpha,
                                          float scale_surf, int32_t surf_decrease, float surf_thresh, float alpha_bound,
                                          float surf_bound, float *grad_alpha, float *grad_surf, void *stream);
/* surface_normal_grad_sparse, :1572-1622 (eikonal_scale and the ndc coefficients of the reference are unused there).
 * accel: as for asurf_surf_tv_grad_sparse. */
int asurf_surface_normal_grad_sparse(const int32_t *links, const int32_t size[3], const float *surf,
                                     const int32_t *rand_cells, int64_t n_cells, uint8_t *mask_out, float lv_set,
                                     int32_t start_dim, int32_t end_dim, float scale, int32_t con_check,
                                     int32_t ignore_empty, int32_t use_l1, float *grad_data, const uint64_t *accel,
                                     void *stream);

/* ---- multi-sphere-image background, render_lerp_kernel_surf_trav.cu:2914-3137, :3370-3455 (the same code serves the cuvol
 *      renderer, render_lerp_kernel_cuvol.cu:540-760) ----
 * The render entries above run these passes themselves when the grid carries a background (forward: after the foreground
 * pass, onto its colours; backward / fused: after the foreground backward).  They are exported for callers that hold the
 * per-ray state of a foreground pass: log_transmit (Q,) = log-transmittance behind the grid, accum (Q,) = what the foreground
 * backward left of its running sum (NaN: the foreground never visited the ray; the pass then starts from
 * sum_c colour_c * dL/dcolour_c - beta_loss; +Inf: the ray missed the grid: the same without the beta term). */
/* render_background_kernel, :3370-3387: rgb_out (Q,3) += background colours + exp(log_transmit_final) * background_brightness */
int asurf_msi_forward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, const float *log_transmit,
                      float *rgb_out, void *stream);
/* render_background_image_kernel, :3389-3411: the same for the pixels of a pinhole camera (c2w: 12 host floats) */
int asurf_msi_forward_image(const asurf_grid_t *grid, const float *c2w_host, float fx, float fy, float cx, float cy,
                            int32_t width, int32_t height, const asurf_opt_t *opt, const float *log_transmit, float *rgb_out,
                            void *stream);
/* render_background_backward_kernel, :3413-3455: grad_is_rgb != 0: grad_in is rgb_gt, dL/dRGB = (color_cache - rgb_gt) * 2 /
 * (3 norm_rays) (norm_rays = 0: this call's Q); grads->grad_background (+ mask_background) receive the gradients */
int asurf_msi_backward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, const float *grad_in,
                       const float *color_cache, int32_t grad_is_rgb, int64_t norm_rays, const float *log_transmit,
                       const float *accum, float beta_loss, float sparsity_loss, const asurf_grads_t *grads, void *stream);
/* msi_tv_grad_sparse, loss_kernel.cu:979-1064, :1624-1659: TV over (texel, layer) cells of the background; links is
 * (links_x, links_y), msi (n, nlayers, n_channels); the last channel (sigma) uses scale_last; rand_cells holds
 * (x * links_y + y) * nlayers + z; mask_out (n, nlayers) bool or NULL */
int asurf_msi_tv_grad_sparse(const int32_t *links, int32_t links_x, int32_t links_y, const float *msi, int32_t nlayers,
                             int32_t n_channels, const int32_t *rand_cells, int64_t n_cells, uint8_t *mask_out, float scale,
                             float scale_last, float *grad_msi, void *stream);
/* test hook (synchronises): per-ray state the last foreground pass with a background left behind: log_transmit (Q,) and
 * accum (Q,) (either may be NULL) */
int asurf_debug_bg_state(float *log_transmit_out, float *accum_out, int64_t n_rays);

/* ---- multi-GPU gradient exchange helpers (ours; alphasurf_b200/dist.py) ----
 * rows: device int64 (n_rows,), ascending row indices touched on some rank, optionally padded with negative entries (a list
 * of fixed capacity: their bucket rows are zero-filled and ignored); bucket: device (n_rows, 2 + sh_dim) floats,
 * row = [density.grad, surface.grad, sh.grad...].  pack copies the rows into the bucket (and clears them in the gradient
 * tensors when clear_rows != 0); unpack_add adds the (all-reduced) bucket back. */
int asurf_rows_pack(const int64_t *rows, int64_t n_rows, float *grad_density, float *grad_surface, float *grad_sh,
                    int32_t sh_dim, float *bucket, int32_t clear_rows, void *stream);
int asurf_rows_unpack_add(const int64_t *rows, int64_t n_rows, float *grad_density, float *grad_surface, float *grad_sh,
                          int32_t sh_dim, const float *bucket, void *stream);

/* Touched-row masks for the same exchange, bit-packed (NCCL has no bitwise OR): pack a (n,) bool mask into (n + 31) / 32
 * words; unpack_or writes mask[i] = OR over the `world` word arrays laid end to end in words_all (an all-gather's output). */
int asurf_mask_pack(const uint8_t *mask, int64_t n, uint32_t *words, void *stream);
int asurf_mask_unpack_or(const uint32_t *words_all, int32_t world, int64_t n, uint8_t *mask, void *stream);

/* ---- per-kernel timing (ours; feeds bench.py's roofline) ----
 * After asurf_profile_enable(capacity > 0) every asurf_surf_trav_fused call records CUDA events around its forward
 * (work pyramid build, pre-march, shading) and its backward pass on the launching stream (up to `capacity` calls are kept).  asurf_profile_read waits for
 * the recorded events, returns the number of calls and the summed kernel durations in ms, and resets the ring.
 * asurf_profile_enable(0) switches it off and destroys the events. */
int asurf_profile_enable(int32_t capacity);
int asurf_profile_read(int32_t *n_calls, float *fwd_ms_sum, float *bwd_ms_sum);
/* same, split by stage: [0] work pyramid build, [1] pre-march, [2] forward shading, [3] backward (4 floats) */
int asurf_profile_read_stages(int32_t *n_calls, float *stage_ms_sum);

/* number of CUDA kernels this library has enqueued since the last reset (bench.py's gpu_launches) */
uint64_t asurf_launch_count(int32_t reset);

/* release the library's device workspaces (arena, scratch) */
void asurf_release(void);

#ifdef __cplusplus
}
#endif
#endif /* ASURF_H */
