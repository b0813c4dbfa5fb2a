/*
 * TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the reference svox2 SparseGrid render hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call anything in oracle/.  The product (alphasurf_b200/) never links or imports it.
 *
 * Parity pinning: the reference ships NO golden vectors for this path; this oracle is
 * pinned (a) against fixtures generated from the reference's own pure-PyTorch renderer
 * (tests/golden/, generator oracle/gen_golden.py) and (b) on the GPU box against the UNMODIFIED
 * reference CUDA kernels compiled by oracle/build_ref_cuda.sh into oracle/_ref/.
 *
 * All citations are relative to /root/reference/svox2/csrc/.
 */
#ifndef ASURF_ORACLE_COMMON_H
#define ASURF_ORACLE_COMMON_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* include/data_spec.hpp:39-56 + include/data_spec_packed.cuh:11-83 (packed view) */
typedef struct {
    int32_t size[3];
    const int32_t *links;       /* (X,Y,Z) int32, contiguous */
    const float *density;       /* (N,1) */
    const float *surface;       /* (N,1) or NULL */
    const float *sh;            /* (N,D) */
    const float *level_set;     /* (L,) */
    int32_t level_set_num;
    int32_t basis_dim;
    int32_t sh_dim;             /* D = 3*basis_dim */
    float offset[3];            /* _offset  (svox2.py:6250-6253, already times grid size) */
    float scaling[3];           /* _scaling */
    float fake_sample_std;
    float truncated_vol_render_a;
} OGrid;

/* include/data_spec.hpp:168-201 */
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
} OOpt;

/* scalars of volume_render_surf_trav_fused (render_lerp_kernel_surf_trav.cu:3802-3828) */
typedef struct {
    float beta_loss;
    float sparsity_loss;
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
} OFused;

typedef struct {
    float *grad_density;          /* (N,1) += */
    float *grad_surface;          /* (N,1) += or NULL */
    float *grad_sh;               /* (N,D) += */
    float *grad_fake_sample_std;  /* (1,) += or NULL */
    uint8_t *mask;                /* (N,) bool or NULL */
} OGrads;

/* single-ray state: include/data_spec_packed.cuh:137-160 */
typedef struct {
    float origin[3];
    float dir[3];
    float tmin, tmax, world_step;
    float pos[3];
    int32_t l[3];
} ORay;

/* Per-ray trace (test hook): composited samples in march order. */
typedef struct {
    int32_t max_hits;     /* capacity per ray */
    int32_t *hit_count;   /* (Q,) number of composited samples (may exceed max_hits) */
    int32_t *hit_cell;    /* (Q,max_hits) flat link offset of voxel_l */
    int32_t *hit_kind;    /* (Q,max_hits) st_id 0..2, or 3 for a fake sample; +8*intersect_i */
    float *hit_t;         /* (Q,max_hits) t_close + st */
    int64_t *counters;    /* (Q,4) Nv, Nl, Na, S    or NULL */
} OTrace;

/* cuda_util.cuh:74-77 */
static inline float o_lerp(float a, float b, float w) { return fmaf(w, b - a, a); }
static inline float o_maxf(float a, float b) { return a > b ? a : b; } /* CUDA max(): NaN-dropping not replicated */
static inline float o_minf(float a, float b) { return a < b ? a : b; }
static inline int o_maxi(int a, int b) { return a > b ? a : b; }
static inline int o_mini(int a, int b) { return a < b ? a : b; }
#define O_SQR(x) ((x) * (x))
#define O_CUBIC(x) ((x) * (x) * (x))
#define O_PI 3.1415926535897931e+0

/* include/data_spec.hpp:11-31 */
enum {
    O_CUBIC_TYPE_NO_ROOT = 200,
    O_CUBIC_TYPE_LINEAR = 201,
    O_CUBIC_TYPE_POLY_ONE_R = 202,
    O_CUBIC_TYPE_POLY = 203,
    O_CUBIC_TYPE_CUBIC_ONE_R = 204,
    O_CUBIC_TYPE_CUBIC_THREE_R = 205,
    O_CUBIC_TYPE_CUBIC_ONE_R_ = 206,
};
enum { O_SIGMOID_FN = 0, O_EXP_FN = 1 };

/* atomic float add so the oracle can run its rays on all host cores (bench cpu_baseline) */
static inline void o_atomic_add(float *p, float v) {
#ifdef _OPENMP
#pragma omp atomic
#endif
    *p += v;
}

void o_calc_sh(int basis_dim, const float *dir, float *out);
void o_ray_find_bounds(ORay *ray, const OGrid *g, const OOpt *opt);
float o_trilerp_cuvol_one(const int32_t *links, const float *data, int offx, int offy, size_t stride,
                          const int32_t *l, const float *pos, int idx);
void o_trilerp_backward_cuvol_one(const int32_t *links, float *grad_data, int offx, int offy, size_t stride,
                                  const int32_t *l, const float *pos, float grad_out, int idx);
void o_trilerp_backward_cuvol_one_density(const int32_t *links, float *grad_data, uint8_t *mask, int offx,
                                          int offy, const int32_t *l, const float *pos, float grad_out);
float o_seg_sum(const float *v, int n);

#endif
