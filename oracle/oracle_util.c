/*
 * TEST INFRASTRUCTURE ONLY -- see oracle_common.h.
 * Device-math helpers of include/render_util.cuh and include/cuda_util.cuh restated for the CPU.
 */
#include "oracle_common.h"

/* render_util.cuh:373-436 (calc_sh).  Note out[6] mixes a double literal: computed in double. */
void o_calc_sh(int basis_dim, const float *dir, float *out) {
    const float C0 = 0.28209479177387814f;
    const float C1 = 0.4886025119029199f;
    const float C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                         -1.0925484305920792f, 0.5462742152960396f};
    out[0] = C0;
    const float x = dir[0], y = dir[1], z = dir[2];
    const float xx = x * x, yy = y * y, zz = z * z;
    const float xy = x * y, yz = y * z, xz = x * z;
    if (basis_dim == 9) {
        out[4] = C2[0] * xy;
        out[5] = C2[1] * yz;
        out[6] = (float)((double)C2[2] * (2.0 * (double)zz - (double)xx - (double)yy));
        out[7] = C2[3] * xz;
        out[8] = C2[4] * (xx - yy);
    }
    if (basis_dim == 9 || basis_dim == 4) {
        out[1] = -C1 * y;
        out[2] = C1 * z;
        out[3] = -C1 * x;
    }
}

/* render_util.cuh:651-701 (ray_find_bounds), cuda_util.cuh:63-69 (transform_coord),
 * render_util.cuh:536-547 (_get_delta_scale).  rnorm3df is a CUDA libm routine (<=1 ulp); the CPU uses
 * the correctly rounded value, so tests feed GPU-transformed rays when bit-exact traversal is compared. */
void o_ray_find_bounds(ORay *ray, const OGrid *g, const OOpt *opt) {
    for (int i = 0; i < 3; ++i) ray->origin[i] = fmaf(ray->origin[i], g->scaling[i], g->offset[i]);
    for (int i = 0; i < 3; ++i) ray->dir[i] *= g->scaling[i];
    const double n2 = (double)ray->dir[0] * ray->dir[0] + (double)ray->dir[1] * ray->dir[1] +
                      (double)ray->dir[2] * ray->dir[2];
    const float delta_scale = (float)(1.0 / sqrt(n2));
    for (int i = 0; i < 3; ++i) ray->dir[i] *= delta_scale;
    ray->world_step = delta_scale * opt->step_size;

    if (opt->use_spheric_clip) {
        float sph_origin[3], sph_dir[3];
        for (int i = 0; i < 3; ++i) {
            const float ss = 2.f / (float)g->size[i];
            sph_origin[i] = fmaf(ray->origin[i] + 0.5f, ss, -1.f);
            sph_dir[i] = ray->dir[i] * ss;
        }
        /* ConcentricSpheresIntersector, render_util.cuh:619-649 */
        const float q2a = 2 * (sph_dir[0] * sph_dir[0] + sph_dir[1] * sph_dir[1] + sph_dir[2] * sph_dir[2]);
        const float qb = 2 * (sph_origin[0] * sph_dir[0] + sph_origin[1] * sph_dir[1] + sph_origin[2] * sph_dir[2]);
        const float f = qb * qb - 2 * q2a * (sph_origin[0] * sph_origin[0] + sph_origin[1] * sph_origin[1] +
                                             sph_origin[2] * sph_origin[2]);
        const float r1 = 1.f, r2 = 1.f - opt->near_clip;
        const float det1 = f + 2 * q2a * r1 * r1, det2 = f + 2 * q2a * r2 * r2;
        int ok = 1;
        if (det1 < 0) ok = 0; else ray->tmax = (-qb + sqrtf(det1)) / q2a;
        if (ok) { if (det2 < 0) ok = 0; else ray->tmin = (-qb - sqrtf(det2)) / q2a; }
        if (!ok) { ray->tmin = 1e-9f; ray->tmax = 0.f; }
    } else {
        ray->tmin = opt->near_clip / ray->world_step * opt->step_size;
        ray->tmax = 2e3f;
        for (int i = 0; i < 3; ++i) {
            const float invdir = (float)(1.0 / (double)ray->dir[i]);
            const float t1 = (-0.5f - ray->origin[i]) * invdir;
            const float t2 = ((float)g->size[i] - 0.5f - ray->origin[i]) * invdir;
            if (ray->dir[i] != 0.f) {
                ray->tmin = o_maxf(ray->tmin, o_minf(t1, t2));
                ray->tmax = o_minf(ray->tmax, o_maxf(t1, t2));
            }
        }
    }
}

/* render_util.cuh:72-92 */
float o_trilerp_cuvol_one(const int32_t *links, const float *data, int offx, int offy, size_t stride,
                          const int32_t *l, const float *pos, int idx) {
    const int32_t *lp = links + ((int64_t)offx * l[0] + (int64_t)offy * l[1] + l[2]);
#define RD(u) ((lp[u] >= 0) ? data[(size_t)lp[u] * stride + idx] : 0.f)
    const float ix0y0 = o_lerp(RD(0), RD(1), pos[2]);
    const float ix0y1 = o_lerp(RD(offy), RD(offy + 1), pos[2]);
    const float ix0 = o_lerp(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = o_lerp(RD(offx), RD(offx + 1), pos[2]);
    const float ix1y1 = o_lerp(RD(offy + offx), RD(offy + offx + 1), pos[2]);
    const float ix1 = o_lerp(ix1y0, ix1y1, pos[1]);
    return o_lerp(ix0, ix1, pos[0]);
#undef RD
}

/* render_util.cuh:94-122 */
void o_trilerp_backward_cuvol_one(const int32_t *links, float *grad_data, int offx, int offy, size_t stride,
                                  const int32_t *l, const float *pos, float grad_out, int idx) {
    const float ay = 1.f - pos[1], az = 1.f - pos[2];
    float xo = (1.0f - pos[0]) * grad_out;
    const int32_t *lp = links + ((int64_t)offx * l[0] + (int64_t)offy * l[1] + l[2]);
#define ADD(u, val) if (lp[u] >= 0) o_atomic_add(&grad_data[(size_t)lp[u] * stride + idx], (val))
    ADD(0, ay * az * xo);
    ADD(1, ay * pos[2] * xo);
    ADD(offy, pos[1] * az * xo);
    ADD(offy + 1, pos[1] * pos[2] * xo);
    xo = pos[0] * grad_out;
    ADD(offx + 0, ay * az * xo);
    ADD(offx + 1, ay * pos[2] * xo);
    ADD(offx + offy, pos[1] * az * xo);
    ADD(offx + offy + 1, pos[1] * pos[2] * xo);
#undef ADD
}

/* render_util.cuh:124-154 */
void o_trilerp_backward_cuvol_one_density(const int32_t *links, float *grad_data, uint8_t *mask, int offx,
                                          int offy, const int32_t *l, const float *pos, float grad_out) {
    const float ay = 1.f - pos[1], az = 1.f - pos[2];
    float xo = (1.0f - pos[0]) * grad_out;
    const int32_t *lp = links + ((int64_t)offx * l[0] + (int64_t)offy * l[1] + l[2]);
#define ADD(u, val) if (lp[u] >= 0) { o_atomic_add(&grad_data[lp[u]], (val)); if (mask) mask[lp[u]] = 1; }
    ADD(0, ay * az * xo);
    ADD(1, ay * pos[2] * xo);
    ADD(offy, pos[1] * az * xo);
    ADD(offy + 1, pos[1] * pos[2] * xo);
    xo = pos[0] * grad_out;
    ADD(offx + 0, ay * az * xo);
    ADD(offx + 1, ay * pos[2] * xo);
    ADD(offx + offy, pos[1] * az * xo);
    ADD(offx + offy + 1, pos[1] * pos[2] * xo);
#undef ADD
}

/* cub::WarpReduce<float>::HeadSegmentedSum as used at render_lerp_kernel_surf_trav.cu:402 -- shuffle-down
 * tree inside a segment of n lanes: lane i adds lane i+off when i+off is still inside the segment. */
float o_seg_sum(const float *v, int n) {
    float a[32], b[32];
    for (int i = 0; i < n; ++i) a[i] = v[i];
    for (int off = 1; off < 32; off <<= 1) {
        for (int i = 0; i < n; ++i) b[i] = (i + off < n) ? a[i] + a[i + off] : a[i];
        for (int i = 0; i < n; ++i) a[i] = b[i];
    }
    return a[0];
}
