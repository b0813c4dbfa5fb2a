/*
 * TEST INFRASTRUCTURE ONLY -- see oracle_common.h.
 *
 * CPU restatement of the alpha-Surf "surf_trav" renderer, following the CUDA semantics of
 *   render_lerp_kernel_surf_trav.cu:37-562   (trace_ray_surf_trav, forward)
 *   render_lerp_kernel_surf_trav.cu:1710-2911 (trace_ray_surf_trav_backward)
 *   render_lerp_kernel_surf_trav.cu:3139-3193, 3241-3368 (kernels: ray setup, dL/dRGB)
 *   render_lerp_kernel_surf_trav.cu:3802-3942 (fused host: lambda / Q scaling)
 * and the helpers of include/render_util.cuh cited at each function.
 * One "warp" of the reference is one scalar loop here; per-lane work is a loop over the D SH lanes.
 */
#include <string.h>
#include "oracle_common.h"

/* ---- include/render_util.cuh:789-848 surface_to_cubic_equation_01 ---- */
static void surface_to_cubic_equation_01(const double *s, const double *o, const double *d, double *outs) {
    double const m00 = s[0] * (1 - o[2]) + s[1] * (o[2]);
    double const m01 = s[2] * (1 - o[2]) + s[3] * (o[2]);
    double const m10 = s[4] * (1 - o[2]) + s[5] * (o[2]);
    double const m11 = s[6] * (1 - o[2]) + s[7] * (o[2]);
    double const k0 = (m01 * d[1] + d[2] * (s[3] - s[2]) * (o[1])) - (m00 * d[1] - d[2] * (s[1] - s[0]) * (1 - o[1]));
    double const k1 = (m11 * d[1] + d[2] * (s[7] - s[6]) * (o[1])) - (m10 * d[1] - d[2] * (s[5] - s[4]) * (1 - o[1]));
    double const h0 = d[1] * d[2] * (s[3] - s[2]) - d[1] * d[2] * (s[1] - s[0]);
    double const h1 = d[1] * d[2] * (s[7] - s[6]) - d[1] * d[2] * (s[5] - s[4]);
    outs[3] = h1 * d[0] - h0 * d[0];
    outs[2] = k1 * d[0] + h1 * (o[0]) - k0 * d[0] + h0 * (1 - o[0]);
    outs[1] = (m10 * (1 - o[1]) + m11 * (o[1])) * d[0] + k1 * (o[0]) - (m00 * (1 - o[1]) + m01 * (o[1])) * d[0] + k0 * (1 - o[0]);
    outs[0] = (m00 * (1 - o[1]) + m01 * (o[1])) * (1 - o[0]) + (m10 * (1 - o[1]) + m11 * (o[1])) * (o[0]);
}

/* ---- include/render_util.cuh:850-934 calc_surface_grad_01 ---- */
static void calc_surface_grad_01(const float *o, const float *d, const float *g, float *gs) {
    gs[0] = g[0] * ((1 - o[0]) * (1 - o[1]) * (1 - o[2]))
          + g[1] * (d[0] * (1 - o[1]) * (o[2] - 1) + (-d[1] * (1 - o[2]) - d[2] * (1 - o[1])) * (1 - o[0]))
          + g[2] * (d[0] * (d[1] * (1 - o[2]) + d[2] * (1 - o[1])) + d[1] * d[2] * (1 - o[0]))
          + g[3] * (-d[0] * d[1] * d[2]);
    gs[1] = g[0] * ((o[2]) * (1 - o[0]) * (1 - o[1]))
          + g[1] * (d[0] * (-o[2]) * (1 - o[1]) + (-d[1] * (o[2]) + d[2] * (1 - o[1])) * (1 - o[0]))
          + g[2] * (d[0] * (d[1] * (o[2]) - d[2] * (1 - o[1])) - d[1] * d[2] * (1 - o[0]))
          + g[3] * (d[0] * d[1] * d[2]);
    gs[2] = g[0] * ((o[1]) * (1 - o[0]) * (1 - o[2]))
          + g[1] * (d[0] * (-o[1]) * (1 - o[2]) + (d[1] * (1 - o[2]) - d[2] * (o[1])) * (1 - o[0]))
          + g[2] * (d[0] * (-d[1] * (1 - o[2]) + d[2] * (o[1])) - d[1] * d[2] * (1 - o[0]))
          + g[3] * (d[0] * d[1] * d[2]);
    gs[3] = g[0] * ((o[1]) * (o[2]) * (1 - o[0]))
          + g[1] * (d[0] * (-o[1]) * (o[2]) + (d[1] * (o[2]) + d[2] * (o[1])) * (1 - o[0]))
          + g[2] * (d[0] * (-d[1] * (o[2]) - d[2] * (o[1])) + d[1] * d[2] * (1 - o[0]))
          + g[3] * (-d[0] * d[1] * d[2]);
    gs[4] = g[0] * ((o[0]) * (1 - o[1]) * (1 - o[2]))
          + g[1] * (d[0] * (1 - o[1]) * (1 - o[2]) + (o[0]) * (-d[1] * (1 - o[2]) - d[2] * (1 - o[1])))
          + g[2] * (d[0] * (-d[1] * (1 - o[2]) - d[2] * (1 - o[1])) + d[1] * d[2] * (o[0]))
          + g[3] * (d[0] * d[1] * d[2]);
    gs[5] = g[0] * ((o[0]) * (o[2]) * (1 - o[1]))
          + g[1] * (d[0] * (o[2]) * (1 - o[1]) + (o[0]) * (-d[1] * (o[2]) + d[2] * (1 - o[1])))
          + g[2] * (d[0] * (-d[1] * (o[2]) + d[2] * (1 - o[1])) - d[1] * d[2] * (o[0]))
          + g[3] * (-d[0] * d[1] * d[2]);
    gs[6] = g[0] * ((o[0]) * (o[1]) * (1 - o[2]))
          + g[1] * (d[0] * (o[1]) * (1 - o[2]) + (o[0]) * (d[1] * (1 - o[2]) - d[2] * (o[1])))
          + g[2] * (d[0] * (d[1] * (1 - o[2]) - d[2] * (o[1])) - d[1] * d[2] * (o[0]))
          + g[3] * (-d[0] * d[1] * d[2]);
    gs[7] = g[0] * ((o[0]) * (o[1]) * (o[2]))
          + g[1] * (d[0] * (o[1]) * (o[2]) + (o[0]) * (d[1] * (o[2]) + d[2] * (o[1])))
          + g[2] * (d[0] * (d[1] * (o[2]) + d[2] * (o[1])) + d[1] * d[2] * (o[0]))
          + g[3] * (d[0] * d[1] * d[2]);
}

#define CLOSE0(x, eps) (fabs(x) < (eps))
#define SIGN_D(x) (((x) > 0.) ? 1. : -1.)

/* ---- include/render_util.cuh:1126-1203 cubic_equation_solver_vieta ---- */
static int cubic_equation_solver_vieta(double f0, double f1, double f2, double f3, double eps_double, double *outs) {
    if (CLOSE0(f3, eps_double)) {
        if (CLOSE0(f2, eps_double)) {
            if (CLOSE0(f1, eps_double)) return O_CUBIC_TYPE_NO_ROOT;
            outs[0] = -f0 / f1;
            return O_CUBIC_TYPE_LINEAR;
        } else {
            double const D = O_SQR(f1) - 4.0 * f2 * f0;
            double const sqrt_D = sqrt(D);
            if (D > 0) {
                if (f2 > 0) {
                    outs[0] = (-f1 - sqrt_D) / (2 * f2);
                    outs[1] = (-f1 + sqrt_D) / (2 * f2);
                } else {
                    outs[0] = (-f1 + sqrt_D) / (2 * f2);
                    outs[1] = (-f1 - sqrt_D) / (2 * f2);
                }
                if (CLOSE0(outs[0] - outs[1], eps_double)) {
                    outs[1] = -1;
                    return O_CUBIC_TYPE_POLY_ONE_R;
                }
                return O_CUBIC_TYPE_POLY;
            }
            return O_CUBIC_TYPE_NO_ROOT;
        }
    } else {
        double const b = f2 / f3, c = f1 / f3, d = f0 / f3;
        double const Q = (O_SQR(b) - 3. * c) / 9.;
        double const R = (2. * O_CUBIC(b) - 9. * b * c + 27. * d) / 54.;
        if (O_SQR(R) < O_CUBIC(Q)) {
            double const theta = acos(R / sqrt(O_CUBIC(Q)));
            outs[0] = -2. * sqrt(Q) * cos(theta / 3.) - b / 3.;
            outs[1] = -2. * sqrt(Q) * cos((theta - 2. * O_PI) / 3.) - b / 3.;
            outs[2] = -2. * sqrt(Q) * cos((theta + 2. * O_PI) / 3.) - b / 3.;
            return O_CUBIC_TYPE_CUBIC_THREE_R;
        } else {
            double const A = -SIGN_D(R) * cbrt(fabs(R) + sqrt(O_SQR(R) - O_CUBIC(Q)));
            double const B = (A == 0.) ? 0. : Q / A;
            outs[0] = (A + B) - b / 3.;
            return O_CUBIC_TYPE_CUBIC_ONE_R_;
        }
    }
}

static inline double dmax(double a, double b) { return a > b ? a : b; }

/* ---- include/render_util.cuh:1206-1415 calc_cubic_root_grad_vieta ---- */
static void calc_cubic_root_grad_vieta(int type, int st_id, const double *fs, float *grad_fs) {
    if (type == O_CUBIC_TYPE_LINEAR) {
        grad_fs[0] *= (float)(-1. / fs[1]);
        grad_fs[1] *= (float)(fs[0] / O_SQR(fs[1]));
        grad_fs[2] = 0.f;
        grad_fs[3] = 0.f;
    } else if (type == O_CUBIC_TYPE_POLY_ONE_R) {
        double const D = O_SQR(fs[1]) - 4. * fs[2] * fs[0];
        double const sqrt_D = sqrt(D);
        double const dt0_dD = 1 / (4. * fs[2] * sqrt_D);
        grad_fs[0] *= (float)(-1 / sqrt_D);
        grad_fs[1] *= (float)(((-1) / (2 * fs[2]) + (dt0_dD * 2 * fs[1])));
        grad_fs[2] *= (float)(((fs[1] - sqrt_D) / (4 * O_SQR(fs[2])) + (dt0_dD * (-4) * fs[0])));
        grad_fs[3] = 0.f;
    } else if (type == O_CUBIC_TYPE_POLY) {
        double const D = O_SQR(fs[1]) - 4.0 * fs[2] * fs[0];
        double const sqrt_D = sqrt(D);
        double const sqr_f2 = O_SQR(fs[2]);
        /* "S" (smaller root) when st_id == 0, "L" otherwise, for either sign of f2 (:1246-1262) */
        if (st_id == 0) {
            double const dt_dD = -1 / (4 * fs[2] * sqrt_D);
            grad_fs[0] *= (float)(1 / sqrt_D);
            grad_fs[1] *= (float)(((-1) / (2 * fs[2]) + (dt_dD * 2 * fs[1])));
            grad_fs[2] *= (float)(((fs[1] + sqrt_D) / (2 * sqr_f2) + (dt_dD * (-4) * fs[0])));
        } else {
            double const dt_dD = 1 / (4 * fs[2] * sqrt_D);
            grad_fs[0] *= (float)(-1 / sqrt_D);
            grad_fs[1] *= (float)(((-1) / (2 * fs[2]) + (dt_dD * 2 * fs[1])));
            grad_fs[2] *= (float)(((fs[1] - sqrt_D) / (2 * sqr_f2) + (dt_dD * (-4) * fs[0])));
        }
        grad_fs[3] = 0.f;
    } else {
        double const b = fs[2] / fs[3], c = fs[1] / fs[3], d = fs[0] / fs[3];
        double const Q = (O_SQR(b) - 3. * c) / 9.;
        double const R = (2. * O_CUBIC(b) - 9. * b * c + 27. * d) / 54.;
        double const DQ[4] = {0., -1. / (3. * fs[3]), 2. * fs[2] / (9. * O_SQR(fs[3])),
                              fs[1] / (3. * O_SQR(fs[3])) - 2. * O_SQR(fs[2]) / (9 * O_CUBIC(fs[3]))};
        double const DR[4] = {1. / (2. * fs[3]), -fs[2] / (6. * O_SQR(fs[3])),
                              -fs[1] / (6. * O_SQR(fs[3])) + O_SQR(fs[2]) / (9 * O_CUBIC(fs[3])),
                              -fs[0] / (2. * O_SQR(fs[3])) + fs[1] * fs[2] / (3. * O_CUBIC(fs[3])) -
                                  O_CUBIC(fs[2]) / (9 * (fs[3] * fs[3] * fs[3] * fs[3]))};
        double const Db[4] = {0., 0., 1. / fs[3], -fs[2] / O_SQR(fs[3])};
        double const Dst_Db = -1. / 3.;
        if (type == O_CUBIC_TYPE_CUBIC_THREE_R) {
            double const theta = acos(R / sqrt(O_CUBIC(Q)));
            double Dst_DQ, Dst_Dtheta;
            double const Dtheta_DQ = 3. * R / (2. * Q * sqrt(1. - O_SQR(R) / O_CUBIC(Q)) * sqrt(O_CUBIC(Q)));
            double const Dtheta_DR = -1 / (sqrt(1 - O_SQR(R) / O_CUBIC(Q)) * sqrt(O_CUBIC(Q)));
            if (st_id == 0) {
                Dst_DQ = -cos(theta / 3.) / sqrt(Q);
                Dst_Dtheta = 2. * sqrt(Q) * sin(theta / 3.) / 3.;
            } else if (st_id == 1) {
                Dst_DQ = cos(theta / 3. + O_PI / 3.) / sqrt(Q);
                Dst_Dtheta = -2. * sqrt(Q) * sin(theta / 3. + O_PI / 3.) / 3.;
            } else {
                Dst_DQ = sin(theta / 3. + O_PI / 6.) / sqrt(Q);
                Dst_Dtheta = 2. * sqrt(Q) * cos(theta / 3. + O_PI / 6.) / 3.;
            }
            for (int k = 0; k < 4; ++k)
                grad_fs[k] *= (float)(Dst_Dtheta * (Dtheta_DQ * DQ[k] + Dtheta_DR * DR[k]) + Dst_DQ * DQ[k] + Dst_Db * Db[k]);
        } else if (type == O_CUBIC_TYPE_CUBIC_ONE_R_) {
            double const A = -SIGN_D(R) * cbrt(fabs(R) + sqrt(O_SQR(R) - O_CUBIC(Q)));
            double const sq = dmax(sqrt(-O_CUBIC(Q) + O_SQR(R)), 1e-10);
            double const DA_DR = (R >= 0.) ? (-(R / (3. * sq) + 1. / 3.) / dmax(cbrt(O_SQR(R + sq)), 1e-10))
                                           : ((R / (3. * sq) - 1. / 3.) / dmax(cbrt(O_SQR(-R + sq)), 1e-10));
            double const DA_DQ = (R >= 0.) ? (O_SQR(Q) / (2. * sq * cbrt(O_SQR(R + sq))))
                                           : (-O_SQR(Q) / (2. * sq * cbrt(O_SQR(-R + sq))));
            double const DB_DA = (A == 0.) ? 0. : -Q / O_SQR(A);
            double const DB_DQ = (A == 0.) ? 0. : 1. / A;
            for (int k = 0; k < 4; ++k)
                grad_fs[k] *= (float)((DB_DA + 1.) * (DA_DQ * DQ[k] + DA_DR * DR[k]) + DB_DQ * DQ[k] + Dst_Db * Db[k]);
        }
    }
}

/* ---- include/render_util.cuh:2138-2188 ---- */
static float surf_alpha_act(float raw, int type) {
    if (type == O_SIGMOID_FN) return (float)(1. / (1. + (double)expf(-raw)));
    return (raw >= 0.f) ? 1.f - expf(-raw) : 0.f;
}
static float surf_alpha_act_grad(float alpha, int type) {
    if (type == O_SIGMOID_FN) return alpha * (1 - alpha);
    return (alpha > 0.f) ? 1 - alpha : 0.f;
}
static float truncated_vol_render_rw(float x, float a, float clamp_min) {
    const float arg = (float)(O_PI * (double)o_minf(o_maxf(a - x, 0.f), 1.f));
    return o_maxf(.5f * (1.f - cosf(arg)), clamp_min);
}

/* ---- include/render_util.cuh:2190-2236 compute_field_grad ---- */
static void compute_field_grad(const int32_t *links, const float *data, int offx, int offy, const int32_t *l,
                               const float *pos, float *out) {
    const int32_t *lp = links + ((int64_t)offx * l[0] + (int64_t)offy * l[1] + l[2]);
#define RD(u) ((lp[u] >= 0) ? data[lp[u]] : 0.f)
    const float ix0y0 = o_lerp(RD(0), RD(1), pos[2]);
    const float ix0y1 = o_lerp(RD(offy), RD(offy + 1), pos[2]);
    const float ix0 = o_lerp(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = o_lerp(RD(offx), RD(offx + 1), pos[2]);
    const float ix1y1 = o_lerp(RD(offy + offx), RD(offy + offx + 1), pos[2]);
    const float ix1 = o_lerp(ix1y0, ix1y1, pos[1]);
    out[0] = ix1 - ix0;
    out[1] = (float)((double)(pos[0] * (-ix1y0 + ix1y1)) + (1. - (double)pos[0]) * (double)(-ix0y0 + ix0y1));
    out[2] = pos[0] * (pos[1] * (-RD(offx + offy) + RD(offx + offy + 1)) + (1 - pos[1]) * (-RD(offx) + RD(offx + 1))) +
             (1 - pos[0]) * (pos[1] * (-RD(offy) + RD(offy + 1)) + (1 - pos[1]) * (-RD(0) + RD(1)));
#undef RD
}

/* ---- include/render_util.cuh:156-204 trilerp_backward_one_pos ---- */
static void trilerp_backward_one_pos(const int32_t *links, const float *data, int offx, int offy, size_t stride,
                                     const int32_t *l, const float *pos, int idx, float grad_in, float *grad_out) {
    const int32_t *lp = links + ((int64_t)offx * l[0] + (int64_t)offy * l[1] + l[2]);
#define RD(u) (data[(size_t)lp[u] * stride + idx])
    const float ix0y0 = o_lerp(RD(0), RD(1), pos[2]);
    const float ix0y1 = o_lerp(RD(offy), RD(offy + 1), pos[2]);
    const float ix0 = o_lerp(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = o_lerp(RD(offx), RD(offx + 1), pos[2]);
    const float ix1y1 = o_lerp(RD(offy + offx), RD(offy + offx + 1), pos[2]);
    const float ix1 = o_lerp(ix1y0, ix1y1, pos[1]);
    grad_out[0] += grad_in * (ix1 - ix0);
    grad_out[1] += grad_in * ((1 - pos[0]) * (ix0y1 - ix0y0) + (pos[0]) * (ix1y1 - ix1y0));
    grad_out[2] += grad_in * ((1 - pos[0]) * ((1 - pos[1]) * (RD(1) - RD(0)) + (pos[1]) * (RD(offy + 1) - RD(offy))) +
                              (pos[0]) * ((1 - pos[1]) * (RD(offx + 1) - RD(offx)) + (pos[1]) * (RD(offx + offy + 1) - RD(offx + offy))));
#undef RD
}

/* ---- include/render_util.cuh:1794-1821 assign_surface_grad ---- */
static void assign_surface_grad(const int32_t *links, float *grad_surface_out, uint8_t *mask, int offx, int offy,
                                const int32_t *l, const float *gs) {
    const int32_t *lp = links + ((int64_t)offx * l[0] + (int64_t)offy * l[1] + l[2]);
    const int u[8] = {0, 1, offy, offy + 1, offx, offx + 1, offx + offy, offx + offy + 1};
    for (int k = 0; k < 8; ++k)
        if (lp[u[k]] >= 0) {
            o_atomic_add(&grad_surface_out[lp[u[k]]], gs[k]);
            if (mask) mask[lp[u[k]]] = 1;
        }
}

/* fake-sample distance estimate shared by forward (:463-490) and backward (:2505-2532) */
static float fake_sample_dist_fn(const OGrid *g, const OOpt *opt, const double *surface, const float *pos,
                                 double *surf_miu_out, double *surf_std_out) {
    double const surf_miu = (surface[0] + surface[1] + surface[2] + surface[3] + surface[4] + surface[5] + surface[6] + surface[7]) / 8;
    double var = 0;
    for (int k = 0; k < 8; ++k) var += O_SQR(surface[k] - surf_miu);
    var /= 8;
    /* max(1e-9f, double) -> double; sqrtf() takes float */
    double surf_std = (double)sqrtf((float)(var > (double)1e-9f ? var : (double)1e-9f));
    if (!opt->fake_sample_normalize_surf) surf_std = 1.;
#define NS(x) ((float)(surface[x] / surf_std))
    const float ix0y0 = o_lerp(NS(0), NS(1), pos[2]);
    const float ix0y1 = o_lerp(NS(2), NS(3), pos[2]);
    const float ix0 = o_lerp(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = o_lerp(NS(4), NS(5), pos[2]);
    const float ix1y1 = o_lerp(NS(6), NS(7), pos[2]);
    const float ix1 = o_lerp(ix1y0, ix1y1, pos[1]);
    const float s = o_lerp(ix0, ix1, pos[0]);
#undef NS
    float dist = INFINITY;
    for (int i = 0; i < g->level_set_num; ++i)
        dist = fabsf(s - g->level_set[i]) < fabsf(dist) ? (s - g->level_set[i]) : dist;
    *surf_miu_out = surf_miu;
    *surf_std_out = surf_std;
    return dist;
}

/* Per-ray set-up shared by the forward and backward kernels (:3151-3172, :3276-3324) */
static void setup_ray(const OGrid *g, const OOpt *opt, const float *origin, const float *dir, const float *xf,
                      ORay *ray, float *sphfunc) {
    for (int i = 0; i < 3; ++i) { ray->origin[i] = origin[i]; ray->dir[i] = dir[i]; }
    o_calc_sh(g->basis_dim, ray->dir, sphfunc); /* world-space direction */
    if (xf) { /* test hook: rays already transformed by the device (origin,dir,tmin,tmax,world_step) */
        for (int i = 0; i < 3; ++i) { ray->origin[i] = xf[i]; ray->dir[i] = xf[3 + i]; }
        ray->tmin = xf[6]; ray->tmax = xf[7]; ray->world_step = xf[8];
    } else {
        o_ray_find_bounds(ray, g, opt);
    }
}

/* One step of the reference DDA (:88-197 == :1836-1920).  Returns the voxel to process. */
typedef struct { int32_t voxel_l[3]; float t_close, t_far; } OStep;
static void dda_step(const OGrid *g, const ORay *ray, int32_t *next_voxel, float *t, OStep *s) {
    for (int k = 0; k < 3; ++k) s->voxel_l[k] = next_voxel[k];
    float tc[3], tf[3];
    for (int k = 0; k < 3; ++k) {
        const int32_t close_plane = ray->dir[k] > 0.f ? s->voxel_l[k] : s->voxel_l[k] + 1;
        const int32_t far_plane = ray->dir[k] > 0.f ? s->voxel_l[k] + 1 : s->voxel_l[k];
        tc[k] = ((float)close_plane - ray->origin[k]) / ray->dir[k];
        tf[k] = ((float)far_plane - ray->origin[k]) / ray->dir[k];
    }
    s->t_close = o_maxf(o_maxf(o_maxf(tc[0], tc[1]), tc[2]), 0.f);
    s->t_far = o_minf(o_minf(tf[0], tf[1]), tf[2]);
    *t = s->t_far;
    int a = (s->t_far == tf[0]) ? 0 : ((s->t_far == tf[1]) ? 1 : 2);
    next_voxel[a] += (ray->dir[a] > 0.f) ? 1 : -1;
    if ((next_voxel[a] < 0) || (next_voxel[a] >= g->size[a] - 1)) *t = ray->tmax + 1.f;
}

static int voxel_links_ok(const OGrid *g, const int32_t *v, const int32_t *lp, int offx, int offy) {
    if ((v[0] + 1 >= g->size[0]) || (v[1] + 1 >= g->size[1]) || (v[2] + 1 >= g->size[2])) return 0;
    return !((lp[0] < 0) || (lp[1] < 0) || (lp[offy] < 0) || (lp[offy + 1] < 0) || (lp[offx] < 0) ||
             (lp[offx + 1] < 0) || (lp[offx + offy] < 0) || (lp[offx + offy + 1] < 0));
}
static int voxel_density_all_below(const OGrid *g, const int32_t *lp, int offx, int offy, float thr) {
    const int u[8] = {0, 1, offy, offy + 1, offx, offx + 1, offx + offy, offx + offy + 1};
    for (int k = 0; k < 8; ++k) if (!(g->density[lp[u[k]]] < thr)) return 0;
    return 1;
}

static void trace_push(const OTrace *tr, int64_t ray_id, int32_t cell, int32_t kind, float t) {
    if (!tr || !tr->hit_count) return;
    const int32_t n = tr->hit_count[ray_id]++;
    if (n < tr->max_hits) {
        tr->hit_cell[ray_id * tr->max_hits + n] = cell;
        tr->hit_kind[ray_id * tr->max_hits + n] = kind;
        tr->hit_t[ray_id * tr->max_hits + n] = t;
    }
}

/* ===================== forward: trace_ray_surf_trav (:37-562) ===================== */
static void trace_ray_forward(const OGrid *g, ORay *ray, const OOpt *opt, const float *sphfunc, float *out,
                              float *out_log_transmit, int l_dist_max_sample, float *sample_alphas,
                              float *sample_weights, float *sample_ts, const OTrace *tr, int64_t ray_id) {
    const int D = g->sh_dim, bd = g->basis_dim;
    int sample_i = 0, intersect_i = -1;
    int64_t Nv = 0, Nl = 0, Na = 0, S = 0;
    double const ray_dir_d[3] = {ray->dir[0], ray->dir[1], ray->dir[2]};
    float outv[3] = {0.f, 0.f, 0.f};
    float lane_color[32];

    if (ray->tmin > ray->tmax) {
        for (int c = 0; c < 3; ++c) out[c] = opt->background_brightness;
        if (out_log_transmit) *out_log_transmit = 0.f;
        return;
    }
    float t = ray->tmin;
    float log_transmit = 0.f;
    int32_t next_voxel[3];
    for (int j = 0; j < 3; ++j) {
        next_voxel[j] = (int32_t)fmaf(t, ray->dir[j], ray->origin[j]);
        next_voxel[j] = o_mini(o_maxi(next_voxel[j], 0), g->size[j] - 2);
    }
    const int offx = g->size[1] * g->size[2], offy = g->size[2];

    while (t <= ray->tmax) {
        OStep s;
        dda_step(g, ray, next_voxel, &t, &s);
        const int32_t *voxel_l = s.voxel_l;
        const float t_close = s.t_close, t_far = s.t_far;
        ++Nv;
        const int32_t cell = offx * voxel_l[0] + offy * voxel_l[1] + voxel_l[2];
        const int32_t *lp = g->links + cell;
        if (!voxel_links_ok(g, voxel_l, lp, offx, offy)) continue;
        ++Nl;
        if (voxel_density_all_below(g, lp, offx, offy, opt->sigma_thresh)) continue;
        ++Na;

        float new_origin_f[3];
        double new_origin[3], new_norm_origin[3];
        for (int k = 0; k < 3; ++k) {
            new_origin_f[k] = fmaf(t_close, ray->dir[k], ray->origin[k]); /* nvcc contracts o + t*d (:243) */
            new_origin[k] = new_origin_f[k];
            new_norm_origin[k] = new_origin[k] - voxel_l[k];
        }
        const int u[8] = {0, 1, offy, offy + 1, offx, offx + 1, offx + offy, offx + offy + 1};
        double surface[8];
        for (int k = 0; k < 8; ++k) surface[k] = g->surface[lp[u[k]]];
        double fs[4];
        surface_to_cubic_equation_01(surface, new_norm_origin, ray_dir_d, fs);

        int vox_has_sample = 0, vox_has_surf = 0;
        double smin = surface[0], smax = surface[0];
        for (int k = 1; k < 8; ++k) { if (surface[k] < smin) smin = surface[k]; if (surface[k] > smax) smax = surface[k]; }

        for (int i = 0; i < g->level_set_num; ++i) {
            double const lv_set = g->level_set[i];
            if ((lv_set < smin) || (lv_set > smax)) continue;
            vox_has_surf = 1;
            double st[3] = {-1, -1, -1};
            cubic_equation_solver_vieta(fs[0] - lv_set, fs[1], fs[2], fs[3], 1e-10, st);
            for (int j = 0; j < 3; ++j) {
                if (st[j] <= 0) continue;
                for (int k = 0; k < 3; ++k) {
                    ray->pos[k] = fmaf((float)st[j], ray->dir[k], (float)new_origin[k]);
                    ray->l[k] = o_mini(voxel_l[k], g->size[k] - 2);
                    ray->pos[k] -= (float)ray->l[k];
                }
                if ((ray->pos[0] < 0) | (ray->pos[0] > 1) | (ray->pos[1] < 0) | (ray->pos[1] > 1) | (ray->pos[2] < 0) | (ray->pos[2] > 1)) continue;
                vox_has_sample = 1;
                if (opt->only_outward_intersect) {
                    float sg[3];
                    compute_field_grad(g->links, g->surface, offx, offy, ray->l, ray->pos, sg);
                    float const norm_dir_dot = -(sg[0] * ray->dir[0] + sg[1] * ray->dir[1] + sg[2] * ray->dir[2]);
                    if (norm_dir_dot >= 0.f) continue;
                }
                ++intersect_i;
                float const raw_alpha = o_trilerp_cuvol_one(g->links, g->density, offx, offy, 1, ray->l, ray->pos, 0);
                if (raw_alpha > opt->sigma_thresh) {
                    float const alpha = surf_alpha_act(raw_alpha, opt->alpha_activation_type);
                    float const trunc_reweight = opt->truncated_vol_render ?
                        truncated_vol_render_rw((float)intersect_i, g->truncated_vol_render_a, opt->trunc_vol_weight_min) : 1.f;
                    float const rwalpha = alpha * trunc_reweight;
                    for (int lane = 0; lane < D; ++lane)
                        lane_color[lane] = o_trilerp_cuvol_one(g->links, g->sh, offx, offy, D, ray->l, ray->pos, lane) * sphfunc[lane % bd];
                    const float pcnt = -1 * logf(1 - rwalpha);
                    const float weight = expf(log_transmit) * (1.f - expf(-pcnt));
                    log_transmit -= pcnt;
                    for (int c = 0; c < 3; ++c)
                        outv[c] += weight * fmaxf(o_seg_sum(lane_color + c * bd, bd) + 0.5f, 0.f);
                    ++S;
                    trace_push(tr, ray_id, cell, j + 8 * intersect_i, (float)((double)t_close + st[j]));
                    if (sample_weights && (sample_i < l_dist_max_sample)) {
                        sample_alphas[sample_i] = rwalpha;
                        sample_weights[sample_i] = weight;
                        sample_ts[sample_i] = (float)((double)t_close + st[j]);
                        sample_i += 1;
                    }
                }
            }
        }

        /* fake sampling (:423-541) */
        if (opt->surf_fake_sample && !vox_has_sample && (!opt->limited_fake_sample || vox_has_surf)) {
            if ((t_far - t_close) > opt->surf_fake_sample_min_vox_len) {
                for (int k = 0; k < 3; ++k) {
                    ray->pos[k] = fmaf((t_far + t_close) / 2.f, ray->dir[k], ray->origin[k]);
                    ray->l[k] = o_mini(voxel_l[k], g->size[k] - 2);
                    ray->pos[k] -= (float)ray->l[k];
                }
                float alpha = o_trilerp_cuvol_one(g->links, g->density, offx, offy, 1, ray->l, ray->pos, 0);
                if (alpha > opt->sigma_thresh) {
                    alpha = surf_alpha_act(alpha, opt->alpha_activation_type);
                    double miu, sd;
                    const float fake_sample_dist = fake_sample_dist_fn(g, opt, surface, ray->pos, &miu, &sd);
                    alpha = alpha * expf(-.5f * O_SQR(fake_sample_dist / g->fake_sample_std));
                    float const trunc_reweight = opt->truncated_vol_render ?
                        truncated_vol_render_rw((float)intersect_i, g->truncated_vol_render_a, opt->trunc_vol_weight_min) : 1.f;
                    alpha = alpha * trunc_reweight;
                    for (int lane = 0; lane < D; ++lane)
                        lane_color[lane] = o_trilerp_cuvol_one(g->links, g->sh, offx, offy, D, ray->l, ray->pos, lane) * sphfunc[lane % bd];
                    const float pcnt = -1 * logf(1 - alpha);
                    const float weight = expf(log_transmit) * (1.f - expf(-pcnt));
                    log_transmit -= pcnt;
                    for (int c = 0; c < 3; ++c)
                        outv[c] += weight * fmaxf(o_seg_sum(lane_color + c * bd, bd) + 0.5f, 0.f);
                    ++S;
                    trace_push(tr, ray_id, cell, 3 + 8 * intersect_i, (t_far + t_close) / 2.f);
                    if (sample_weights && (sample_i < l_dist_max_sample) && opt->fake_sample_l_dist) {
                        sample_alphas[sample_i] = alpha;
                        sample_weights[sample_i] = weight;
                        sample_ts[sample_i] = (t_far + t_close) / 2.f;
                        sample_i += 1;
                    }
                }
            }
        }
        if (expf(log_transmit) < opt->stop_thresh) {
            log_transmit = -1e3f;
            break;
        }
    }
    for (int c = 0; c < 3; ++c) out[c] = outv[c] + expf(log_transmit) * opt->background_brightness;
    if (out_log_transmit) *out_log_transmit = log_transmit;
    if (tr && tr->counters) {
        tr->counters[ray_id * 4 + 0] = Nv; tr->counters[ray_id * 4 + 1] = Nl;
        tr->counters[ray_id * 4 + 2] = Na; tr->counters[ray_id * 4 + 3] = S;
    }
}

/* ===================== backward: trace_ray_surf_trav_backward (:1710-2911) ===================== */
typedef struct {
    float sample_alpha_sum, sample_weight_sum, Den_Dasum, Den_Dwsum, sample_t_mean, shared_Dmeant_sign;
    int valid_sample_n, max_sample_id;
} OPre;

/* lane-0 additions to d/d(rwalpha) from l_dist(w) and l_entropy(w) (:2141-2210 == :2602-2666) */
static float extra_grad_rwalpha_w(const OFused *f, const OPre *p, int sample_i, float log_transmit, float rwalpha,
                                  const float *sa, const float *sw, const float *sts) {
    float add = 0.f;
    const int M = f->l_dist_max_sample;
    if (f->lambda_l_dist > 0.f) {
        float Dldist_Dai = 0.f, log_Tk = log_transmit;
        for (int k = sample_i; k < p->valid_sample_n; ++k) {
            float Dldist_Dwk = 0.f;
            for (int j = 0; j < M; ++j) Dldist_Dwk += sw[j] * fabsf(sts[k] - sts[j]);
            if (k == sample_i) {
                Dldist_Dai += Dldist_Dwk * expf(log_transmit);
            } else {
                log_Tk += logf(o_maxf(1.f - sa[k - 1], 1e-8f));
                Dldist_Dai += Dldist_Dwk * expf(log_Tk) * sa[k] / o_minf(rwalpha - 1.f, -1e-8f);
            }
        }
        add += f->lambda_l_dist * Dldist_Dai;
    }
    if (f->lambda_l_entropy > 0.f) {
        float Den_Dai, log_Tk = log_transmit;
        if (f->no_norm_weight_l_entropy) {
            float const Den_Dwi = -(logf(o_maxf(sw[sample_i], 1e-8f)) + 1.f);
            Den_Dai = Den_Dwi * expf(log_transmit);
            for (int k = sample_i + 1; k < p->valid_sample_n; ++k) {
                float const Den_Dwk = -(logf(o_maxf(sw[k], 1e-8f)) + 1.f);
                log_Tk += logf(o_maxf(1.f - sa[k - 1], 1e-8f));
                Den_Dai += Den_Dwk * expf(log_Tk) * sa[k] / o_minf(rwalpha - 1.f, -1e-8f);
            }
        } else {
            float const Den_Dwi = -(logf(o_maxf(sw[sample_i], 1e-8f) / p->sample_weight_sum) + 1.f) / p->sample_weight_sum + p->Den_Dwsum;
            Den_Dai = Den_Dwi * expf(log_transmit);
            for (int k = sample_i + 1; k < p->valid_sample_n; ++k) {
                float const Den_Dwk = -(logf(o_maxf(sw[k], 1e-8f) / p->sample_weight_sum) + 1.f) / p->sample_weight_sum + p->Den_Dwsum;
                log_Tk += logf(o_maxf(1.f - sa[k - 1], 1e-8f));
                Den_Dai += Den_Dwk * expf(log_Tk) * sa[k] / o_minf(rwalpha - 1.f, -1e-8f);
            }
        }
        add += f->lambda_l_entropy * Den_Dai;
    }
    return add;
}

static void trace_ray_backward(const OGrid *g, const float *grad_output, const float *color_cache, ORay *ray,
                               const OOpt *opt, const float *sphfunc, const OFused *f, const float *sa,
                               const float *sw, const float *sts, OGrads *grads) {
    const int D = g->sh_dim, bd = g->basis_dim, M = f->l_dist_max_sample;
    double const ray_dir_d[3] = {ray->dir[0], ray->dir[1], ray->dir[2]};
    int sample_i = 0, intersect_i = -1;
    OPre p;
    memset(&p, 0, sizeof(p));
    /* preamble (:1756-1793) */
    for (int i = 0; i < M; ++i) { p.sample_alpha_sum += sa[i]; p.sample_weight_sum += sw[i]; }
    p.sample_alpha_sum = o_maxf(p.sample_alpha_sum, 1e-8f);
    p.sample_weight_sum = o_maxf(p.sample_weight_sum, 1e-8f);
    for (int i = 0; i < M; ++i) {
        p.Den_Dasum += sa[i] * (logf(o_maxf(sa[i], 1e-8f) / p.sample_alpha_sum) + 1.f) / O_SQR(p.sample_alpha_sum);
        p.Den_Dwsum += sw[i] * (logf(o_maxf(sw[i], 1e-8f) / p.sample_weight_sum) + 1.f) / O_SQR(p.sample_weight_sum);
        p.sample_t_mean += sw[i] / p.sample_weight_sum * sts[i];
    }
    for (int i = 0; i < M; ++i)
        if (sts[i] > 0.f) {
            p.valid_sample_n++;
            p.shared_Dmeant_sign += (p.sample_t_mean > sts[i]) ? 1.f : ((p.sample_t_mean < sts[i]) ? -1.f : 0.f);
        }
    float max_weight = 0.f;
    for (int i = 0; i < M; ++i) if (sw[i] > max_weight) { max_weight = sw[i]; p.max_sample_id = i; }

    float accum = fmaf(color_cache[0], grad_output[0], fmaf(color_cache[1], grad_output[1], color_cache[2] * grad_output[2]));
    if (ray->tmin > ray->tmax) return;
    float t = ray->tmin;
    float log_transmit = 0.f;
    int32_t next_voxel[3];
    for (int j = 0; j < 3; ++j) {
        next_voxel[j] = (int32_t)fmaf(t, ray->dir[j], ray->origin[j]);
        next_voxel[j] = o_mini(o_maxi(next_voxel[j], 0), g->size[j] - 2);
    }
    const int offx = g->size[1] * g->size[2], offy = g->size[2];
    float lane_color[32], curr_grad_color[32];

    while (t <= ray->tmax) {
        OStep s;
        dda_step(g, ray, next_voxel, &t, &s);
        const int32_t *voxel_l = s.voxel_l;
        const float t_close = s.t_close, t_far = s.t_far;
        const int32_t *lp = g->links + ((int64_t)offx * voxel_l[0] + offy * voxel_l[1] + voxel_l[2]);
        if (!voxel_links_ok(g, voxel_l, lp, offx, offy)) {
            t += opt->step_size; /* reference quirk (:1935): backward only */
            continue;
        }
        if (voxel_density_all_below(g, lp, offx, offy, opt->sigma_thresh)) continue;

        float new_origin_f[3];
        double new_origin[3], new_norm_origin[3];
        for (int k = 0; k < 3; ++k) {
            new_origin_f[k] = fmaf(t_close, ray->dir[k], ray->origin[k]);
            new_origin[k] = new_origin_f[k];
            new_norm_origin[k] = new_origin[k] - voxel_l[k];
        }
        const int u[8] = {0, 1, offy, offy + 1, offx, offx + 1, offx + offy, offx + offy + 1};
        double surface[8];
        for (int k = 0; k < 8; ++k) surface[k] = g->surface[lp[u[k]]];
        double fs[4];
        surface_to_cubic_equation_01(surface, new_norm_origin, ray_dir_d, fs);
        double const fs0_original = fs[0];
        int vox_has_sample = 0, vox_has_surf = 0;
        double smin = surface[0], smax = surface[0];
        for (int k = 1; k < 8; ++k) { if (surface[k] < smin) smin = surface[k]; if (surface[k] > smax) smax = surface[k]; }

        for (int i = 0; i < g->level_set_num; ++i) {
            double const lv_set = g->level_set[i];
            if ((lv_set < smin) || (lv_set > smax)) continue;
            vox_has_surf = 1;
            fs[0] = fs0_original - lv_set;
            double st[3] = {-1, -1, -1};
            const int cubic_root_type = cubic_equation_solver_vieta(fs[0], fs[1], fs[2], fs[3], 1e-10, st);
            for (int st_id = 0; st_id < 3; ++st_id) {
                if (st[st_id] <= 0) continue;
                for (int k = 0; k < 3; ++k) {
                    ray->pos[k] = fmaf((float)st[st_id], ray->dir[k], (float)new_origin[k]);
                    ray->l[k] = o_mini(voxel_l[k], g->size[k] - 2);
                    ray->pos[k] -= (float)ray->l[k];
                }
                if ((ray->pos[0] < 0) | (ray->pos[0] > 1) | (ray->pos[1] < 0) | (ray->pos[1] > 1) | (ray->pos[2] < 0) | (ray->pos[2] > 1)) continue;
                vox_has_sample = 1;
                if (opt->only_outward_intersect) {
                    float sg[3];
                    compute_field_grad(g->links, g->surface, offx, offy, ray->l, ray->pos, sg);
                    float const norm_dir_dot = -(sg[0] * ray->dir[0] + sg[1] * ray->dir[1] + sg[2] * ray->dir[2]);
                    if (norm_dir_dot >= 0.f) continue;
                }
                ++intersect_i;
                float const raw_alpha = o_trilerp_cuvol_one(g->links, g->density, offx, offy, 1, ray->l, ray->pos, 0);
                if (!(raw_alpha > opt->sigma_thresh)) continue;
                float const alpha = surf_alpha_act(raw_alpha, opt->alpha_activation_type);
                float const trunc_reweight = opt->truncated_vol_render ?
                    truncated_vol_render_rw((float)intersect_i, g->truncated_vol_render_a, opt->trunc_vol_weight_min) : 1.f;
                float const rwalpha = alpha * trunc_reweight;
                for (int lane = 0; lane < D; ++lane)
                    lane_color[lane] = o_trilerp_cuvol_one(g->links, g->sh, offx, offy, D, ray->l, ray->pos, lane) * sphfunc[lane % bd];
                const float pcnt = -logf(o_maxf(1.f - rwalpha, 1e-8f));
                const float weight = expf(log_transmit) * (1.f - expf(-pcnt));
                float tc[3], in01[3];
                for (int c = 0; c < 3; ++c) {
                    const float lct = o_seg_sum(lane_color + c * bd, bd) + 0.5f;
                    const float tcol = fmaxf(lct, 0.f);
                    in01[c] = (tcol == lct) ? 1.f : 0.f;
                    tc[c] = tcol * grad_output[c];
                }
                float total_color = tc[0];            /* shuffle order (:2112-2115): (c0 + c2) + c1 */
                total_color += tc[2];
                total_color += tc[1];
                for (int lane = 0; lane < D; ++lane) {
                    const int c = lane / bd;
                    const float grad_common = weight * in01[c] * grad_output[c];
                    curr_grad_color[lane] = sphfunc[lane % bd] * grad_common;
                }
                accum -= weight * total_color;
                float curr_grad_rwalpha = accum / o_minf(rwalpha - 1.f, -1e-8f) + total_color * expf(log_transmit);
                curr_grad_rwalpha += extra_grad_rwalpha_w(f, &p, sample_i, log_transmit, rwalpha, sa, sw, sts);
                log_transmit -= pcnt;

                for (int lane = 0; lane < D; ++lane)
                    o_trilerp_backward_cuvol_one(g->links, grads->grad_sh, offx, offy, D, ray->l, ray->pos, curr_grad_color[lane], lane);

                float grad_xyz[3] = {0, 0, 0};
                if (!opt->no_surf_grad_from_sh) {
                    for (int lane = 0; lane < D; ++lane) {
                        float gl[3] = {0, 0, 0};
                        trilerp_backward_one_pos(g->links, g->sh, offx, offy, D, ray->l, ray->pos, lane, curr_grad_color[lane], gl);
                        grad_xyz[0] += gl[0]; grad_xyz[1] += gl[1]; grad_xyz[2] += gl[2];
                    }
                }
                /* lane 0 (:2266-2448) */
                if (f->lambda_l_dist_a > 0.f) {
                    float a = 0.f;
                    for (int j = 0; j < M; ++j) a += sa[j] * fabsf(sts[sample_i] - sts[j]);
                    curr_grad_rwalpha += f->lambda_l_dist_a * a;
                }
                if (f->lambda_l_entropy_a > 0.f) {
                    float const Den_Dai = -(logf(o_maxf(sa[sample_i], 1e-8f) / p.sample_alpha_sum) + 1.f) / p.sample_alpha_sum;
                    curr_grad_rwalpha += f->lambda_l_entropy_a * (Den_Dai + p.Den_Dasum);
                }
                if ((f->sparsity_loss > 0.f) && (raw_alpha > 0.f)) {
                    float const _1_a = o_maxf(1.f - alpha, 1e-8f);
                    curr_grad_rwalpha += -f->sparsity_loss * (1.f / o_minf(_1_a * logf(_1_a), -1e-8f)) * (1.f - weight / p.sample_weight_sum);
                }
                if (f->lambda_inwards_norm_loss > 0.f) {
                    float sg[3];
                    compute_field_grad(g->links, g->surface, offx, offy, ray->l, ray->pos, sg);
                    float const surf_n = o_maxf(sqrtf(O_SQR(sg[0]) + O_SQR(sg[1]) + O_SQR(sg[2])), 1e-8f);
                    float const nd = (-sg[0] / surf_n) * ray->dir[0] + (-sg[1] / surf_n) * ray->dir[1] + (-sg[2] / surf_n) * ray->dir[2];
                    if (nd > 0.f) curr_grad_rwalpha += f->lambda_inwards_norm_loss * O_SQR(nd);
                }
                float const curr_grad_alpha = curr_grad_rwalpha * trunc_reweight;
                float curr_grad_raw_alpha = curr_grad_alpha * surf_alpha_act_grad(alpha, opt->alpha_activation_type);
                o_trilerp_backward_cuvol_one_density(g->links, grads->grad_density, grads->mask, offx, offy, ray->l, ray->pos, curr_grad_raw_alpha);
                if ((f->lambda_l_di > 0.f) && (alpha < f->l_di_alpha_thresh))
                    curr_grad_raw_alpha += f->lambda_l_di * (-1.f) * surf_alpha_act_grad(alpha, opt->alpha_activation_type);
                trilerp_backward_one_pos(g->links, g->density, offx, offy, 1, ray->l, ray->pos, 0, curr_grad_raw_alpha, grad_xyz);
                float grad_st = grad_xyz[0] * ray->dir[0] + grad_xyz[1] * ray->dir[1] + grad_xyz[2] * ray->dir[2];
                if ((f->lambda_l_dist > 0.f) || (f->lambda_l_dist_a > 0.f)) {
                    float gw = 0.f, ga = 0.f;
                    for (int j = 0; j < M; ++j) {
                        float const sg_ = (sts[sample_i] > sts[j]) ? 1.f : ((sts[sample_i] < sts[j]) ? -1.f : 0.f);
                        gw += sg_ * sw[sample_i] * sw[j];
                        ga += sg_ * sa[sample_i] * sa[j];
                    }
                    grad_st += f->lambda_l_dist * gw + f->lambda_l_dist_a * ga;
                }
                if (f->lambda_l_samp_dist > 0.f) {
                    float const sg_ = (p.sample_t_mean > sts[sample_i]) ? 1.f : ((p.sample_t_mean < sts[sample_i]) ? -1.f : 0.f);
                    grad_st += f->lambda_l_samp_dist * (p.shared_Dmeant_sign * sa[sample_i] / p.sample_weight_sum + sg_ * (-1.f));
                }
                if ((f->lambda_conv_mode_samp > 0.f) && (trunc_reweight > opt->trunc_vol_weight_min)) {
                    float const gc = (sts[sample_i] > sts[p.max_sample_id]) ? 1.f : ((sts[sample_i] < sts[p.max_sample_id]) ? -1.f : 0.f);
                    grad_st += f->lambda_conv_mode_samp * gc;
                }
                float grad_fs[4] = {grad_st, grad_st, grad_st, grad_st};
                calc_cubic_root_grad_vieta(cubic_root_type, st_id, fs, grad_fs);
                float grad_surface[8];
                float const nno_f[3] = {(float)new_norm_origin[0], (float)new_norm_origin[1], (float)new_norm_origin[2]};
                calc_surface_grad_01(nno_f, ray->dir, grad_fs, grad_surface);
                if (grads->grad_surface) assign_surface_grad(g->links, grads->grad_surface, grads->mask, offx, offy, ray->l, grad_surface);
                if (alpha < f->surf_sparse_alpha_thresh && grads->grad_surface)
                    o_trilerp_backward_cuvol_one_density(g->links, grads->grad_surface, grads->mask, offx, offy, ray->l, ray->pos, f->lambda_inplace_surf_sparse);
                if (sample_i < M - 1) sample_i += 1;
            }
        }

        /* fake sample gradient (:2460-2866) */
        if (opt->surf_fake_sample && !vox_has_sample && (!opt->limited_fake_sample || vox_has_surf)) {
            if ((t_far - t_close) > opt->surf_fake_sample_min_vox_len) {
                for (int k = 0; k < 3; ++k) {
                    ray->pos[k] = fmaf((t_far + t_close) / 2, ray->dir[k], ray->origin[k]);
                    ray->l[k] = o_mini(voxel_l[k], g->size[k] - 2);
                    ray->pos[k] -= (float)ray->l[k];
                }
                float const raw_alpha = o_trilerp_cuvol_one(g->links, g->density, offx, offy, 1, ray->l, ray->pos, 0);
                if (raw_alpha > opt->sigma_thresh) {
                    float const alpha = surf_alpha_act(raw_alpha, opt->alpha_activation_type);
                    double surf_miu, surf_std;
                    const float fake_sample_dist = fake_sample_dist_fn(g, opt, surface, ray->pos, &surf_miu, &surf_std);
                    float const reweight = expf(-.5f * O_SQR(fake_sample_dist / g->fake_sample_std));
                    float const trunc_reweight = opt->truncated_vol_render ?
                        truncated_vol_render_rw((float)intersect_i, g->truncated_vol_render_a, opt->trunc_vol_weight_min) : 1.f;
                    float const rw_alpha = alpha * reweight * trunc_reweight;
                    for (int lane = 0; lane < D; ++lane)
                        lane_color[lane] = o_trilerp_cuvol_one(g->links, g->sh, offx, offy, D, ray->l, ray->pos, lane) * sphfunc[lane % bd];
                    const float pcnt = -1 * logf(o_maxf(1.f - rw_alpha, 1e-8f));
                    const float weight = expf(log_transmit) * (1.f - expf(-pcnt));
                    float tc[3], in01[3];
                    for (int c = 0; c < 3; ++c) {
                        const float lct = o_seg_sum(lane_color + c * bd, bd) + 0.5f;
                        const float tcol = fmaxf(lct, 0.f);
                        in01[c] = (tcol == lct) ? 1.f : 0.f;
                        tc[c] = tcol * grad_output[c];
                    }
                    float total_color_fs = tc[0];
                    total_color_fs += tc[2];
                    total_color_fs += tc[1];
                    for (int lane = 0; lane < D; ++lane) {
                        const int c = lane / bd;
                        curr_grad_color[lane] = sphfunc[lane % bd] * (weight * in01[c] * grad_output[c]);
                    }
                    accum -= weight * total_color_fs;
                    float curr_grad_rwalpha = accum / o_minf(rw_alpha - 1.f, -1e-8f) + total_color_fs * expf(log_transmit);
                    if (opt->fake_sample_l_dist)
                        curr_grad_rwalpha += extra_grad_rwalpha_w(f, &p, sample_i, log_transmit, rw_alpha, sa, sw, sts);
                    log_transmit -= pcnt;
                    for (int lane = 0; lane < D; ++lane)
                        o_trilerp_backward_cuvol_one(g->links, grads->grad_sh, offx, offy, D, ray->l, ray->pos, curr_grad_color[lane], lane);
                    if (opt->fake_sample_l_dist) {
                        if (f->lambda_l_dist_a > 0.f) {
                            float a = 0.f;
                            for (int j = 0; j < M; ++j) a += sa[j] * fabsf(sts[sample_i] - sts[j]);
                            curr_grad_rwalpha += f->lambda_l_dist_a * a;
                        }
                        if (f->lambda_l_entropy_a > 0.f) {
                            float const Den_Dai = -(logf(o_maxf(sa[sample_i], 1e-8f) / p.sample_alpha_sum) + 1.f) / p.sample_alpha_sum;
                            curr_grad_rwalpha += f->lambda_l_entropy_a * (Den_Dai + p.Den_Dasum);
                        }
                        if (sample_i < M - 1) sample_i += 1;
                    }
                    if ((f->sparsity_loss > 0.f) && (raw_alpha > 0.f)) {
                        float const _1_a = o_maxf(1.f - rw_alpha, 1e-8f);
                        curr_grad_rwalpha += -f->sparsity_loss * (1.f / o_minf(_1_a * logf(_1_a), -1e-8f)) * (1.f - weight / p.sample_weight_sum);
                    }
                    if (f->lambda_inwards_norm_loss > 0.f) {
                        float sg[3];
                        compute_field_grad(g->links, g->surface, offx, offy, ray->l, ray->pos, sg);
                        float const surf_n = o_maxf(sqrtf(O_SQR(sg[0]) + O_SQR(sg[1]) + O_SQR(sg[2])), 1e-8f);
                        float const nd = (-sg[0] / surf_n) * ray->dir[0] + (-sg[1] / surf_n) * ray->dir[1] + (-sg[2] / surf_n) * ray->dir[2];
                        if (nd > 0.f) curr_grad_rwalpha += f->lambda_inwards_norm_loss * O_SQR(nd);
                    }
                    float curr_grad_alpha = curr_grad_rwalpha * reweight * trunc_reweight;
                    float curr_grad_raw_alpha = curr_grad_alpha * surf_alpha_act_grad(alpha, opt->alpha_activation_type);
                    o_trilerp_backward_cuvol_one_density(g->links, grads->grad_density, grads->mask, offx, offy, ray->l, ray->pos, curr_grad_raw_alpha);
                    float grad_fake_dist = curr_grad_rwalpha * (-alpha * trunc_reweight * fake_sample_dist * reweight / O_SQR(g->fake_sample_std));
                    float grad_ns[8];
                    const float ay = 1.f - ray->pos[1], az = 1.f - ray->pos[2];
                    float xo = (1.0f - ray->pos[0]) * grad_fake_dist;
                    grad_ns[0] = ay * az * xo; grad_ns[1] = ay * ray->pos[2] * xo;
                    grad_ns[2] = ray->pos[1] * az * xo; grad_ns[3] = ray->pos[1] * ray->pos[2] * xo;
                    xo = ray->pos[0] * grad_fake_dist;
                    grad_ns[4] = ay * az * xo; grad_ns[5] = ay * ray->pos[2] * xo;
                    grad_ns[6] = ray->pos[1] * az * xo; grad_ns[7] = ray->pos[1] * ray->pos[2] * xo;
                    float grad_surface[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    if (!opt->fake_sample_normalize_surf) {
                        for (int ks = 0; ks < 8; ++ks) grad_surface[ks] = grad_ns[ks];
                    } else {
                        for (int ks = 0; ks < 8; ++ks)
                            for (int kn = 0; kn < 8; ++kn) {
                                if (ks == kn)
                                    grad_surface[ks] += (float)(grad_ns[kn] * (surface[ks] * (surf_miu - surface[ks]) / 8.f / O_CUBIC(surf_std) + 1.f / surf_std));
                                else
                                    grad_surface[ks] += (float)(grad_ns[kn] * (surface[kn] * (surf_miu - surface[ks]) / 8.f / O_CUBIC(surf_std)));
                            }
                    }
                    if (grads->grad_surface) assign_surface_grad(g->links, grads->grad_surface, grads->mask, offx, offy, ray->l, grad_surface);
                    if (grads->grad_fake_sample_std) {
                        float grad_std = curr_grad_rwalpha * alpha * O_SQR(fake_sample_dist) * reweight * trunc_reweight / O_CUBIC(g->fake_sample_std);
                        o_atomic_add(grads->grad_fake_sample_std, grad_std);
                    }
                }
            }
        }
        if (expf(log_transmit) < opt->stop_thresh) break;
    }
}

/* ===================== host-level entry points ===================== */

/* volume_render_surf_trav (:3596-3654) when l_dist_max_sample == 0, else the forward half of the fused call.
 * xf: optional (Q,9) device-transformed rays (origin3, dir3, tmin, tmax, world_step). */
void oracle_surf_trav_forward(const OGrid *g, const OOpt *opt, const float *origins, const float *dirs,
                              const float *xf, int64_t Q, float *rgb_out, int l_dist_max_sample,
                              float *sample_alphas, float *sample_weights, float *sample_ts, const OTrace *tr) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int64_t r = 0; r < Q; ++r) {
        ORay ray;
        float sph[9];
        setup_ray(g, opt, origins + 3 * r, dirs + 3 * r, xf ? xf + 9 * r : NULL, &ray, sph);
        const int M = l_dist_max_sample;
        trace_ray_forward(g, &ray, opt, sph, rgb_out + 3 * r, NULL, M,
                          sample_alphas ? sample_alphas + (int64_t)M * r : NULL,
                          sample_weights ? sample_weights + (int64_t)M * r : NULL,
                          sample_ts ? sample_ts + (int64_t)M * r : NULL, tr, r);
    }
}

/* volume_render_surf_trav_backward (:3708-3800): grad_out is dL/dRGB, all lambdas 0. */
void oracle_surf_trav_backward(const OGrid *g, const OOpt *opt, const float *origins, const float *dirs,
                               const float *xf, int64_t Q, const float *grad_out, const float *color_cache,
                               OGrads *grads) {
    OFused f;
    memset(&f, 0, sizeof(f));
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int64_t r = 0; r < Q; ++r) {
        ORay ray;
        float sph[9];
        setup_ray(g, opt, origins + 3 * r, dirs + 3 * r, xf ? xf + 9 * r : NULL, &ray, sph);
        trace_ray_backward(g, grad_out + 3 * r, color_cache + 3 * r, &ray, opt, sph, &f, NULL, NULL, NULL, grads);
    }
}

/* volume_render_surf_trav_fused (:3802-3942).  q_norm: the Q used for loss normalisation (== Q in the
 * reference; the multi-GPU wrapper passes the global batch size). */
void oracle_surf_trav_fused(const OGrid *g, const OOpt *opt, const float *origins, const float *dirs,
                            const float *xf, int64_t Q, int64_t q_norm, const float *rgb_gt, const OFused *fused,
                            float *rgb_out, OGrads *grads, const OTrace *tr) {
    const int M = fused->l_dist_max_sample;
    float *sa = (float *)calloc((size_t)Q * (M > 0 ? M : 1), sizeof(float));
    float *sw = (float *)calloc((size_t)Q * (M > 0 ? M : 1), sizeof(float));
    float *sts = (float *)calloc((size_t)Q * (M > 0 ? M : 1), sizeof(float));
    oracle_surf_trav_forward(g, opt, origins, dirs, xf, Q, rgb_out, M, sa, sw, sts, tr);
    OFused f = *fused; /* lambda scaling at launch (:3896-3914) */
    const float Qf = (float)q_norm;
    f.beta_loss /= Qf;
    f.lambda_l_dist /= Qf;
    f.lambda_l_entropy /= Qf;
    f.lambda_l_dist_a /= Qf;
    f.lambda_l_entropy_a /= Qf;
    f.lambda_l_samp_dist /= Qf;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int64_t r = 0; r < Q; ++r) {
        ORay ray;
        float sph[9];
        setup_ray(g, opt, origins + 3 * r, dirs + 3 * r, xf ? xf + 9 * r : NULL, &ray, sph);
        /* dL/dRGB (:3306-3316) */
        const float norm_factor_l2 = 2.f / (3 * (int)q_norm);
        const float norm_factor_l1 = 1.f / (3 * (int)q_norm);
        float grad_out[3];
        for (int i = 0; i < 3; ++i) {
            const float resid = rgb_out[r * 3 + i] - rgb_gt[r * 3 + i];
            grad_out[i] = resid * norm_factor_l2 * f.lambda_l2;
            grad_out[i] += (resid > 0.f) ? (norm_factor_l1 * f.lambda_l1) : (-norm_factor_l1 * f.lambda_l1);
        }
        trace_ray_backward(g, grad_out, rgb_out + 3 * r, &ray, opt, sph, &f, sa + (int64_t)M * r, sw + (int64_t)M * r,
                           sts + (int64_t)M * r, grads);
    }
    free(sa); free(sw); free(sts);
}

/* test hook: expose the cubic solver and its gradient (analogue of the reference's test_cuda.cu:108-122) */
/* ===================== scalar renders: trace_ray_expected_term (:564-794), _mode_term_surf_trav (:796-1001),
 * _sigma_thresh (:1003-1168), _alpha (:1170-1337), _normal (:1339-1534) =====================
 * Same traversal as the colour renderer but: no density gate, no outward test, no truncated re-weighting, no fake samples,
 * every in-range root composited.  mode: 0 expected depth, 1 mode depth (param = weight_thresh), 2 depth of the first
 * sample with alpha > param, 3 alpha of that sample, 4 surface gradient at the first sample with alpha > 0 (3 floats),
 * 5 trace_ray_extract_pt (:1536-1708): depths (out) and alphas (out2) of the first max_sample samples with alpha > param. */
static void trace_ray_scalar(const OGrid *g, ORay *ray, const OOpt *opt, int mode, float param, int max_sample, float *out,
                             float *out2) {
    const int nout = (mode == 4) ? 3 : ((mode == 5) ? max_sample : 1);
    int sample_id = 0;
    for (int c = 0; c < nout; ++c) out[c] = 0.f;
    if (mode == 5) for (int c = 0; c < nout; ++c) out2[c] = 0.f;
    if (ray->tmin > ray->tmax) return;
    double const ray_dir_d[3] = {ray->dir[0], ray->dir[1], ray->dir[2]};
    float t = ray->tmin, outv = 0.f, log_transmit = 0.f, max_weight = 0.f, weight_acc = 0.f;
    int32_t next_voxel[3];
    for (int j = 0; j < 3; ++j) {
        next_voxel[j] = (int32_t)fmaf(t, ray->dir[j], ray->origin[j]);
        next_voxel[j] = o_mini(o_maxi(next_voxel[j], 0), g->size[j] - 2);
    }
    const int offx = g->size[1] * g->size[2], offy = g->size[2];
    while (t <= ray->tmax) {
        OStep s;
        dda_step(g, ray, next_voxel, &t, &s);
        const int32_t *voxel_l = s.voxel_l;
        const float t_close = s.t_close;
        const int32_t *lp = g->links + (offx * voxel_l[0] + offy * voxel_l[1] + voxel_l[2]);
        if (!voxel_links_ok(g, voxel_l, lp, offx, offy)) continue;
        double new_origin[3], new_norm_origin[3];
        for (int k = 0; k < 3; ++k) {
            new_origin[k] = fmaf(t_close, ray->dir[k], ray->origin[k]);
            new_norm_origin[k] = new_origin[k] - voxel_l[k];
        }
        const int u[8] = {0, 1, offy, offy + 1, offx, offx + 1, offx + offy, offx + offy + 1};
        double surface[8];
        for (int k = 0; k < 8; ++k) surface[k] = g->surface[lp[u[k]]];
        double fs[4];
        surface_to_cubic_equation_01(surface, new_norm_origin, ray_dir_d, fs);
        double smin = surface[0], smax = surface[0];
        for (int k = 1; k < 8; ++k) { if (surface[k] < smin) smin = surface[k]; if (surface[k] > smax) smax = surface[k]; }
        for (int i = 0; i < g->level_set_num; ++i) {
            double const lv_set = g->level_set[i];
            if ((lv_set < smin) || (lv_set > smax)) continue;
            double st[3] = {-1, -1, -1};
            cubic_equation_solver_vieta(fs[0] - lv_set, fs[1], fs[2], fs[3], 1e-10, st);
            for (int j = 0; j < 3; ++j) {
                if (st[j] <= 0) continue;
                for (int k = 0; k < 3; ++k) {
                    ray->pos[k] = fmaf((float)st[j], ray->dir[k], (float)new_origin[k]);
                    ray->l[k] = o_mini(voxel_l[k], g->size[k] - 2);
                    ray->pos[k] -= (float)ray->l[k];
                }
                if ((ray->pos[0] < 0) | (ray->pos[0] > 1) | (ray->pos[1] < 0) | (ray->pos[1] > 1) | (ray->pos[2] < 0) |
                    (ray->pos[2] > 1))
                    continue;
                const float alpha = surf_alpha_act(o_trilerp_cuvol_one(g->links, g->density, offx, offy, 1, ray->l, ray->pos, 0),
                                                   opt->alpha_activation_type);
                const float depth = (float)(((st[j] + (double)t_close) / (double)opt->step_size) * (double)ray->world_step);
                if (mode <= 1) {
                    const float pcnt = -1 * logf(1 - alpha);
                    const float weight = expf(log_transmit) * (1.f - expf(-pcnt));
                    log_transmit -= pcnt;
                    if (mode == 0) {
                        outv = (float)((double)outv + (double)weight * (st[j] + (double)t_close) / (double)opt->step_size *
                                                          (double)ray->world_step);
                    } else {
                        weight_acc += weight;
                        if (weight > max_weight) { max_weight = weight; outv = depth; }
                    }
                } else if (mode == 2 || mode == 3) {
                    if (alpha > param) { out[0] = (mode == 2) ? depth : alpha; return; }
                } else if (mode == 5) {
                    if (alpha > param) {
                        out[sample_id] = depth;
                        out2[sample_id] = alpha;
                        sample_id += 1;
                        if (sample_id >= max_sample) return;
                    }
                } else {
                    if (alpha > 0) { compute_field_grad(g->links, g->surface, offx, offy, ray->l, ray->pos, out); return; }
                }
            }
        }
        if (mode <= 1 && expf(log_transmit) < opt->stop_thresh) {
            log_transmit = -1e3f;
            break;
        }
    }
    if (mode == 0) out[0] = outv;
    else if (mode == 1) out[0] = (weight_acc > param) ? outv : 0.f;
}

void oracle_surf_trav_scalar(const OGrid *g, const OOpt *opt, const float *origins, const float *dirs, const float *xf,
                             int64_t Q, int mode, float param, int max_sample, float *out, float *out2) {
    const int nout = (mode == 4) ? 3 : ((mode == 5) ? max_sample : 1);
    if (mode == 5 && max_sample <= 0) return;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t q = 0; q < Q; ++q) {
        ORay ray;
        float sphfunc[16];
        setup_ray(g, opt, origins + q * 3, dirs + q * 3, xf ? xf + q * 9 : NULL, &ray, sphfunc);
        trace_ray_scalar(g, &ray, opt, mode, param, max_sample, out + q * nout, out2 ? out2 + q * nout : NULL);
    }
}

/* sparse_grid_visbility_trace_ray_surf, misc_kernel.cu:511-719: from t = 0 (no near clip), every voxel the DDA visits adds
 * 1 to its stored corner rows, until the first in-voxel root of a level set.  xf: (Q,9) origin, dir, t, tmax, - as the GPU
 * derived them, or NULL.  The reference leaves `ray.tmax` uninitialised where the DDA steps off the grid (:600-609); like
 * the CUDA side here the march ends there. */
void oracle_visibility_surf(const OGrid *g, const float *origins, const float *dirs, const float *xf, int64_t Q,
                            float *visibility) {
    const int offx = g->size[1] * g->size[2], offy = g->size[2];
    for (int64_t q = 0; q < Q; ++q) {
        ORay ray;
        for (int i = 0; i < 3; ++i) { ray.origin[i] = origins[q * 3 + i]; ray.dir[i] = dirs[q * 3 + i]; }
        if (xf) {
            for (int i = 0; i < 3; ++i) { ray.origin[i] = xf[q * 9 + i]; ray.dir[i] = xf[q * 9 + 3 + i]; }
            ray.tmin = xf[q * 9 + 6]; ray.tmax = xf[q * 9 + 7];
        } else {
            OOpt o0;
            memset(&o0, 0, sizeof(o0));
            o0.step_size = 1.f;
            o_ray_find_bounds(&ray, g, &o0);   /* near_clip 0: tmin = max(0, AABB entry) */
        }
        if (ray.tmin > ray.tmax) continue;
        double const ray_dir_d[3] = {ray.dir[0], ray.dir[1], ray.dir[2]};
        float t = ray.tmin;
        int32_t next_voxel[3];
        for (int j = 0; j < 3; ++j) {
            next_voxel[j] = (int32_t)fmaf(t, ray.dir[j], ray.origin[j]);
            next_voxel[j] = o_mini(o_maxi(next_voxel[j], 0), g->size[j] - 2);
        }
        int hit = 0;
        while (t <= ray.tmax && !hit) {
            OStep s;
            dda_step(g, &ray, next_voxel, &t, &s);
            const int32_t *voxel_l = s.voxel_l;
            const int32_t *lp = g->links + (offx * voxel_l[0] + offy * voxel_l[1] + voxel_l[2]);
            const int u[8] = {0, 1, offy, offy + 1, offx, offx + 1, offx + offy, offx + offy + 1};
            for (int k = 0; k < 8; ++k) if (lp[u[k]] >= 0) visibility[lp[u[k]]] += 1.f;
            if (!voxel_links_ok(g, voxel_l, lp, offx, offy)) continue;
            double new_origin[3], new_norm_origin[3], surface[8];
            for (int k = 0; k < 3; ++k) {
                new_origin[k] = fmaf(s.t_close, ray.dir[k], ray.origin[k]);
                new_norm_origin[k] = new_origin[k] - voxel_l[k];
            }
            for (int k = 0; k < 8; ++k) surface[k] = g->surface[lp[u[k]]];
            double fs[4];
            surface_to_cubic_equation_01(surface, new_norm_origin, ray_dir_d, fs);
            double smin = surface[0], smax = surface[0];
            for (int k = 1; k < 8; ++k) { if (surface[k] < smin) smin = surface[k]; if (surface[k] > smax) smax = surface[k]; }
            for (int i = 0; i < g->level_set_num && !hit; ++i) {
                double const lv_set = g->level_set[i];
                if ((lv_set < smin) || (lv_set > smax)) continue;
                double st[3] = {-1, -1, -1};
                cubic_equation_solver_vieta(fs[0] - lv_set, fs[1], fs[2], fs[3], 1e-10, st);
                for (int j = 0; j < 3 && !hit; ++j) {
                    if (st[j] <= 0) continue;
                    float pos[3];
                    for (int k = 0; k < 3; ++k) pos[k] = fmaf((float)st[j], ray.dir[k], (float)new_origin[k]) - (float)voxel_l[k];
                    if ((pos[0] < 0) | (pos[0] > 1) | (pos[1] < 0) | (pos[1] > 1) | (pos[2] < 0) | (pos[2] > 1)) continue;
                    hit = 1;
                }
            }
        }
    }
}

/* cubic_extract_iso_pts_kernel, svox2_kernel.cu:248-376 */
void oracle_cubic_extract_iso_pts(const int32_t *links, const int32_t *size, const float *level, const float *maskv,
                                  const int32_t *cell_ids, int64_t n_cells, int n_sample, float density_thresh, float *out) {
    const int offy = size[2], offx = size[1] * size[2];
    for (int64_t tid = 0; tid < n_cells; ++tid) {
        const int xyz = cell_ids[tid];
        const int z = xyz % size[2], xy = xyz / size[2], y = xy % size[1], x = xy / size[1];
        if ((x >= size[0] - 1) || (y >= size[1] - 1) || (z >= size[2] - 1)) continue;
        const int32_t *lp = links + ((int64_t)offx * x + (int64_t)offy * y + z);
        const int u[8] = {0, 1, offy, offy + 1, offx, offx + 1, offx + offy, offx + offy + 1};
        int ok = 1;
        for (int k = 0; k < 8; ++k) ok &= (lp[u[k]] >= 0);
        if (!ok) continue;
        double surface[8];
        float mv[8];
        for (int k = 0; k < 8; ++k) { surface[k] = level[lp[u[k]]]; mv[k] = maskv[lp[u[k]]]; }
        const float step_size = 1.f / (n_sample - 1);
        for (int i = 0; i < n_sample; ++i) {
            const float pos1 = i * step_size;
            for (int j = 0; j < n_sample; ++j) {
                const float pos2 = j * step_size;
                for (int dir_id = 0; dir_id < 3; ++dir_id) {
                    double dirs[3] = {0., 0., 0.}, origin[3] = {0., 0., 0.};
                    if (dir_id == 0) { dirs[0] = 1.; origin[1] = pos1; origin[2] = pos2; }
                    else if (dir_id == 1) { dirs[1] = 1.; origin[0] = pos1; origin[2] = pos2; }
                    else { dirs[2] = 1.; origin[0] = pos1; origin[1] = pos2; }
                    double fs[4], st[3] = {-1, -1, -1};
                    surface_to_cubic_equation_01(surface, origin, dirs, fs);
                    cubic_equation_solver_vieta(fs[0], fs[1], fs[2], fs[3], 1e-10, st);
                    for (int st_i = 0; st_i < 3; ++st_i) {
                        if ((st[st_i] >= 0.) && (st[st_i] <= 1.)) {
                            const float pt[3] = {(float)(origin[0] + dirs[0] * st[st_i]), (float)(origin[1] + dirs[1] * st[st_i]),
                                                 (float)(origin[2] + dirs[2] * st[st_i])};
                            const float ix0y0 = o_lerp(mv[0], mv[1], pt[2]), ix0y1 = o_lerp(mv[2], mv[3], pt[2]);
                            const float ix0 = o_lerp(ix0y0, ix0y1, pt[1]);
                            const float ix1y0 = o_lerp(mv[4], mv[5], pt[2]), ix1y1 = o_lerp(mv[6], mv[7], pt[2]);
                            const float ix1 = o_lerp(ix1y0, ix1y1, pt[1]);
                            if (o_lerp(ix0, ix1, pt[0]) >= density_thresh) {
                                float *o = out + (tid * 3 * n_sample * n_sample + i * n_sample * 3 + j * 3 + dir_id) * 3;
                                o[0] = pt[0] + x; o[1] = pt[1] + y; o[2] = pt[2] + z;
                                break;
                            }
                        }
                    }
                }
            }
        }
    }
}

int oracle_cubic_solve(const double *fs, double *st) {
    st[0] = st[1] = st[2] = -1;
    return cubic_equation_solver_vieta(fs[0], fs[1], fs[2], fs[3], 1e-10, st);
}
void oracle_cubic_root_grad(int type, int st_id, const double *fs, float *grad_fs) {
    calc_cubic_root_grad_vieta(type, st_id, fs, grad_fs);
}
