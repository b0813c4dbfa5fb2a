// alphasurf_b200: device math of the alpha-Surf ray / level-set intersection.
//
// Formulas restate /root/reference/svox2/csrc/include/render_util.cuh (cited per function) with the same
// operand order and precision (fp64 where the reference uses double, fast intrinsics where it does), because
// hit selection has to agree bit for bit with the reference kernels.  Everything else -- who evaluates them,
// when, and on which data -- is this repo's own design (see surf_trav.cu).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace asurf {

#define ASURF_PI 3.1415926535897931e+0

__device__ __forceinline__ float lerpf(float a, float b, float w) { return fmaf(w, b - a, a); }

// cubic root classes, values of include/data_spec.hpp:25-31
enum : int {
    ROOT_NONE = 200, ROOT_LINEAR = 201, ROOT_POLY_ONE = 202, ROOT_POLY = 203,
    ROOT_CUBIC_THREE = 205, ROOT_CUBIC_ONE = 206
};

// SH basis for one ray from its world-space unit direction (render_util.cuh:373-436).
// out[6] mixes a double literal in the reference, hence the fp64 detour.
__device__ __forceinline__ void eval_sh(int basis_dim, float x, float y, float z, float *__restrict__ out) {
    out[0] = 0.28209479177387814f;
    if (basis_dim >= 4) {
        const float C1 = 0.4886025119029199f;
        out[1] = -C1 * y;
        out[2] = C1 * z;
        out[3] = -C1 * x;
    }
    if (basis_dim >= 9) {
        const float xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, yz = y * z, xz = x * z;
        out[4] = 1.0925484305920792f * xy;
        out[5] = -1.0925484305920792f * yz;
        out[6] = (float)((double)0.31539156525252005f * (2.0 * (double)zz - (double)xx - (double)yy));
        out[7] = -1.0925484305920792f * xz;
        out[8] = 0.5462742152960396f * (xx - yy);
    }
}

// Trilinear field restricted to the ray: f(t) = f0 + f1 t + f2 t^2 + f3 t^3 (render_util.cuh:789-848).
// s: 8 corner scalars (z fastest), o: entry point in [0,1]^3, d: direction.
__device__ __forceinline__ void field_to_cubic(const double *__restrict__ s, const double *__restrict__ o,
                                               const double *__restrict__ d, double *__restrict__ f) {
    double const m00 = s[0] * (1 - o[2]) + s[1] * (o[2]);
    double const m01 = s[2] * (1 - o[2]) + s[3] * (o[2]);
    double const m10 = s[4] * (1 - o[2]) + s[5] * (o[2]);
    double const m11 = s[6] * (1 - o[2]) + s[7] * (o[2]);
    double const k0 = (m01 * d[1] + d[2] * (s[3] - s[2]) * (o[1])) - (m00 * d[1] - d[2] * (s[1] - s[0]) * (1 - o[1]));
    double const k1 = (m11 * d[1] + d[2] * (s[7] - s[6]) * (o[1])) - (m10 * d[1] - d[2] * (s[5] - s[4]) * (1 - o[1]));
    double const h0 = d[1] * d[2] * (s[3] - s[2]) - d[1] * d[2] * (s[1] - s[0]);
    double const h1 = d[1] * d[2] * (s[7] - s[6]) - d[1] * d[2] * (s[5] - s[4]);
    f[3] = h1 * d[0] - h0 * d[0];
    f[2] = k1 * d[0] + h1 * (o[0]) - k0 * d[0] + h0 * (1 - o[0]);
    f[1] = (m10 * (1 - o[1]) + m11 * (o[1])) * d[0] + k1 * (o[0]) - (m00 * (1 - o[1]) + m01 * (o[1])) * d[0] + k0 * (1 - o[0]);
    f[0] = (m00 * (1 - o[1]) + m01 * (o[1])) * (1 - o[0]) + (m10 * (1 - o[1]) + m11 * (o[1])) * (o[0]);
}

// Real roots, ascending, of f0 + f1 t + f2 t^2 + f3 t^3 (Vieta / trigonometric; render_util.cuh:1126-1203).
// st must be pre-set to -1.  Returns the root class.
__device__ __forceinline__ int solve_cubic(double f0, double f1, double f2, double f3, double *__restrict__ st) {
    const double eps = 1e-10;
    if (fabs(f3) < eps) {
        if (fabs(f2) < eps) {
            if (fabs(f1) < eps) return ROOT_NONE;
            st[0] = -f0 / f1;
            return ROOT_LINEAR;
        }
        double const D = f1 * f1 - 4.0 * f2 * f0;
        double const sqrt_D = sqrt(D);
        if (D > 0) {
            if (f2 > 0) {
                st[0] = (-f1 - sqrt_D) / (2 * f2);
                st[1] = (-f1 + sqrt_D) / (2 * f2);
            } else {
                st[0] = (-f1 + sqrt_D) / (2 * f2);
                st[1] = (-f1 - sqrt_D) / (2 * f2);
            }
            if (fabs(st[0] - st[1]) < eps) {
                st[1] = -1;
                return ROOT_POLY_ONE;
            }
            return ROOT_POLY;
        }
        return ROOT_NONE;
    }
    double const b = f2 / f3;
    double const c = f1 / f3;
    double const d = f0 / f3;
    double const Q = ((b) * (b) - 3. * c) / 9.;
    double const R = (2. * ((b) * (b) * (b)) - 9. * b * c + 27. * d) / 54.;
    if (((R) * (R)) < ((Q) * (Q) * (Q))) {
        double const theta = acos(R / sqrt(((Q) * (Q) * (Q))));
        st[0] = -2. * sqrt(Q) * cos(theta / 3.) - b / 3.;
        st[1] = -2. * sqrt(Q) * cos((theta - 2. * ASURF_PI) / 3.) - b / 3.;
        st[2] = -2. * sqrt(Q) * cos((theta + 2. * ASURF_PI) / 3.) - b / 3.;
        return ROOT_CUBIC_THREE;
    }
    double const A = -((R > 0.) ? 1. : -1.) * cbrt(fabs(R) + sqrt(((R) * (R)) - ((Q) * (Q) * (Q))));
    double const B = (A == 0.) ? 0. : Q / A;
    st[0] = (A + B) - b / 3.;
    return ROOT_CUBIC_ONE;
}

// d(root)/d(f0..f3), multiplied into g[0..3] (render_util.cuh:1206-1415).
__device__ __forceinline__ void root_grad(int type, int st_id, const double *__restrict__ fs, float *__restrict__ g) {
#define SQR_(x) ((x) * (x))
#define CUB_(x) ((x) * (x) * (x))
    if (type == ROOT_LINEAR) {
        g[0] *= static_cast<float>(-1. / fs[1]);
        g[1] *= static_cast<float>(fs[0] / SQR_(fs[1]));
        g[2] = 0.f;
        g[3] = 0.f;
    } else if (type == ROOT_POLY_ONE) {
        double const D = SQR_(fs[1]) - 4. * fs[2] * fs[0];
        double const sqrt_D = sqrt(D);
        double const dt0_dD = 1 / (4. * fs[2] * sqrt_D);
        g[0] *= static_cast<float>(-1 / sqrt_D);
        g[1] *= static_cast<float>(((-1) / (2 * fs[2]) + (dt0_dD * 2 * fs[1])));
        g[2] *= static_cast<float>(((fs[1] - sqrt_D) / (4 * SQR_(fs[2])) + (dt0_dD * (-4) * fs[0])));
        g[3] = 0.f;
    } else if (type == ROOT_POLY) {
        double const D = SQR_(fs[1]) - 4.0 * fs[2] * fs[0];
        double const sqrt_D = sqrt(D);
        double const sqr_f2 = SQR_(fs[2]);
        if (st_id == 0) {   // smaller-index root, either sign of f2 (:1246-1262)
            double const dt_dD = -1 / (4 * fs[2] * sqrt_D);
            g[0] *= static_cast<float>(1 / sqrt_D);
            g[1] *= static_cast<float>(((-1) / (2 * fs[2]) + (dt_dD * 2 * fs[1])));
            g[2] *= static_cast<float>(((fs[1] + sqrt_D) / (2 * sqr_f2) + (dt_dD * (-4) * fs[0])));
        } else {
            double const dt_dD = 1 / (4 * fs[2] * sqrt_D);
            g[0] *= static_cast<float>(-1 / sqrt_D);
            g[1] *= static_cast<float>(((-1) / (2 * fs[2]) + (dt_dD * 2 * fs[1])));
            g[2] *= static_cast<float>(((fs[1] - sqrt_D) / (2 * sqr_f2) + (dt_dD * (-4) * fs[0])));
        }
        g[3] = 0.f;
    } else {
        double const norm_term = fs[3];
        double const b = fs[2] / norm_term;
        double const c = fs[1] / norm_term;
        double const d = fs[0] / norm_term;
        double const Q = (SQR_(b) - 3. * c) / 9.;
        double const R = (2. * CUB_(b) - 9. * b * c + 27. * d) / 54.;

        double const DQ3 = fs[1] / (3. * SQR_(fs[3])) - 2. * SQR_(fs[2]) / (9 * CUB_(fs[3]));
        double const DQ2 = 2. * fs[2] / (9. * SQR_(fs[3]));
        double const DQ1 = -1. / (3. * fs[3]);
        double const DQ0 = 0.;
        double const DR3 = -fs[0] / (2. * SQR_(fs[3])) + fs[1] * fs[2] / (3. * CUB_(fs[3])) - CUB_(fs[2]) / (9 * (fs[3] * fs[3] * fs[3] * fs[3]));
        double const DR2 = -fs[1] / (6. * SQR_(fs[3])) + SQR_(fs[2]) / (9 * CUB_(fs[3]));
        double const DR1 = -fs[2] / (6. * SQR_(fs[3]));
        double const DR0 = 1. / (2. * fs[3]);
        double const Db3 = -fs[2] / SQR_(fs[3]);
        double const Db2 = 1. / fs[3];
        double const Db1 = 0.;
        double const Db0 = 0.;
        double const Dst_Db = -1. / 3.;

        if (type == ROOT_CUBIC_THREE) {
            double const theta = acos(R / sqrt(CUB_(Q)));
            double Dst_DQ, Dst_Dtheta;
            double const Dtheta_DQ = 3. * R / (2. * Q * sqrt(1. - SQR_(R) / CUB_(Q)) * sqrt(CUB_(Q)));
            double const Dtheta_DR = -1 / (sqrt(1 - SQR_(R) / CUB_(Q)) * sqrt(CUB_(Q)));
            if (st_id == 0) {
                Dst_DQ = -cos(theta / 3.) / sqrt(Q);
                Dst_Dtheta = 2. * sqrt(Q) * sin(theta / 3.) / 3.;
            } else if (st_id == 1) {
                Dst_DQ = cos(theta / 3. + ASURF_PI / 3.) / sqrt(Q);
                Dst_Dtheta = -2. * sqrt(Q) * sin(theta / 3. + ASURF_PI / 3.) / 3.;
            } else {
                Dst_DQ = sin(theta / 3. + ASURF_PI / 6.) / sqrt(Q);
                Dst_Dtheta = 2. * sqrt(Q) * cos(theta / 3. + ASURF_PI / 6.) / 3.;
            }
            g[0] *= static_cast<float>(Dst_Dtheta * (Dtheta_DQ * DQ0 + Dtheta_DR * DR0) + Dst_DQ * DQ0 + Dst_Db * Db0);
            g[1] *= static_cast<float>(Dst_Dtheta * (Dtheta_DQ * DQ1 + Dtheta_DR * DR1) + Dst_DQ * DQ1 + Dst_Db * Db1);
            g[2] *= static_cast<float>(Dst_Dtheta * (Dtheta_DQ * DQ2 + Dtheta_DR * DR2) + Dst_DQ * DQ2 + Dst_Db * Db2);
            g[3] *= static_cast<float>(Dst_Dtheta * (Dtheta_DQ * DQ3 + Dtheta_DR * DR3) + Dst_DQ * DQ3 + Dst_Db * Db3);
        } else if (type == ROOT_CUBIC_ONE) {
            double const A = -((R > 0.) ? 1. : -1.) * cbrt(fabs(R) + sqrt(SQR_(R) - CUB_(Q)));
            double const sq = fmax(sqrt(-CUB_(Q) + SQR_(R)), 1e-10);
            double const DA_DR = (R >= 0.) ? (-(R / (3. * sq) + 1. / 3.) / fmax(cbrt(SQR_(R + sq)), 1e-10))
                                           : ((R / (3. * sq) - 1. / 3.) / fmax(cbrt(SQR_(-R + sq)), 1e-10));
            double const DA_DQ = (R >= 0.) ? (SQR_(Q) / (2. * sq * cbrt(SQR_(R + sq))))
                                           : (-SQR_(Q) / (2. * sq * cbrt(SQR_(-R + sq))));
            double const DB_DA = (A == 0.) ? 0. : -Q / SQR_(A);
            double const DB_DQ = (A == 0.) ? 0. : 1. / A;
            g[0] *= static_cast<float>((DB_DA + 1.) * (DA_DQ * DQ0 + DA_DR * DR0) + DB_DQ * DQ0 + Dst_Db * Db0);
            g[1] *= static_cast<float>((DB_DA + 1.) * (DA_DQ * DQ1 + DA_DR * DR1) + DB_DQ * DQ1 + Dst_Db * Db1);
            g[2] *= static_cast<float>((DB_DA + 1.) * (DA_DQ * DQ2 + DA_DR * DR2) + DB_DQ * DQ2 + Dst_Db * Db2);
            g[3] *= static_cast<float>((DB_DA + 1.) * (DA_DQ * DQ3 + DA_DR * DR3) + DB_DQ * DQ3 + Dst_Db * Db3);
        }
    }
#undef SQR_
#undef CUB_
}

// d(f0..f3)/d(8 corner scalars) contracted with g (render_util.cuh:850-934). o in [0,1]^3 (float), d dir.
__device__ __forceinline__ void cubic_to_corner_grad(const float *__restrict__ o, const float *__restrict__ d,
                                                     const float *__restrict__ g, float *__restrict__ gs) {
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

// opacity activation and its derivative (render_util.cuh:2138-2188): sigmoid or 1-exp(-x) for x >= 0.
// The EXP branch uses the precise expf, as the reference's `exp(float)` does.
__device__ __forceinline__ float alpha_act(float raw, int type) {
    if (type == 0) return (float)(1. / (1. + (double)__expf(-raw)));
    return (raw >= 0.f) ? 1.f - expf(-raw) : 0.f;
}
__device__ __forceinline__ float alpha_act_grad(float alpha, int type) {
    if (type == 0) return alpha * (1 - alpha);
    return (alpha > 0.f) ? 1 - alpha : 0.f;
}

// truncated Hann re-weighting of the i-th intersection (render_util.cuh:2157-2169)
__device__ __forceinline__ float trunc_rw(int intersect_i, float a, float clamp_min) {
    const float x = (float)intersect_i;
    const float arg = (float)(ASURF_PI * (double)fminf(fmaxf(a - x, 0.f), 1.f));
    return fmaxf(.5f * (1.f - __cosf(arg)), clamp_min);
}

// trilinear interpolation of 8 corner values (z fastest), same lerp order as render_util.cuh:83-90
__device__ __forceinline__ float trilerp8(const float *__restrict__ v, const float *__restrict__ pos) {
    const float ix0y0 = lerpf(v[0], v[1], pos[2]);
    const float ix0y1 = lerpf(v[2], v[3], pos[2]);
    const float ix0 = lerpf(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = lerpf(v[4], v[5], pos[2]);
    const float ix1y1 = lerpf(v[6], v[7], pos[2]);
    const float ix1 = lerpf(ix1y0, ix1y1, pos[1]);
    return lerpf(ix0, ix1, pos[0]);
}

// gradient of the trilinear field at pos (render_util.cuh:2190-2236); out[1] goes through fp64 as there.
__device__ __forceinline__ void field_grad8(const float *__restrict__ v, const float *__restrict__ pos,
                                            float *__restrict__ out) {
    const float ix0y0 = lerpf(v[0], v[1], pos[2]);
    const float ix0y1 = lerpf(v[2], v[3], pos[2]);
    const float ix0 = lerpf(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = lerpf(v[4], v[5], pos[2]);
    const float ix1y1 = lerpf(v[6], v[7], pos[2]);
    const float ix1 = lerpf(ix1y0, ix1y1, pos[1]);
    out[0] = ix1 - ix0;
    out[1] = pos[0] * (-ix1y0 + ix1y1) + (1. - pos[0]) * (-ix0y0 + ix0y1);
    out[2] = pos[0] * (pos[1] * (-v[6] + v[7]) + (1 - pos[1]) * (-v[4] + v[5])) +
             (1 - pos[0]) * (pos[1] * (-v[2] + v[3]) + (1 - pos[1]) * (-v[0] + v[1]));
}

// d(trilerp)/d(pos) accumulated into acc[3] with weight w (render_util.cuh:156-204)
__device__ __forceinline__ void trilerp8_pos_grad(const float *__restrict__ v, const float *__restrict__ pos, float w,
                                                  float *__restrict__ acc) {
    const float ix0y0 = lerpf(v[0], v[1], pos[2]);
    const float ix0y1 = lerpf(v[2], v[3], pos[2]);
    const float ix0 = lerpf(ix0y0, ix0y1, pos[1]);
    const float ix1y0 = lerpf(v[4], v[5], pos[2]);
    const float ix1y1 = lerpf(v[6], v[7], pos[2]);
    const float ix1 = lerpf(ix1y0, ix1y1, pos[1]);
    acc[0] += w * (ix1 - ix0);
    acc[1] += w * ((1 - pos[0]) * (ix0y1 - ix0y0) + (pos[0]) * (ix1y1 - ix1y0));
    acc[2] += w * ((1 - pos[0]) * ((1 - pos[1]) * (v[1] - v[0]) + (pos[1]) * (v[3] - v[2])) +
                   (pos[0]) * ((1 - pos[1]) * (v[5] - v[4]) + (pos[1]) * (v[7] - v[6])));
}

// the 8 trilinear corner weights times g, in the reference's evaluation order (render_util.cuh:103-120)
__device__ __forceinline__ void corner_weights(const float *__restrict__ pos, float g, float *__restrict__ w) {
    const float ay = 1.f - pos[1], az = 1.f - pos[2];
    float xo = (1.0f - pos[0]) * g;
    w[0] = ay * az * xo;
    w[1] = ay * pos[2] * xo;
    w[2] = pos[1] * az * xo;
    w[3] = pos[1] * pos[2] * xo;
    xo = pos[0] * g;
    w[4] = ay * az * xo;
    w[5] = ay * pos[2] * xo;
    w[6] = pos[1] * az * xo;
    w[7] = pos[1] * pos[2] * xo;
}

}  // namespace asurf
