/* TEST INFRASTRUCTURE ONLY (see oracle/README.md): CPU restatement of the grid-maintenance renders of
 * /root/reference/svox2/csrc/misc_kernel.cu -- dilate (:24-54), grid_trace_ray (:187-284), sprase_grid_trace_ray (:287-401),
 * sprase_grid_mask_trace_ray (:403-509); cam2world_ray (include/render_util.cuh:599-617).
 * The level-set visibility pass (:511-719) lives in oracle_surf_trav.c next to the cubic solver it needs.
 * Parity pinned on the GPU against the UNMODIFIED reference kernels (tests/test_gridtools_gpu.py). */
#include "oracle_common.h"

/* cam2world_ray for every pixel, raster order; no NDC */
void oracle_cam_rays(const float *c2w, float fx, float fy, float cx, float cy, int width, int height, float *origins,
                     float *dirs) {
    for (int iy = 0; iy < height; ++iy)
        for (int ix = 0; ix < width; ++ix) {
            float x = ((float)ix + 0.5f - cx) / fx;
            float y = ((float)iy + 0.5f - cy) / fy;
            float z = sqrtf((float)((double)(x * x + y * y) + 1.0));
            x /= z; y /= z; z = 1.0f / z;
            float *d = dirs + ((int64_t)iy * width + ix) * 3, *o = origins + ((int64_t)iy * width + ix) * 3;
            d[0] = c2w[0] * x + c2w[1] * y + c2w[2] * z;
            d[1] = c2w[4] * x + c2w[5] * y + c2w[6] * z;
            d[2] = c2w[8] * x + c2w[9] * y + c2w[10] * z;
            o[0] = c2w[3]; o[1] = c2w[7]; o[2] = c2w[11];
        }
}

void oracle_dilate(const uint8_t *in, const int32_t *size, uint8_t *out) {
    const int sx = size[0], sy = size[1], sz = size[2];
    for (int x = 0; x < sx; ++x)
        for (int y = 0; y < sy; ++y)
            for (int z = 0; z < sz; ++z) {
                int any = 0;
                for (int a = o_maxi(x - 1, 0); a <= o_mini(x + 1, sx - 1); ++a)
                    for (int b = o_maxi(y - 1, 0); b <= o_mini(y + 1, sy - 1); ++b)
                        for (int c = o_maxi(z - 1, 0); c <= o_mini(z + 1, sz - 1); ++c)
                            any |= in[((int64_t)a * sy + b) * sz + c];
                out[((int64_t)x * sy + y) * sz + z] = any ? 1 : 0;
            }
}

/* world -> grid and the AABB bounds from t = 0 (:198-219); returns delta_scale.  xf (9 floats: origin, dir, t, tmax,
 * delta_scale) overrides the computation with what the GPU derived (rnorm3df vs 1/sqrt differ in the last bit). */
static float gt_bounds(const int32_t *size, const float *offset, const float *scaling, const float *xf, float *o, float *d,
                       float *t, float *tmax) {
    if (xf) {
        for (int i = 0; i < 3; ++i) { o[i] = xf[i]; d[i] = xf[3 + i]; }
        *t = xf[6]; *tmax = xf[7];
        return xf[8];
    }
    for (int i = 0; i < 3; ++i) { o[i] = fmaf(o[i], scaling[i], offset[i]); d[i] *= scaling[i]; }
    const double n2 = (double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2];
    const float delta_scale = (float)(1.0 / sqrt(n2));
    for (int i = 0; i < 3; ++i) d[i] *= delta_scale;
    *t = 0.f;
    *tmax = 2e3f;
    for (int i = 0; i < 3; ++i) {
        const float invdir = (float)(1.0 / (double)d[i]);
        const float t1 = (-0.5f - o[i]) * invdir, t2 = ((float)size[i] - 0.5f - o[i]) * invdir;
        if (d[i] != 0.f) { *t = o_maxf(*t, o_minf(t1, t2)); *tmax = o_minf(*tmax, o_maxf(t1, t2)); }
    }
    return delta_scale;
}

static void gt_sample(const int32_t *size, const float *o, const float *d, float t, int32_t *l, float *pos) {
    for (int j = 0; j < 3; ++j) {
        pos[j] = fmaf(t, d[j], o[j]);
        pos[j] = o_minf(o_maxf(pos[j], 0.f), size[j] - 1.f);
        l[j] = o_mini((int32_t)pos[j], size[j] - 2);
        pos[j] -= (float)l[j];
    }
}

static void max8(float *q, int64_t s0, int s1, float w) {
    const int64_t u[8] = {0, 1, s1, s1 + 1, s0, s0 + 1, s0 + s1, s0 + s1 + 1};
    for (int c = 0; c < 8; ++c) q[u[c]] = fmaxf(w, q[u[c]]);
}

/* sparse == 0: data is a dense (X,Y,Z) volume, vertices take the max weight; sparse != 0: data is (N,1) behind links,
 * vertices take the max transmittance in front of the sample. */
void oracle_weight_render(int sparse, const float *data, const int32_t *links, const int32_t *size, const float *offset,
                          const float *scaling, const float *origins, const float *dirs, const float *xf, int64_t Q,
                          float step_size, float stop_thresh, int last_sample_opaque, float *grid_weight) {
    const int64_t s0 = (int64_t)size[1] * size[2];
    const int s1 = size[2];
    for (int64_t q = 0; q < Q; ++q) {
        float o[3] = {origins[q * 3], origins[q * 3 + 1], origins[q * 3 + 2]}, d[3] = {dirs[q * 3], dirs[q * 3 + 1], dirs[q * 3 + 2]};
        float t, tmax;
        const float world_step = gt_bounds(size, offset, scaling, xf ? xf + q * 9 : NULL, o, d, &t, &tmax) * step_size;
        if (t > tmax) continue;
        float log_light = 0.f;
        while (t <= tmax) {
            int32_t l[3];
            float pos[3];
            gt_sample(size, o, d, t, l, pos);
            const int64_t idx = l[0] * s0 + (int64_t)l[1] * s1 + l[2];
            float sigma;
            if (sparse) {
                sigma = o_trilerp_cuvol_one(links, data, (int)s0, s1, 1, l, pos, 0);
            } else {
                const float *p = data + idx;
                const float ix0y0 = o_lerp(p[0], p[1], pos[2]), ix0y1 = o_lerp(p[s1], p[s1 + 1], pos[2]);
                const float ix1y0 = o_lerp(p[s0], p[s0 + 1], pos[2]), ix1y1 = o_lerp(p[s0 + s1], p[s0 + s1 + 1], pos[2]);
                sigma = o_lerp(o_lerp(ix0y0, ix0y1, pos[1]), o_lerp(ix1y0, ix1y1, pos[1]), pos[0]);
                if (last_sample_opaque && t + step_size > tmax) { sigma += 1e9f; log_light = 0.f; }
            }
            if (sigma > 1e-8f) {
                const float log_att = -world_step * sigma;
                const float w = sparse ? expf(log_light) : expf(log_light) * (1.f - expf(log_att));
                max8(grid_weight + idx, s0, s1, w);
                log_light += log_att;
                if (expf(log_light) < stop_thresh) break;
            }
            t += step_size;
        }
    }
}

void oracle_mask_render(const int32_t *links, const int32_t *size, const float *offset, const float *scaling,
                        const float *origins, const float *dirs, const float *xf, int64_t Q, float near_clip,
                        float *grid_mask) {
    const int64_t s0 = (int64_t)size[1] * size[2];
    const int s1 = size[2];
    const float step_size = 0.1f;
    for (int64_t q = 0; q < Q; ++q) {
        float o[3] = {origins[q * 3], origins[q * 3 + 1], origins[q * 3 + 2]}, d[3] = {dirs[q * 3], dirs[q * 3 + 1], dirs[q * 3 + 2]};
        float t, tmax;
        gt_bounds(size, offset, scaling, xf ? xf + q * 9 : NULL, o, d, &t, &tmax);
        if (t < near_clip) t = near_clip;
        if (t > tmax) continue;
        while (t <= tmax) {
            int32_t l[3];
            float pos[3];
            gt_sample(size, o, d, t, l, pos);
            const int32_t *lp = links + (l[0] * s0 + (int64_t)l[1] * s1 + l[2]);
            const int64_t u[8] = {0, 1, s1, s1 + 1, s0, s0 + 1, s0 + s1, s0 + s1 + 1};
            for (int c = 0; c < 8; ++c)
                if (lp[u[c]] >= 0) grid_mask[lp[u[c]]] = fmaxf(1.f, grid_mask[lp[u[c]]]);
            t += step_size;
        }
    }
}

/* ---- point queries, svox2_kernel.cu:11-246 ---- */
static void sp_locate(const int32_t *size, const float *offset, const float *scaling, const float *pt, int32_t *l, float *w) {
    for (int i = 0; i < 3; ++i) {
        float p = fmaf(pt[i], scaling[i], offset[i]);
        p = fminf(fmaxf(p, 0.f), size[i] - 1.f);
        l[i] = o_mini((int32_t)p, size[i] - 2);
        w[i] = p - (float)l[i];
    }
}

void oracle_sample_grid(const int32_t *links, const int32_t *size, const float *offset, const float *scaling, const float *data,
                        int n_cols, float missing, const float *points, int64_t n_points, float *out) {
    const int offy = size[2], offx = size[1] * size[2];
    for (int64_t p = 0; p < n_points; ++p) {
        int32_t l[3];
        float w[3];
        sp_locate(size, offset, scaling, points + p * 3, l, w);
        const int32_t *lp = links + ((int64_t)offx * l[0] + (int64_t)offy * l[1] + l[2]);
        for (int idx = 0; idx < n_cols; ++idx) {
#define RD(u) ((lp[u] >= 0) ? data[(int64_t)lp[u] * n_cols + idx] : missing)
            const float ix0y0 = o_lerp(RD(0), RD(1), w[2]), ix0y1 = o_lerp(RD(offy), RD(offy + 1), w[2]);
            const float ix0 = o_lerp(ix0y0, ix0y1, w[1]);
            const float ix1y0 = o_lerp(RD(offx), RD(offx + 1), w[2]), ix1y1 = o_lerp(RD(offy + offx), RD(offy + offx + 1), w[2]);
            const float ix1 = o_lerp(ix1y0, ix1y1, w[1]);
#undef RD
            out[p * n_cols + idx] = o_lerp(ix0, ix1, w[0]);
        }
    }
}

void oracle_sample_grid_backward(const int32_t *links, const int32_t *size, const float *offset, const float *scaling,
                                 const float *points, int64_t n_points, const float *grad_out, int n_cols, float *grad_data) {
    const int offy = size[2], offx = size[1] * size[2];
    for (int64_t p = 0; p < n_points; ++p) {
        int32_t l[3];
        float w[3];
        sp_locate(size, offset, scaling, points + p * 3, l, w);
        const int32_t *lp = links + ((int64_t)offx * l[0] + (int64_t)offy * l[1] + l[2]);
        const int u[8] = {0, 1, offy, offy + 1, offx, offx + 1, offx + offy, offx + offy + 1};
        for (int idx = 0; idx < n_cols; ++idx) {
            const float go = grad_out[p * n_cols + idx];
            const float xb = w[0], yb = w[1], zb = w[2], xa = 1.f - w[0], ya = 1.f - w[1], za = 1.f - w[2];
            const float xago = xa * go, xbgo = xb * go;
            const float t00 = ya * xago, t01 = yb * xago, t10 = ya * xbgo, t11 = yb * xbgo;
            const float c[8] = {t00 * za, t00 * zb, t01 * za, t01 * zb, t10 * za, t10 * zb, t11 * za, t11 * zb};
            for (int q = 0; q < 8; ++q)
                if (lp[u[q]] >= 0) grad_data[(int64_t)lp[u[q]] * n_cols + idx] += c[q];
        }
    }
}
