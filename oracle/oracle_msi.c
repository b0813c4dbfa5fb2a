/*
 * TEST INFRASTRUCTURE ONLY -- see oracle_common.h.
 *
 * CPU restatement of the multi-sphere-image background of /root/reference/svox2/csrc:
 *   render_background_forward / _backward   render_lerp_kernel_surf_trav.cu:2914-3137 (kernels :3370-3455)
 *   ray_find_bounds_bg, ConcentricSpheresIntersector, _unitvec2equirect   include/render_util.cuh:555-563, :619-649, :703-745
 *   trilerp_bg_one / trilerp_backward_bg_one   include/render_util.cuh:207-283
 *   msi_tv_grad_sparse                         loss_kernel.cu:979-1064, :1624-1659
 * One loop iteration == one CUDA thread (ray) of the reference.  The per-ray inputs a foreground pass leaves (final
 * log-transmittance, leftover accum) are arguments: the foreground renderers have their own restatements.
 * Pinning: tests/test_msi_gpu.py compares this file, our kernels and the UNMODIFIED reference CUDA build on the GPU box.
 */
#include "oracle_common.h"

#define MSI_C0 0.28209479177387814f

typedef struct {
    const int32_t *links;
    const float *data;
    int reso, nlayers;
    int32_t size[3];
    float offset[3], scaling[3];
} OMsi;

static float rnorm3(float a, float b, float c) { return 1.f / sqrtf(a * a + b * b + c * c); }

static void msi_ray(const OMsi *m, float *o, float *d, float *world_step) { /* ray_find_bounds_bg :703-730 */
    for (int i = 0; i < 3; ++i) {
        o[i] = fmaf(o[i], m->scaling[i], m->offset[i]);
        d[i] *= m->scaling[i];
    }
    const float ds = rnorm3(d[0], d[1], d[2]);
    for (int i = 0; i < 3; ++i) d[i] *= ds;
    *world_step = ds;
    for (int i = 0; i < 3; ++i) {
        const float ss = 2.f / (float)m->size[i];
        o[i] = fmaf(o[i] + 0.5f, ss, -1.f);
        d[i] = d[i] * ss;
    }
    const float inorm = rnorm3(d[0], d[1], d[2]);
    *world_step *= inorm;
    for (int i = 0; i < 3; ++i) d[i] *= inorm;
}

static int msi_pos(const OMsi *m, const float *o, const float *d, float q2a, float qb, float f, float radius, float inner,
                   int *l, float *pos, float *invr_mid) {
    const float det = f + 2 * q2a * radius * radius;
    if (radius < inner || det < 0) return 0;
    const float t = (-qb + sqrtf(det)) / q2a;
    for (int j = 0; j < 3; ++j) pos[j] = fmaf(t, d[j], o[j]);
    *invr_mid = rnorm3(pos[0], pos[1], pos[2]);
    for (int j = 0; j < 3; ++j) pos[j] *= *invr_mid;
    const float lat = asinf(pos[1]), lon = atan2f(pos[0], pos[2]);
    pos[0] = (float)(m->reso * 2 * (0.5 + lon * 0.5 * 0.318309886183790671538));
    pos[1] = (float)(m->reso * (0.5 - lat * 0.318309886183790671538));
    pos[2] = o_minf(o_maxf((1.f - *invr_mid) * m->nlayers - 0.5f, 0.f), (float)(m->nlayers - 1));
    for (int j = 0; j < 3; ++j) l[j] = (int)pos[j];
    if (l[0] > m->reso * 2 - 1) l[0] = m->reso * 2 - 1;
    if (l[1] > m->reso - 1) l[1] = m->reso - 1;
    if (l[2] > m->nlayers - 2) l[2] = m->nlayers - 2;
    for (int j = 0; j < 3; ++j) pos[j] -= (float)l[j];
    return 1;
}

static void corners(const OMsi *m, const int *l, int *u) {
    const int ny = l[1] < (m->reso - 1) ? (l[1] + 1) : 0;
    const int nx = l[0] < (2 * m->reso - 1) ? (l[0] + 1) : 0;
    u[0] = m->reso * l[0] + l[1]; u[1] = m->reso * l[0] + ny; u[2] = m->reso * nx + l[1]; u[3] = m->reso * nx + ny;
}

static float msi_trilerp(const OMsi *m, const int *l, const float *pos, int idx) { /* trilerp_bg_one :207-240 */
    int u[4];
    float v[4];
    corners(m, l, u);
    for (int c = 0; c < 4; ++c) {
        const int link = m->links[u[c]];
        if (link >= 0) {
            const float *dp = m->data + ((int64_t)link * m->nlayers + l[2]) * 4 + idx;
            v[c] = o_lerp(dp[0], dp[4], pos[2]);
        } else v[c] = 0.f;
    }
    return o_lerp(o_lerp(v[0], v[1], pos[1]), o_lerp(v[2], v[3], pos[1]), pos[0]);
}

static void msi_trilerp_backward(const OMsi *m, float *grad, uint8_t *mask, const int *l, const float *pos, float g, int idx) {
    int u[4];
    corners(m, l, u);
    const float ay = 1.f - pos[1], az = 1.f - pos[2];
    const float xo0 = (1.0f - pos[0]) * g, xo1 = pos[0] * g;
    const float w[4] = {ay * xo0, pos[1] * xo0, ay * xo1, pos[1] * xo1};
    for (int c = 0; c < 4; ++c) {
        const int link = m->links[u[c]];
        if (link >= 0) {
            const int64_t row = (int64_t)link * m->nlayers + l[2];
            grad[row * 4 + idx] += w[c] * az;
            grad[(row + 1) * 4 + idx] += w[c] * pos[2];
            if (mask) { mask[row] = 1; mask[row + 1] = 1; }
        }
    }
}

static OMsi make(const int32_t *links, const float *data, int reso, int nlayers, const int32_t *size, const float *offset,
                 const float *scaling) {
    OMsi m;
    m.links = links; m.data = data; m.reso = reso; m.nlayers = nlayers;
    for (int i = 0; i < 3; ++i) { m.size[i] = size[i]; m.offset[i] = offset[i]; m.scaling[i] = scaling[i]; }
    return m;
}

/* render_background_kernel :3370-3387 + render_background_forward :2914-3004: rgb (Q,3) += */
void oracle_msi_forward(const int32_t *links, const float *data, int reso, int nlayers, const int32_t *size, const float *offset,
                        const float *scaling, const OOpt *opt, const float *origins, const float *dirs, int64_t Q,
                        const float *log_transmit_in, float *rgb) {
    const OMsi m = make(links, data, reso, nlayers, size, offset, scaling);
    for (int64_t r = 0; r < Q; ++r) {
        float lt = log_transmit_in[r];
        if (lt < -25.f) continue;
        float o[3] = {origins[r * 3], origins[r * 3 + 1], origins[r * 3 + 2]}, d[3] = {dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]};
        float ws;
        msi_ray(&m, o, d, &ws);
        const float q2a = 2 * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), qb = 2 * (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]);
        const float f = qb * qb - 2 * q2a * (o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
        const float c0 = o[1] * d[2] - o[2] * d[1], c1 = o[2] * d[0] - o[0] * d[2], c2 = o[0] * d[1] - o[1] * d[0];
        const float inner = o_maxf(sqrtf(c0 * c0 + c1 * c1 + c2 * c2) + 1e-3f, 1.f);
        float invr_last = 1.f / inner;
        const int n_steps = (int)(nlayers / opt->step_size) + 2;
        float outv[3] = {0, 0, 0};
        for (int i = 0; i < n_steps; ++i) {
            const float radius = (float)(n_steps / (n_steps - i - 0.5));
            int l[3];
            float pos[3], invr_mid;
            if (!msi_pos(&m, o, d, q2a, qb, f, radius, inner, l, pos, &invr_mid)) continue;
            const float sigma = msi_trilerp(&m, l, pos, 3);
            if (sigma > 0.f) {
                const float pcnt = (invr_last - invr_mid) * ws * sigma;
                const float weight = expf(lt) * (1.f - expf(-pcnt));
                lt -= pcnt;
                for (int c = 0; c < 3; ++c) outv[c] += weight * o_maxf(msi_trilerp(&m, l, pos, c) * MSI_C0 + 0.5f, 0.f);
                if (expf(lt) < opt->stop_thresh) break;
            }
            invr_last = invr_mid;
        }
        for (int c = 0; c < 3; ++c) rgb[r * 3 + c] += outv[c] + expf(lt) * opt->background_brightness;
    }
}

/* render_background_backward_kernel :3413-3455 + render_background_backward :3006-3137.  accum: leftover of the foreground
 * backward per ray (the caller supplies it; no sentinels here). */
void oracle_msi_backward(const int32_t *links, const float *data, int reso, int nlayers, const int32_t *size,
                         const float *offset, const float *scaling, const OOpt *opt, const float *origins, const float *dirs,
                         int64_t Q, const float *grad_in, const float *color_cache, int grad_is_rgb, const float *log_transmit_in,
                         const float *accum_in, float sparsity_loss, float *grad_bg, uint8_t *mask_bg) {
    const OMsi m = make(links, data, reso, nlayers, size, offset, scaling);
    const float norm_factor = 2.f / (float)(3 * (int)Q);
    for (int64_t r = 0; r < Q; ++r) {
        float lt = log_transmit_in[r];
        if (lt < -25.f) continue;
        float o[3] = {origins[r * 3], origins[r * 3 + 1], origins[r * 3 + 2]}, d[3] = {dirs[r * 3], dirs[r * 3 + 1], dirs[r * 3 + 2]};
        float ws;
        msi_ray(&m, o, d, &ws);
        float go[3];
        for (int c = 0; c < 3; ++c)
            go[c] = grad_is_rgb ? (color_cache[r * 3 + c] - grad_in[r * 3 + c]) * norm_factor : grad_in[r * 3 + c];
        float accum = accum_in[r];
        const float q2a = 2 * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), qb = 2 * (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]);
        const float f = qb * qb - 2 * q2a * (o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
        const float c0 = o[1] * d[2] - o[2] * d[1], c1 = o[2] * d[0] - o[0] * d[2], c2 = o[0] * d[1] - o[1] * d[0];
        const float inner = o_maxf(sqrtf(c0 * c0 + c1 * c1 + c2 * c2) + 1e-3f, 1.f);
        float invr_last = 1.f / inner;
        const int n_steps = (int)(nlayers / opt->step_size) + 2;
        for (int i = 0; i < n_steps; ++i) {
            const float radius = (float)(n_steps / (n_steps - i - 0.5));
            int l[3];
            float pos[3], invr_mid;
            if (!msi_pos(&m, o, d, q2a, qb, f, radius, inner, l, pos, &invr_mid)) continue;
            const float sigma = msi_trilerp(&m, l, pos, 3);
            if (sigma > 0.f) {
                float total_color = 0.f;
                const float pcnt = ws * (invr_last - invr_mid) * sigma;
                const float weight = expf(lt) * (1.f - expf(-pcnt));
                lt -= pcnt;
                for (int c = 0; c < 3; ++c) {
                    const float color = msi_trilerp(&m, l, pos, c) * MSI_C0 + 0.5f;
                    total_color += o_maxf(color, 0.f) * go[c];
                    if (color > 0.f) msi_trilerp_backward(&m, grad_bg, NULL, l, pos, MSI_C0 * weight * go[c], c);
                }
                accum -= weight * total_color;
                float gs = ws * (invr_last - invr_mid) * (total_color * expf(lt) - accum);
                if (sparsity_loss > 0.f) gs += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
                msi_trilerp_backward(&m, grad_bg, mask_bg, l, pos, gs, 3);
                if (expf(lt) < opt->stop_thresh) break;
            }
            invr_last = invr_mid;
        }
    }
}

/* msi_tv_grad_sparse_kernel :979-1064 + host :1624-1659 */
void oracle_msi_tv_grad_sparse(const int32_t *links, int lx, int ly, const float *msi, int nlayers, int nch,
                               const int32_t *cells, int64_t n_cells, uint8_t *mask, float scale, float scale_last, float *grad) {
    const float nl = (float)(int)n_cells;
    scale /= nl;
    scale_last /= nl;
    for (int64_t i = 0; i < n_cells; ++i)
        for (int ch = 0; ch < nch; ++ch) {
            const int idx = cells[i];
            const int z = idx % nlayers, tmp = idx / nlayers, y = tmp % ly, x = tmp / ly;
            const int nx = (x == lx - 1) ? 0 : x + 1, ny = (y == ly - 1) ? 0 : y + 1;
            const int l00 = links[x * ly + y], l01 = links[x * ly + ny], l10 = links[nx * ly + y];
#define MSIV(l, zz) msi[((int64_t)(l) * nlayers + (zz)) * nch + ch]
            const float v00 = l00 >= 0 ? MSIV(l00, z) : 0.f;
            const float v_nxl = (l00 >= 0 && z + 1 < nlayers) ? MSIV(l00, z + 1) : ((ch == nch - 1) ? 0.f : v00);
            const float v01 = l01 >= 0 ? MSIV(l01, z) : 0.f;
            const float v10 = l10 >= 0 ? MSIV(l10, z) : 0.f;
            const float sc = (ch == nch - 1) ? scale_last : scale;
            float dx = v10 - v00, dy = v01 - v00, dz = v_nxl - v00;
            const float idelta = sc * (1.f / sqrtf(1e-9f + dx * dx + dy * dy + dz * dz));
            dx *= lx * (1.f / 256.f);
            dy *= ly * (1.f / 256.f);
            dz *= nlayers * (1.f / 256.f);
            const float sm = -(dx + dy + dz);
#define ADDSET(l, zz, val) if ((l) >= 0 && (val) != 0.f) { grad[((int64_t)(l) * nlayers + (zz)) * nch + ch] += (val) * idelta; if (mask) mask[(int64_t)(l) * nlayers + (zz)] = 1; }
            ADDSET(l00, z, sm);
            if (z + 1 < nlayers) { ADDSET(l00, z + 1, dz); }
            ADDSET(l01, z, dy);
            ADDSET(l10, z, dx);
#undef ADDSET
#undef MSIV
        }
}
