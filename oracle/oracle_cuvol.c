/*
 * TEST INFRASTRUCTURE ONLY -- see oracle_common.h.
 *
 * CPU restatement of the Plenoxels "cuvol" renderer of /root/reference/svox2/csrc/render_lerp_kernel_cuvol.cu:
 * trace_ray_cuvol (:30-125), trace_ray_cuvol_backward (:371-535), the kernels' per-ray set-up (:762-919) and
 * compute_skip_dist (include/render_util.cuh:286-368).  One loop iteration == one warp (ray) of the reference; the SH
 * lanes become an inner loop over the D coefficients.  expf / logf stand in for the fast intrinsics (1e-4 tolerance).
 * Pinning: tests/golden/l0_cuvol_*.npz (the reference's pure-PyTorch renderer, svox2/svox2.py:1215-1441, via
 * oracle/gen_golden.py) and, on the GPU box, the UNMODIFIED reference kernels (oracle/_ref).
 */
#include "oracle_common.h"

/* include/render_util.cuh:286-368 (pos_offset = 0) */
static float compute_skip_dist(const ORay *ray, const int32_t *links, int offx, int offy) {
    const int32_t link_val = links[(int64_t)offx * ray->l[0] + (int64_t)offy * ray->l[1] + ray->l[2]];
    if (link_val >= -1) return 0.f;
    const uint32_t dist = (uint32_t)(-link_val);
    const uint32_t cell_ul_shift = dist - 1;
    const uint32_t cell_side_len = (uint32_t)((float)(1 << cell_ul_shift) - 1.f);
    float tmin = 0.f, tmax = 1e9f;
    for (int i = 0; i < 3; ++i) {
        int ul = ((ray->l[i] >> cell_ul_shift) << cell_ul_shift);
        ul -= ray->l[i];
        const float invdir = (float)(1.0 / (double)ray->dir[i]);
        const float t1 = ((float)ul - ray->pos[i] + 0.f) * invdir;
        const float t2 = ((float)((uint32_t)ul + cell_side_len) - ray->pos[i] + 0.f) * invdir;
        if (ray->dir[i] != 0.f) {
            tmin = o_maxf(tmin, o_minf(t1, t2));
            if (o_maxf(t1, t2) < tmax) tmax = o_maxf(t1, t2);
        }
    }
    if (tmin > 0.f) return 0.f;
    return tmax;
}

static void sample_position(ORay *ray, const OGrid *g, float t) { /* :57-63 */
    for (int j = 0; j < 3; ++j) {
        ray->pos[j] = fmaf(t, ray->dir[j], ray->origin[j]);
        ray->pos[j] = o_minf(o_maxf(ray->pos[j], 0.f), g->size[j] - 1.f);
        ray->l[j] = o_mini((int32_t)ray->pos[j], g->size[j] - 2);
        ray->pos[j] -= (float)ray->l[j];
    }
}

/* per-channel sums of the D lane colours in the reference's segmented shuffle order */
static void channel_sums(const OGrid *g, const ORay *ray, const float *sph, int offx, int offy, float *c) {
    const int bd = g->basis_dim;
    float lane[32];
    for (int k = 0; k < g->sh_dim; ++k)
        lane[k] = o_trilerp_cuvol_one(g->links, g->sh, offx, offy, (size_t)g->sh_dim, ray->l, ray->pos, k) * sph[k % bd];
    for (int ch = 0; ch < 3; ++ch) c[ch] = o_seg_sum(lane + ch * bd, bd);
}

/* render_ray_kernel :762-800 + trace_ray_cuvol :30-125 */
/* xf: optional (Q,9) grid-space rays (origin3, dir3, tmin, tmax, world_step) as the GPU kernels see them; it decouples
 * the parity tests from the last-bit difference between rnorm3df (GPU) and 1/sqrtf (host) in ray_find_bounds. */
static void ray_setup(ORay *ray, const OGrid *g, const OOpt *opt, const float *origins, const float *dirs, const float *xf,
                      int64_t q, float *sph) {
    for (int i = 0; i < 3; ++i) { ray->origin[i] = origins[q * 3 + i]; ray->dir[i] = dirs[q * 3 + i]; }
    o_calc_sh(g->basis_dim, ray->dir, sph);
    if (xf) {
        for (int i = 0; i < 3; ++i) { ray->origin[i] = xf[q * 9 + i]; ray->dir[i] = xf[q * 9 + 3 + i]; }
        ray->tmin = xf[q * 9 + 6]; ray->tmax = xf[q * 9 + 7]; ray->world_step = xf[q * 9 + 8];
    } else {
        o_ray_find_bounds(ray, g, opt);
    }
}

void oracle_cuvol_forward(const OGrid *g, const OOpt *opt, const float *origins, const float *dirs, const float *xf, int64_t Q,
                          float *rgb_out, float *log_transmit_out) {
    const int offx = g->size[1] * g->size[2], offy = g->size[2];
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t q = 0; q < Q; ++q) {
        ORay ray;
        float sph[9];
        ray_setup(&ray, g, opt, origins, dirs, xf, q, sph);
        float *out = rgb_out + q * 3;
        if (ray.tmin > ray.tmax) {
            out[0] = out[1] = out[2] = opt->background_brightness;
            if (log_transmit_out) log_transmit_out[q] = 0.f;
            continue;
        }
        float t = ray.tmin, outv[3] = {0.f, 0.f, 0.f}, log_transmit = 0.f;
        while (t <= ray.tmax) {
            sample_position(&ray, g, t);
            const float skip = compute_skip_dist(&ray, g->links, offx, offy);
            if (skip >= opt->step_size) {
                t += ceilf(skip / opt->step_size) * opt->step_size;
                continue;
            }
            const float sigma = o_trilerp_cuvol_one(g->links, g->density, offx, offy, 1, ray.l, ray.pos, 0);
            if (opt->last_sample_opaque && t + opt->step_size > ray.tmax) ray.world_step = 1e9f;
            if (sigma > opt->sigma_thresh) {
                float c[3];
                channel_sums(g, &ray, sph, offx, offy, c);
                const float pcnt = ray.world_step * sigma;
                const float weight = expf(log_transmit) * (1.f - expf(-pcnt));
                log_transmit -= pcnt;
                for (int ch = 0; ch < 3; ++ch) outv[ch] += weight * o_maxf(c[ch] + 0.5f, 0.f);
                if (expf(log_transmit) < opt->stop_thresh) {
                    log_transmit = -1e3f;
                    break;
                }
            }
            t += opt->step_size;
        }
        for (int ch = 0; ch < 3; ++ch) out[ch] = outv[ch] + expf(log_transmit) * opt->background_brightness;
        if (log_transmit_out) log_transmit_out[q] = log_transmit;
    }
}

/* Depth renders of the cuvol backend: trace_ray_expected_term (render_lerp_kernel_cuvol.cu:127-188), trace_ray_mode_term
 * (:190-257), trace_ray_med_term (:259-319), trace_ray_sigma_thresh (:322-369).
 * mode 0 expected depth (param = weight_thresh), 1 depth of the heaviest sample (param = weight_thresh), 2 per-sample
 * depths and sigmas of the first max_sample samples (out, out2: (Q, max_sample), zero-filled), 3 depth of the first sample
 * with sigma > param. */
void oracle_cuvol_scalar(const OGrid *g, const OOpt *opt, const float *origins, const float *dirs, const float *xf, int64_t Q,
                         int mode, float param, int max_sample, float *out, float *out2) {
    const int offx = g->size[1] * g->size[2], offy = g->size[2];
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t q = 0; q < Q; ++q) {
        ORay ray;
        float sph[9];
        ray_setup(&ray, g, opt, origins, dirs, xf, q, sph);
        if (mode == 2) {
            for (int i = 0; i < max_sample; ++i) out[q * max_sample + i] = out2[q * max_sample + i] = 0.f;
        } else {
            out[q] = 0.f;
        }
        if (ray.tmin > ray.tmax) continue;
        float t = ray.tmin, outv = 0.f, weight_acc = 0.f, max_weight = -1.f, log_transmit = 0.f;
        int sample_i = 0, found = 0;
        while (t <= ray.tmax) {
            sample_position(&ray, g, t);
            const float skip = compute_skip_dist(&ray, g->links, offx, offy);
            if (skip >= opt->step_size) {
                t += ceilf(skip / opt->step_size) * opt->step_size;
                continue;
            }
            const float sigma = o_trilerp_cuvol_one(g->links, g->density, offx, offy, 1, ray.l, ray.pos, 0);
            if (mode == 3) {
                if (sigma > param) { out[q] = (t / opt->step_size) * ray.world_step; found = 1; break; }
            } else if (sigma > opt->sigma_thresh) {
                const float pcnt = ray.world_step * sigma;
                const float weight = expf(log_transmit) * (1.f - expf(-pcnt));
                log_transmit -= pcnt;
                if (mode == 0) {
                    outv += weight * (t / opt->step_size) * ray.world_step;
                    weight_acc += weight;
                } else if (mode == 1) {
                    weight_acc += weight;
                    if (weight > max_weight) { max_weight = weight; outv = (t / opt->step_size) * ray.world_step; }
                } else if (sample_i < max_sample) {
                    out[q * max_sample + sample_i] = (t / opt->step_size) * ray.world_step;
                    out2[q * max_sample + sample_i] = sigma;
                    sample_i += 1;
                }
                if (expf(log_transmit) < opt->stop_thresh) break;
            }
            t += opt->step_size;
        }
        (void)found;
        if (mode <= 1) out[q] = (weight_acc > param) ? outv : 0.f;
    }
}

/* render_ray_backward_kernel :844-919 + trace_ray_cuvol_backward :371-535.
 * grad_is_rgb: grad_in is rgb_gt and dL/dRGB = (colour - gt) * norm_factor (fused, :873-880). */
void oracle_cuvol_backward(const OGrid *g, const OOpt *opt, const float *origins, const float *dirs, const float *xf, int64_t Q,
                           const float *grad_in, const float *color_cache, int grad_is_rgb, float norm_factor,
                           const float *log_transmit_in, float beta_loss_in, float sparsity_loss, const OGrads *grads) {
    const int offx = g->size[1] * g->size[2], offy = g->size[2];
    const int bd = g->basis_dim;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t q = 0; q < Q; ++q) {
        ORay ray;
        float sph[9], gout[3];
        ray_setup(&ray, g, opt, origins, dirs, xf, q, sph);
        const float *cc = color_cache + q * 3;
        for (int i = 0; i < 3; ++i) gout[i] = grad_is_rgb ? (cc[i] - grad_in[q * 3 + i]) * norm_factor : grad_in[q * 3 + i];
        float accum = fmaf(cc[0], gout[0], fmaf(cc[1], gout[1], cc[2] * gout[2]));
        float beta_loss = beta_loss_in;
        if (beta_loss > 0.f) {
            const float transmit_in = expf(log_transmit_in ? log_transmit_in[q] : 0.f);
            beta_loss *= (1 - transmit_in / (1 - transmit_in + 1e-3));
            accum += beta_loss;
        }
        if (ray.tmin > ray.tmax) continue;
        float t = ray.tmin, log_transmit = 0.f;
        while (t <= ray.tmax) {
            sample_position(&ray, g, t);
            const float skip = compute_skip_dist(&ray, g->links, offx, offy);
            if (skip >= opt->step_size) {
                t += ceilf(skip / opt->step_size) * opt->step_size;
                continue;
            }
            const float sigma = o_trilerp_cuvol_one(g->links, g->density, offx, offy, 1, ray.l, ray.pos, 0);
            if (opt->last_sample_opaque && t + opt->step_size > ray.tmax) ray.world_step = 1e9f;
            if (sigma > opt->sigma_thresh) {
                float c[3];
                channel_sums(g, &ray, sph, offx, offy, c);
                const float pcnt = ray.world_step * sigma;
                const float weight = expf(log_transmit) * (1.f - expf(-pcnt));
                log_transmit -= pcnt;
                float tc[3], in01[3];
                for (int ch = 0; ch < 3; ++ch) {
                    const float l = c[ch] + 0.5f;
                    tc[ch] = o_maxf(l, 0.f);
                    in01[ch] = (tc[ch] == l) ? 1.f : 0.f;
                    tc[ch] *= gout[ch];
                }
                float total_color = tc[0];   /* shuffle order :466-469: (c0 + c2) + c1 */
                total_color += tc[2];
                total_color += tc[1];
                for (int k = 0; k < g->sh_dim; ++k) {
                    const int ch = k / bd;
                    const float grad_common = weight * in01[ch] * gout[ch];
                    const float curr_grad_color = sph[k % bd] * grad_common;
                    o_trilerp_backward_cuvol_one(g->links, grads->grad_sh, offx, offy, (size_t)g->sh_dim, ray.l, ray.pos,
                                                 curr_grad_color, k);
                }
                accum -= weight * total_color;
                float curr_grad_sigma = ray.world_step * (total_color * expf(log_transmit) - accum);
                if (sparsity_loss > 0.f) curr_grad_sigma += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
                o_trilerp_backward_cuvol_one_density(g->links, grads->grad_density, grads->mask, offx, offy, ray.l, ray.pos,
                                                     curr_grad_sigma);
                if (expf(log_transmit) < opt->stop_thresh) break;
            }
            t += opt->step_size;
        }
    }
}
