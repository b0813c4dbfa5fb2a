// alphasurf_b200: the Plenoxels "cuvol" volume renderer (forward, backward, fused, image) for sm_100a.
//
// Replaces volume_render_cuvol / _backward / _fused / _image of
// /root/reference/svox2/csrc/render_lerp_kernel_cuvol.cu:1120-1354 (kernels :762-919, marchers trace_ray_cuvol :30-125 and
// trace_ray_cuvol_backward :371-535, skipping compute_skip_dist include/render_util.cuh:286-368).
//
// Semantics are the reference's sample by sample: t advances by step_size (or by ceil(skip/step)*step over an empty
// block announced by a negative link), every sample gathers 8 links + 8 densities, and where sigma > sigma_thresh the
// 8 x D SH coefficients.  The execution differs:
//   * one warp per ray (the SH gather is warp-wide), but the DENSITY phase runs a batch of 8 consecutive sample
//     positions at once on lanes 0..7 -- their t are produced by the same repeated `t += step` additions, so they are
//     the reference's positions exactly; samples behind the first skipping one are discarded and the batch restarts
//     at the skip target.  That puts 8 x (1 + 8 + 8) independent gathers in flight per warp instead of one dependent
//     chain per sample.
//   * per-ray scalars live in registers (no shared SingleRaySpec), channel sums use the reference's
//     HeadSegmentedSum add order, gradients leave as coalesced red.global.add (27 consecutive floats per corner row).
#include "common.cuh"
#include "surf_math.cuh"

namespace asurf {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CV_THREADS = 32;   // one warp = one ray per CTA: rays differ in length, and a 5000-ray batch is ~1.2 waves of the
                                 // machine -- single-warp CTAs let the block scheduler refill every slot the moment a ray ends
constexpr int CV_WARPS = CV_THREADS / 32;
constexpr int CV_BATCH = 8;

struct CvGrid {
    const int32_t *links;
    const float *density, *sh;
    int size[3];
    int basis_dim, sh_dim;
    float offset[3], scaling[3];
};

struct CvCam {          // include/data_spec.hpp:124-137 (CameraSpec), c2w row-major 3x4
    float c2w[12];
    float fx, fy, cx, cy;
    int width, height;
};

struct CvRay {
    float o[3], d[3];
    float tmin, tmax, world_step;
};

// ray_find_bounds (include/render_util.cuh:651-701) + transform_coord (cuda_util.cuh:63-69); AABB clip only
__device__ __forceinline__ void cv_ray_bounds(const CvGrid &g, const asurf_opt_t &opt, CvRay &r) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.o[i] = fmaf(r.o[i], g.scaling[i], g.offset[i]);
        r.d[i] *= g.scaling[i];
    }
    const float delta_scale = rnorm3df(r.d[0], r.d[1], r.d[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) r.d[i] *= delta_scale;
    r.world_step = delta_scale * opt.step_size;
    r.tmin = opt.near_clip / r.world_step * opt.step_size;
    r.tmax = 2e3f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float inv = (float)(1.0 / (double)r.d[i]);
        const float t1 = (-0.5f - r.o[i]) * inv, t2 = ((float)g.size[i] - 0.5f - r.o[i]) * inv;
        if (r.d[i] != 0.f) {
            r.tmin = fmaxf(r.tmin, fminf(t1, t2));
            r.tmax = fminf(r.tmax, fmaxf(t1, t2));
        }
    }
}

// compute_skip_dist (include/render_util.cuh:286-368) for the voxel l, interpolation offsets pos
__device__ __forceinline__ float cv_skip_dist(const CvRay &r, const int *l, const float *pos, int32_t link_val) {
    if (link_val >= -1) return 0.f;
    const uint32_t dist = (uint32_t)(-link_val);
    const uint32_t shift = dist - 1;
    const uint32_t side = (uint32_t)((float)(1 << shift) - 1.f);
    float tmin = 0.f, tmax = 1e9f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int ul = ((l[i] >> shift) << shift);
        ul -= l[i];
        const float invdir = (float)(1.0 / (double)r.d[i]);
        const float t1 = ((float)ul - pos[i] + 0.f) * invdir;
        const float t2 = ((float)(ul + side) - pos[i] + 0.f) * invdir;
        if (r.d[i] != 0.f) {
            tmin = fmaxf(tmin, fminf(t1, t2));
            if (fmaxf(t1, t2) < tmax) tmax = fmaxf(t1, t2);
        }
    }
    if (tmin > 0.f) return 0.f;
    return tmax;
}

// cam2world_ray (include/render_util.cuh:599-617), no NDC
__device__ __forceinline__ void cv_cam_ray(const CvCam &cam, int ix, int iy, float *o, float *d) {
    float x = ((float)ix + 0.5f - cam.cx) / cam.fx;
    float y = ((float)iy + 0.5f - cam.cy) / cam.fy;
    float z = sqrtf((float)((double)(x * x + y * y) + 1.0));
    x /= z; y /= z; z = 1.0f / z;
    d[0] = cam.c2w[0] * x + cam.c2w[1] * y + cam.c2w[2] * z;
    d[1] = cam.c2w[4] * x + cam.c2w[5] * y + cam.c2w[6] * z;
    d[2] = cam.c2w[8] * x + cam.c2w[9] * y + cam.c2w[10] * z;
    o[0] = cam.c2w[3]; o[1] = cam.c2w[7]; o[2] = cam.c2w[11];
}

__device__ __forceinline__ float cv_segment_sum(float v, int pos_in_seg, int bd) {   // HeadSegmentedSum order
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        const float o = __shfl_down_sync(FULL, v, off);
        if (pos_in_seg + off < bd) v += o;
    }
    return v;
}

struct CvSample {      // lane-private result of the density phase (lanes 0..CV_BATCH-1)
    float t, sigma, skip;
    float pos[3];
    int lk[8];
    bool in_range;
};

// One sample position on the calling lane: voxel, skip test, 8 links, trilinear sigma (trace_ray_cuvol :57-84).
__device__ __forceinline__ void cv_density_phase(const CvGrid &g, const CvRay &r, CvSample &s) {
    int l[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float p = fmaf(s.t, r.d[j], r.o[j]);
        p = fminf(fmaxf(p, 0.f), (float)g.size[j] - 1.f);
        l[j] = min((int)p, g.size[j] - 2);
        s.pos[j] = p - (float)l[j];
    }
    const int offy = g.size[2];
    const int64_t offx = (int64_t)g.size[1] * g.size[2];
    const int32_t *lp = g.links + (offx * l[0] + (int64_t)offy * l[1] + l[2]);
    s.lk[0] = __ldg(lp);
    s.skip = cv_skip_dist(r, l, s.pos, s.lk[0]);
    s.lk[1] = __ldg(lp + 1);
    s.lk[2] = __ldg(lp + offy);
    s.lk[3] = __ldg(lp + offy + 1);
    s.lk[4] = __ldg(lp + offx);
    s.lk[5] = __ldg(lp + offx + 1);
    s.lk[6] = __ldg(lp + offx + offy);
    s.lk[7] = __ldg(lp + offx + offy + 1);
    float dn[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) dn[c] = (s.lk[c] >= 0) ? __ldg(g.density + s.lk[c]) : 0.f;
    s.sigma = trilerp8(dn, s.pos);
}

// BWD = false: colour;  BWD = true: gradients.  IMAGE: rays from the camera (forward only).
// 28 resident warps per SM: the march is a chain of dependent gathers per ray, so warps in flight are what hides its latency
// (measured: 72 -> 80 registers for the backward cost 13 % of the fused call)
template <bool BWD, bool IMAGE>
__global__ void __launch_bounds__(CV_THREADS, 28)
cuvol_kernel(const CvGrid g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
             const CvCam cam, const int64_t Q, float *__restrict__ rgb_out, float *__restrict__ log_transmit_out,
             const float *__restrict__ grad_in, const float *__restrict__ color_cache, int grad_is_rgb, float norm_factor,
             const float *__restrict__ log_transmit_in, float beta_loss, float sparsity_loss, const asurf_grads_t grads,
             const int has_bg, float *__restrict__ accum_out) {
    // has_bg: an MSI background follows (msi.cu): no background_brightness term here (:113-115); the backward leaves what
    // remains of `accum` (minus the beta term, :524-529) in accum_out and, when log_transmit_out is given, its own final
    // log-transmittance
    __shared__ float s_sph[CV_WARPS][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ray_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ray_id >= Q) return;
    const int D = g.sh_dim, bd = g.basis_dim;
    CvRay r;
    float wd[3];
    if (IMAGE) {
        cv_cam_ray(cam, (int)(ray_id % cam.width), (int)(ray_id / cam.width), r.o, r.d);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            r.o[i] = origins[ray_id * 3 + i];
            r.d[i] = dirs[ray_id * 3 + i];
        }
    }
    wd[0] = r.d[0]; wd[1] = r.d[1]; wd[2] = r.d[2];
    if (lane == 0) eval_sh(bd, wd[0], wd[1], wd[2], s_sph[warp]);   // world-space direction
    __syncwarp();
    cv_ray_bounds(g, opt, r);
    const int kb = (lane < D) ? (lane % bd) : 0;
    const float sph = (lane < D) ? s_sph[warp][kb] : 0.f;

    float g0 = 0.f, g1 = 0.f, g2 = 0.f, accum = 0.f;
    if (BWD) {
        if (grad_is_rgb) {   // fused: grad_in is rgb_gt (:873-880)
            g0 = (color_cache[ray_id * 3 + 0] - grad_in[ray_id * 3 + 0]) * norm_factor;
            g1 = (color_cache[ray_id * 3 + 1] - grad_in[ray_id * 3 + 1]) * norm_factor;
            g2 = (color_cache[ray_id * 3 + 2] - grad_in[ray_id * 3 + 2]) * norm_factor;
        } else {
            g0 = grad_in[ray_id * 3 + 0]; g1 = grad_in[ray_id * 3 + 1]; g2 = grad_in[ray_id * 3 + 2];
        }
        accum = fmaf(color_cache[ray_id * 3 + 0], g0, fmaf(color_cache[ray_id * 3 + 1], g1, color_cache[ray_id * 3 + 2] * g2));
        if (beta_loss > 0.f) {
            const float transmit_in = __expf(log_transmit_in ? log_transmit_in[ray_id] : 0.f);
            beta_loss *= (1 - transmit_in / (1 - transmit_in + 1e-3));
            accum += beta_loss;
        }
    }
    float out0 = 0.f, out1 = 0.f, out2 = 0.f;
    float log_transmit = 0.f;
    if (r.tmin > r.tmax) {
        if (lane == 0) {
            if (!BWD) {
                const float bg = has_bg ? 0.f : opt.background_brightness;
                rgb_out[ray_id * 3 + 0] = bg;
                rgb_out[ray_id * 3 + 1] = bg;
                rgb_out[ray_id * 3 + 2] = bg;
            } else if (accum_out) {
                accum_out[ray_id] = accum;     // (:405-409: before the beta term is cancelled)
            }
            if (log_transmit_out) log_transmit_out[ray_id] = 0.f;
        }
        return;
    }
    float t = r.tmin;
    bool stop = false;
    while (!stop && (t <= r.tmax)) {
        // ---- density phase: lanes 0..7 take the next 8 sample positions t, t+step, (t+step)+step, ... ----
        CvSample s;
        s.t = t;
        for (int i = 0; i < lane && i < CV_BATCH - 1; ++i) s.t += opt.step_size;
        s.in_range = (lane < CV_BATCH) && (s.t <= r.tmax);
        s.sigma = 0.f;
        s.skip = 0.f;
        if (s.in_range) cv_density_phase(g, r, s);
        const unsigned m_in = __ballot_sync(FULL, s.in_range);
        const unsigned m_skip = __ballot_sync(FULL, s.in_range && (s.skip >= opt.step_size));
        // samples up to (excluding) the first skipping one are the reference's samples
        const int n_in = __popc(m_in);
        const int first_skip = m_skip ? (__ffs(m_skip) - 1) : n_in;
        for (int i = 0; i < first_skip; ++i) {
            const float sigma = __shfl_sync(FULL, s.sigma, i);
            const float ts = __shfl_sync(FULL, s.t, i);
            float world_step = r.world_step;
            if (opt.last_sample_opaque && ts + opt.step_size > r.tmax) {
                world_step = 1e9f;
                r.world_step = 1e9f;   // sticky, as in the reference (:85-87)
            }
            if (sigma > opt.sigma_thresh) {
                int lk[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) lk[c] = __shfl_sync(FULL, s.lk[c], i);
                float pos[3];
                pos[0] = __shfl_sync(FULL, s.pos[0], i);
                pos[1] = __shfl_sync(FULL, s.pos[1], i);
                pos[2] = __shfl_sync(FULL, s.pos[2], i);
                float v[8];
                float lane_color = 0.f;
                if (lane < D) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = (lk[c] >= 0) ? __ldg(g.sh + (int64_t)lk[c] * D + lane) : 0.f;
                    lane_color = trilerp8(v, pos) * sph;
                }
                const float pcnt = world_step * sigma;
                const float weight = __expf(log_transmit) * (1.f - __expf(-pcnt));
                log_transmit -= pcnt;
                const float seg = cv_segment_sum(lane_color, (lane < D) ? kb : 32, bd);
                const float c0 = __shfl_sync(FULL, seg, 0);
                const float c1 = __shfl_sync(FULL, seg, bd);
                const float c2 = __shfl_sync(FULL, seg, 2 * bd);
                if (!BWD) {
                    out0 += weight * fmaxf(c0 + 0.5f, 0.f);
                    out1 += weight * fmaxf(c1 + 0.5f, 0.f);
                    out2 += weight * fmaxf(c2 + 0.5f, 0.f);
                    if (__expf(log_transmit) < opt.stop_thresh) {
                        log_transmit = -1e3f;
                        stop = true;
                        break;
                    }
                } else {
                    const float l0 = c0 + 0.5f, l1 = c1 + 0.5f, l2 = c2 + 0.5f;
                    const float t0 = fmaxf(l0, 0.f), t1 = fmaxf(l1, 0.f), t2 = fmaxf(l2, 0.f);
                    float total_color = t0 * g0;   // (c0 + c2) + c1, the reference's shuffle order (:466-469)
                    total_color += t2 * g2;
                    total_color += t1 * g1;
                    if (lane < D) {
                        const int ch = lane / bd;
                        const float in01 = (ch == 0) ? ((t0 == l0) ? 1.f : 0.f)
                                                     : ((ch == 1) ? ((t1 == l1) ? 1.f : 0.f) : ((t2 == l2) ? 1.f : 0.f));
                        const float gch = (ch == 0) ? g0 : ((ch == 1) ? g1 : g2);
                        const float curr_grad_color = sph * (weight * in01 * gch);
                        float w[8];
                        corner_weights(pos, curr_grad_color, w);
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (lk[c] >= 0) atomicAdd(grads.grad_sh + (int64_t)lk[c] * D + lane, w[c]);
                    }
                    accum -= weight * total_color;
                    float curr_grad_sigma = world_step * (total_color * __expf(log_transmit) - accum);
                    if (sparsity_loss > 0.f) curr_grad_sigma += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
                    {   // density scatter: corner c on lane c (trilerp_backward_cuvol_one_density, render_util.cuh:124-154)
                        float w[8];
                        corner_weights(pos, curr_grad_sigma, w);
                        if (lane < 8) {
                            float wc = w[0];
                            int lc = lk[0];
#pragma unroll
                            for (int c = 1; c < 8; ++c)
                                if (lane == c) { wc = w[c]; lc = lk[c]; }
                            if (lc >= 0) {
                                atomicAdd(grads.grad_density + lc, wc);
                                if (grads.mask) grads.mask[lc] = 1;
                            }
                        }
                    }
                    if (__expf(log_transmit) < opt.stop_thresh) {
                        stop = true;
                        break;
                    }
                }
            }
        }
        if (stop) break;
        // ---- next batch start: behind the last sample, or at the skip target ----
        if (m_skip && first_skip < n_in) {
            const float ts = __shfl_sync(FULL, s.t, first_skip);
            const float sk = __shfl_sync(FULL, s.skip, first_skip);
            t = ts + ceilf(sk / opt.step_size) * opt.step_size;
        } else if (n_in > 0) {
            t = __shfl_sync(FULL, s.t, n_in - 1) + opt.step_size;
        } else {
            break;
        }
    }
    if (lane == 0) {
        if (!BWD) {
            const float bg = has_bg ? 0.f : __expf(log_transmit) * opt.background_brightness;
            rgb_out[ray_id * 3 + 0] = out0 + bg;
            rgb_out[ray_id * 3 + 1] = out1 + bg;
            rgb_out[ray_id * 3 + 2] = out2 + bg;
        } else if (accum_out) {
            accum_out[ray_id] = accum - beta_loss;
        }
        if (log_transmit_out) log_transmit_out[ray_id] = log_transmit;
    }
}

Workspace g_ws_lt;

// ---- depth renders (evaluation): trace_ray_expected_term :127-188, trace_ray_mode_term :190-257, trace_ray_med_term
//      :259-319, trace_ray_sigma_thresh :322-369; kernels :921-1005.  Thread per ray like the reference: no SH gather, so
//      a warp per ray would idle; the sample positions are the reference's (same repeated t += step additions).
template <int MODE>
__global__ void __launch_bounds__(128)
cuvol_scalar_kernel(const CvGrid g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                    const int64_t Q, const float param, const int max_sample, float *__restrict__ out,
                    float *__restrict__ out2) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray_id >= Q) return;
    CvRay r;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.o[i] = origins[ray_id * 3 + i];
        r.d[i] = dirs[ray_id * 3 + i];
    }
    cv_ray_bounds(g, opt, r);
    float res = 0.f;
    if (!(r.tmin > r.tmax)) {
        float t = r.tmin, outv = 0.f, weight_acc = 0.f, max_weight = -1.f, log_transmit = 0.f;
        int sample_i = 0;
        while (t <= r.tmax) {
            CvSample s;
            s.t = t;
            cv_density_phase(g, r, s);
            if (s.skip >= opt.step_size) {
                t += ceilf(s.skip / opt.step_size) * opt.step_size;
                continue;
            }
            if (MODE == ASURF_CUVOL_SIGMA_THRESH) {
                if (s.sigma > param) {
                    res = (t / opt.step_size) * r.world_step;
                    break;
                }
            } else if (s.sigma > opt.sigma_thresh) {
                const float pcnt = r.world_step * s.sigma;
                const float weight = __expf(log_transmit) * (1.f - __expf(-pcnt));
                log_transmit -= pcnt;
                if (MODE == ASURF_CUVOL_EXPECTED_TERM) {
                    outv += weight * (t / opt.step_size) * r.world_step;
                    weight_acc += weight;
                } else if (MODE == ASURF_CUVOL_MODE_TERM) {
                    weight_acc += weight;
                    if (weight > max_weight) {
                        max_weight = weight;
                        outv = (t / opt.step_size) * r.world_step;
                    }
                } else if (sample_i < max_sample) {
                    out[ray_id * max_sample + sample_i] = (t / opt.step_size) * r.world_step;
                    out2[ray_id * max_sample + sample_i] = s.sigma;
                    sample_i += 1;
                }
                if (__expf(log_transmit) < opt.stop_thresh) break;
            }
            t += opt.step_size;
        }
        if (MODE == ASURF_CUVOL_EXPECTED_TERM || MODE == ASURF_CUVOL_MODE_TERM) res = (weight_acc > param) ? outv : 0.f;
    }
    if (MODE != ASURF_CUVOL_MED_TERM) out[ray_id] = res;
}

// march counters of SURVEY.md 8(d) for the cuvol flavour (bench.py's algorithmic bytes): thread per ray, the sample
// positions and the early stop of trace_ray_cuvol (:30-125)
__global__ void __launch_bounds__(128)
cuvol_count_kernel(const CvGrid g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                   const int64_t Q, unsigned long long *__restrict__ stats) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n_pos = 0, n_skip = 0, n_samp = 0, n_contrib = 0;
    if (ray_id < Q) {
        CvRay r;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            r.o[i] = origins[ray_id * 3 + i];
            r.d[i] = dirs[ray_id * 3 + i];
        }
        cv_ray_bounds(g, opt, r);
        if (!(r.tmin > r.tmax)) {
            float t = r.tmin, log_transmit = 0.f;
            while (t <= r.tmax) {
                CvSample s;
                s.t = t;
                cv_density_phase(g, r, s);
                ++n_pos;
                if (s.skip >= opt.step_size) {
                    ++n_skip;
                    t += ceilf(s.skip / opt.step_size) * opt.step_size;
                    continue;
                }
                ++n_samp;
                float world_step = r.world_step;
                if (opt.last_sample_opaque && t + opt.step_size > r.tmax) world_step = r.world_step = 1e9f;
                if (s.sigma > opt.sigma_thresh) {
                    ++n_contrib;
                    log_transmit -= world_step * s.sigma;
                    if (__expf(log_transmit) < opt.stop_thresh) break;
                }
                t += opt.step_size;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        n_pos += __shfl_xor_sync(FULL, n_pos, off);
        n_skip += __shfl_xor_sync(FULL, n_skip, off);
        n_samp += __shfl_xor_sync(FULL, n_samp, off);
        n_contrib += __shfl_xor_sync(FULL, n_contrib, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + 0, n_pos);       // asurf_stats_t.n_steps: sample positions (one skip-link read each)
        atomicAdd(stats + 1, n_skip);      // n_skips: positions that jump over an empty block
        atomicAdd(stats + 2, n_samp);      // n_linked: samples whose 8 links + 8 densities are gathered
        atomicAdd(stats + 3, n_samp);      // n_active
        atomicAdd(stats + 4, n_contrib);   // n_samples: samples with sigma > sigma_thresh (SH gather, gradients)
    }
}

int cv_make_grid(const asurf_grid_t *grid, CvGrid &g, const char *who) {
    ASURF_REQUIRE(grid, ASURF_E_INVALID, "%s: null grid", who);
    ASURF_REQUIRE(grid->links && grid->density && grid->sh, ASURF_E_INVALID, "%s: null grid tensor", who);
    ASURF_REQUIRE(grid->size[0] >= 2 && grid->size[1] >= 2 && grid->size[2] >= 2, ASURF_E_INVALID, "%s: grid smaller than 2^3", who);
    ASURF_REQUIRE(grid->basis_dim == 1 || grid->basis_dim == 4 || grid->basis_dim == 9, ASURF_E_UNSUPPORTED,
                  "%s: basis_dim %d not supported (SH with 1, 4 or 9 functions)", who, grid->basis_dim);
    ASURF_REQUIRE(grid->sh_dim == 3 * grid->basis_dim, ASURF_E_INVALID, "%s: sh_dim must be 3*basis_dim", who);
    g.links = grid->links;
    g.density = grid->density;
    g.sh = grid->sh;
    for (int i = 0; i < 3; ++i) {
        g.size[i] = grid->size[i];
        g.offset[i] = grid->offset[i];
        g.scaling[i] = grid->scaling[i];
    }
    g.basis_dim = grid->basis_dim;
    g.sh_dim = grid->sh_dim;
    return 0;
}

inline int cv_blocks(int64_t Q) { return (int)((Q * 32 + CV_THREADS - 1) / CV_THREADS); }

}  // namespace
void cuvol_release() { g_ws_lt.release(); }
}  // namespace asurf

using namespace asurf;

extern "C" int asurf_cuvol_forward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, float *rgb_out,
                                   float *log_transmit_out, void *stream) {
    ASURF_REQUIRE(rays && opt && rgb_out, ASURF_E_INVALID, "cuvol_forward: null argument");
    ASURF_REQUIRE(!opt->use_spheric_clip, ASURF_E_UNSUPPORTED, "cuvol_forward: spheric clip is not on the hot path");
    const int64_t Q = rays->n_rays;
    if (Q <= 0) return 0;
    CvGrid g;
    int rc = cv_make_grid(grid, g, "cuvol_forward");
    if (rc) return rc;
    asurf_grads_t nog = {};
    CvCam cam = {};
    const int has_bg = grid_has_background(grid) ? 1 : 0;
    float *lt = log_transmit_out;
    if (has_bg) {
        float *acc = nullptr;
        rc = bg_state_reserve(Q, &lt, &acc);
        if (rc) return rc;
    }
    cuvol_kernel<false, false><<<cv_blocks(Q), CV_THREADS, 0, (cudaStream_t)stream>>>(
        g, *opt, rays->origins, rays->dirs, cam, Q, rgb_out, lt, nullptr, nullptr, 0, 0.f, nullptr, 0.f, 0.f, nog, has_bg, nullptr);
    note_launches(1);
    rc = check_cuda(cudaGetLastError(), "cuvol_forward launch");
    if (rc) return rc;
    if (has_bg) {
        rc = asurf_msi_forward(grid, rays, opt, lt, rgb_out, stream);
        if (rc) return rc;
        if (log_transmit_out) ASURF_CUDA(cudaMemcpyAsync(log_transmit_out, lt, (size_t)Q * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return 0;
}

extern "C" int asurf_cuvol_stats(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                 asurf_stats_t *stats_dev, void *stream) {
    ASURF_REQUIRE(rays && opt && stats_dev, ASURF_E_INVALID, "cuvol_stats: null argument");
    const int64_t Q = rays->n_rays;
    if (Q <= 0) return 0;
    CvGrid g;
    int rc = cv_make_grid(grid, g, "cuvol_stats");
    if (rc) return rc;
    cuvol_count_kernel<<<(int)((Q + 127) / 128), 128, 0, (cudaStream_t)stream>>>(g, *opt, rays->origins, rays->dirs, Q,
                                                                                 (unsigned long long *)stats_dev);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "cuvol_stats launch");
}

extern "C" int asurf_cuvol_scalar(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, int32_t mode,
                                  float param, int32_t max_sample, float *out, float *out2, void *stream) {
    ASURF_REQUIRE(rays && opt, ASURF_E_INVALID, "cuvol_scalar: null argument");
    ASURF_REQUIRE(mode >= ASURF_CUVOL_EXPECTED_TERM && mode <= ASURF_CUVOL_SIGMA_THRESH, ASURF_E_INVALID,
                  "cuvol_scalar: unknown mode %d", mode);
    ASURF_REQUIRE(!opt->use_spheric_clip, ASURF_E_UNSUPPORTED, "cuvol_scalar: spheric clip is not on the hot path");
    const int64_t Q = rays->n_rays;
    if (Q <= 0) return 0;
    ASURF_REQUIRE(rays->origins && rays->dirs && out, ASURF_E_INVALID, "cuvol_scalar: null tensor");
    CvGrid g;
    int rc = cv_make_grid(grid, g, "cuvol_scalar");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (int)((Q + 127) / 128);
    if (mode == ASURF_CUVOL_MED_TERM) {
        ASURF_REQUIRE(out2 && max_sample >= 0, ASURF_E_INVALID, "cuvol_scalar: med_term needs the sigma output and max_sample >= 0");
        if (max_sample == 0) return 0;
        ASURF_CUDA(cudaMemsetAsync(out, 0, (size_t)Q * max_sample * sizeof(float), st));
        ASURF_CUDA(cudaMemsetAsync(out2, 0, (size_t)Q * max_sample * sizeof(float), st));
        cuvol_scalar_kernel<ASURF_CUVOL_MED_TERM><<<blocks, 128, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, param,
                                                                          max_sample, out, out2);
    } else if (mode == ASURF_CUVOL_EXPECTED_TERM) {
        cuvol_scalar_kernel<ASURF_CUVOL_EXPECTED_TERM><<<blocks, 128, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, param, 0,
                                                                               out, nullptr);
    } else if (mode == ASURF_CUVOL_MODE_TERM) {
        cuvol_scalar_kernel<ASURF_CUVOL_MODE_TERM><<<blocks, 128, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, param, 0, out,
                                                                           nullptr);
    } else {
        cuvol_scalar_kernel<ASURF_CUVOL_SIGMA_THRESH><<<blocks, 128, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, param, 0,
                                                                              out, nullptr);
    }
    note_launches(1);
    return check_cuda(cudaGetLastError(), "cuvol_scalar launch");
}

extern "C" int asurf_cuvol_image(const asurf_grid_t *grid, const float *c2w_host, float fx, float fy, float cx, float cy,
                                 int32_t width, int32_t height, const asurf_opt_t *opt, float *rgb_out, void *stream) {
    ASURF_REQUIRE(c2w_host && opt && rgb_out, ASURF_E_INVALID, "cuvol_image: null argument");
    ASURF_REQUIRE(!opt->use_spheric_clip, ASURF_E_UNSUPPORTED, "cuvol_image: spheric clip is not on the hot path");
    const int64_t Q = (int64_t)width * height;
    if (Q <= 0) return 0;
    CvGrid g;
    int rc = cv_make_grid(grid, g, "cuvol_image");
    if (rc) return rc;
    asurf_grads_t nog = {};
    CvCam cam;
    for (int i = 0; i < 12; ++i) cam.c2w[i] = c2w_host[i];
    cam.fx = fx; cam.fy = fy; cam.cx = cx; cam.cy = cy;
    cam.width = width; cam.height = height;
    const int has_bg = grid_has_background(grid) ? 1 : 0;
    float *lt = nullptr;
    if (has_bg) {
        float *acc = nullptr;
        rc = bg_state_reserve(Q, &lt, &acc);
        if (rc) return rc;
    }
    cuvol_kernel<false, true><<<cv_blocks(Q), CV_THREADS, 0, (cudaStream_t)stream>>>(
        g, *opt, nullptr, nullptr, cam, Q, rgb_out, lt, nullptr, nullptr, 0, 0.f, nullptr, 0.f, 0.f, nog, has_bg, nullptr);
    note_launches(1);
    rc = check_cuda(cudaGetLastError(), "cuvol_image launch");
    if (rc) return rc;
    if (has_bg) return asurf_msi_forward_image(grid, c2w_host, fx, fy, cx, cy, width, height, opt, lt, rgb_out, stream);
    return 0;
}

extern "C" int asurf_cuvol_backward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                    const float *grad_out, const float *color_cache, const asurf_grads_t *grads,
                                    void *stream) {
    ASURF_REQUIRE(rays && opt && grad_out && color_cache && grads, ASURF_E_INVALID, "cuvol_backward: null argument");
    ASURF_REQUIRE(grads->grad_density && grads->grad_sh, ASURF_E_INVALID, "cuvol_backward: null gradient buffer");
    ASURF_REQUIRE(!opt->use_spheric_clip, ASURF_E_UNSUPPORTED, "cuvol_backward: spheric clip is not on the hot path");
    const int64_t Q = rays->n_rays;
    if (Q <= 0) return 0;
    CvGrid g;
    int rc = cv_make_grid(grid, g, "cuvol_backward");
    if (rc) return rc;
    CvCam cam = {};
    const int has_bg = grid_has_background(grid) ? 1 : 0;
    float *lt = nullptr, *acc = nullptr;
    if (has_bg) {   // the stand-alone backward hands its own log-transmittance and accum to the background (:1226-1262)
        ASURF_REQUIRE(grads->grad_background, ASURF_E_INVALID, "cuvol_backward: the grid has a background but no gradient buffer for it");
        rc = bg_state_reserve(Q, &lt, &acc);
        if (rc) return rc;
    }
    cuvol_kernel<true, false><<<cv_blocks(Q), CV_THREADS, 0, (cudaStream_t)stream>>>(
        g, *opt, rays->origins, rays->dirs, cam, Q, nullptr, lt, grad_out, color_cache, 0, 0.f, nullptr, 0.f, 0.f, *grads, has_bg,
        acc);
    note_launches(1);
    rc = check_cuda(cudaGetLastError(), "cuvol_backward launch");
    if (rc) return rc;
    if (has_bg) return asurf_msi_backward(grid, rays, opt, grad_out, color_cache, 0, 0, lt, acc, 0.f, 0.f, grads, stream);
    return 0;
}

extern "C" int asurf_cuvol_fused(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                 const float *rgb_gt, float beta_loss, float sparsity_loss, int64_t norm_rays, float *rgb_out,
                                 const asurf_grads_t *grads, void *stream) {
    ASURF_REQUIRE(rays && opt && rgb_gt && rgb_out && grads, ASURF_E_INVALID, "cuvol_fused: null argument");
    ASURF_REQUIRE(grads->grad_density && grads->grad_sh, ASURF_E_INVALID, "cuvol_fused: null gradient buffer");
    ASURF_REQUIRE(!opt->use_spheric_clip, ASURF_E_UNSUPPORTED, "cuvol_fused: spheric clip is not on the hot path");
    const int64_t Q = rays->n_rays;
    if (Q <= 0) return 0;
    CvGrid g;
    int rc = cv_make_grid(grid, g, "cuvol_fused");
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float *lt = nullptr, *acc = nullptr;
    const int has_bg = grid_has_background(grid) ? 1 : 0;
    if (has_bg) {            // forward log-transmittance + the backward's leftover accum for the background pass (:1289-1350)
        ASURF_REQUIRE(grads->grad_background, ASURF_E_INVALID, "cuvol_fused: the grid has a background but no gradient buffer for it");
        rc = bg_state_reserve(Q, &lt, &acc);
        if (rc) return rc;
    } else if (beta_loss > 0.f) {   // the backward needs the forward's final log-transmittance (:1291-1296)
        rc = g_ws_lt.reserve((size_t)Q * sizeof(float));
        if (rc) return rc;
        lt = (float *)g_ws_lt.ptr;
    }
    const int64_t qn = norm_rays > 0 ? norm_rays : Q;
    asurf_grads_t nog = {};
    CvCam cam = {};
    cuvol_kernel<false, false><<<cv_blocks(Q), CV_THREADS, 0, st>>>(g, *opt, rays->origins, rays->dirs, cam, Q, rgb_out, lt,
                                                                    nullptr, nullptr, 0, 0.f, nullptr, 0.f, 0.f, nog, has_bg, nullptr);
    note_launches(1);
    if (has_bg) {
        rc = check_cuda(cudaGetLastError(), "cuvol_fused forward launch");
        if (rc) return rc;
        rc = asurf_msi_forward(grid, rays, opt, lt, rgb_out, stream);
        if (rc) return rc;
    }
    cuvol_kernel<true, false><<<cv_blocks(Q), CV_THREADS, 0, st>>>(g, *opt, rays->origins, rays->dirs, cam, Q, nullptr, nullptr,
                                                                   rgb_gt, rgb_out, 1, 2.f / (float)(3 * (int)qn),
                                                                   beta_loss > 0.f ? lt : nullptr, beta_loss / (float)qn,
                                                                   sparsity_loss, *grads, has_bg, acc);
    note_launches(1);
    rc = check_cuda(cudaGetLastError(), "cuvol_fused launch");
    if (rc) return rc;
    if (has_bg) return asurf_msi_backward(grid, rays, opt, rgb_gt, rgb_out, 1, qn, lt, acc, 0.f, sparsity_loss, grads, stream);
    return 0;
}
