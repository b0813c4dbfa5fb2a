// alphasurf_b200: scalar renders of the surf_trav backend (depth / alpha / normal images for evaluation).  Included by
// surf_trav.cu inside namespace asurf::{anon}.
//
// Reference: trace_ray_expected_term (render_lerp_kernel_surf_trav.cu:564-794), trace_ray_mode_term_surf_trav (:796-1001),
// trace_ray_sigma_thresh_surf_trav (:1003-1168), trace_ray_alpha_surf_trav (:1170-1337), trace_ray_normal (:1339-1534),
// trace_ray_extract_pt (:1536-1708); kernels :3458-3594, one thread per ray.  They share the DDA of the colour renderer but none of its gates: a voxel counts
// when its 8 links are stored and a level set lies within the range of its corner values, every root inside the voxel is
// a sample, alpha = surf_alpha_act(trilerp(density)) with no threshold, no outward test, no truncated re-weighting and no
// fake samples.
//
// Here: thread per ray as well (a render of this kind happens once per evaluation image, not per training step), but the
// march walks a work pyramid built for exactly that voxel predicate (asurf_work_build with the density gate disabled), so a
// ray jumps over empty 16^3 / 64^3 blocks with the exact landing rule of march_step and only stops at voxels the level set
// crosses; the reference visits every voxel along the ray.

// mode: ASURF_SCALAR_* of include/asurf.h
template <int MODE>
__global__ void __launch_bounds__(128)
scalar_render_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                     const int64_t Q, const float param, const int max_sample, float *__restrict__ out,
                     float *__restrict__ out2) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray_id >= Q) return;
    constexpr int NOUT = (MODE == ASURF_SCALAR_NORMAL) ? 3 : 1;
    float res[3] = {0.f, 0.f, 0.f};
    Lane L;
    L.ray_id = ray_id;
    L.state = ST_IDLE;
    L.ray_done = false;
    L.ox = origins[ray_id * 3 + 0]; L.oy = origins[ray_id * 3 + 1]; L.oz = origins[ray_id * 3 + 2];
    L.dx = dirs[ray_id * 3 + 0]; L.dy = dirs[ray_id * 3 + 1]; L.dz = dirs[ray_id * 3 + 2];
    float world_step;
    ray_bounds(g, opt, L, world_step);
    float logT = 0.f, outv = 0.f, max_weight = 0.f, weight_acc = 0.f;
    bool found = false;
    int sample_id = 0;   // EXTRACT_PTS: samples written so far
    Counters cnt;
    if (!(L.tmin > L.tmax)) {
        dda_init(g, L);
        while (!L.ray_done && !found) {
            march_step<false, false, false>(g, opt, L, cnt);
            if (L.state != ST_VOXEL) continue;
            L.state = ST_MARCH;
            // voxel (vx,vy,vz): near crossing, corner data
            const float tcx = PT_X(L, L.vx + (L.dx > 0.f ? 0 : 1));
            const float tcy = PT_Y(L, L.vy + (L.dy > 0.f ? 0 : 1));
            const float tcz = PT_Z(L, L.vz + (L.dz > 0.f ? 0 : 1));
            const float t_close = fmaxf(fmaxf(fmaxf(tcx, tcy), tcz), 0.f);
            const float nof[3] = {fmaf(t_close, L.dx, L.ox), fmaf(t_close, L.dy, L.oy), fmaf(t_close, L.dz, L.oz)};
            const double nno[3] = {(double)nof[0] - L.vx, (double)nof[1] - L.vy, (double)nof[2] - L.vz};
            wave_links(g, L);
            float sf[8], dn[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) sf[c] = __ldg(g.surface + L.lk[c]);
            float smin = sf[0], smax = sf[0];
#pragma unroll
            for (int c = 1; c < 8; ++c) {
                smin = fminf(smin, sf[c]);
                smax = fmaxf(smax, sf[c]);
            }
            bool dn_loaded = false;
            double fs[4];
            bool fs_ready = false;
            for (int i = 0; i < g.level_set_num && !found; ++i) {
                const float lv = __ldg(g.level_set + i);
                if ((lv < smin) || (lv > smax)) continue;
                if (!fs_ready) {
                    double s[8], dd[3] = {(double)L.dx, (double)L.dy, (double)L.dz};
#pragma unroll
                    for (int c = 0; c < 8; ++c) s[c] = (double)sf[c];
                    field_to_cubic(s, nno, dd, fs);
                    fs_ready = true;
                }
                double st[3] = {-1, -1, -1};
                solve_cubic(fs[0] - (double)lv, fs[1], fs[2], fs[3], st);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const double stj = st[j];
                    if (stj <= 0 || found) continue;
                    const float stf = (float)stj;
                    const float pos[3] = {fmaf(stf, L.dx, nof[0]) - (float)L.vx, fmaf(stf, L.dy, nof[1]) - (float)L.vy,
                                          fmaf(stf, L.dz, nof[2]) - (float)L.vz};
                    if ((pos[0] < 0) | (pos[0] > 1) | (pos[1] < 0) | (pos[1] > 1) | (pos[2] < 0) | (pos[2] > 1)) continue;
                    if (!dn_loaded) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) dn[c] = __ldg(g.density + L.lk[c]);
                        dn_loaded = true;
                    }
                    const float alpha = alpha_act(trilerp8(dn, pos), opt.alpha_activation_type);
                    // depth in world units, in the reference's double arithmetic (:785, :983, :1160)
                    if (MODE == ASURF_SCALAR_EXPECTED_TERM || MODE == ASURF_SCALAR_MODE_TERM) {
                        const float pcnt = -1 * __logf(1 - alpha);
                        const float weight = __expf(logT) * (1.f - __expf(-pcnt));
                        logT -= pcnt;
                        if (MODE == ASURF_SCALAR_EXPECTED_TERM) {
                            outv = (float)((double)outv +
                                           (double)weight * (stj + (double)t_close) / (double)opt.step_size * (double)world_step);
                        } else {
                            weight_acc += weight;
                            if (weight > max_weight) {
                                max_weight = weight;
                                outv = (float)(((stj + (double)t_close) / (double)opt.step_size) * (double)world_step);
                            }
                        }
                    } else if (MODE == ASURF_SCALAR_THRESH_DEPTH || MODE == ASURF_SCALAR_THRESH_ALPHA) {
                        if (alpha > param) {
                            res[0] = (MODE == ASURF_SCALAR_THRESH_DEPTH)
                                         ? (float)(((stj + (double)t_close) / (double)opt.step_size) * (double)world_step)
                                         : alpha;
                            found = true;
                        }
                    } else if (MODE == ASURF_SCALAR_EXTRACT_PTS) {
                        if (alpha > param) {   // every sample above the threshold, up to max_sample per ray (:1689-1697)
                            out[ray_id * max_sample + sample_id] =
                                (float)(((stj + (double)t_close) / (double)opt.step_size) * (double)world_step);
                            out2[ray_id * max_sample + sample_id] = alpha;
                            sample_id += 1;
                            found = sample_id >= max_sample;
                        }
                    } else {
                        if (alpha > 0) {
                            field_grad8(sf, pos, res);
                            found = true;
                        }
                    }
                }
            }
            if (MODE == ASURF_SCALAR_EXPECTED_TERM || MODE == ASURF_SCALAR_MODE_TERM) {
                if (__expf(logT) < opt.stop_thresh) break;   // :788-791
            }
        }
        if (MODE == ASURF_SCALAR_EXPECTED_TERM) res[0] = outv;
        if (MODE == ASURF_SCALAR_MODE_TERM) res[0] = (weight_acc > param) ? outv : 0.f;
    }
    if (MODE == ASURF_SCALAR_EXTRACT_PTS) return;   // (Q, max_sample) outputs are zero-filled by the host side
#pragma unroll
    for (int c = 0; c < NOUT; ++c) out[ray_id * NOUT + c] = res[c];
}
