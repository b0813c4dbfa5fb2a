// alphasurf_b200: wavefront kernels of the surf_trav renderer.  Included by surf_trav.cu inside namespace asurf::{anon}.
//
// The persistent shading kernel (surf_trav_kernel) interleaves four kinds of work per warp; on the training workload
// (a thin level-set sheet: ~4 listed voxels and ~1 sample per ray that hits it) its lanes mostly wait for each other.
// Rays whose whole march fits the pre-march list ("short" rays: no continuation, one level set) take this path instead:
// every stage is its own kernel over a compact queue, so each runs with full warps and no state machine.
//
//   wave_eval       thread per (ray, listed voxel): 8 links / surface / density, fp64 cubic + Vieta solve, hit filtering
//                   (trace_ray_surf_trav :212-541 without the compositing) -> up to 3 entries per voxel + a hit queue
//   wave_color      warp per sample: 8 x D SH gather, per-channel sums in the reference's HeadSegmentedSum order (:383-411)
//   wave_composite  thread per ray: the sequential part (intersect_i, truncated re-weighting, log-transmittance,
//                   sample caches, early stop, colour; :370-547) in forward AND backward arithmetic (:2097-2101)
//   wave_bwd_wide   warp per sample: SH gradient scatter + d(colour)/d(position) (:2212-2263)
//   wave_bwd_hit    thread per sample: everything the reference does on lane 0 (:2137-2448, :2590-2866), started from the
//                   loop state the composite stage left in the sample record
//
// Entry order per ray = (listed voxel, root index), i.e. the order trace_ray_surf_trav meets them.

enum { ENT_NONE = 0, ENT_COUNT = 1, ENT_SAMPLE = 2, ENT_FAKE = 3 };
constexpr int WAVE_ENT = 3;   // entries per voxel: 3 roots of the one level set, or one fake sample

struct ItemRec {              // per queued (ray, voxel)
    double fs[4];             // cubic coefficients, level set subtracted
    double surf_miu, surf_std;
    int32_t n_ent;            // bits 0-7 count; entry e: kind at bits 8+8e .. 9+8e, root index at bits 10+8e .. 11+8e
    int32_t root_type;
};

struct HitRec {               // per entry
    float px, py, pz, ts;
    float raw_alpha, alpha;
    float fake_dist, reweight;
    float c0, c1, c2;         // per-channel SH sums (before + 0.5 and the clamp)
    float weight_f, weight_b; // compositing weight in forward / backward arithmetic
    float gx, gy, gz;         // d(colour loss)/d(position) from the SH part
    int32_t live_b;           // the backward loop reaches this entry
    // state of the backward loop when it reaches this entry (written by the composite stage)
    float accum_b, logT_b;
    int32_t isect_b, sample_b;
};

struct WaveP {
    ItemRec *items;
    HitRec *hits;
    int32_t *hitq;                 // compact queue of sample entries: item index * 4 + entry
    unsigned long long *n_hits;
    Pre *ray_pre;                  // (Q,) per-ray constants of the fused losses (fused_preamble)
    uint32_t *ray_mask;            // (Q, mask_words(K)): which of the ray's listed voxels produced entries (zeroed per call)
};

__device__ __forceinline__ int ent_kind(int32_t n_ent, int e) { return (n_ent >> (8 + 8 * e)) & 3; }
__device__ __forceinline__ int ent_root(int32_t n_ent, int e) { return (n_ent >> (10 + 8 * e)) & 3; }

// ray set-up shared by all stages: grid-space ray + reciprocals
__device__ __forceinline__ void wave_ray(const GridP &g, const asurf_opt_t &opt, const float *__restrict__ origins,
                                         const float *__restrict__ dirs, int64_t ray_id, Lane &L) {
    L.ray_id = ray_id;
    L.ox = origins[ray_id * 3 + 0]; L.oy = origins[ray_id * 3 + 1]; L.oz = origins[ray_id * 3 + 2];
    L.dx = dirs[ray_id * 3 + 0]; L.dy = dirs[ray_id * 3 + 1]; L.dz = dirs[ray_id * 3 + 2];
    float world_step;
    ray_bounds(g, opt, L, world_step);
}

// voxel of a listed cell + its far / near crossing times (next_from_list, voxel_advance PH_ENTER)
__device__ __forceinline__ void wave_voxel(const GridP &g, int32_t cell, Lane &L) {
    L.vz = cell % g.size[2];
    const int xy = cell / g.size[2];
    L.vy = xy % g.size[1];
    L.vx = xy / g.size[1];
    const float tfx = PT_X(L, L.vx + (L.dx > 0.f ? 1 : 0));
    const float tfy = PT_Y(L, L.vy + (L.dy > 0.f ? 1 : 0));
    const float tfz = PT_Z(L, L.vz + (L.dz > 0.f ? 1 : 0));
    L.t_far = fminf(fminf(tfx, tfy), tfz);
    const float tcx = PT_X(L, L.vx + (L.dx > 0.f ? 0 : 1));
    const float tcy = PT_Y(L, L.vy + (L.dy > 0.f ? 0 : 1));
    const float tcz = PT_Z(L, L.vz + (L.dz > 0.f ? 0 : 1));
    L.t_close = fmaxf(fmaxf(fmaxf(tcx, tcy), tcz), 0.f);
    L.nof[0] = fmaf(L.t_close, L.dx, L.ox);
    L.nof[1] = fmaf(L.t_close, L.dy, L.oy);
    L.nof[2] = fmaf(L.t_close, L.dz, L.oz);
    L.nno[0] = (double)L.nof[0] - L.vx;
    L.nno[1] = (double)L.nof[1] - L.vy;
    L.nno[2] = (double)L.nof[2] - L.vz;
}

__device__ __forceinline__ void wave_links(const GridP &g, Lane &L) {
    const int offy = g.size[2];
    const int64_t offx = (int64_t)g.size[1] * g.size[2];
    const int32_t *lp = g.links + (offx * L.vx + (int64_t)offy * L.vy + L.vz);
    L.lk[0] = __ldg(lp);
    L.lk[1] = __ldg(lp + 1);
    L.lk[2] = __ldg(lp + offy);
    L.lk[3] = __ldg(lp + offy + 1);
    L.lk[4] = __ldg(lp + offx);
    L.lk[5] = __ldg(lp + offx + 1);
    L.lk[6] = __ldg(lp + offx + offy);
    L.lk[7] = __ldg(lp + offx + offy + 1);
}

// ---- stage 1: per (ray, voxel) ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
wave_eval_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                 const PreP pre, const WaveP wv) {
    const int64_t n_all = (int64_t)*pre.n_items;
    const int64_t n_items = n_all < pre.item_cap ? n_all : pre.item_cap;
    const int lane = threadIdx.x & 31;
    for (int64_t t0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; t0 < n_items;
         t0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = t0 + lane;
        int n_ent = 0, n_samp = 0;
        const int32_t code = (t < n_items) ? __ldg(pre.itemq + t) : -1;
        if (code >= 0) {   // -1: slot of a ray that did not fit the queue (it stays with the persistent kernels)
            const int64_t ray_id = code / pre.K;
            Lane L;
            wave_ray(g, opt, origins, dirs, ray_id, L);
            wave_voxel(g, __ldg(pre.cells + code), L);
            wave_links(g, L);
#pragma unroll
            for (int c = 0; c < 8; ++c) L.sf[c] = __ldg(g.surface + L.lk[c]);
#pragma unroll
            for (int c = 0; c < 8; ++c) L.dn[c] = __ldg(g.density + L.lk[c]);
            float smin = L.sf[0], smax = L.sf[0];
#pragma unroll
            for (int c = 1; c < 8; ++c) {
                smin = fminf(smin, L.sf[c]);
                smax = fmaxf(smax, L.sf[c]);
            }
            ItemRec it;
            it.fs[0] = it.fs[1] = it.fs[2] = it.fs[3] = 0.;
            it.surf_miu = 0.;
            it.surf_std = 1.;
            it.root_type = ROOT_NONE;
            HitRec *hr = wv.hits + t * WAVE_ENT;
            bool has_sample = false, has_surf = false;
            int32_t packed = 0;
            const float lv = __ldg(g.level_set);
            if (!((lv < smin) || (lv > smax))) {   // the level set crosses this voxel (:273-277)
                has_surf = true;
                double s[8], dd[3] = {(double)L.dx, (double)L.dy, (double)L.dz};
#pragma unroll
                for (int c = 0; c < 8; ++c) s[c] = (double)L.sf[c];
                field_to_cubic(s, L.nno, dd, it.fs);
                it.fs[0] = it.fs[0] - (double)lv;
                double st[3] = {-1, -1, -1};
                it.root_type = solve_cubic(it.fs[0], it.fs[1], it.fs[2], it.fs[3], st);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const double stj = st[j];
                    if (stj <= 0) continue;
                    const float stf = (float)stj;
                    const float px = fmaf(stf, L.dx, L.nof[0]) - (float)L.vx;
                    const float py = fmaf(stf, L.dy, L.nof[1]) - (float)L.vy;
                    const float pz = fmaf(stf, L.dz, L.nof[2]) - (float)L.vz;
                    if ((px < 0) | (px > 1) | (py < 0) | (py > 1) | (pz < 0) | (pz > 1)) continue;
                    has_sample = true;
                    const float pos[3] = {px, py, pz};
                    if (opt.only_outward_intersect) {
                        float sg[3];
                        field_grad8(L.sf, pos, sg);
                        const float norm_dir_dot = -(sg[0] * L.dx + sg[1] * L.dy + sg[2] * L.dz);
                        if (norm_dir_dot >= 0.f) continue;
                    }
                    const float raw_alpha = trilerp8(L.dn, pos);
                    const bool samp = raw_alpha > opt.sigma_thresh;
                    HitRec h;
                    h.px = px; h.py = py; h.pz = pz;
                    h.ts = (float)((double)L.t_close + stj);
                    h.raw_alpha = raw_alpha;
                    h.alpha = samp ? alpha_act(raw_alpha, opt.alpha_activation_type) : 0.f;
                    h.fake_dist = 0.f; h.reweight = 1.f;
                    h.c0 = h.c1 = h.c2 = 0.f;
                    h.weight_f = h.weight_b = 0.f;
                    h.gx = h.gy = h.gz = 0.f;
                    h.live_b = 0; h.accum_b = 0.f; h.logT_b = 0.f; h.isect_b = 0; h.sample_b = 0;
                    hr[n_ent] = h;
                    n_samp |= (samp ? 1 : 0) << n_ent;
                    packed |= ((samp ? ENT_SAMPLE : ENT_COUNT) | (j << 2)) << (8 + 8 * n_ent);
                    n_ent += 1;
                }
            }
            // fake sample at the voxel midpoint (:423-541)
            if (opt.surf_fake_sample && !has_sample && (!opt.limited_fake_sample || has_surf) &&
                ((L.t_far - L.t_close) > opt.surf_fake_sample_min_vox_len)) {
                const float tm = (L.t_far + L.t_close) / 2.f;
                const float px = fmaf(tm, L.dx, L.ox) - (float)L.vx;
                const float py = fmaf(tm, L.dy, L.oy) - (float)L.vy;
                const float pz = fmaf(tm, L.dz, L.oz) - (float)L.vz;
                const float pos[3] = {px, py, pz};
                const float raw_alpha = trilerp8(L.dn, pos);
                if (raw_alpha > opt.sigma_thresh) {
                    double s[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) s[c] = (double)L.sf[c];
                    double const surf_miu = (s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7]) / 8;
                    double const var = (((s[0] - surf_miu) * (s[0] - surf_miu)) + ((s[1] - surf_miu) * (s[1] - surf_miu)) +
                                        ((s[2] - surf_miu) * (s[2] - surf_miu)) + ((s[3] - surf_miu) * (s[3] - surf_miu)) +
                                        ((s[4] - surf_miu) * (s[4] - surf_miu)) + ((s[5] - surf_miu) * (s[5] - surf_miu)) +
                                        ((s[6] - surf_miu) * (s[6] - surf_miu)) + ((s[7] - surf_miu) * (s[7] - surf_miu))) / 8;
                    double surf_std = (double)sqrtf((float)fmax((double)1e-9f, var));
                    if (!opt.fake_sample_normalize_surf) surf_std = 1.;
                    float ns[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) ns[c] = (float)(s[c] / surf_std);
                    const float fake_s = trilerp8(ns, pos);
                    float fake_dist = INFINITY;
                    fake_dist = fabsf(fake_s - lv) < fabsf(fake_dist) ? (fake_s - lv) : fake_dist;
                    const float q = fake_dist / g.fake_sample_std;
                    HitRec h;
                    h.px = px; h.py = py; h.pz = pz;
                    h.ts = tm;
                    h.raw_alpha = raw_alpha;
                    h.alpha = alpha_act(raw_alpha, opt.alpha_activation_type);
                    h.fake_dist = fake_dist;
                    h.reweight = __expf((float)(-.5 * (double)(q * q)));
                    h.c0 = h.c1 = h.c2 = 0.f;
                    h.weight_f = h.weight_b = 0.f;
                    h.gx = h.gy = h.gz = 0.f;
                    h.live_b = 0; h.accum_b = 0.f; h.logT_b = 0.f; h.isect_b = 0; h.sample_b = 0;
                    hr[0] = h;
                    it.surf_miu = surf_miu;
                    it.surf_std = surf_std;
                    n_ent = 1;
                    n_samp = 1;
                    packed = ENT_FAKE << 8;
                }
            }
            it.n_ent = n_ent | packed;
            wv.items[t] = it;
            if (n_ent > 0) {   // most listed voxels yield no entry: the per-ray stages only visit the flagged ones
                const int k = (int)(code - ray_id * pre.K);
                atomicOr(wv.ray_mask + ray_id * mask_words(pre.K) + (k >> 5), 1u << (k & 31));
            }
        }
        // queue the entries that need the wide (SH) stages: warp-aggregated append
        const int mine = __popc((unsigned)n_samp);
        int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += o;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        if (total) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(wv.n_hits, (unsigned long long)total);
            base = __shfl_sync(FULL, base, 0);
            int64_t w = (int64_t)base + incl - mine;
#pragma unroll
            for (int e = 0; e < WAVE_ENT; ++e)
                if ((n_samp >> e) & 1) wv.hitq[w++] = (int32_t)(t * 4 + e);
        }
    }
}

// ---- stage 2 / 4: warp per sample ------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(128)
wave_wide_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ dirs, const PreP pre, const WaveP wv,
                 const float *__restrict__ grad_in, const float *__restrict__ color_cache, const FusedP f,
                 const asurf_grads_t grads) {
    __shared__ float s_sph[4][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = g.sh_dim, bd = g.basis_dim;
    const int64_t n_hits = (int64_t)*wv.n_hits;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int offy = g.size[2];
    const int64_t offx = (int64_t)g.size[1] * g.size[2];
    for (int64_t h = warp0; h < n_hits; h += n_warps) {
        const int32_t hc = __ldg(wv.hitq + h);
        const int64_t t = hc >> 2;
        const int e = hc & 3;
        HitRec *hr = wv.hits + t * WAVE_ENT + e;
        if (BWD && !hr->live_b) continue;
        const int32_t code = __ldg(pre.itemq + t);
        const int64_t ray_id = code / pre.K;
        const int32_t cell = __ldg(pre.cells + code);
        const int vz = cell % g.size[2];
        const int xy = cell / g.size[2];
        const int vy = xy % g.size[1], vx = xy / g.size[1];
        const int32_t *lp = g.links + (offx * vx + (int64_t)offy * vy + vz);
        int lk[8];
        lk[0] = __ldg(lp); lk[1] = __ldg(lp + 1); lk[2] = __ldg(lp + offy); lk[3] = __ldg(lp + offy + 1);
        lk[4] = __ldg(lp + offx); lk[5] = __ldg(lp + offx + 1); lk[6] = __ldg(lp + offx + offy);
        lk[7] = __ldg(lp + offx + offy + 1);
        __syncwarp();
        if (lane == 0)   // SH basis from the world-space direction (:3165-3171)
            eval_sh(bd, dirs[ray_id * 3 + 0], dirs[ray_id * 3 + 1], dirs[ray_id * 3 + 2], s_sph[warp]);
        __syncwarp();
        const float pos[3] = {hr->px, hr->py, hr->pz};
        float v[8];
        float lane_color = 0.f, sph = 0.f;
        const int kb = (lane < D) ? (lane % bd) : 0;
        if (lane < D) {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = __ldg(g.sh + (int64_t)lk[c] * D + lane);
            sph = s_sph[warp][kb];
            lane_color = trilerp8(v, pos) * sph;
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = 0.f;
        }
        if (!BWD) {
            const float seg = segment_sum(lane_color, (lane < D) ? kb : 32, bd);
            const float c0 = __shfl_sync(FULL, seg, 0);
            const float c1 = __shfl_sync(FULL, seg, bd);
            const float c2 = __shfl_sync(FULL, seg, 2 * bd);
            if (lane == 0) { hr->c0 = c0; hr->c1 = c1; hr->c2 = c2; }
        } else {
            float g0, g1, g2;
            if (f.grad_is_rgb) {   // fused: dL/dRGB from the L2 / L1 mix (:3306-3316)
                float gi[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float resid = color_cache[ray_id * 3 + i] - grad_in[ray_id * 3 + i];
                    gi[i] = resid * f.norm_l2 * f.lambda_l2;
                    gi[i] += (resid > 0.f) ? (f.norm_l1 * f.lambda_l1) : (-f.norm_l1 * f.lambda_l1);
                }
                g0 = gi[0]; g1 = gi[1]; g2 = gi[2];
            } else {
                g0 = grad_in[ray_id * 3 + 0]; g1 = grad_in[ray_id * 3 + 1]; g2 = grad_in[ray_id * 3 + 2];
            }
            const float weight = hr->weight_b;
            const float l0 = hr->c0 + 0.5f, l1 = hr->c1 + 0.5f, l2 = hr->c2 + 0.5f;
            const float t0 = fmaxf(l0, 0.f), t1 = fmaxf(l1, 0.f), t2 = fmaxf(l2, 0.f);
            float gacc[3] = {0.f, 0.f, 0.f};
            if (lane < D) {
                const int ch = lane / bd;
                const float in01 = (ch == 0) ? ((t0 == l0) ? 1.f : 0.f)
                                             : ((ch == 1) ? ((t1 == l1) ? 1.f : 0.f) : ((t2 == l2) ? 1.f : 0.f));
                const float gch = (ch == 0) ? g0 : ((ch == 1) ? g1 : g2);
                const float grad_common = weight * in01 * gch;
                const float curr_grad_color = sph * grad_common;
                float w[8];
                corner_weights(pos, curr_grad_color, w);
#pragma unroll
                for (int c = 0; c < 8; ++c) atomicAdd(grads.grad_sh + (int64_t)lk[c] * D + lane, w[c]);
                if (!opt.no_surf_grad_from_sh) trilerp8_pos_grad(v, pos, curr_grad_color, gacc);
            }
            if (!opt.no_surf_grad_from_sh) {
                const float sx = warp_sum_down(gacc[0], lane);
                const float sy = warp_sum_down(gacc[1], lane);
                const float sz = warp_sum_down(gacc[2], lane);
                if (lane == 0) { hr->gx = sx; hr->gy = sy; hr->gz = sz; }
            }
        }
    }
}

// ---- stage 3: thread per ray, the sequential compositing ---------------------------------------------------------------------
// Forward: colours, sample caches.  With a gradient source (grad_in != NULL) it also walks the backward loop's scalar
// recurrences -- accum (:1795-1797, :2118), log-transmittance, intersect_i, cache index -- and leaves their value AT each
// entry in the entry's record, plus the per-ray loss constants, so that the backward of every entry can run on its own.
__global__ void __launch_bounds__(128)
wave_composite_kernel(const GridP g, const asurf_opt_t opt, const PreP pre, const WaveP wv, const CacheP cache, int M,
                      float *__restrict__ rgb_out, const float *__restrict__ grad_in,
                      const float *__restrict__ color_cache, const FusedP f) {
    const int64_t n_short = (int64_t)*pre.n_short;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_short; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ray_id = __ldg(pre.rays_short + s);
        const int32_t code = __ldg(pre.code + ray_id);
        const int n = code & CODE_CNT_MASK, n_bwd = (code >> CODE_CNT_BITS) & CODE_CNT_MASK;
        const int64_t base = __ldg(pre.item_base + ray_id);
        float logT = 0.f, logT_b = 0.f, out0 = 0.f, out1 = 0.f, out2 = 0.f;
        int intersect_i = -1, sample_i = 0;
        bool alive_f = true, alive_b = true;
        const uint32_t *rm = wv.ray_mask + ray_id * mask_words(pre.K);
        const int n_words = (n + 31) >> 5;
        // voxels without entries change nothing (the early-stop test after them sees the same log-transmittance as after the
        // previous voxel): visit only the flagged ones, in order
        for (int w = 0; w < n_words; ++w)
        for (uint32_t mm = rm[w]; mm && (alive_f || alive_b); mm &= mm - 1) {
            const int k = w * 32 + __ffs(mm) - 1;
            if (k >= n) break;
            if (k >= n_bwd) alive_b = false;
            const int32_t ne = wv.items[base + k].n_ent;
            const int cnt = ne & 255;
            for (int e = 0; e < cnt; ++e) {
                const int kind = ent_kind(ne, e);
                HitRec *hr = wv.hits + (base + k) * WAVE_ENT + e;
                if (kind != ENT_FAKE) ++intersect_i;
                if (kind == ENT_COUNT) continue;
                const float trw = opt.truncated_vol_render ? trunc_rw(intersect_i, g.trunc_a, opt.trunc_vol_weight_min) : 1.f;
                float rwalpha;
                if (kind == ENT_SAMPLE) {
                    rwalpha = hr->alpha * trw;
                } else {
                    rwalpha = hr->alpha * hr->reweight;
                    rwalpha = rwalpha * trw;
                }
                float wf = 0.f, wb = 0.f;
                if (alive_f) {
                    const float pcnt = -1 * __logf(1 - rwalpha);
                    wf = __expf(logT) * (1.f - __expf(-pcnt));
                    logT -= pcnt;
                    if ((sample_i < M) && (kind == ENT_SAMPLE || opt.fake_sample_l_dist)) {
                        cache.sa[ray_id * M + sample_i] = rwalpha;
                        cache.sw[ray_id * M + sample_i] = wf;
                        cache.st[ray_id * M + sample_i] = hr->ts;
                        sample_i += 1;
                    }
                    out0 += wf * fmaxf(hr->c0 + 0.5f, 0.f);
                    out1 += wf * fmaxf(hr->c1 + 0.5f, 0.f);
                    out2 += wf * fmaxf(hr->c2 + 0.5f, 0.f);
                }
                if (alive_b) {
                    hr->logT_b = logT_b;
                    hr->isect_b = intersect_i;
                    const float pcnt = -__logf(fmaxf(1.f - rwalpha, 1e-8f));
                    wb = __expf(logT_b) * (1.f - __expf(-pcnt));
                    logT_b -= pcnt;
                }
                hr->weight_f = wf;
                hr->weight_b = wb;
                hr->live_b = alive_b ? 1 : 0;
            }
            // early stop after the voxel (:544-547 forward, :2893-2895 backward)
            if (alive_f && (__expf(logT) < opt.stop_thresh)) {
                logT = -1e3f;
                alive_f = false;
            }
            if (alive_b && (__expf(logT_b) < opt.stop_thresh)) alive_b = false;
        }
        const float bg = g.has_bg ? 0.f : __expf(logT) * opt.background_brightness;
        float cc[3] = {out0 + bg, out1 + bg, out2 + bg};
        if (g.bg_lt) g.bg_lt[ray_id] = rgb_out ? logT : logT_b;   // forward state, or a stand-alone backward's own (msi.cu)
        if (rgb_out) {
            rgb_out[ray_id * 3 + 0] = cc[0];
            rgb_out[ray_id * 3 + 1] = cc[1];
            rgb_out[ray_id * 3 + 2] = cc[2];
            if (M > 0) cache.n[ray_id] = sample_i;
        }
        if (grad_in == nullptr) continue;
        // ---- the backward loop's recurrences ----
        if (color_cache) {
            cc[0] = color_cache[ray_id * 3 + 0]; cc[1] = color_cache[ray_id * 3 + 1]; cc[2] = color_cache[ray_id * 3 + 2];
        }
        float gi[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (f.grad_is_rgb) {   // fused: dL/dRGB from the L2 / L1 mix (:3306-3316)
                const float resid = cc[i] - grad_in[ray_id * 3 + i];
                gi[i] = resid * f.norm_l2 * f.lambda_l2;
                gi[i] += (resid > 0.f) ? (f.norm_l1 * f.lambda_l1) : (-f.norm_l1 * f.lambda_l1);
            } else {
                gi[i] = grad_in[ray_id * 3 + i];
            }
        }
        CacheView cv;
        cv.sa = cv.sw = cv.st = nullptr;
        cv.n = 0;
        if (M > 0) {
            cv.sa = cache.sa + ray_id * M; cv.sw = cache.sw + ray_id * M; cv.st = cache.st + ray_id * M;
            cv.n = rgb_out ? sample_i : cache.n[ray_id];
        }
        Pre lossc;
        fused_preamble(cv, lossc);
        wv.ray_pre[ray_id] = lossc;
        float accum = fmaf(cc[0], gi[0], fmaf(cc[1], gi[1], cc[2] * gi[2]));
        int sample_b = 0;
        bool live = true;
        for (int w = 0; w < n_words; ++w)
        for (uint32_t mm = rm[w]; mm && live; mm &= mm - 1) {
            const int k = w * 32 + __ffs(mm) - 1;
            if (k >= n_bwd) { live = false; break; }
            const int32_t ne = wv.items[base + k].n_ent;
            const int cnt = ne & 255;
            for (int e = 0; e < cnt; ++e) {
                const int kind = ent_kind(ne, e);
                if (kind == ENT_COUNT) continue;
                HitRec *hr = wv.hits + (base + k) * WAVE_ENT + e;
                if (!hr->live_b) { live = false; break; }
                hr->accum_b = accum;
                hr->sample_b = sample_b;
                const float l0 = hr->c0 + 0.5f, l1 = hr->c1 + 0.5f, l2 = hr->c2 + 0.5f;
                float total_color = fmaxf(l0, 0.f) * gi[0];   // shuffle order of the reference (:2112-2115): (c0 + c2) + c1
                total_color += fmaxf(l2, 0.f) * gi[2];
                total_color += fmaxf(l1, 0.f) * gi[1];
                accum -= hr->weight_b * total_color;
                if ((kind == ENT_SAMPLE || opt.fake_sample_l_dist) && (sample_b < M - 1)) sample_b += 1;
            }
        }
        if (g.bg_accum) g.bg_accum[ray_id] = accum - g.bg_beta;
    }
}

// ---- stage 5: thread per sample, the scalar part of its backward (the reference's lane-0 code, :2137-2448, :2590-2866) --------
__global__ void __launch_bounds__(128)
wave_bwd_hit_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ origins,
                    const float *__restrict__ dirs, const PreP pre, const WaveP wv, const float *__restrict__ grad_in,
                    const float *__restrict__ color_cache, const FusedP f, const CacheP cache,
                    const asurf_grads_t grads) {
    const int64_t n_hits = (int64_t)*wv.n_hits;
    const int M = f.M;
    for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < n_hits; h += (int64_t)gridDim.x * blockDim.x) {
        const int32_t hc = __ldg(wv.hitq + h);
        const int64_t t = hc >> 2;
        const int e = hc & 3;
        const HitRec hr = wv.hits[t * WAVE_ENT + e];
        if (!hr.live_b) continue;
        const int32_t code = __ldg(pre.itemq + t);
        const int64_t ray_id = code / pre.K;
        const ItemRec it = wv.items[t];
        const int kind = ent_kind(it.n_ent, e);
        Lane L;
        wave_ray(g, opt, origins, dirs, ray_id, L);
        wave_voxel(g, __ldg(pre.cells + code), L);
        wave_links(g, L);
#pragma unroll
        for (int c = 0; c < 8; ++c) L.sf[c] = __ldg(g.surface + L.lk[c]);
#pragma unroll
        for (int c = 0; c < 8; ++c) L.dn[c] = __ldg(g.density + L.lk[c]);
        L.fs[0] = it.fs[0]; L.fs[1] = it.fs[1]; L.fs[2] = it.fs[2]; L.fs[3] = it.fs[3];
        L.root_type = it.root_type;
        L.surf_miu = it.surf_miu;
        L.surf_std = it.surf_std;
        float g0, g1, g2;
        if (f.grad_is_rgb) {
            float gi[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float resid = color_cache[ray_id * 3 + i] - grad_in[ray_id * 3 + i];
                gi[i] = resid * f.norm_l2 * f.lambda_l2;
                gi[i] += (resid > 0.f) ? (f.norm_l1 * f.lambda_l1) : (-f.norm_l1 * f.lambda_l1);
            }
            g0 = gi[0]; g1 = gi[1]; g2 = gi[2];
        } else {
            g0 = grad_in[ray_id * 3 + 0]; g1 = grad_in[ray_id * 3 + 1]; g2 = grad_in[ray_id * 3 + 2];
        }
        CacheView cv;
        cv.sa = cv.sw = cv.st = nullptr;
        cv.n = 0;
        if (M > 0) {
            cv.sa = cache.sa + ray_id * M; cv.sw = cache.sw + ray_id * M; cv.st = cache.st + ray_id * M;
            cv.n = cache.n[ray_id];
        }
        const Pre lossc = wv.ray_pre[ray_id];
        L.logT = hr.logT_b;
        L.intersect_i = hr.isect_b;
        L.sample_i = hr.sample_b;
        float accum = hr.accum_b;
        L.px = hr.px; L.py = hr.py; L.pz = hr.pz;
        L.ts = hr.ts;
        L.raw_alpha = hr.raw_alpha;
        L.alpha = hr.alpha;
        L.trunc_rw_ = opt.truncated_vol_render ? trunc_rw(L.intersect_i, g.trunc_a, opt.trunc_vol_weight_min) : 1.f;
        L.fake = (kind == ENT_FAKE);
        if (!L.fake) {
            L.st_id = ent_root(it.n_ent, e);
            L.rwalpha = L.alpha * L.trunc_rw_;
            L.pcnt = -__logf(fmaxf(1.f - L.rwalpha, 1e-8f));
        } else {
            L.reweight = hr.reweight;
            L.fake_dist = hr.fake_dist;
            L.rwalpha = L.alpha * L.reweight * L.trunc_rw_;
            L.pcnt = -1 * __logf(fmaxf(1.f - L.rwalpha, 1e-8f));
        }
        L.weight = __expf(L.logT) * (1.f - __expf(-L.pcnt));
        const float l0 = hr.c0 + 0.5f, l1 = hr.c1 + 0.5f, l2 = hr.c2 + 0.5f;
        float total_color = fmaxf(l0, 0.f) * g0;   // shuffle order of the reference (:2112-2115): (c0 + c2) + c1
        total_color += fmaxf(l2, 0.f) * g2;
        total_color += fmaxf(l1, 0.f) * g1;
        if (!L.fake) finish_real_bwd(g, opt, f, lossc, cv, grads, L, accum, total_color, hr.gx, hr.gy, hr.gz);
        else finish_fake_bwd(g, opt, f, lossc, cv, grads, L, accum, total_color);
    }
}
