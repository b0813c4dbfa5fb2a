// alphasurf_b200: the alpha-Surf "surf_trav" renderer (forward, backward, fused) for sm_100a.
//
// Replaces volume_render_surf_trav / _backward / _fused of
// /root/reference/svox2/csrc/render_lerp_kernel_surf_trav.cu:3596-3942 (kernels :3139-3368, ray marchers
// :37-562 and :1710-2911).  Results follow the reference semantics voxel by voxel;
// the execution model is this repo's own:
//
//  * LANE PER RAY for everything that is scalar per ray -- the DDA, the 8-corner gates, the fp64 cubic
//    solve, compositing and all the lane-0-only loss terms of the reference.  The reference runs those
//    redundantly on 27 lanes of a warp that owns a single ray; here 32 rays advance per warp.
//  * The DDA tests ONE BIT of an occupancy bitmap (accel.cu; 16.7 MB at 512^3, L2 resident) per visited
//    voxel instead of gathering 8 scattered int32 links; the visited-voxel sequence and every t_close /
//    t_far stay bit-identical to the reference's incremental DDA (same IEEE divides, same tie-breaks).
//  * WARP PER SAMPLE for the wide part: when lanes have found a sample to composite, the warp walks the
//    pending lanes; for each, lane k < D gathers coefficient k of the 8 corner SH rows (8 coalesced
//    108-byte row reads), and in the backward scatters the 8xD gradient with coalesced red.global.add.
//  * The forward keeps the per-ray (rwalpha, weight, t) sample caches the fused losses need, but writes
//    only the entries a ray produced (plus a count) instead of zero-filling 3 x (Q,64) floats per call.
#include "common.cuh"
#include "surf_math.cuh"

namespace asurf {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CTA_THREADS = 128;
constexpr int CTA_WARPS = CTA_THREADS / 32;

struct GridP {
    const int32_t *links;
    const float *density, *surface, *sh, *level_set;
    const uint64_t *accel;  // occupancy pyramid: bit = all 8 links >= 0
    const uint64_t *work;   // work pyramid: bit = voxel can contribute a sample (asurf_work_build)
    int size[3];
    int level_set_num, basis_dim, sh_dim;
    int ab1, ab2;  // level-0 bitmap block counts along y and z
    AccelLayout lay;
    int use_skip;   // hierarchical empty-block skipping (off for the counting kernel)
    float offset[3], scaling[3];
    float fake_sample_std, trunc_a;
    // MSI background behind the grid (msi.cu): the pass then ends WITHOUT the background_brightness term (:550-552) and leaves
    // its per-ray state for the background pass -- bg_lt: final log-transmittance (written by the forward pass, or by a
    // backward pass that is not part of a fused call, :3776-3781 vs :3905); bg_accum: what the backward loop left of
    // `accum`, minus the beta term (:2901-2905).  Either pointer may be NULL.
    int has_bg;
    float *bg_lt, *bg_accum;
    float bg_beta;
};

// fused-loss scalars after the launch-time scaling of render_lerp_kernel_surf_trav.cu:3896-3914
struct FusedP {
    float sparsity_loss, lambda_l2, lambda_l1, lambda_l_dist, lambda_l_entropy, lambda_l_dist_a, lambda_l_entropy_a,
        lambda_l_samp_dist, lambda_l_di, l_di_alpha_thresh, surf_sparse_alpha_thresh, lambda_inplace_surf_sparse,
        lambda_inwards_norm_loss, lambda_conv_mode_samp;
    float norm_l2, norm_l1;  // 2/(3Q), 1/(3Q)
    int no_norm_weight_l_entropy;
    int M;                   // l_dist_max_sample
    int grad_is_rgb;         // fused: grad_in is rgb_gt
};

struct CacheP {
    float *sa, *sw, *st;  // (Q, M)
    int *n;               // (Q,) entries written
};

struct DebugP {
    int32_t max_hits;
    int32_t *hit_count, *hit_cell, *hit_kind;
    float *hit_t;
    float *xf;                    // (Q,9)
    unsigned long long *stats;    // asurf_stats_t
};

// lane state: what the lane needs next; phase: where it is inside a work voxel
// Output of the pre-march (premarch_kernel): per ray the first PRE_K voxels whose work bit is set, in march order,
// and where to resume the march if there are more.
constexpr int PRE_K_MAX = 512;   // list capacity per ray (PreP::K <= PRE_K_MAX; the counts are stored in 10 bits)
constexpr int PRE_K_DEFAULT = 128;   // ... used unless the previous call listed many voxels per ray (see premarch())
// words of the per-ray "voxel produced entries" mask of the wavefront path
__host__ __device__ __forceinline__ int mask_words(int K) { return (K + 31) >> 5; }
// PreP::code layout
constexpr int CODE_CNT_BITS = 10, CODE_CNT_MASK = (1 << CODE_CNT_BITS) - 1;
constexpr int CODE_CONT = 2 * CODE_CNT_BITS, CODE_BWD_ALIVE = CODE_CONT + 1, CODE_FORCE_FINE = CODE_CONT + 2;
static_assert(PRE_K_MAX <= CODE_CNT_MASK, "list counts must fit their bit field");
struct PreP {
    int32_t *cells;        // (Q, K) linear voxel index x*Y*Z + y*Z + z
    int K;                 // list capacity per ray
    int32_t *code;         // (Q,)  bits 0-9 count, 10-19 count visible to the backward loop, 20 continuation,
                           //       21 backward loop still alive at the continuation, 22 force_fine at the continuation
    float *cont_t;         // (Q,)  t at the continuation
    int32_t *cont_vox;     // (Q,)  next voxel at the continuation, 10 bits per axis
    int32_t *rays;         // compact list of the rays the persistent shading kernels serve
    unsigned long long *n_rays;   // its length (device counter)
    int enabled;
    // wavefront path (see "wavefront kernels" below): rays whose whole march fits the list ("short" rays)
    int wave;                     // 1: short rays go to the wavefront kernels, `rays` holds only the long ones
    int32_t *rays_short;          // compact list of short rays
    unsigned long long *n_short;  // its length
    int32_t *item_base;           // (Q,) first entry of the ray's voxels in the compact item queue
    int32_t *itemq;               // compact item queue: ray_id * K + k  (-1: slot of a ray that did not fit)
    int64_t item_cap;             // capacity of the item queue; rays that do not fit stay with the persistent kernels
    unsigned long long *n_items;  // its length
};

enum { ST_IDLE = 0, ST_MARCH = 1, ST_VOXEL = 2, ST_SAMPLE = 3 };
enum { PM_DONE = 0, PM_LOOKUP = 1, PM_JUMP = 2, PM_FINE = 3 };   // pre-march phases
constexpr int PM_FINE_STEPS = 16, PM_JUMP_STEPS = 3;
enum { PH_ENTER = 0, PH_ROOTS = 1, PH_FAKE = 2, PH_POST = 3 };

struct Lane {
    // ray in grid space; r* = RN(1/d*) for the exact fast divide, slow_div: a direction component is (almost) zero
    float ox, oy, oz, dx, dy, dz, tmin, tmax;
    float rx, ry, rz;
    bool slow_div;
    // pre-marched work-voxel list (premarch_kernel) being consumed, then an optional continuation of the march
    int list_pos, list_cnt, list_code;
    bool in_list, bwd_alive;
    // DDA
    int nx, ny, nz;
    float tfx, tfy, tfz, t;
    // current voxel
    int vx, vy, vz;
    float t_close, t_far;
    uint64_t word, w1, w2;   // cached pyramid words: level 0 (4^3 voxels), level 1 (16^3), level 2 (64^3)
    int wkey, k1, k2;
    int lk[8];
    float sf[8], dn[8];
    float smin, smax;
    double fs[4];
    double fs0;          // fs[0] before the level set is subtracted
    double st0, st1, st2;
    double nno[3];       // entry point relative to the voxel
    float nof[3];        // entry point, float
    int root_type, lv_i, j;
    int phase, state;
    bool ray_done, force_fine, has_sample, has_surf, dn_loaded, fs_ready;
    int64_t ray_id;
    int n_hits;
    // compositing state
    float logT;
    int intersect_i, sample_i;
    // pending sample
    float px, py, pz;
    float weight;
    bool fake;
    // backward-only pending data
    int st_id;
    float raw_alpha, alpha, trunc_rw_, rwalpha, pcnt;
    float fake_dist, reweight;
    double surf_miu, surf_std;
    float ts;            // t of the sample
};

// (plane - o) / d rounded exactly like the reference's IEEE divide, from the per-ray reciprocal r = RN(1/d):
// q = a*r followed by two residual corrections with exact FMA residuals (Markstein's correction step; checked against
// the true divide on 2.4e9 random (plane, o, d) of this range, including all-ones significands).  Lanes whose
// direction has an (almost) zero component take the true divide.
__device__ __forceinline__ float plane_t(int plane, float o, float d, float r, bool slow) {
    const float a = (float)plane - o;
    if (slow) return a / d;
    float q = a * r;
    float e = fmaf(-d, q, a);
    q = fmaf(e, r, q);
    e = fmaf(-d, q, a);
    return fmaf(e, r, q);
}
#define PT_X(L, p) plane_t((p), (L).ox, (L).dx, (L).rx, (L).slow_div)
#define PT_Y(L, p) plane_t((p), (L).oy, (L).dy, (L).ry, (L).slow_div)
#define PT_Z(L, p) plane_t((p), (L).oz, (L).dz, (L).rz, (L).slow_div)

__device__ __forceinline__ void load_density(const GridP &g, Lane &L) {
    if (!L.dn_loaded) {
#pragma unroll
        for (int c = 0; c < 8; ++c) L.dn[c] = __ldg(g.density + L.lk[c]);
        L.dn_loaded = true;
    }
}

__device__ __forceinline__ void div_init(Lane &L) {
    L.slow_div = !((fabsf(L.dx) >= 1e-18f) && (fabsf(L.dy) >= 1e-18f) && (fabsf(L.dz) >= 1e-18f));
    L.rx = 1.0f / L.dx;
    L.ry = 1.0f / L.dy;
    L.rz = 1.0f / L.dz;
}

// Ray set-up: world -> grid transform and AABB bounds (ray_find_bounds, include/render_util.cuh:651-701;
// transform_coord, include/cuda_util.cuh:63-69; _get_delta_scale, render_util.cuh:536-547).
__device__ __forceinline__ void ray_bounds(const GridP &g, const asurf_opt_t &opt, Lane &L, float &world_step) {
    L.ox = fmaf(L.ox, g.scaling[0], g.offset[0]);
    L.oy = fmaf(L.oy, g.scaling[1], g.offset[1]);
    L.oz = fmaf(L.oz, g.scaling[2], g.offset[2]);
    L.dx *= g.scaling[0];
    L.dy *= g.scaling[1];
    L.dz *= g.scaling[2];
    const float delta_scale = rnorm3df(L.dx, L.dy, L.dz);
    L.dx *= delta_scale;
    L.dy *= delta_scale;
    L.dz *= delta_scale;
    world_step = delta_scale * opt.step_size;
    if (opt.use_spheric_clip) {
        // ConcentricSpheresIntersector, render_util.cuh:619-649
        const float s0 = 2.f / (float)g.size[0], s1 = 2.f / (float)g.size[1], s2 = 2.f / (float)g.size[2];
        const float so0 = fmaf(L.ox + 0.5f, s0, -1.f), so1 = fmaf(L.oy + 0.5f, s1, -1.f), so2 = fmaf(L.oz + 0.5f, s2, -1.f);
        const float sd0 = L.dx * s0, sd1 = L.dy * s1, sd2 = L.dz * s2;
        const float q2a = 2 * (sd0 * sd0 + sd1 * sd1 + sd2 * sd2);
        const float qb = 2 * (so0 * sd0 + so1 * sd1 + so2 * sd2);
        const float f = qb * qb - 2 * q2a * (so0 * so0 + so1 * so1 + so2 * so2);
        const float r2 = 1.f - opt.near_clip;
        const float det1 = f + 2 * q2a * 1.f * 1.f, det2 = f + 2 * q2a * r2 * r2;
        bool ok = true;
        if (det1 < 0) ok = false; else L.tmax = (-qb + sqrtf(det1)) / q2a;
        if (ok) { if (det2 < 0) ok = false; else L.tmin = (-qb - sqrtf(det2)) / q2a; }
        if (!ok) { L.tmin = 1e-9f; L.tmax = 0.f; }
    } else {
        L.tmin = opt.near_clip / world_step * opt.step_size;
        L.tmax = 2e3f;
        {
            const float inv = (float)(1.0 / (double)L.dx);
            const float t1 = (-0.5f - L.ox) * inv, t2 = ((float)g.size[0] - 0.5f - L.ox) * inv;
            if (L.dx != 0.f) { L.tmin = fmaxf(L.tmin, fminf(t1, t2)); L.tmax = fminf(L.tmax, fmaxf(t1, t2)); }
        }
        {
            const float inv = (float)(1.0 / (double)L.dy);
            const float t1 = (-0.5f - L.oy) * inv, t2 = ((float)g.size[1] - 0.5f - L.oy) * inv;
            if (L.dy != 0.f) { L.tmin = fmaxf(L.tmin, fminf(t1, t2)); L.tmax = fminf(L.tmax, fmaxf(t1, t2)); }
        }
        {
            const float inv = (float)(1.0 / (double)L.dz);
            const float t1 = (-0.5f - L.oz) * inv, t2 = ((float)g.size[2] - 0.5f - L.oz) * inv;
            if (L.dz != 0.f) { L.tmin = fmaxf(L.tmin, fminf(t1, t2)); L.tmax = fminf(L.tmax, fmaxf(t1, t2)); }
        }
    }
    div_init(L);
}

__device__ __forceinline__ void dda_restart(Lane &L) {   // far-plane times of the next voxel + empty pyramid cache
    L.tfx = PT_X(L, L.nx + (L.dx > 0.f ? 1 : 0));
    L.tfy = PT_Y(L, L.ny + (L.dy > 0.f ? 1 : 0));
    L.tfz = PT_Z(L, L.nz + (L.dz > 0.f ? 1 : 0));
    L.wkey = -1;
    L.k1 = -1;
    L.k2 = -1;
    L.word = 0;
    L.w1 = 0;
    L.w2 = 0;
    L.in_list = false;
    L.state = ST_MARCH;
}

__device__ __forceinline__ void dda_init(const GridP &g, Lane &L) {
    L.t = L.tmin;
    L.nx = min(max((int)fmaf(L.t, L.dx, L.ox), 0), g.size[0] - 2);
    L.ny = min(max((int)fmaf(L.t, L.dy, L.oy), 0), g.size[1] - 2);
    L.nz = min(max((int)fmaf(L.t, L.dz, L.oz), 0), g.size[2] - 2);
    L.force_fine = false;
    L.bwd_alive = true;
    dda_restart(L);
}
// ---- exact hierarchical skipping -----------------------------------------------------------------------------------
// The reference DDA is a 3-way merge of the per-axis plane-crossing events ordered by (t, axis): each step takes the
// smallest t_far, ties going to the lowest axis (:188-197), and every t is the correctly rounded (plane - o) / d.
// Jumping over an empty aligned block therefore lands on EXACTLY the voxel the voxel-by-voxel DDA would reach, if we
// (a) find the first event (T, A) that leaves the block and (b) for the other two axes count the planes whose event
// precedes (T, A) in that order -- estimated from o + T*d and then corrected with the same divides the DDA uses.
__device__ __forceinline__ bool crossed(int p, float o, float d, float r, bool slow, float T, bool tie_before) {
    const float tp = plane_t(p, o, d, r, slow);
    return (tp < T) || ((tp == T) && tie_before);
}
// Voxel index on a non-exit axis after the jump.  floor(o + T*d) is within one voxel of the answer (its error is
// ~1e-4 voxel), so one exact test of the plane above and one of the plane below settle it; straight-line code so
// that the lanes of a warp stay together.
__device__ __forceinline__ int axis_after(int v, float o, float d, float r, bool slow, float T, bool tie_before, int lo,
                                          int hi) {
    const int est = (int)floorf(fmaf(T, d, o));
    const bool fwd = d > 0.f;   // fwd: crossing plane p enters voxel p; otherwise it enters voxel p - 1
    const int n = fwd ? min(max(est, v), hi - 1) : min(max(est, lo), v);
    const bool cu = crossed(n + 1, o, d, r, slow, T, tie_before);
    const bool cl = crossed(n, o, d, r, slow, T, tie_before);
    int delta;
    if (fwd) delta = ((n < hi - 1) && cu) ? 1 : (((n > v) && !cl) ? -1 : 0);
    else delta = ((n > lo) && cl) ? -1 : (((n < v) && !cu) ? 1 : 0);
    return n + delta;
}

struct Counters {
    unsigned long long steps, linked, active, samples, skips;
};

__device__ __forceinline__ void record_hit(const DebugP &dbg, const GridP &g, Lane &L) {
    if (dbg.hit_count) {
        if (L.n_hits < dbg.max_hits) {
            const int64_t o = L.ray_id * dbg.max_hits + L.n_hits;
            dbg.hit_cell[o] = (int32_t)(((int64_t)L.vx * g.size[1] + L.vy) * g.size[2] + L.vz);
            dbg.hit_kind[o] = (L.fake ? 3 : L.st_id) + 8 * L.intersect_i;
            dbg.hit_t[o] = L.ts;
        }
        ++L.n_hits;
    }
}

// One generalized DDA step of a marching lane -- the same instruction stream whether the lane walks one voxel (s = 0)
// or jumps over an empty aligned block of 2^s voxels per side (s = 4, 6), so that the rays of a warp stay converged.
// s = 0 reproduces one iteration of the reference loop (:86-221 / :1834-1937) for voxel (nx,ny,nz): on a set bit the
// lane goes to ST_VOXEL; the backward keeps the reference's `t += step_size` on unlinked voxels (:1935).
// TRACK (forward pre-march): also follow where the BACKWARD loop would stop because of that quirk (L.bwd_alive).
template <bool BWD, bool DEBUG, bool TRACK>
__device__ __forceinline__ void march_step(const GridP &g, const asurf_opt_t &opt, Lane &L, Counters &cnt) {
    if (!(L.t <= L.tmax)) {   // `while (t <= ray.tmax)`
        L.state = ST_IDLE;
        L.ray_done = true;
        return;
    }
    // DEBUG kernels walk the occupancy pyramid (bit = all 8 corner links >= 0) and apply the gates themselves;
    // the production kernels walk the work pyramid, whose bits already include the gates.
    const uint64_t *bm = DEBUG ? g.accel : g.work;
    int s = 0;
    const int k0 = ((L.nx >> 2) * g.ab1 + (L.ny >> 2)) * g.ab2 + (L.nz >> 2);
    if (k0 != L.wkey) {   // entering another 4^3 block: consult the pyramid top-down
        if (g.use_skip && !L.force_fine) {
            const int k2 = ((L.nx >> 6) * g.lay.b[2][1] + (L.ny >> 6)) * g.lay.b[2][2] + (L.nz >> 6);
            if (k2 != L.k2) {
                L.k2 = k2;
                L.w2 = __ldg(bm + g.lay.off[2] + k2);
            }
            const int bit1 = (((L.nx >> 4) & 3) << 4) | (((L.ny >> 4) & 3) << 2) | ((L.nz >> 4) & 3);
            if (L.w2 == 0) s = 6;
            else if (!((L.w2 >> bit1) & 1ull)) s = 4;
            // an empty 4^3 block is walked voxel by voxel: ~5 cheap steps beat one jump
        }
        if (s == 0) {
            L.wkey = k0;
            L.word = __ldg(bm + k0);
        }
    }
    // the block [lo, hi) that is left by this step; for s == 0 it is the voxel itself and T* are its cached t_far
    const int lox = (L.nx >> s) << s, loy = (L.ny >> s) << s, loz = (L.nz >> s) << s;
    const int hix = min(lox + (1 << s), g.size[0] - 1), hiy = min(loy + (1 << s), g.size[1] - 1),
              hiz = min(loz + (1 << s), g.size[2] - 1);
    const int Px = (L.dx > 0.f) ? hix : lox, Py = (L.dy > 0.f) ? hiy : loy, Pz = (L.dz > 0.f) ? hiz : loz;
    float Tx = L.tfx, Ty = L.tfy, Tz = L.tfz;
    if (s) {
        Tx = PT_X(L, Px);
        Ty = PT_Y(L, Py);
        Tz = PT_Z(L, Pz);
    }
    const float T = fminf(fminf(Tx, Ty), Tz);
    if (s) {
        if ((BWD || (TRACK && L.bwd_alive)) && !(T + opt.step_size <= L.tmax)) {
            L.force_fine = true;   // the backward quirk may end the loop in here: walk voxel by voxel from now on
            return;
        }
        if (!(T <= L.tmax)) {   // `while (t <= tmax)` fails at a voxel of this (empty) block
            L.state = ST_IDLE;
            L.ray_done = true;
            return;
        }
    }
    const int A = (T == Tx) ? 0 : ((T == Ty) ? 1 : 2);   // exit axis; ties go to the lowest axis (:188-197)
    const int vx = L.nx, vy = L.ny, vz = L.nz;
    int nx = vx, ny = vy, nz = vz;
    if (s) {   // the other two axes: planes whose crossing precedes the exit event (T, A)
        nx = axis_after(vx, L.ox, L.dx, L.rx, L.slow_div, T, A > 0, lox, hix);
        ny = axis_after(vy, L.oy, L.dy, L.ry, L.slow_div, T, A > 1, loy, hiy);
        nz = axis_after(vz, L.oz, L.dz, L.rz, L.slow_div, T, false, loz, hiz);
    }
    bool out;
    if (A == 0) {
        nx = (L.dx > 0.f) ? Px : Px - 1;
        out = (nx < 0) || (nx >= g.size[0] - 1);
    } else if (A == 1) {
        ny = (L.dy > 0.f) ? Py : Py - 1;
        out = (ny < 0) || (ny >= g.size[1] - 1);
    } else {
        nz = (L.dz > 0.f) ? Pz : Pz - 1;
        out = (nz < 0) || (nz >= g.size[2] - 1);
    }
    if (s && out) {   // the ray leaves the grid through this empty block
        L.state = ST_IDLE;
        L.ray_done = true;
        return;
    }
    if (!out) {
        L.nx = nx; L.ny = ny; L.nz = nz;
        L.tfx = PT_X(L, nx + (L.dx > 0.f ? 1 : 0));
        L.tfy = PT_Y(L, ny + (L.dy > 0.f ? 1 : 0));
        L.tfz = PT_Z(L, nz + (L.dz > 0.f ? 1 : 0));
    }
    L.t = out ? L.tmax + 1.f : T;
    if (s == 0) {
        if (DEBUG) ++cnt.steps;
        const int bit = ((vx & 3) << 4) | ((vy & 3) << 2) | (vz & 3);
        if ((L.word >> bit) & 1ull) {
            L.vx = vx; L.vy = vy; L.vz = vz;
            L.t_far = T;
            L.state = ST_VOXEL;
            L.phase = PH_ENTER;
        } else if (BWD) {
            // reference quirk (:1935): an UNLINKED voxel adds step_size to t before the loop test.  t is overwritten by
            // the next step, so this only matters when it ends the loop.
            if (DEBUG) {
                L.t += opt.step_size;
            } else if (!(L.t + opt.step_size <= L.tmax)) {
                if (!((__ldg(g.accel + k0) >> bit) & 1ull)) L.t += opt.step_size;
            }
        } else if (TRACK) {
            if (L.bwd_alive && !(L.t + opt.step_size <= L.tmax)) {
                if (!((__ldg(g.accel + k0) >> bit) & 1ull)) L.bwd_alive = false;   // the backward loop ends here
            }
        }
    } else if (DEBUG) {
        ++cnt.skips;
    }
}

// Next unit of work of a lane that consumes a pre-marched list: the next listed voxel, else the continuation of the
// march, else the ray is finished.
template <bool BWD>
__device__ __forceinline__ void next_from_list(const GridP &g, const PreP &pre, Lane &L) {
    if (L.list_pos < L.list_cnt) {
        const int32_t cell = __ldg(pre.cells + L.ray_id * pre.K + L.list_pos);
        ++L.list_pos;
        L.vz = cell % g.size[2];
        const int xy = cell / g.size[2];
        L.vy = xy % g.size[1];
        L.vx = xy / g.size[1];
        const float tfx = PT_X(L, L.vx + (L.dx > 0.f ? 1 : 0));
        const float tfy = PT_Y(L, L.vy + (L.dy > 0.f ? 1 : 0));
        const float tfz = PT_Z(L, L.vz + (L.dz > 0.f ? 1 : 0));
        L.t_far = fminf(fminf(tfx, tfy), tfz);
        L.state = ST_VOXEL;
        L.phase = PH_ENTER;
        return;
    }
    const bool cont = ((L.list_code >> CODE_CONT) & 1) && (!BWD || ((L.list_code >> CODE_BWD_ALIVE) & 1));
    if (cont) {
        const int32_t pv = __ldg(pre.cont_vox + L.ray_id);
        L.nx = pv & 1023;
        L.ny = (pv >> 10) & 1023;
        L.nz = (pv >> 20) & 1023;
        L.t = __ldg(pre.cont_t + L.ray_id);
        L.force_fine = (L.list_code >> CODE_FORCE_FINE) & 1;
        L.bwd_alive = true;
        dda_restart(L);   // -> ST_MARCH, in_list = false
        return;
    }
    L.state = ST_IDLE;
    L.ray_done = true;
}

// Work of a lane inside a voxel whose bit is set: 8-corner loads, level-set / root iteration, fake sample, early stop
// (trace_ray_surf_trav :212-547, backward :1921-2895).  Runs until the lane has a sample to composite (ST_SAMPLE; the
// phase is kept so that it resumes behind that sample), returns to marching (ST_MARCH) or ends the ray.
template <bool BWD, bool DEBUG>
__device__ __forceinline__ void voxel_advance(const GridP &g, const asurf_opt_t &opt, Lane &L, const CacheP &cache,
                                              int M, Counters &cnt, const DebugP &dbg, const PreP &pre) {
    const int offy = g.size[2];
    const int64_t offx = (int64_t)g.size[1] * g.size[2];
    const int64_t ray_id = L.ray_id;
    for (;;) {
        if (L.phase == PH_ENTER) {
            if (DEBUG) ++cnt.linked;
            const int32_t *lp = g.links + (offx * L.vx + (int64_t)offy * L.vy + L.vz);
            L.lk[0] = __ldg(lp);
            L.lk[1] = __ldg(lp + 1);
            L.lk[2] = __ldg(lp + offy);
            L.lk[3] = __ldg(lp + offy + 1);
            L.lk[4] = __ldg(lp + offx);
            L.lk[5] = __ldg(lp + offx + 1);
            L.lk[6] = __ldg(lp + offx + offy);
            L.lk[7] = __ldg(lp + offx + offy + 1);
            // 8-corner density gate (:230-239): skip only if ALL corners are below the threshold
            if (DEBUG) {
                bool pass = false;
#pragma unroll
                for (int c = 0; c < 8; ++c) pass |= !(__ldg(g.density + L.lk[c]) < opt.sigma_thresh);
                if (!pass) {
                    L.state = ST_MARCH;   // DEBUG kernels never consume a list
                    return;
                }
                ++cnt.active;
            }
            L.dn_loaded = false;
#pragma unroll
            for (int c = 0; c < 8; ++c) L.sf[c] = __ldg(g.surface + L.lk[c]);
            L.smin = L.sf[0];
            L.smax = L.sf[0];
#pragma unroll
            for (int c = 1; c < 8; ++c) {
                L.smin = fminf(L.smin, L.sf[c]);
                L.smax = fmaxf(L.smax, L.sf[c]);
            }
            const float tcx = PT_X(L, L.vx + (L.dx > 0.f ? 0 : 1));
            const float tcy = PT_Y(L, L.vy + (L.dy > 0.f ? 0 : 1));
            const float tcz = PT_Z(L, L.vz + (L.dz > 0.f ? 0 : 1));
            L.t_close = fmaxf(fmaxf(fmaxf(tcx, tcy), tcz), 0.f);
            L.nof[0] = fmaf(L.t_close, L.dx, L.ox);
            L.nof[1] = fmaf(L.t_close, L.dy, L.oy);
            L.nof[2] = fmaf(L.t_close, L.dz, L.oz);
            L.fs_ready = false;
            L.has_sample = false;
            L.has_surf = false;
            L.lv_i = 0;
            L.j = 3;
            L.phase = PH_ROOTS;
        }
        if (L.phase == PH_ROOTS) {
            bool found = false;
            for (;;) {
                if (L.j >= 3) {
                    // next level set that crosses this voxel (:273-277)
                    bool have_lv = false;
                    double lv_set = 0.;
                    while (L.lv_i < g.level_set_num) {
                        const float lv = __ldg(g.level_set + L.lv_i);
                        ++L.lv_i;
                        if ((lv < L.smin) || (lv > L.smax)) continue;
                        lv_set = (double)lv;
                        have_lv = true;
                        break;
                    }
                    if (!have_lv) break;
                    L.has_surf = true;
                    if (!L.fs_ready) {
                        double s[8], dd[3];
#pragma unroll
                        for (int c = 0; c < 8; ++c) s[c] = (double)L.sf[c];
                        L.nno[0] = (double)L.nof[0] - L.vx;
                        L.nno[1] = (double)L.nof[1] - L.vy;
                        L.nno[2] = (double)L.nof[2] - L.vz;
                        dd[0] = (double)L.dx; dd[1] = (double)L.dy; dd[2] = (double)L.dz;
                        field_to_cubic(s, L.nno, dd, L.fs);
                        L.fs0 = L.fs[0];
                        L.fs_ready = true;
                    }
                    L.fs[0] = L.fs0 - lv_set;
                    double st[3] = {-1, -1, -1};
                    L.root_type = solve_cubic(L.fs[0], L.fs[1], L.fs[2], L.fs[3], st);
                    L.st0 = st[0]; L.st1 = st[1]; L.st2 = st[2];
                    L.j = 0;
                }
                while (L.j < 3) {
                    const int j = L.j++;
                    const double stj = (j == 0) ? L.st0 : ((j == 1) ? L.st1 : L.st2);
                    if (stj <= 0) continue;
                    const float stf = (float)stj;
                    const float px = fmaf(stf, L.dx, L.nof[0]) - (float)L.vx;
                    const float py = fmaf(stf, L.dy, L.nof[1]) - (float)L.vy;
                    const float pz = fmaf(stf, L.dz, L.nof[2]) - (float)L.vz;
                    if ((px < 0) | (px > 1) | (py < 0) | (py > 1) | (pz < 0) | (pz > 1)) continue;
                    L.has_sample = true;
                    const float pos[3] = {px, py, pz};
                    if (opt.only_outward_intersect) {
                        float sg[3];
                        field_grad8(L.sf, pos, sg);
                        const float norm_dir_dot = -(sg[0] * L.dx + sg[1] * L.dy + sg[2] * L.dz);
                        if (norm_dir_dot >= 0.f) continue;
                    }
                    ++L.intersect_i;
                    load_density(g, L);
                    const float raw_alpha = trilerp8(L.dn, pos);
                    if (!(raw_alpha > opt.sigma_thresh)) continue;
                    const float alpha = alpha_act(raw_alpha, opt.alpha_activation_type);
                    const float trw = opt.truncated_vol_render ? trunc_rw(L.intersect_i, g.trunc_a, opt.trunc_vol_weight_min) : 1.f;
                    const float rwalpha = alpha * trw;
                    L.px = px; L.py = py; L.pz = pz;
                    L.fake = false;
                    L.st_id = j;
                    L.ts = (float)((double)L.t_close + stj);
                    if (!BWD) {
                        const float pcnt = -1 * __logf(1 - rwalpha);
                        L.weight = __expf(L.logT) * (1.f - __expf(-pcnt));
                        L.logT -= pcnt;
                        if (L.sample_i < M) {
                            cache.sa[ray_id * M + L.sample_i] = rwalpha;
                            cache.sw[ray_id * M + L.sample_i] = L.weight;
                            cache.st[ray_id * M + L.sample_i] = L.ts;
                            L.sample_i += 1;
                        }
                    } else {
                        L.raw_alpha = raw_alpha;
                        L.alpha = alpha;
                        L.trunc_rw_ = trw;
                        L.rwalpha = rwalpha;
                        L.pcnt = -__logf(fmaxf(1.f - rwalpha, 1e-8f));
                        L.weight = __expf(L.logT) * (1.f - __expf(-L.pcnt));
                    }
                    if (DEBUG) {
                        ++cnt.samples;
                        record_hit(dbg, g, L);
                    }
                    found = true;
                    break;
                }
                if (found) break;
            }
            if (found) {
                L.state = ST_SAMPLE;
                return;
            }
            L.phase = PH_FAKE;
        }
        if (L.phase == PH_FAKE) {
            L.phase = PH_POST;
            // fake sample at the voxel midpoint (:423-541 / :2460-2866)
            if (opt.surf_fake_sample && !L.has_sample && (!opt.limited_fake_sample || L.has_surf) &&
                ((L.t_far - L.t_close) > opt.surf_fake_sample_min_vox_len)) {
                const float tm = (L.t_far + L.t_close) / 2.f;
                const float px = fmaf(tm, L.dx, L.ox) - (float)L.vx;
                const float py = fmaf(tm, L.dy, L.oy) - (float)L.vy;
                const float pz = fmaf(tm, L.dz, L.oz) - (float)L.vz;
                const float pos[3] = {px, py, pz};
                load_density(g, L);
                const float raw_alpha = trilerp8(L.dn, pos);
                if (raw_alpha > opt.sigma_thresh) {
                    const float alpha = alpha_act(raw_alpha, opt.alpha_activation_type);
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
                    for (int i = 0; i < g.level_set_num; ++i) {
                        const float lv = __ldg(g.level_set + i);
                        fake_dist = fabsf(fake_s - lv) < fabsf(fake_dist) ? (fake_s - lv) : fake_dist;
                    }
                    const float q = fake_dist / g.fake_sample_std;
                    const float reweight = __expf((float)(-.5 * (double)(q * q)));
                    const float trw = opt.truncated_vol_render ? trunc_rw(L.intersect_i, g.trunc_a, opt.trunc_vol_weight_min) : 1.f;
                    L.px = px; L.py = py; L.pz = pz;
                    L.fake = true;
                    L.ts = tm;
                    if (!BWD) {
                        float a = alpha * reweight;
                        a = a * trw;
                        const float pcnt = -1 * __logf(1 - a);
                        L.weight = __expf(L.logT) * (1.f - __expf(-pcnt));
                        L.logT -= pcnt;
                        if ((L.sample_i < M) && opt.fake_sample_l_dist) {
                            cache.sa[ray_id * M + L.sample_i] = a;
                            cache.sw[ray_id * M + L.sample_i] = L.weight;
                            cache.st[ray_id * M + L.sample_i] = tm;
                            L.sample_i += 1;
                        }
                    } else {
                        L.raw_alpha = raw_alpha;
                        L.alpha = alpha;
                        L.trunc_rw_ = trw;
                        L.reweight = reweight;
                        L.fake_dist = fake_dist;
                        L.surf_miu = surf_miu;
                        L.surf_std = surf_std;
                        L.rwalpha = alpha * reweight * trw;
                        L.pcnt = -1 * __logf(fmaxf(1.f - L.rwalpha, 1e-8f));
                        L.weight = __expf(L.logT) * (1.f - __expf(-L.pcnt));
                    }
                    if (DEBUG) {
                        ++cnt.samples;
                        record_hit(dbg, g, L);
                    }
                    L.state = ST_SAMPLE;
                    return;
                }
            }
        }
        // PH_POST: early stop (:544-547 / :2893-2895)
        if (__expf(L.logT) < opt.stop_thresh) {
            if (!BWD) L.logT = -1e3f;
            L.state = ST_IDLE;
            L.ray_done = true;
            return;
        }
        if (!L.in_list) {
            L.state = ST_MARCH;
            return;
        }
        next_from_list<BWD>(g, pre, L);
        if (L.state != ST_VOXEL) return;
    }
}

// ---- pre-march: nothing but the generalized DDA ------------------------------------------------------------------------
// Finds, for every ray, the voxels whose work bit is set, in march order (up to PreP::K, with the state to resume from if
// there are more), writes the background colour of the rays that have no work at all, and compacts the others into queues
// for the shading kernels.  Two shapes: one thread per ray (premarch_kernel), or -- for large batches -- the two-level
// march below (premarch_coarse_kernel / premarch_fine_kernel / premarch_merge_kernel).

// The march of one lane from its current DDA state until the ray ends, the list is full (cont), or -- box_shift > 0 -- the
// ray leaves the aligned block of 2^box_shift voxels per side it started in.
__device__ __forceinline__ void premarch_march(const GridP &g, const asurf_opt_t &opt, Lane &L, int32_t *__restrict__ out_cells,
                                               int out_cap, int &n, int &n_bwd, bool &cont, int box_shift) {
    // The march runs in warp-wide phases so that lanes doing the same kind of step execute together: a JUMP phase
    // (look up which pyramid level is empty around the next voxel, leave empty 16^3 / 64^3 blocks: ~250 instructions per
    // jump) repeated while lanes keep jumping, then a FINE phase (voxel-by-voxel steps inside non-empty 16^3 blocks,
    // branch-free, ~40 instructions each, re-loading the 4^3 word when the lane crosses into the next one).  The
    // arithmetic of each step is that of march_step<false, false, true>.
    int mode = (L.state == ST_MARCH) ? PM_LOOKUP : PM_DONE;
    int jump_s = 0;
    const uint64_t *bm = g.work;
    const int sgx = (L.dx > 0.f) ? 1 : -1, sgy = (L.dy > 0.f) ? 1 : -1, sgz = (L.dz > 0.f) ? 1 : -1;
    // pyramid decision for the voxel (nx,ny,nz): FINE with its 4^3 word loaded, or JUMP over an empty block
    auto lookup = [&]() {
        int s = 0;
        if (g.use_skip && !L.force_fine) {
            const int k2 = ((L.nx >> 6) * g.lay.b[2][1] + (L.ny >> 6)) * g.lay.b[2][2] + (L.nz >> 6);
            if (k2 != L.k2) {
                L.k2 = k2;
                L.w2 = __ldg(bm + g.lay.off[2] + k2);
            }
            const int bit1 = (((L.nx >> 4) & 3) << 4) | (((L.ny >> 4) & 3) << 2) | ((L.nz >> 4) & 3);
            if (L.w2 == 0) s = 6;
            else if (!((L.w2 >> bit1) & 1ull)) s = 4;
        }
        if (s == 0) {
            const int k0 = ((L.nx >> 2) * g.ab1 + (L.ny >> 2)) * g.ab2 + (L.nz >> 2);
            L.wkey = k0;
            L.word = __ldg(bm + k0);
            mode = PM_FINE;
        } else {
            jump_s = s;
            mode = PM_JUMP;
        }
    };
    while (__any_sync(FULL, mode != PM_DONE)) {
        // ---------------- JUMP phase ----------------
#pragma unroll 1
        for (int jt = 0; jt < PM_JUMP_STEPS; ++jt) {
            if (mode == PM_LOOKUP) {
                if (!(L.t <= L.tmax)) mode = PM_DONE;
                else lookup();
            }
            if (!__any_sync(FULL, mode == PM_JUMP)) break;
            if (mode == PM_JUMP) {
                const int sft = jump_s;
                const int lox = (L.nx >> sft) << sft, loy = (L.ny >> sft) << sft, loz = (L.nz >> sft) << sft;
                const int hix = min(lox + (1 << sft), g.size[0] - 1), hiy = min(loy + (1 << sft), g.size[1] - 1),
                          hiz = min(loz + (1 << sft), g.size[2] - 1);
                const int Px = (L.dx > 0.f) ? hix : lox, Py = (L.dy > 0.f) ? hiy : loy, Pz = (L.dz > 0.f) ? hiz : loz;
                const float Tx = PT_X(L, Px), Ty = PT_Y(L, Py), Tz = PT_Z(L, Pz);
                const float T = fminf(fminf(Tx, Ty), Tz);
                if (L.bwd_alive && !(T + opt.step_size <= L.tmax)) {
                    L.force_fine = true;   // the backward quirk may end the loop in here: voxel by voxel from now on
                    mode = PM_LOOKUP;
                } else if (!(T <= L.tmax)) {
                    mode = PM_DONE;        // `while (t <= tmax)` fails at a voxel of this (empty) block
                } else {
                    const int A = (T == Tx) ? 0 : ((T == Ty) ? 1 : 2);
                    int nx = axis_after(L.nx, L.ox, L.dx, L.rx, L.slow_div, T, A > 0, lox, hix);
                    int ny = axis_after(L.ny, L.oy, L.dy, L.ry, L.slow_div, T, A > 1, loy, hiy);
                    int nz = axis_after(L.nz, L.oz, L.dz, L.rz, L.slow_div, T, false, loz, hiz);
                    nx = (A == 0) ? ((L.dx > 0.f) ? Px : Px - 1) : nx;
                    ny = (A == 1) ? ((L.dy > 0.f) ? Py : Py - 1) : ny;
                    nz = (A == 2) ? ((L.dz > 0.f) ? Pz : Pz - 1) : nz;
                    const int na = (A == 0) ? nx : ((A == 1) ? ny : nz);
                    const int lim = (A == 0) ? g.size[0] : ((A == 1) ? g.size[1] : g.size[2]);
                    if ((na < 0) || (na >= lim - 1)) {
                        mode = PM_DONE;    // the ray leaves the grid through this empty block
                    } else {
                        L.nx = nx; L.ny = ny; L.nz = nz;
                        L.tfx = PT_X(L, nx + (L.dx > 0.f ? 1 : 0));
                        L.tfy = PT_Y(L, ny + (L.dy > 0.f ? 1 : 0));
                        L.tfz = PT_Z(L, nz + (L.dz > 0.f ? 1 : 0));
                        L.t = T;
                        mode = PM_LOOKUP;
                    }
                }
            }
        }
        // ---------------- FINE phase: voxel-by-voxel steps inside non-empty 16^3 blocks ----------------
#pragma unroll 1
        for (int it = 0; it < PM_FINE_STEPS; ++it) {
            if (!__any_sync(FULL, mode == PM_FINE)) break;
            if (mode == PM_FINE) {
                if (!(L.t <= L.tmax)) {
                    mode = PM_DONE;
                } else {
                    const float T = fminf(fminf(L.tfx, L.tfy), L.tfz);
                    const int vx = L.nx, vy = L.ny, vz = L.nz;
                    // exit axis (ties go to the lowest axis, :188-197); everything below is select-based
                    const bool ax = (T == L.tfx), ay = (!ax) && (T == L.tfy);
                    const int m_old = ax ? vx : (ay ? vy : vz);
                    const int sg = ax ? sgx : (ay ? sgy : sgz);
                    const int lim = ax ? g.size[0] : (ay ? g.size[1] : g.size[2]);
                    const int m_new = m_old + sg;
                    const bool out = (m_new < 0) || (m_new >= lim - 1);
                    const float oa = ax ? L.ox : (ay ? L.oy : L.oz), da = ax ? L.dx : (ay ? L.dy : L.dz),
                                ra = ax ? L.rx : (ay ? L.ry : L.rz);
                    const float tnew = plane_t(m_new + (sg > 0 ? 1 : 0), oa, da, ra, L.slow_div);
                    if (!out) {
                        L.nx = ax ? m_new : vx;
                        L.ny = ay ? m_new : vy;
                        L.nz = (ax || ay) ? vz : m_new;
                        L.tfx = ax ? tnew : L.tfx;
                        L.tfy = ay ? tnew : L.tfy;
                        L.tfz = (ax || ay) ? L.tfz : tnew;
                    }
                    L.t = out ? L.tmax + 1.f : T;
                    const int bit = ((vx & 3) << 4) | ((vy & 3) << 2) | (vz & 3);
                    if ((L.word >> bit) & 1ull) {
                        out_cells[n] = (int32_t)(((int64_t)vx * g.size[1] + vy) * g.size[2] + vz);
                        ++n;
                        if (L.bwd_alive) ++n_bwd;
                        if (n == out_cap) {
                            cont = true;   // resume here (the shading kernels re-check `t <= tmax`)
                            mode = PM_DONE;
                        }
                    } else if (L.bwd_alive && !(L.t + opt.step_size <= L.tmax)) {
                        // backward quirk (:1935): an UNLINKED voxel this close to tmax ends the backward loop
                        if (!((__ldg(g.accel + L.wkey) >> bit) & 1ull)) L.bwd_alive = false;
                    }
                    if (mode == PM_FINE) {
                        if (out) mode = PM_DONE;
                        else if (box_shift && ((m_old ^ m_new) >> box_shift)) mode = PM_DONE;   // left the item's block
                        else if ((m_old ^ m_new) >> 2) lookup();   // next 4^3 block: new word, or an empty block to jump
                    }
                }
            }
        }
    }
}

// Per-ray epilogue: list header, background for rays without work, short / long classification and the compact queues.
__device__ __forceinline__ void premarch_finish(const GridP &g, const asurf_opt_t &opt, const PreP &pre, const Lane &L,
                                                int64_t ray_id, bool valid, int n, int n_bwd, bool cont,
                                                float *__restrict__ rgb_out, int *__restrict__ cache_n) {
    const int lane = threadIdx.x & 31;
    bool has_work = false;
    if (valid) {
        int code = n | (n_bwd << CODE_CNT_BITS);
        if (cont) {
            code |= (1 << CODE_CONT) | ((L.bwd_alive ? 1 : 0) << CODE_BWD_ALIVE) | ((L.force_fine ? 1 : 0) << CODE_FORCE_FINE);
            pre.cont_t[ray_id] = L.t;
            pre.cont_vox[ray_id] = L.nx | (L.ny << 10) | (L.nz << 20);
        }
        pre.code[ray_id] = code;
        has_work = (n > 0) || cont;
        if (!has_work) {
            if (rgb_out) {   // forward: the ray composites nothing -> background (:59-65, :553-555)
                const float bg = g.has_bg ? 0.f : opt.background_brightness;
                rgb_out[ray_id * 3 + 0] = bg; rgb_out[ray_id * 3 + 1] = bg; rgb_out[ray_id * 3 + 2] = bg;
            }
            if (g.bg_lt) g.bg_lt[ray_id] = 0.f;
            // bg_accum keeps its NaN = "never visited by the backward loop" (msi_backward_kernel starts such a ray from the
            // initial accum minus the beta term, :2901-2905); a ray that misses the grid returns before that (:1810-1815): +Inf
            if (g.bg_accum && (L.tmin > L.tmax)) g.bg_accum[ray_id] = INFINITY;
            if (cache_n) cache_n[ray_id] = 0;
        }
    }
    // short rays (whole march listed, and room left in the item queue) feed the wavefront kernels, the others the
    // persistent shading kernels
    const bool want_short = has_work && pre.wave && !cont;
    bool is_short = false;
    const unsigned mw = __ballot_sync(FULL, want_short);
    if (mw) {
        // item queue: the ray's n voxels get consecutive entries (exclusive prefix over the warp + one atomic)
        int incl = want_short ? n : 0;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += o;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        unsigned long long ibase = 0;
        if (lane == 0) ibase = atomicAdd(pre.n_items, (unsigned long long)total);
        ibase = __shfl_sync(FULL, ibase, 0);
        if (want_short) {
            const int64_t first = (int64_t)ibase + incl - n;
            is_short = (first + n <= pre.item_cap);
            if (is_short) pre.item_base[ray_id] = (int32_t)first;
            for (int k = 0; k < n; ++k)
                if (first + k < pre.item_cap) pre.itemq[first + k] = is_short ? (int32_t)(ray_id * pre.K + k) : -1;
        }
    }
    const bool is_long = has_work && !is_short;
    const unsigned m = __ballot_sync(FULL, is_long);
    if (m) {
        unsigned long long base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(pre.n_rays, (unsigned long long)__popc(m));
        base = __shfl_sync(FULL, base, __ffs(m) - 1);
        if (is_long) pre.rays[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)ray_id;
    }
    const unsigned ms = __ballot_sync(FULL, is_short);
    if (ms) {
        unsigned long long base = 0;
        if (lane == __ffs(ms) - 1) base = atomicAdd(pre.n_short, (unsigned long long)__popc(ms));
        base = __shfl_sync(FULL, base, __ffs(ms) - 1);
        if (is_short) pre.rays_short[base + __popc(ms & ((1u << lane) - 1u))] = (int32_t)ray_id;
    }
}

__device__ __forceinline__ void premarch_ray_setup(const GridP &g, const asurf_opt_t &opt, const float *__restrict__ origins,
                                                   const float *__restrict__ dirs, int64_t ray_id, Lane &L) {
    L.ox = origins[ray_id * 3 + 0]; L.oy = origins[ray_id * 3 + 1]; L.oz = origins[ray_id * 3 + 2];
    L.dx = dirs[ray_id * 3 + 0]; L.dy = dirs[ray_id * 3 + 1]; L.dz = dirs[ray_id * 3 + 2];
    float world_step;
    ray_bounds(g, opt, L, world_step);
    if (!(L.tmin > L.tmax)) dda_init(g, L);
}

__global__ void __launch_bounds__(256)
premarch_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                const int64_t Q, const PreP pre, float *__restrict__ rgb_out, int *__restrict__ cache_n) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    Lane L;
    L.state = ST_IDLE;
    L.ray_done = false;
    int n = 0, n_bwd = 0;
    bool cont = false;
    if (ray_id < Q) premarch_ray_setup(g, opt, origins, dirs, ray_id, L);
    premarch_march(g, opt, L, pre.cells + (ray_id < Q ? ray_id : 0) * pre.K, pre.K, n, n_bwd, cont, 0);
    premarch_finish(g, opt, pre, L, ray_id, ray_id < Q, n, n_bwd, cont, rgb_out, cache_n);
}

// ---- two-level pre-march -----------------------------------------------------------------------------------------------------
// coarse: thread per ray, block jumps ONLY -- every 16^3 / 64^3 block is left with the exact landing rule whether it is empty
//         or not (the rule does not depend on the block's content); a non-empty 16^3 block (or any block in the last
//         step_size of the ray, where the backward's early exit has to be tracked voxel by voxel) is queued as an item
//         (ray, entry voxel, entry t).  All lanes of a warp do the same kind of step.
// fine:   thread per item, voxel-by-voxel steps inside that one block, listing its work voxels.  Items are homogeneous
//         (<= 46 steps), so warps stay full -- the long fine stretches of grazing rays no longer hold 31 other rays up.
// merge:  thread per ray, concatenates its items' lists in order and runs the per-ray epilogue.
constexpr int CI_MAX = 96;       // items per ray (non-empty 16^3 blocks it crosses); more -> the ray goes to the persistent kernels
constexpr int FI_K = 48;         // work voxels listed per item (a ray crosses at most 46 voxels of a 16^3 block)
struct TwoP {
    int32_t *ray_items;          // (Q, CI_MAX, 2): packed voxel (10 bits per axis), t bits
    int32_t *ray_nitems;         // (Q,) number of items, or -1 when there were more than CI_MAX
    int32_t *item_first;         // (Q,) first entry of the ray's items in the item queue
    int32_t *fq;                 // fine-item queue: ray_id * CI_MAX + i
    unsigned long long *n_fq;
    int32_t *fi_cells;           // (Q * CI_MAX budget, FI_K) indexed by queue position
    int32_t *fi_meta;            // count | visible-to-backward count << 8 | alive at the end << 16 | overflow << 17
    int64_t fq_cap;
};

__global__ void __launch_bounds__(256)
premarch_coarse_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                       const int64_t Q, const TwoP tw) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    Lane L;
    L.state = ST_IDLE;
    L.ray_done = false;
    if (ray_id < Q) premarch_ray_setup(g, opt, origins, dirs, ray_id, L);
    const uint64_t *bm = g.work;
    int n_items = 0;
    bool tail = false;   // within step_size of tmax: every further block becomes an item (backward quirk, :1935)
    bool active = (L.state == ST_MARCH);
    while (__any_sync(FULL, active)) {
        if (active) {
            if (!(L.t <= L.tmax)) {
                active = false;
            } else {
                int sft = 4;
                bool emit = tail || !g.use_skip;
                if (g.use_skip) {
                    const int k2 = ((L.nx >> 6) * g.lay.b[2][1] + (L.ny >> 6)) * g.lay.b[2][2] + (L.nz >> 6);
                    if (k2 != L.k2) {
                        L.k2 = k2;
                        L.w2 = __ldg(bm + g.lay.off[2] + k2);
                    }
                    const int bit1 = (((L.nx >> 4) & 3) << 4) | (((L.ny >> 4) & 3) << 2) | ((L.nz >> 4) & 3);
                    if (L.w2 == 0) sft = tail ? 4 : 6;
                    else if ((L.w2 >> bit1) & 1ull) emit = true;
                }
                const int lox = (L.nx >> sft) << sft, loy = (L.ny >> sft) << sft, loz = (L.nz >> sft) << sft;
                const int hix = min(lox + (1 << sft), g.size[0] - 1), hiy = min(loy + (1 << sft), g.size[1] - 1),
                          hiz = min(loz + (1 << sft), g.size[2] - 1);
                const int Px = (L.dx > 0.f) ? hix : lox, Py = (L.dy > 0.f) ? hiy : loy, Pz = (L.dz > 0.f) ? hiz : loz;
                const float Tx = PT_X(L, Px), Ty = PT_Y(L, Py), Tz = PT_Z(L, Pz);
                const float T = fminf(fminf(Tx, Ty), Tz);
                if (!tail && !(T + opt.step_size <= L.tmax)) {
                    tail = true;       // re-evaluate this position with 16^3 granularity and an item
                } else {
                    if (emit) {
                        if (n_items < CI_MAX) {
                            tw.ray_items[(ray_id * CI_MAX + n_items) * 2 + 0] = L.nx | (L.ny << 10) | (L.nz << 20);
                            tw.ray_items[(ray_id * CI_MAX + n_items) * 2 + 1] = __float_as_int(L.t);
                        }
                        ++n_items;
                    }
                    if (!(T <= L.tmax)) {
                        active = false;    // `while (t <= tmax)` fails inside this block (the item, if any, handles it)
                    } else {
                        const int A = (T == Tx) ? 0 : ((T == Ty) ? 1 : 2);
                        int nx = axis_after(L.nx, L.ox, L.dx, L.rx, L.slow_div, T, A > 0, lox, hix);
                        int ny = axis_after(L.ny, L.oy, L.dy, L.ry, L.slow_div, T, A > 1, loy, hiy);
                        int nz = axis_after(L.nz, L.oz, L.dz, L.rz, L.slow_div, T, false, loz, hiz);
                        nx = (A == 0) ? ((L.dx > 0.f) ? Px : Px - 1) : nx;
                        ny = (A == 1) ? ((L.dy > 0.f) ? Py : Py - 1) : ny;
                        nz = (A == 2) ? ((L.dz > 0.f) ? Pz : Pz - 1) : nz;
                        const int na = (A == 0) ? nx : ((A == 1) ? ny : nz);
                        const int lim = (A == 0) ? g.size[0] : ((A == 1) ? g.size[1] : g.size[2]);
                        if ((na < 0) || (na >= lim - 1)) {
                            active = false;   // the ray leaves the grid through this block
                        } else {
                            L.nx = nx; L.ny = ny; L.nz = nz;
                            L.t = T;
                        }
                    }
                }
            }
        }
    }
    // queue the items: a ray's items get consecutive queue entries
    const bool ok = (ray_id < Q) && (n_items > 0) && (n_items <= CI_MAX);
    int incl = ok ? n_items : 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += o;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    unsigned long long base = 0;
    if (total) {
        if (lane == 0) base = atomicAdd(tw.n_fq, (unsigned long long)total);
        base = __shfl_sync(FULL, base, 0);
    }
    if (ray_id < Q) {
        int32_t ni = (n_items > CI_MAX) ? -1 : n_items;
        if (ok) {
            const int64_t first = (int64_t)base + incl - n_items;
            if (first + n_items <= tw.fq_cap) {
                tw.item_first[ray_id] = (int32_t)first;
                for (int i = 0; i < n_items; ++i) tw.fq[first + i] = (int32_t)(ray_id * CI_MAX + i);
            } else {
                ni = -1;   // queue full: the persistent kernels march this ray
                for (int i = 0; i < n_items; ++i)
                    if (first + i < tw.fq_cap) tw.fq[first + i] = -1;
            }
        }
        tw.ray_nitems[ray_id] = ni;
    }
}

__global__ void __launch_bounds__(256)
premarch_fine_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                     const TwoP tw) {
    const int64_t n_all = (int64_t)*tw.n_fq;
    const int64_t n_fq = n_all < tw.fq_cap ? n_all : tw.fq_cap;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    Lane L;
    L.state = ST_IDLE;
    L.ray_done = false;
    int n = 0, n_bwd = 0;
    bool cont = false;
    const int32_t code = (q < n_fq) ? __ldg(tw.fq + q) : -1;
    if (code >= 0) {
        const int64_t ray_id = code / CI_MAX;
        premarch_ray_setup(g, opt, origins, dirs, ray_id, L);   // bounds + reciprocals (the start voxel is overwritten)
        const int32_t pv = tw.ray_items[(int64_t)code * 2 + 0];
        L.nx = pv & 1023;
        L.ny = (pv >> 10) & 1023;
        L.nz = (pv >> 20) & 1023;
        L.t = __int_as_float(tw.ray_items[(int64_t)code * 2 + 1]);
        L.force_fine = true;   // inside the item's block every voxel is visited
        L.bwd_alive = true;
        dda_restart(L);
    }
    premarch_march(g, opt, L, tw.fi_cells + (code >= 0 ? q : 0) * FI_K, FI_K, n, n_bwd, cont, 4);
    if (code >= 0) tw.fi_meta[q] = n | (n_bwd << 8) | ((L.bwd_alive ? 1 : 0) << 16) | ((cont ? 1 : 0) << 17);
}

__global__ void __launch_bounds__(256)
premarch_merge_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                      const int64_t Q, const PreP pre, const TwoP tw, float *__restrict__ rgb_out, int *__restrict__ cache_n) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    Lane L;
    L.state = ST_IDLE;
    L.ray_done = false;
    int n = 0, n_bwd = 0;
    bool cont = false;
    if (ray_id < Q) {
        const int ni = tw.ray_nitems[ray_id];
        bool alive = true, too_long = (ni < 0);
        const int64_t first = (ni > 0) ? (int64_t)tw.item_first[ray_id] : 0;
        for (int i = 0; i < ni && !too_long; ++i) {
            const int32_t m = tw.fi_meta[first + i];
            const int c = m & 255;
            if (((m >> 17) & 1) || (n + c > pre.K)) {
                too_long = true;
                break;
            }
            const int32_t *src = tw.fi_cells + (first + i) * FI_K;
            for (int k = 0; k < c; ++k) pre.cells[ray_id * pre.K + n + k] = src[k];
            n += c;
            if (alive) n_bwd += (m >> 8) & 255;
            if (!((m >> 16) & 1)) alive = false;
        }
        if (too_long) {   // more work than the lists hold: the persistent kernels march this ray from its start
            premarch_ray_setup(g, opt, origins, dirs, ray_id, L);
            n = 0;
            n_bwd = 0;
            cont = true;
        }
    }
    premarch_finish(g, opt, pre, L, ray_id, ray_id < Q, n, n_bwd, cont, rgb_out, cache_n);
}

// Sum of `bd` consecutive lanes starting at a segment head, same add order as the reference's
// cub::WarpReduce::HeadSegmentedSum (shuffle-down tree clamped to the segment).
__device__ __forceinline__ float segment_sum(float v, int pos_in_seg, int bd) {
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        const float o = __shfl_down_sync(FULL, v, off);
        if (pos_in_seg + off < bd) v += o;
    }
    return v;
}
// full-warp shuffle-down tree (cub::WarpReduce::Sum order); the total lands in lane 0
__device__ __forceinline__ float warp_sum_down(float v, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float o = __shfl_down_sync(FULL, v, off);
        if (lane + off < 32) v += o;
    }
    return v;
}

// ---- per-ray constants of the fused losses (preamble of trace_ray_surf_trav_backward, :1756-1793) --------
struct Pre {
    float asum, wsum, Den_Dasum, Den_Dwsum, t_mean, Dmeant_sign;
    int valid_n, max_id;
};

struct CacheView {
    const float *sa, *sw, *st;
    int n;
    __device__ __forceinline__ float A(int i) const { return (i < n) ? sa[i] : 0.f; }
    __device__ __forceinline__ float W(int i) const { return (i < n) ? sw[i] : 0.f; }
    __device__ __forceinline__ float T(int i) const { return (i < n) ? st[i] : 0.f; }
};

__device__ __forceinline__ float log_clamped_ratio(float v, float sum) {
    // _LOG(max(v, 1e-8) / sum) with the reference's double promotion
    return __logf((float)(fmax((double)v, 1e-8) / (double)sum));
}

__device__ __forceinline__ void fused_preamble(const CacheView &c, Pre &p) {
    p.asum = 0.f; p.wsum = 0.f;
    for (int i = 0; i < c.n; ++i) { p.asum += c.sa[i]; p.wsum += c.sw[i]; }
    p.asum = fmaxf(p.asum, 1e-8f);
    p.wsum = fmaxf(p.wsum, 1e-8f);
    p.Den_Dasum = 0.f; p.Den_Dwsum = 0.f; p.t_mean = 0.f;
    for (int i = 0; i < c.n; ++i) {
        p.Den_Dasum += c.sa[i] * (log_clamped_ratio(c.sa[i], p.asum) + 1.f) / (p.asum * p.asum);
        p.Den_Dwsum += c.sw[i] * (log_clamped_ratio(c.sw[i], p.wsum) + 1.f) / (p.wsum * p.wsum);
        p.t_mean += c.sw[i] / p.wsum * c.st[i];
    }
    p.valid_n = 0; p.Dmeant_sign = 0.f;
    for (int i = 0; i < c.n; ++i) {
        if (c.st[i] > 0.f) {
            p.valid_n++;
            p.Dmeant_sign += (p.t_mean > c.st[i]) ? 1.f : ((p.t_mean < c.st[i]) ? -1.f : 0.f);
        }
    }
    p.max_id = 0;
    float max_w = 0.f;
    for (int i = 0; i < c.n; ++i)
        if (c.sw[i] > max_w) { max_w = c.sw[i]; p.max_id = i; }
}

// l_dist(w) and l_entropy(w) contributions to d/d(rwalpha) (:2141-2210)
__device__ __forceinline__ float extra_grad_rwalpha_w(const FusedP &f, const Pre &p, const CacheView &c, int sample_i, float logT,
                                      float rwalpha) {
    float add = 0.f;
    const float denom = fminf(rwalpha - 1.f, -1e-8f);
    if (f.lambda_l_dist > 0.f) {
        float Dldist_Dai = 0.f, log_Tk = logT;
        for (int k = sample_i; k < p.valid_n; ++k) {
            float Dldist_Dwk = 0.f;
            for (int j = 0; j < c.n; ++j) Dldist_Dwk += c.sw[j] * fabsf(c.T(k) - c.st[j]);
            // entries j >= n are zero-weight in the reference's zero-filled cache
            if (k == sample_i) {
                Dldist_Dai += Dldist_Dwk * __expf(logT);
            } else {
                log_Tk += __logf(fmaxf(1.f - c.A(k - 1), 1e-8f));
                Dldist_Dai += Dldist_Dwk * __expf(log_Tk) * c.A(k) / denom;
            }
        }
        add += f.lambda_l_dist * Dldist_Dai;
    }
    if (f.lambda_l_entropy > 0.f) {
        float Den_Dai, log_Tk = logT;
        if (f.no_norm_weight_l_entropy) {
            const float Den_Dwi = -(__logf(fmaxf(c.W(sample_i), 1e-8f)) + 1.f);
            Den_Dai = Den_Dwi * __expf(logT);
            for (int k = sample_i + 1; k < p.valid_n; ++k) {
                const float Den_Dwk = -(__logf(fmaxf(c.W(k), 1e-8f)) + 1.f);
                log_Tk += __logf(fmaxf(1.f - c.A(k - 1), 1e-8f));
                Den_Dai += Den_Dwk * __expf(log_Tk) * c.A(k) / denom;
            }
        } else {
            const float Den_Dwi = -(log_clamped_ratio(c.W(sample_i), p.wsum) + 1.f) / p.wsum + p.Den_Dwsum;
            Den_Dai = Den_Dwi * __expf(logT);
            for (int k = sample_i + 1; k < p.valid_n; ++k) {
                const float Den_Dwk = -(log_clamped_ratio(c.W(k), p.wsum) + 1.f) / p.wsum + p.Den_Dwsum;
                log_Tk += __logf(fmaxf(1.f - c.A(k - 1), 1e-8f));
                Den_Dai += Den_Dwk * __expf(log_Tk) * c.A(k) / denom;
            }
        }
        add += f.lambda_l_entropy * Den_Dai;
    }
    return add;
}

__device__ __forceinline__ void scatter8(float *__restrict__ grad, uint8_t *__restrict__ mask, const int *lk,
                                         const float *pos, float gval) {
    float w[8];
    corner_weights(pos, gval, w);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        atomicAdd(grad + lk[c], w[c]);
        if (mask) mask[lk[c]] = 1;
    }
}

// Everything of a REAL sample's backward that is scalar per ray (:2137-2448, lane-0 parts).
__device__ __forceinline__ void finish_real_bwd(const GridP &g, const asurf_opt_t &opt, const FusedP &f, const Pre &p, const CacheView &c,
                                const asurf_grads_t &grads, Lane &L, float &accum, float total_color, float gx, float gy,
                                float gz) {
    const float pos[3] = {L.px, L.py, L.pz};
    accum -= L.weight * total_color;
    float curr_grad_rwalpha = accum / fminf(L.rwalpha - 1.f, -1e-8f) + total_color * __expf(L.logT);
    curr_grad_rwalpha += extra_grad_rwalpha_w(f, p, c, L.sample_i, L.logT, L.rwalpha);
    L.logT -= L.pcnt;
    if (f.lambda_l_dist_a > 0.f) {
        float a = 0.f;
        for (int j = 0; j < c.n; ++j) a += c.sa[j] * fabsf(c.T(L.sample_i) - c.st[j]);
        curr_grad_rwalpha += f.lambda_l_dist_a * a;
    }
    if (f.lambda_l_entropy_a > 0.f) {
        const float Den_Dai = -(log_clamped_ratio(c.A(L.sample_i), p.asum) + 1.f) / p.asum;
        curr_grad_rwalpha += f.lambda_l_entropy_a * (Den_Dai + p.Den_Dasum);
    }
    if ((f.sparsity_loss > 0.f) && (L.raw_alpha > 0.f)) {
        const float _1_a = fmaxf(1.f - L.alpha, 1e-8f);
        const double m = fmin((double)(_1_a * __logf(_1_a)), -1e-8);
        curr_grad_rwalpha = (float)((double)curr_grad_rwalpha +
                                    (double)(-f.sparsity_loss) * (1.0 / m) * (double)(1.f - L.weight / p.wsum));
    }
    if (f.lambda_inwards_norm_loss > 0.f) {
        float sg[3];
        field_grad8(L.sf, pos, sg);
        const float surf_n = fmaxf(sqrtf(sg[0] * sg[0] + sg[1] * sg[1] + sg[2] * sg[2]), 1e-8f);
        const float nd = (-sg[0] / surf_n) * L.dx + (-sg[1] / surf_n) * L.dy + (-sg[2] / surf_n) * L.dz;
        if (nd > 0.f) curr_grad_rwalpha += f.lambda_inwards_norm_loss * (nd * nd);
    }
    const float curr_grad_alpha = curr_grad_rwalpha * L.trunc_rw_;
    const float act_g = alpha_act_grad(L.alpha, opt.alpha_activation_type);
    float curr_grad_raw_alpha = curr_grad_alpha * act_g;
    scatter8(grads.grad_density, grads.mask, L.lk, pos, curr_grad_raw_alpha);
    if ((f.lambda_l_di > 0.f) && (L.alpha < f.l_di_alpha_thresh)) curr_grad_raw_alpha += f.lambda_l_di * -1.f * act_g;
    float gxyz[3] = {gx, gy, gz};
    trilerp8_pos_grad(L.dn, pos, curr_grad_raw_alpha, gxyz);
    float grad_st = gxyz[0] * L.dx + gxyz[1] * L.dy + gxyz[2] * L.dz;
    if ((f.lambda_l_dist > 0.f) || (f.lambda_l_dist_a > 0.f)) {
        float gw = 0.f, ga = 0.f;
        const float ti = c.T(L.sample_i);
        for (int j = 0; j < f.M; ++j) {
            const float tj = c.T(j);
            const float sgn = (ti > tj) ? 1.f : ((ti < tj) ? -1.f : 0.f);
            gw += sgn * c.W(L.sample_i) * c.W(j);
            ga += sgn * c.A(L.sample_i) * c.A(j);
        }
        grad_st += f.lambda_l_dist * gw + f.lambda_l_dist_a * ga;
    }
    if (f.lambda_l_samp_dist > 0.f) {
        const float ti = c.T(L.sample_i);
        const float sgn = (p.t_mean > ti) ? 1.f : ((p.t_mean < ti) ? -1.f : 0.f);
        grad_st += f.lambda_l_samp_dist * (p.Dmeant_sign * c.A(L.sample_i) / p.wsum + sgn * (-1.f));
    }
    if ((f.lambda_conv_mode_samp > 0.f) && (L.trunc_rw_ > opt.trunc_vol_weight_min)) {
        const float ti = c.T(L.sample_i), tm = c.T(p.max_id);
        grad_st += f.lambda_conv_mode_samp * ((ti > tm) ? 1.f : ((ti < tm) ? -1.f : 0.f));
    }
    if (grads.grad_surface) {
        float grad_fs[4] = {grad_st, grad_st, grad_st, grad_st};
        root_grad(L.root_type, L.st_id, L.fs, grad_fs);
        const float nno_f[3] = {(float)L.nno[0], (float)L.nno[1], (float)L.nno[2]};
        const float dir[3] = {L.dx, L.dy, L.dz};
        float gs[8];
        cubic_to_corner_grad(nno_f, dir, grad_fs, gs);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            atomicAdd(grads.grad_surface + L.lk[k], gs[k]);
            if (grads.mask) grads.mask[L.lk[k]] = 1;
        }
        if (L.alpha < f.surf_sparse_alpha_thresh)
            scatter8(grads.grad_surface, grads.mask, L.lk, pos, f.lambda_inplace_surf_sparse);
    }
    if (L.sample_i < f.M - 1) L.sample_i += 1;
}

// Scalar part of a FAKE sample's backward (:2590-2866).
__device__ __forceinline__ void finish_fake_bwd(const GridP &g, const asurf_opt_t &opt, const FusedP &f, const Pre &p, const CacheView &c,
                                const asurf_grads_t &grads, Lane &L, float &accum, float total_color) {
    const float pos[3] = {L.px, L.py, L.pz};
    accum -= L.weight * total_color;
    float curr_grad_rwalpha = accum / fminf(L.rwalpha - 1.f, -1e-8f) + total_color * __expf(L.logT);
    if (opt.fake_sample_l_dist) curr_grad_rwalpha += extra_grad_rwalpha_w(f, p, c, L.sample_i, L.logT, L.rwalpha);
    L.logT -= L.pcnt;
    if (opt.fake_sample_l_dist) {
        if (f.lambda_l_dist_a > 0.f) {
            float a = 0.f;
            for (int j = 0; j < c.n; ++j) a += c.sa[j] * fabsf(c.T(L.sample_i) - c.st[j]);
            curr_grad_rwalpha += f.lambda_l_dist_a * a;
        }
        if (f.lambda_l_entropy_a > 0.f) {
            const float Den_Dai = -(log_clamped_ratio(c.A(L.sample_i), p.asum) + 1.f) / p.asum;
            curr_grad_rwalpha += f.lambda_l_entropy_a * (Den_Dai + p.Den_Dasum);
        }
        if (L.sample_i < f.M - 1) L.sample_i += 1;
    }
    if ((f.sparsity_loss > 0.f) && (L.raw_alpha > 0.f)) {
        const float _1_a = fmaxf(1.f - L.rwalpha, 1e-8f);
        const double m = fmin((double)(_1_a * __logf(_1_a)), -1e-8);
        curr_grad_rwalpha = (float)((double)curr_grad_rwalpha +
                                    (double)(-f.sparsity_loss) * (1.0 / m) * (double)(1.f - L.weight / p.wsum));
    }
    if (f.lambda_inwards_norm_loss > 0.f) {
        float sg[3];
        field_grad8(L.sf, pos, sg);
        const float surf_n = fmaxf(sqrtf(sg[0] * sg[0] + sg[1] * sg[1] + sg[2] * sg[2]), 1e-8f);
        const float nd = (-sg[0] / surf_n) * L.dx + (-sg[1] / surf_n) * L.dy + (-sg[2] / surf_n) * L.dz;
        if (nd > 0.f) curr_grad_rwalpha += f.lambda_inwards_norm_loss * (nd * nd);
    }
    const float curr_grad_alpha = curr_grad_rwalpha * L.reweight * L.trunc_rw_;
    const float curr_grad_raw_alpha = curr_grad_alpha * alpha_act_grad(L.alpha, opt.alpha_activation_type);
    scatter8(grads.grad_density, grads.mask, L.lk, pos, curr_grad_raw_alpha);
    const float std_ = g.fake_sample_std;
    const float grad_fake_dist = curr_grad_rwalpha * (-L.alpha * L.trunc_rw_ * L.fake_dist * L.reweight / (std_ * std_));
    if (grads.grad_surface) {
        float gns[8];
        corner_weights(pos, grad_fake_dist, gns);
        float gs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (!opt.fake_sample_normalize_surf) {
#pragma unroll
            for (int k = 0; k < 8; ++k) gs[k] = gns[k];
        } else {
            const double sd3 = L.surf_std * L.surf_std * L.surf_std;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const double s_ks = (double)L.sf[ks];
#pragma unroll
                for (int kn = 0; kn < 8; ++kn) {
                    const double s_kn = (double)L.sf[kn];
                    if (ks == kn)
                        gs[ks] += (float)(gns[kn] * (s_ks * (L.surf_miu - s_ks) / 8.f / sd3 + 1.f / L.surf_std));
                    else
                        gs[ks] += (float)(gns[kn] * (s_kn * (L.surf_miu - s_ks) / 8.f / sd3));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            atomicAdd(grads.grad_surface + L.lk[k], gs[k]);
            if (grads.mask) grads.mask[L.lk[k]] = 1;
        }
    }
    if (grads.grad_fake_sample_std) {
        const float grad_std = curr_grad_rwalpha * L.alpha * (L.fake_dist * L.fake_dist) * L.reweight * L.trunc_rw_ /
                               (std_ * std_ * std_);
        atomicAdd(grads.grad_fake_sample_std, grad_std);
    }
}

// -------------------------------------------------------------------------------------------------------------
template <bool BWD, bool DEBUG>
__global__ void __launch_bounds__(CTA_THREADS)
surf_trav_kernel(const GridP g, const asurf_opt_t opt, const float *__restrict__ origins,
                 const float *__restrict__ dirs, const int64_t Q, float *__restrict__ rgb_out,
                 const float *__restrict__ grad_in, const float *__restrict__ color_cache, const FusedP f,
                 const CacheP cache, const asurf_grads_t grads, const DebugP dbg,
                 unsigned long long *__restrict__ ray_counter, const PreP pre) {
    __shared__ float s_sph[CTA_WARPS][32][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = g.sh_dim, bd = g.basis_dim;
    const int M = f.M;

    Lane L;
    L.state = ST_IDLE;
    L.ray_done = false;
    L.ray_id = -1;
    float out0 = 0.f, out1 = 0.f, out2 = 0.f;  // FWD: colour; BWD: dL/dRGB
    float accum = 0.f;
    Pre lossc;
    CacheView cv;
    cv.sa = cv.sw = cv.st = nullptr;
    cv.n = 0;
    Counters cnt = {0, 0, 0, 0, 0};
    bool rays_left = true;   // warp-uniform
    // rays to serve: the compact list of the pre-march, or all Q rays when there was no pre-march
    const int64_t n_serve = pre.enabled ? (int64_t)*pre.n_rays : Q;
    // Rays in flight per warp: with few rays to serve (the long rays left over by the wavefront path) they are spread
    // over all warps instead of being packed 32 to a warp, where they would serialise each other's divergent work.
    const int64_t total_warps = (int64_t)gridDim.x * CTA_WARPS;
    const int quota = (int)min((int64_t)32, max((int64_t)1, (n_serve + total_warps - 1) / total_warps));

    // Persistent warp: lanes pull rays from a global counter as they become free, then the warp repeatedly runs the
    // kind of work most of its lanes are waiting for (march step / voxel work / sample shading / ray set-up).
    for (;;) {
        const unsigned m_idle = __ballot_sync(FULL, L.state == ST_IDLE);
        const unsigned m_march = __ballot_sync(FULL, L.state == ST_MARCH);
        const unsigned m_vox = __ballot_sync(FULL, L.state == ST_VOXEL);
        const unsigned m_samp = __ballot_sync(FULL, L.state == ST_SAMPLE);
        const int can_take = rays_left ? max(0, min(__popc(m_idle), quota - (32 - __popc(m_idle)))) : 0;
        const int n_idle = can_take;
        const int n_march = __popc(m_march), n_vox = __popc(m_vox), n_samp = __popc(m_samp);
        if ((n_idle | n_march | n_vox | n_samp) == 0) break;

        if (n_march > 0 && n_march >= n_vox && n_march >= n_samp && n_march >= n_idle) {
            // ---------------- march: a few generalized DDA steps ----------------
#pragma unroll 1
            for (int it = 0; it < 4; ++it) {
                if (L.state == ST_MARCH) march_step<BWD, DEBUG, false>(g, opt, L, cnt);
            }
        } else if (n_idle > 0 && n_idle >= n_vox && n_idle >= n_samp) {
            // ---------------- ray set-up for the free lanes ----------------
            unsigned long long base = 0;
            if (lane == __ffs(m_idle) - 1) base = atomicAdd(ray_counter, (unsigned long long)can_take);
            base = __shfl_sync(FULL, base, __ffs(m_idle) - 1);
            const int rank = __popc(m_idle & ((1u << lane) - 1u));
            const bool take = (L.state == ST_IDLE) && (rank < can_take);
            const int64_t serve_id = (int64_t)base + rank;
            if (__any_sync(FULL, take && (serve_id >= n_serve))) rays_left = false;
            if (take && serve_id < n_serve) {
                const int64_t ray_id = pre.enabled ? (int64_t)__ldg(pre.rays + serve_id) : serve_id;
                L.ray_id = ray_id;
                L.logT = 0.f;
                L.intersect_i = -1;
                L.sample_i = 0;
                L.n_hits = 0;
                out0 = out1 = out2 = 0.f;
                L.ox = origins[ray_id * 3 + 0]; L.oy = origins[ray_id * 3 + 1]; L.oz = origins[ray_id * 3 + 2];
                L.dx = dirs[ray_id * 3 + 0]; L.dy = dirs[ray_id * 3 + 1]; L.dz = dirs[ray_id * 3 + 2];
                eval_sh(bd, L.dx, L.dy, L.dz, s_sph[warp][lane]);  // world-space direction (:3165-3171)
                float world_step;
                ray_bounds(g, opt, L, world_step);
                if (DEBUG && dbg.xf) {
                    float *x = dbg.xf + ray_id * 9;
                    x[0] = L.ox; x[1] = L.oy; x[2] = L.oz; x[3] = L.dx; x[4] = L.dy; x[5] = L.dz;
                    x[6] = L.tmin; x[7] = L.tmax; x[8] = world_step;
                }
                if (BWD) {
                    if (f.grad_is_rgb) {  // fused: dL/dRGB from the L2 / L1 mix (:3306-3316)
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const float resid = color_cache[ray_id * 3 + i] - grad_in[ray_id * 3 + i];
                            float gi = resid * f.norm_l2 * f.lambda_l2;
                            gi += (resid > 0.f) ? (f.norm_l1 * f.lambda_l1) : (-f.norm_l1 * f.lambda_l1);
                            if (i == 0) out0 = gi; else if (i == 1) out1 = gi; else out2 = gi;
                        }
                    } else {
                        out0 = grad_in[ray_id * 3 + 0]; out1 = grad_in[ray_id * 3 + 1]; out2 = grad_in[ray_id * 3 + 2];
                    }
                    cv.n = 0;
                    if (M > 0) {
                        cv.sa = cache.sa + ray_id * M; cv.sw = cache.sw + ray_id * M; cv.st = cache.st + ray_id * M;
                        cv.n = cache.n[ray_id];
                    }
                    fused_preamble(cv, lossc);
                    accum = fmaf(color_cache[ray_id * 3 + 0], out0,
                                 fmaf(color_cache[ray_id * 3 + 1], out1, color_cache[ray_id * 3 + 2] * out2));
                }
                if (L.tmin > L.tmax) {
                    L.ray_done = true;   // misses the grid: background colour / no gradient (:59-65, :1811-1816)
                } else if (pre.enabled) {
                    L.list_code = __ldg(pre.code + ray_id);
                    L.list_cnt = BWD ? ((L.list_code >> CODE_CNT_BITS) & CODE_CNT_MASK) : (L.list_code & CODE_CNT_MASK);
                    L.list_pos = 0;
                    L.in_list = true;
                    next_from_list<BWD>(g, pre, L);
                } else {
                    dda_init(g, L);   // -> ST_MARCH
                }
            }
            __syncwarp();
        } else if (n_vox > 0 && n_vox >= n_samp) {
            // ---------------- voxel work ----------------
            if (L.state == ST_VOXEL) voxel_advance<BWD, DEBUG>(g, opt, L, cache, M, cnt, dbg, pre);
        } else {
            // ---------------- sample shading: the warp serves its pending lanes one after the other ----------------
            unsigned pend = m_samp;
            const bool have = (L.state == ST_SAMPLE);
            float tot_color = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
            while (pend) {
                const int src = __ffs(pend) - 1;
                pend &= pend - 1;
                int lk[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) lk[c] = __shfl_sync(FULL, L.lk[c], src);
                float pos[3];
                pos[0] = __shfl_sync(FULL, L.px, src);
                pos[1] = __shfl_sync(FULL, L.py, src);
                pos[2] = __shfl_sync(FULL, L.pz, src);
                float v[8];
                float lane_color = 0.f;
                const int kb = (lane < D) ? (lane % bd) : 0;
                float sph = 0.f;
                if (lane < D) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = __ldg(g.sh + (int64_t)lk[c] * D + lane);
                    sph = s_sph[warp][src][kb];
                    lane_color = trilerp8(v, pos) * sph;
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = 0.f;
                }
                const float seg = segment_sum(lane_color, (lane < D) ? kb : 32, bd);
                const float c0 = __shfl_sync(FULL, seg, 0);
                const float c1 = __shfl_sync(FULL, seg, bd);
                const float c2 = __shfl_sync(FULL, seg, 2 * bd);
                if (!BWD) {
                    if (lane == src) {
                        out0 += L.weight * fmaxf(c0 + 0.5f, 0.f);
                        out1 += L.weight * fmaxf(c1 + 0.5f, 0.f);
                        out2 += L.weight * fmaxf(c2 + 0.5f, 0.f);
                    }
                } else {
                    const float g0 = __shfl_sync(FULL, out0, src);
                    const float g1 = __shfl_sync(FULL, out1, src);
                    const float g2 = __shfl_sync(FULL, out2, src);
                    const float weight = __shfl_sync(FULL, L.weight, src);
                    const float l0 = c0 + 0.5f, l1 = c1 + 0.5f, l2 = c2 + 0.5f;
                    const float t0 = fmaxf(l0, 0.f), t1 = fmaxf(l1, 0.f), t2 = fmaxf(l2, 0.f);
                    float total_color = t0 * g0;  // shuffle order of the reference (:2112-2115): (c0 + c2) + c1
                    total_color += t2 * g2;
                    total_color += t1 * g1;
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
                    float sx = 0.f, sy = 0.f, sz = 0.f;
                    if (!opt.no_surf_grad_from_sh) {
                        sx = __shfl_sync(FULL, warp_sum_down(gacc[0], lane), 0);
                        sy = __shfl_sync(FULL, warp_sum_down(gacc[1], lane), 0);
                        sz = __shfl_sync(FULL, warp_sum_down(gacc[2], lane), 0);
                    }
                    if (lane == src) { tot_color = total_color; gx = sx; gy = sy; gz = sz; }
                }
            }
            if (have) {
                if (BWD) {
                    if (!L.fake) finish_real_bwd(g, opt, f, lossc, cv, grads, L, accum, tot_color, gx, gy, gz);
                    else finish_fake_bwd(g, opt, f, lossc, cv, grads, L, accum, tot_color);
                }
                L.state = ST_VOXEL;   // resume behind the sample
            }
        }

        // ---------------- rays that just ended ----------------
        if (L.ray_done) {
            L.ray_done = false;
            L.state = ST_IDLE;
            const int64_t ray_id = L.ray_id;
            if (!BWD) {
                // a miss keeps logT == 0: pure background
                const float bg = g.has_bg ? 0.f : __expf(L.logT) * opt.background_brightness;
                rgb_out[ray_id * 3 + 0] = out0 + bg;
                rgb_out[ray_id * 3 + 1] = out1 + bg;
                rgb_out[ray_id * 3 + 2] = out2 + bg;
                if (M > 0) cache.n[ray_id] = L.sample_i;
            } else if (g.bg_accum) {
                g.bg_accum[ray_id] = accum - g.bg_beta;
            }
            if (g.bg_lt) g.bg_lt[ray_id] = L.logT;
            if (DEBUG && dbg.hit_count) dbg.hit_count[ray_id] = L.n_hits;
        }
    }
    if (DEBUG && dbg.stats) {
        atomicAdd(dbg.stats + 0, cnt.steps);
        atomicAdd(dbg.stats + 1, cnt.skips);
        atomicAdd(dbg.stats + 2, cnt.linked);
        atomicAdd(dbg.stats + 3, cnt.active);
        atomicAdd(dbg.stats + 4, cnt.samples);
    }
}

#include "surf_wave.inl"
#include "surf_scalar.inl"

}  // namespace
}  // namespace asurf

using namespace asurf;

namespace {

Workspace g_ws_wave, g_ws_seg;
int g_wave_enabled = 1;  // asurf_debug_set_wave
int g_seg_enabled = 1;   // asurf_debug_set_seg

Workspace g_ws_accel, g_ws_work, g_ws_cache, g_ws_dbg, g_ws_ctr, g_ws_pre;

int g_skip_enabled = 1;  // asurf_debug_set_skip

// per-kernel timing ring (asurf_profile_*)
struct ProfRing {
    cudaEvent_t *ev = nullptr;  // PROF_EV events per call: start, work pyramid, pre-march, forward, backward
    int cap = 0, n = 0;
} g_prof;
constexpr int PROF_EV = 5;

int make_grid(const asurf_grid_t *grid, const asurf_opt_t *opt, bool need_work, cudaStream_t st, GridP &g) {
    ASURF_REQUIRE(grid, ASURF_E_INVALID, "surf_trav: null grid");
    ASURF_REQUIRE(grid->links && grid->density && grid->sh, ASURF_E_INVALID, "surf_trav: null grid tensor");
    ASURF_REQUIRE(grid->surface && grid->level_set, ASURF_E_INVALID,
                  "surf_trav: the grid has no surface / level set data (surface_type none)");
    ASURF_REQUIRE(grid->size[0] >= 2 && grid->size[1] >= 2 && grid->size[2] >= 2, ASURF_E_INVALID,
                  "surf_trav: grid smaller than 2^3");
    ASURF_REQUIRE(grid->basis_dim == 1 || grid->basis_dim == 4 || grid->basis_dim == 9, ASURF_E_UNSUPPORTED,
                  "surf_trav: basis_dim %d not supported (SH with 1, 4 or 9 functions)", grid->basis_dim);
    ASURF_REQUIRE(grid->sh_dim == 3 * grid->basis_dim, ASURF_E_INVALID, "surf_trav: sh_dim must be 3*basis_dim");
    g.links = grid->links;
    g.density = grid->density;
    g.surface = grid->surface;
    g.sh = grid->sh;
    g.level_set = grid->level_set;
    for (int i = 0; i < 3; ++i) {
        g.size[i] = grid->size[i];
        g.offset[i] = grid->offset[i];
        g.scaling[i] = grid->scaling[i];
    }
    g.level_set_num = grid->level_set_num;
    g.basis_dim = grid->basis_dim;
    g.sh_dim = grid->sh_dim;
    g.fake_sample_std = grid->fake_sample_std;
    g.trunc_a = grid->truncated_vol_render_a;
    g.has_bg = grid_has_background(grid) ? 1 : 0;
    g.bg_lt = g.bg_accum = nullptr;
    g.bg_beta = 0.f;
    AccelLayout lay(grid->size);
    g.ab1 = lay.b[0][1];
    g.ab2 = lay.b[0][2];
    g.lay = lay;
    g.use_skip = g_skip_enabled;
    if (grid->accel) {
        g.accel = grid->accel;
    } else {
        int rc = g_ws_accel.reserve((size_t)asurf_accel_words(grid->size) * sizeof(uint64_t));
        if (rc) return rc;
        rc = asurf_accel_build(grid->links, grid->size, (uint64_t *)g_ws_accel.ptr, st);
        if (rc) return rc;
        g.accel = (const uint64_t *)g_ws_accel.ptr;
    }
    g.work = grid->work;
    if (need_work && !grid->work) {
        asurf_grid_t tmp = *grid;
        tmp.accel = g.accel;
        const uint64_t *w = nullptr;
        int rc = work_pyramid_for_call(&tmp, opt, st, &w);
        if (rc) return rc;
        g.work = w;
    }
    return 0;
}

int check_rays(const asurf_rays_t *rays, const asurf_opt_t *opt) {
    ASURF_REQUIRE(rays && opt, ASURF_E_INVALID, "surf_trav: null rays / options");
    ASURF_REQUIRE(rays->n_rays >= 0, ASURF_E_INVALID, "surf_trav: negative ray count");
    ASURF_REQUIRE(rays->n_rays == 0 || (rays->origins && rays->dirs), ASURF_E_INVALID, "surf_trav: null ray tensor");
    return 0;
}

// Persistent grid: about RAYS_PER_LANE rays per lane so that lanes can pick up new rays while their neighbours are
// still marching, capped at a few resident CTAs per SM for very large batches.
constexpr int RAYS_PER_LANE = 2;
inline int n_ctas(int64_t Q) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int64_t want = (Q + (int64_t)CTA_THREADS * RAYS_PER_LANE - 1) / ((int64_t)CTA_THREADS * RAYS_PER_LANE);
    const int64_t cap = (int64_t)sms * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// zeroed counters for the next launches on `st`: [0] forward ray fetch, [1] backward ray fetch, [2] long-ray list length,
// [3] short-ray list length, [4] item queue length, [5] hit queue length, [6] fine-item queue length of the pre-march
int ray_counters(cudaStream_t st, unsigned long long **ctr) {
    int rc = g_ws_ctr.reserve(8 * sizeof(unsigned long long));
    if (rc) return rc;
    ASURF_CUDA(cudaMemsetAsync(g_ws_ctr.ptr, 0, 8 * sizeof(unsigned long long), st));
    *ctr = (unsigned long long *)g_ws_ctr.ptr;
    return 0;
}

// asynchronous read-back of the item count of a call (see the capacity rule in premarch())
struct ItemProbe {
    unsigned long long *host = nullptr;   // pinned copy of the 8 call counters (ray_counters)
    cudaEvent_t ev = nullptr;
    bool pending = false;
    int64_t q = 0;
    double per_ray = 0.0;        // listed voxels (items of the wavefront queue) per ray of the batch
    double fine_per_ray = 0.0;   // (ray, non-empty 16^3 block) items of the two-level pre-march per ray
    double long_frac = 0.0;      // share of the rays with work that went to the persistent kernels
} g_item_probe;

void item_probe_record(unsigned long long *ctr, int64_t Q, cudaStream_t st) {
    if (!g_item_probe.host) {
        if (cudaMallocHost((void **)&g_item_probe.host, 8 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaEventCreateWithFlags(&g_item_probe.ev, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            g_item_probe.host = nullptr;
            return;
        }
    }
    if (g_item_probe.pending) return;   // the previous probe has not been consumed yet: its buffer is still in flight
    if (cudaMemcpyAsync(g_item_probe.host, ctr, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    cudaEventRecord(g_item_probe.ev, st);
    g_item_probe.pending = true;
    g_item_probe.q = Q;
}

// Runs the pre-march for this call (forward output / cache counts optional) and returns its record.
int premarch(const GridP &g, const asurf_opt_t *opt, const asurf_rays_t *rays, unsigned long long *ctr, float *rgb_out,
             int *cache_n, cudaStream_t st, PreP &pre) {
    pre = PreP();
    if (!g_skip_enabled || g.size[0] > 1024 || g.size[1] > 1024 || g.size[2] > 1024) return 0;   // shading kernels march
    const int64_t Q = rays->n_rays;
    // list capacity per ray: as long as the buffers stay moderate (a grazing ray through a thin sheet lists ~100 voxels)
    if (g_item_probe.host && g_item_probe.pending && cudaEventQuery(g_item_probe.ev) == cudaSuccess) {
        g_item_probe.pending = false;
        const double q = (double)(g_item_probe.q > 0 ? g_item_probe.q : 1);
        g_item_probe.per_ray = (double)g_item_probe.host[4] / q;
        g_item_probe.fine_per_ray = (double)g_item_probe.host[6] / q;
        const double with_work = (double)(g_item_probe.host[2] + g_item_probe.host[3]);
        g_item_probe.long_frac = with_work > 0 ? (double)g_item_probe.host[2] / with_work : 0.0;
    }
    // a grid with a crossing in most voxels (G*: ~90 listed voxels per ray against 3 on the training grids, or most rays
    // overflowing the short lists): long lists
    const bool dense_crossings = g_item_probe.per_ray > 4.5 || g_item_probe.long_frac > 0.1;
    int K = dense_crossings ? PRE_K_MAX : PRE_K_DEFAULT;
    while (K > 16 && (int64_t)Q * K > ((int64_t)1 << 27)) K >>= 1;
    // Item queue capacity: 6 listed voxels per ray serve a thin level-set sheet (1.3 per ray on the training grids); a grid
    // with a crossing in most voxels (G*) lists ~100 per ray.  The library learns it from the previous call: the item
    // count is read back asynchronously (pinned host word + event, never waited for) and the next call on a batch of
    // similar size reserves 1.25 x that, up to the lists' own capacity.  A call that overflows is still correct -- the rays
    // that do not fit are rendered by the persistent kernels -- only slower.
    int64_t item_cap = Q * 6 + 4096;
    if (dense_crossings) {
        int64_t want = (int64_t)(1.25 * g_item_probe.per_ray * (double)Q) + 4096;
        const int64_t hard = (int64_t)Q * K < ((int64_t)24 << 20) ? (int64_t)Q * K : ((int64_t)24 << 20);
        if (want > hard) want = hard;
        if (want > item_cap) item_cap = want;
    }
    const size_t per = (size_t)Q * sizeof(int32_t);
    int rc = g_ws_pre.reserve(per * (K + 6) + (size_t)item_cap * sizeof(int32_t));
    if (rc) return rc;
    char *base = (char *)g_ws_pre.ptr;
    pre.cells = (int32_t *)base;
    pre.K = K;
    pre.code = (int32_t *)(base + per * K);
    pre.cont_t = (float *)(base + per * (K + 1));
    pre.cont_vox = (int32_t *)(base + per * (K + 2));
    pre.rays = (int32_t *)(base + per * (K + 3));
    pre.rays_short = (int32_t *)(base + per * (K + 4));
    pre.item_base = (int32_t *)(base + per * (K + 5));
    pre.itemq = (int32_t *)(base + per * (K + 6));
    pre.item_cap = item_cap;
    pre.n_rays = ctr + 2;
    pre.n_short = ctr + 3;
    pre.n_items = ctr + 4;
    pre.enabled = 1;
    pre.wave = (g_wave_enabled && g.level_set_num == 1) ? 1 : 0;
    if (g_seg_enabled && g.use_skip && Q >= 8192 && Q * CI_MAX < ((int64_t)1 << 31)) {
        // large batch: coarse (block jumps, thread per ray) -> fine (thread per non-empty 16^3 block crossed) -> merge
        TwoP tw;
        tw.fq_cap = Q * 8 + 4096;
        if (g_item_probe.fine_per_ray > 6.0) {   // same feedback for the fine-item queue
            int64_t want = (int64_t)(1.25 * g_item_probe.fine_per_ray * (double)Q) + 4096;
            if (want > Q * CI_MAX) want = Q * CI_MAX;
            if (want > tw.fq_cap) tw.fq_cap = want;
        }
        const size_t b_items = (size_t)Q * CI_MAX * 2 * sizeof(int32_t), b_q = (size_t)Q * sizeof(int32_t),
                     b_fq = (size_t)tw.fq_cap * sizeof(int32_t), b_cells = (size_t)tw.fq_cap * FI_K * sizeof(int32_t);
        rc = g_ws_seg.reserve(b_items + 2 * b_q + 2 * b_fq + b_cells);
        if (rc) return rc;
        char *sb = (char *)g_ws_seg.ptr;
        tw.ray_items = (int32_t *)sb;
        tw.ray_nitems = (int32_t *)(sb + b_items);
        tw.item_first = (int32_t *)(sb + b_items + b_q);
        tw.fq = (int32_t *)(sb + b_items + 2 * b_q);
        tw.fi_meta = (int32_t *)(sb + b_items + 2 * b_q + b_fq);
        tw.fi_cells = (int32_t *)(sb + b_items + 2 * b_q + 2 * b_fq);
        tw.n_fq = ctr + 6;
        premarch_coarse_kernel<<<(int)((Q + 63) / 64), 64, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, tw);
        premarch_fine_kernel<<<(int)((tw.fq_cap + 255) / 256), 256, 0, st>>>(g, *opt, rays->origins, rays->dirs, tw);
        premarch_merge_kernel<<<(int)((Q + 255) / 256), 256, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, pre, tw, rgb_out,
                                                                      cache_n);
        note_launches(3);
        item_probe_record(ctr, Q, st);
        return check_cuda(cudaGetLastError(), "two-level premarch launch");
    }
    premarch_kernel<<<(int)((Q + 255) / 256), 256, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, pre, rgb_out, cache_n);
    note_launches(1);
    item_probe_record(ctr, Q, st);
    return check_cuda(cudaGetLastError(), "premarch launch");
}

// ---- wavefront path: buffers for the item queue and the stage launches -----------------------------------
inline int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
inline int wave_grid(int64_t want_threads, int threads, int ctas_per_sm) {
    const int64_t want = (want_threads + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

int wave_buffers(const PreP &pre, int64_t Q, unsigned long long *ctr, WaveP &wv) {
    const size_t n = (size_t)pre.item_cap;
    const size_t b_items = n * sizeof(ItemRec), b_hits = n * WAVE_ENT * sizeof(HitRec), b_q = n * WAVE_ENT * sizeof(int32_t);
    const size_t b_pre = (size_t)Q * sizeof(Pre), b_mask = (size_t)Q * mask_words(pre.K) * sizeof(uint32_t);
    int rc = g_ws_wave.reserve(b_items + b_hits + b_q + b_pre + b_mask);
    if (rc) return rc;
    char *base = (char *)g_ws_wave.ptr;
    wv.items = (ItemRec *)base;
    wv.hits = (HitRec *)(base + b_items);
    wv.hitq = (int32_t *)(base + b_items + b_hits);
    wv.ray_pre = (Pre *)(base + b_items + b_hits + b_q);
    wv.ray_mask = (uint32_t *)(base + b_items + b_hits + b_q + b_pre);
    wv.n_hits = ctr + 5;
    return 0;
}

// eval -> colour -> composite for the short rays (rgb_out may be NULL: backward-only rematerialisation)
int wave_forward(const GridP &g, const asurf_opt_t *opt, const asurf_rays_t *rays, const PreP &pre, const WaveP &wv,
                 const CacheP &cache, int M, float *rgb_out, const float *grad_in, const float *color_cache,
                 const FusedP &f, cudaStream_t st) {
    const int64_t Q = rays->n_rays;
    FusedP nof = {};
    asurf_grads_t nog = {};
    ASURF_CUDA(cudaMemsetAsync(wv.ray_mask, 0, (size_t)Q * mask_words(pre.K) * sizeof(uint32_t), st));
    wave_eval_kernel<<<wave_grid(Q * 2, 128, 16), 128, 0, st>>>(g, *opt, rays->origins, rays->dirs, pre, wv);
    wave_wide_kernel<false><<<wave_grid(Q * 32, 128, 16), 128, 0, st>>>(g, *opt, rays->dirs, pre, wv, nullptr, nullptr, nof, nog);
    wave_composite_kernel<<<wave_grid(Q, 128, 8), 128, 0, st>>>(g, *opt, pre, wv, cache, M, rgb_out, grad_in, color_cache, f);
    note_launches(3);
    return check_cuda(cudaGetLastError(), "wavefront forward launch");
}

int wave_backward(const GridP &g, const asurf_opt_t *opt, const asurf_rays_t *rays, const PreP &pre, const WaveP &wv,
                  const float *grad_in, const float *color_cache, const FusedP &f, const CacheP &cache,
                  const asurf_grads_t &grads, cudaStream_t st) {
    const int64_t Q = rays->n_rays;
    wave_wide_kernel<true><<<wave_grid(Q * 32, 128, 16), 128, 0, st>>>(g, *opt, rays->dirs, pre, wv, grad_in, color_cache, f, grads);
    wave_bwd_hit_kernel<<<wave_grid(Q, 128, 8), 128, 0, st>>>(g, *opt, rays->origins, rays->dirs, pre, wv, grad_in,
                                                               color_cache, f, cache, grads);
    note_launches(2);
    return check_cuda(cudaGetLastError(), "wavefront backward launch");
}

}  // namespace

extern "C" int asurf_surf_trav_forward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                       float *rgb_out, asurf_stats_t *stats_dev, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rays(rays, opt);
    if (rc) return rc;
    if (rays->n_rays == 0) return 0;
    ASURF_REQUIRE(rgb_out, ASURF_E_INVALID, "surf_trav_forward: null output");
    GridP g;
    rc = make_grid(grid, opt, stats_dev == nullptr, st, g);
    if (rc) return rc;
    FusedP f = {};
    CacheP cache = {};
    asurf_grads_t grads = {};
    DebugP dbg = {};
    unsigned long long *ctr = nullptr;
    rc = ray_counters(st, &ctr);
    if (rc) return rc;
    PreP pre = PreP();
    if (g.has_bg && !stats_dev) {
        float *acc = nullptr;
        rc = bg_state_reserve(rays->n_rays, &g.bg_lt, &acc);
        if (rc) return rc;
    }
    if (stats_dev) {
        dbg.stats = (unsigned long long *)stats_dev;
        g.use_skip = 0;   // count every voxel of the reference DDA
        surf_trav_kernel<false, true><<<n_ctas(rays->n_rays), CTA_THREADS, 0, st>>>(
            g, *opt, rays->origins, rays->dirs, rays->n_rays, rgb_out, nullptr, nullptr, f, cache, grads, dbg, ctr, pre);
    } else {
        rc = premarch(g, opt, rays, ctr, rgb_out, nullptr, st, pre);
        if (rc) return rc;
        if (pre.wave) {
            WaveP wv;
            rc = wave_buffers(pre, rays->n_rays, ctr, wv);
            if (rc) return rc;
            rc = wave_forward(g, opt, rays, pre, wv, cache, 0, rgb_out, nullptr, nullptr, f, st);
            if (rc) return rc;
        }
        surf_trav_kernel<false, false><<<n_ctas(rays->n_rays), CTA_THREADS, 0, st>>>(
            g, *opt, rays->origins, rays->dirs, rays->n_rays, rgb_out, nullptr, nullptr, f, cache, grads, dbg, ctr, pre);
    }
    note_launches(1);
    rc = check_cuda(cudaGetLastError(), "surf_trav_forward launch");
    if (rc) return rc;
    if (g.has_bg && !stats_dev) return asurf_msi_forward(grid, rays, opt, g.bg_lt, rgb_out, stream);   // :3639-3648
    return 0;
}

extern "C" int asurf_surf_trav_scalar(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, int32_t mode,
                                      float param, int32_t max_sample, float *out, float *out2, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rays(rays, opt);
    if (rc) return rc;
    ASURF_REQUIRE(mode >= ASURF_SCALAR_EXPECTED_TERM && mode <= ASURF_SCALAR_EXTRACT_PTS, ASURF_E_INVALID,
                  "surf_trav_scalar: unknown mode %d", mode);
    if (rays->n_rays == 0) return 0;
    ASURF_REQUIRE(out, ASURF_E_INVALID, "surf_trav_scalar: null output");
    if (mode == ASURF_SCALAR_EXTRACT_PTS) {
        ASURF_REQUIRE(out2 && max_sample >= 0, ASURF_E_INVALID, "surf_trav_scalar: extract_pts needs the alpha output and max_sample >= 0");
        if (max_sample == 0) return 0;
        ASURF_CUDA(cudaMemsetAsync(out, 0, (size_t)rays->n_rays * max_sample * sizeof(float), st));
        ASURF_CUDA(cudaMemsetAsync(out2, 0, (size_t)rays->n_rays * max_sample * sizeof(float), st));
    }
    GridP g;
    rc = make_grid(grid, opt, false, st, g);
    if (rc) return rc;
    // work pyramid of the predicate these renders use: 8 stored corners and a level set within their range -- no density
    // gate, no fake samples.  Kept in a cache slot of its own (an image is rendered in many 5000-ray calls; each call only
    // re-validates the pyramid against the data), so the training renderer's cached pyramid stays valid as well.
    {
        asurf_grid_t tmp = *grid;
        tmp.accel = g.accel;
        asurf_opt_t o2 = *opt;
        o2.sigma_thresh = -INFINITY;
        o2.surf_fake_sample = 0;
        const uint64_t *w = nullptr;
        rc = work_pyramid_for_call(&tmp, &o2, st, &w, 1);
        if (rc) return rc;
        g.work = w;
    }
    const int64_t Q = rays->n_rays;
    const int blocks = (int)((Q + 127) / 128);
    switch (mode) {
#define ASURF_SCALAR_CASE(M)                                                                                         \
    case M:                                                                                                          \
        scalar_render_kernel<M><<<blocks, 128, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, param, max_sample, out, \
                                                        out2);                                                       \
        break;
        ASURF_SCALAR_CASE(ASURF_SCALAR_EXPECTED_TERM)
        ASURF_SCALAR_CASE(ASURF_SCALAR_MODE_TERM)
        ASURF_SCALAR_CASE(ASURF_SCALAR_THRESH_DEPTH)
        ASURF_SCALAR_CASE(ASURF_SCALAR_THRESH_ALPHA)
        ASURF_SCALAR_CASE(ASURF_SCALAR_NORMAL)
        ASURF_SCALAR_CASE(ASURF_SCALAR_EXTRACT_PTS)
#undef ASURF_SCALAR_CASE
    }
    note_launches(1);
    return check_cuda(cudaGetLastError(), "surf_trav_scalar launch");
}

extern "C" int asurf_surf_trav_backward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                        const float *grad_out, const float *color_cache, const asurf_grads_t *grads,
                                        void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rays(rays, opt);
    if (rc) return rc;
    if (rays->n_rays == 0) return 0;
    ASURF_REQUIRE(grad_out && color_cache && grads, ASURF_E_INVALID, "surf_trav_backward: null tensor");
    ASURF_REQUIRE(grads->grad_density && grads->grad_sh, ASURF_E_INVALID, "surf_trav_backward: null gradient buffer");
    GridP g;
    rc = make_grid(grid, opt, true, st, g);
    if (rc) return rc;
    FusedP f = {};
    CacheP cache = {};
    DebugP dbg = {};
    unsigned long long *ctr = nullptr;
    rc = ray_counters(st, &ctr);
    if (rc) return rc;
    if (g.has_bg) {   // the stand-alone backward leaves its own log-transmittance and accum for the background (:3776-3796)
        ASURF_REQUIRE(grads->grad_background, ASURF_E_INVALID, "surf_trav_backward: the grid has a background but no gradient buffer for it");
        rc = bg_state_reserve(rays->n_rays, &g.bg_lt, &g.bg_accum);
        if (rc) return rc;
        ASURF_CUDA(cudaMemsetAsync(g.bg_accum, 0xFF, (size_t)rays->n_rays * sizeof(float), st));   // NaN: ray not visited
    }
    PreP pre;
    rc = premarch(g, opt, rays, ctr, nullptr, nullptr, st, pre);
    if (rc) return rc;
    if (pre.wave) {   // rematerialise the short rays' samples, then their backward
        WaveP wv;
        rc = wave_buffers(pre, rays->n_rays, ctr, wv);
        if (rc) return rc;
        rc = wave_forward(g, opt, rays, pre, wv, cache, 0, nullptr, grad_out, color_cache, f, st);
        if (rc) return rc;
        rc = wave_backward(g, opt, rays, pre, wv, grad_out, color_cache, f, cache, *grads, st);
        if (rc) return rc;
    }
    surf_trav_kernel<true, false><<<n_ctas(rays->n_rays), CTA_THREADS, 0, st>>>(
        g, *opt, rays->origins, rays->dirs, rays->n_rays, nullptr, grad_out, color_cache, f, cache, *grads, dbg, ctr + 1,
        pre);
    note_launches(1);
    rc = check_cuda(cudaGetLastError(), "surf_trav_backward launch");
    if (rc) return rc;
    if (g.has_bg)
        return asurf_msi_backward(grid, rays, opt, grad_out, color_cache, 0, 0, g.bg_lt, g.bg_accum, 0.f, 0.f, grads, stream);
    return 0;
}

extern "C" int asurf_surf_trav_fused(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                     const float *rgb_gt, const asurf_fused_t *fu, float *rgb_out,
                                     const asurf_grads_t *grads, asurf_stats_t *stats_dev, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_rays(rays, opt);
    if (rc) return rc;
    ASURF_REQUIRE(fu && grads, ASURF_E_INVALID, "surf_trav_fused: null argument");
    ASURF_REQUIRE(!(fu->fused_surf_norm_reg_scale > 0.f), ASURF_E_UNSUPPORTED,
                  "surf_trav_fused: fused_surf_norm_reg_scale > 0 is not supported (the reference asserts)");
    ASURF_REQUIRE(fu->l_dist_max_sample >= 0, ASURF_E_INVALID, "surf_trav_fused: negative l_dist_max_sample");
    const int64_t Q = rays->n_rays;
    if (Q == 0) return 0;
    ASURF_REQUIRE(rgb_gt && rgb_out, ASURF_E_INVALID, "surf_trav_fused: null colour tensor");
    ASURF_REQUIRE(grads->grad_density && grads->grad_sh, ASURF_E_INVALID, "surf_trav_fused: null gradient buffer");
    const bool prof = g_prof.cap > 0 && g_prof.n < g_prof.cap;
    cudaEvent_t *pe = prof ? g_prof.ev + PROF_EV * g_prof.n : nullptr;
    if (prof) cudaEventRecord(pe[0], st);
    GridP g;
    rc = make_grid(grid, opt, true, st, g);
    if (rc) return rc;
    if (prof) cudaEventRecord(pe[1], st);
    const int M = fu->l_dist_max_sample;
    CacheP cache = {};
    if (M > 0) {
        const size_t per = (size_t)Q * M * sizeof(float);
        rc = g_ws_cache.reserve(3 * per + (size_t)Q * sizeof(int));
        if (rc) return rc;
        char *base = (char *)g_ws_cache.ptr;
        cache.sa = (float *)base;
        cache.sw = (float *)(base + per);
        cache.st = (float *)(base + 2 * per);
        cache.n = (int *)(base + 3 * per);
    }
    const int64_t qn = fu->norm_rays > 0 ? fu->norm_rays : Q;
    const float Qf = (float)qn;
    FusedP f = {};
    f.sparsity_loss = fu->sparsity_loss;
    f.lambda_l2 = fu->lambda_l2;
    f.lambda_l1 = fu->lambda_l1;
    f.lambda_l_dist = fu->lambda_l_dist / Qf;
    f.lambda_l_entropy = fu->lambda_l_entropy / Qf;
    f.lambda_l_dist_a = fu->lambda_l_dist_a / Qf;
    f.lambda_l_entropy_a = fu->lambda_l_entropy_a / Qf;
    f.lambda_l_samp_dist = fu->lambda_l_samp_dist / Qf;
    f.lambda_l_di = fu->lambda_l_di;
    f.l_di_alpha_thresh = fu->l_di_alpha_thresh;
    f.surf_sparse_alpha_thresh = fu->surf_sparse_alpha_thresh;
    f.lambda_inplace_surf_sparse = fu->lambda_inplace_surf_sparse;
    f.lambda_inwards_norm_loss = fu->lambda_inwards_norm_loss;
    f.lambda_conv_mode_samp = fu->lambda_conv_mode_samp;
    f.norm_l2 = 2.f / (float)(3 * (int)qn);
    f.norm_l1 = 1.f / (float)(3 * (int)qn);
    f.no_norm_weight_l_entropy = fu->no_norm_weight_l_entropy;
    f.M = M;
    f.grad_is_rgb = 1;
    asurf_grads_t nog = {};
    DebugP dbg = {};
    FusedP ff = {};
    ff.M = M;
    unsigned long long *ctr = nullptr;
    rc = ray_counters(st, &ctr);
    if (rc) return rc;
    PreP pre = PreP(), nopre = PreP();
    WaveP wv = WaveP();
    float *bg_accum = nullptr;
    if (g.has_bg && !stats_dev) {   // forward state for the background pass; the backward adds its leftover accum (:3862-3940)
        ASURF_REQUIRE(grads->grad_background, ASURF_E_INVALID, "surf_trav_fused: the grid has a background but no gradient buffer for it");
        rc = bg_state_reserve(Q, &g.bg_lt, &bg_accum);
        if (rc) return rc;
        ASURF_CUDA(cudaMemsetAsync(bg_accum, 0xFF, (size_t)Q * sizeof(float), st));   // NaN: ray not visited by the backward
        g.bg_accum = bg_accum;       // (only backward code paths write it)
        g.bg_beta = fu->beta_loss / Qf;
    }
    if (!stats_dev) {
        rc = premarch(g, opt, rays, ctr, rgb_out, cache.n, st, pre);
        if (rc) return rc;
    }
    if (prof) cudaEventRecord(pe[2], st);
    if (stats_dev) {
        dbg.stats = (unsigned long long *)stats_dev;
        GridP gs = g;
        gs.use_skip = 0;
        surf_trav_kernel<false, true><<<n_ctas(Q), CTA_THREADS, 0, st>>>(gs, *opt, rays->origins, rays->dirs, Q, rgb_out,
                                                                          nullptr, nullptr, ff, cache, nog, dbg, ctr, nopre);
    } else {
        if (pre.wave) {
            rc = wave_buffers(pre, rays->n_rays, ctr, wv);
            if (rc) return rc;
            // with a background the loss gradient needs the FINAL colours (foreground + background, :3877-3890): the
            // backward recurrences of the composite stage are then walked in a second pass behind the background forward
            rc = wave_forward(g, opt, rays, pre, wv, cache, M, rgb_out, g.has_bg ? nullptr : rgb_gt, nullptr, f, st);
            if (rc) return rc;
        }
        surf_trav_kernel<false, false><<<n_ctas(Q), CTA_THREADS, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, rgb_out,
                                                                           nullptr, nullptr, ff, cache, nog, dbg, ctr, pre);
    }
    DebugP nodbg = {};
    float *bg_lt = g.bg_lt;
    if (g.has_bg && !stats_dev) {
        rc = check_cuda(cudaGetLastError(), "surf_trav_fused forward launch");
        if (rc) return rc;
        rc = asurf_msi_forward(grid, rays, opt, bg_lt, rgb_out, stream);
        if (rc) return rc;
        g.bg_lt = nullptr;           // the fused backward keeps the FORWARD's log-transmittance (:3934)
        if (pre.wave) {              // second pass of the wavefront stages: same entries, recurrences from the final colours
            ASURF_CUDA(cudaMemsetAsync(ctr + 5, 0, sizeof(unsigned long long), st));
            rc = wave_forward(g, opt, rays, pre, wv, cache, M, nullptr, rgb_gt, rgb_out, f, st);
            if (rc) return rc;
        }
    }
    if (prof) cudaEventRecord(pe[3], st);
    if (pre.wave) {
        rc = wave_backward(g, opt, rays, pre, wv, rgb_gt, rgb_out, f, cache, *grads, st);
        if (rc) return rc;
    }
    surf_trav_kernel<true, false><<<n_ctas(Q), CTA_THREADS, 0, st>>>(g, *opt, rays->origins, rays->dirs, Q, nullptr, rgb_gt,
                                                                      rgb_out, f, cache, *grads, nodbg, ctr + 1, pre);
    if (prof) {
        cudaEventRecord(pe[4], st);
        ++g_prof.n;
    }
    note_launches(2);
    rc = check_cuda(cudaGetLastError(), "surf_trav_fused launch");
    if (rc) return rc;
    if (g.has_bg && !stats_dev)
        return asurf_msi_backward(grid, rays, opt, rgb_gt, rgb_out, 1, qn, bg_lt, bg_accum, fu->beta_loss / Qf,
                                  fu->sparsity_loss, grads, stream);
    return 0;
}

static int debug_launch(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt, DebugP dbg,
                        cudaStream_t st) {
    int rc = check_rays(rays, opt);
    if (rc) return rc;
    if (rays->n_rays == 0) return 0;
    GridP g;
    rc = make_grid(grid, opt, false, st, g);
    if (rc) return rc;
    rc = g_ws_dbg.reserve((size_t)rays->n_rays * 3 * sizeof(float));
    if (rc) return rc;
    FusedP f = {};
    CacheP cache = {};
    asurf_grads_t grads = {};
    unsigned long long *ctr = nullptr;
    rc = ray_counters(st, &ctr);
    if (rc) return rc;
    surf_trav_kernel<false, true><<<n_ctas(rays->n_rays), CTA_THREADS, 0, st>>>(
        g, *opt, rays->origins, rays->dirs, rays->n_rays, (float *)g_ws_dbg.ptr, nullptr, nullptr, f, cache, grads, dbg,
        ctr, PreP());
    note_launches(1);
    return check_cuda(cudaGetLastError(), "surf_trav debug launch");
}

extern "C" int asurf_debug_ray_bounds(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                      float *xf_out, void *stream) {
    ASURF_REQUIRE(xf_out, ASURF_E_INVALID, "debug_ray_bounds: null output");
    DebugP dbg = {};
    dbg.xf = xf_out;
    return debug_launch(grid, rays, opt, dbg, (cudaStream_t)stream);
}

extern "C" int asurf_debug_trace(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                 int32_t max_hits, int32_t *hit_count, int32_t *hit_cell, int32_t *hit_kind, float *hit_t,
                                 void *stream) {
    ASURF_REQUIRE(hit_count && hit_cell && hit_kind && hit_t && max_hits > 0, ASURF_E_INVALID, "debug_trace: bad argument");
    DebugP dbg = {};
    dbg.max_hits = max_hits;
    dbg.hit_count = hit_count;
    dbg.hit_cell = hit_cell;
    dbg.hit_kind = hit_kind;
    dbg.hit_t = hit_t;
    return debug_launch(grid, rays, opt, dbg, (cudaStream_t)stream);
}

extern "C" void asurf_debug_set_skip(int32_t enabled) { if (debug_hooks_enabled()) g_skip_enabled = enabled ? 1 : 0; }
extern "C" void asurf_debug_set_wave(int32_t enabled) { if (debug_hooks_enabled()) g_wave_enabled = enabled ? 1 : 0; }
extern "C" void asurf_debug_set_seg(int32_t enabled) { if (debug_hooks_enabled()) g_seg_enabled = enabled ? 1 : 0; }
extern "C" int asurf_debug_counters(uint64_t *out8) {   // synchronises: queue lengths of the last render call
    ASURF_REQUIRE(out8 && g_ws_ctr.ptr, ASURF_E_INVALID, "debug_counters: nothing to read");
    return check_cuda(cudaMemcpy(out8, g_ws_ctr.ptr, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost), "debug_counters");
}

extern "C" int asurf_profile_enable(int32_t capacity) {
    for (int i = 0; i < PROF_EV * g_prof.cap; ++i) cudaEventDestroy(g_prof.ev[i]);
    delete[] g_prof.ev;
    g_prof = ProfRing();
    if (capacity <= 0) return 0;
    g_prof.ev = new cudaEvent_t[PROF_EV * (size_t)capacity];
    for (int i = 0; i < PROF_EV * capacity; ++i) ASURF_CUDA(cudaEventCreate(&g_prof.ev[i]));
    g_prof.cap = capacity;
    return 0;
}

extern "C" int asurf_profile_read_stages(int32_t *n_calls, float *stage_ms_sum) {
    ASURF_REQUIRE(n_calls && stage_ms_sum, ASURF_E_INVALID, "profile_read_stages: null output");
    for (int k = 0; k < PROF_EV - 1; ++k) stage_ms_sum[k] = 0.f;
    for (int i = 0; i < g_prof.n; ++i) {
        cudaEvent_t *pe = g_prof.ev + PROF_EV * i;
        ASURF_CUDA(cudaEventSynchronize(pe[PROF_EV - 1]));
        for (int k = 0; k < PROF_EV - 1; ++k) {
            float a = 0.f;
            ASURF_CUDA(cudaEventElapsedTime(&a, pe[k], pe[k + 1]));
            stage_ms_sum[k] += a;
        }
    }
    *n_calls = g_prof.n;
    g_prof.n = 0;
    return 0;
}

extern "C" int asurf_profile_read(int32_t *n_calls, float *fwd_ms_sum, float *bwd_ms_sum) {
    ASURF_REQUIRE(n_calls && fwd_ms_sum && bwd_ms_sum, ASURF_E_INVALID, "profile_read: null output");
    float st[PROF_EV - 1];
    int rc = asurf_profile_read_stages(n_calls, st);
    if (rc) return rc;
    *fwd_ms_sum = st[0] + st[1] + st[2];
    *bwd_ms_sum = st[3];
    return 0;
}

extern "C" void asurf_release(void) {
    g_ws_accel.release();
    g_ws_work.release();
    work_cache_release();
    g_ws_cache.release();
    g_ws_dbg.release();
    g_ws_ctr.release();
    g_ws_pre.release();
    g_ws_wave.release();
    g_ws_seg.release();
    loss_release();
    msi_release();
    cuvol_release();
    misc_release();
}
