// alphasurf_b200: grid maintenance renders -- the passes the reference runs before / between training phases to decide
// which voxels to keep.
//
// Replaces, from /root/reference/svox2/csrc/misc_kernel.cu:
//   dilate (:1005-1020, kernel :24-54)                         3x3x3 OR of a bool grid
//   grid_weight_render (:1084-1111, kernel :888-912, :187-284)  max rendering weight per vertex of a DENSE sigma volume
//   sparse_grid_weight_render (:1113-1138, :915-936, :287-401)  max transmittance reaching each vertex, sparse grid
//   sparse_grid_mask_render (:1158-1175, :954-972, :403-509)    rows whose voxel any ray passes through (step 0.1)
//   sparse_grid_visbility_render_surf (:1140-1156, :939-952, :511-719)  per-row count of rays that reach the voxel before
//                                                                        their first level-set intersection
//
// One thread per pixel / ray like the reference (these run a handful of times per training, over the training cameras);
// the float atomicMax is a single integer RED for the non-negative values that occur instead of a CAS loop, the 3x3x3
// dilation reads each byte once per z-row through the read-only path.  Integer / byte work and scattered atomics: HBM / L2
// bound, nothing to stage.
#include "common.cuh"
#include "surf_math.cuh"

namespace asurf {
namespace {

constexpr int GT_THREADS = 256;

struct GtCam {   // include/data_spec.hpp:124-137 (CameraSpec), c2w row-major 3x4
    float c2w[12];
    float fx, fy, cx, cy;
    int width, height;
};

struct GtXf {
    float offset[3], scaling[3];
    int size[3];
};

// cam2world_ray (include/render_util.cuh:599-617), no NDC
__device__ __forceinline__ void gt_cam_ray(const GtCam &cam, int ix, int iy, float *o, float *d) {
    float x = ((float)ix + 0.5f - cam.cx) / cam.fx;
    float y = ((float)iy + 0.5f - cam.cy) / cam.fy;
    float z = sqrtf((float)((double)(x * x + y * y) + 1.0));
    x /= z; y /= z; z = 1.0f / z;
    d[0] = cam.c2w[0] * x + cam.c2w[1] * y + cam.c2w[2] * z;
    d[1] = cam.c2w[4] * x + cam.c2w[5] * y + cam.c2w[6] * z;
    d[2] = cam.c2w[8] * x + cam.c2w[9] * y + cam.c2w[10] * z;
    o[0] = cam.c2w[3]; o[1] = cam.c2w[7]; o[2] = cam.c2w[11];
}

// world -> grid, AABB bounds from t = 0 (misc_kernel.cu:198-219); returns delta_scale
__device__ __forceinline__ float gt_ray_bounds(const GtXf &g, float *o, float *d, float &t, float &tmax) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        o[i] = fmaf(o[i], g.scaling[i], g.offset[i]);
        d[i] *= g.scaling[i];
    }
    const float delta_scale = rnorm3df(d[0], d[1], d[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] *= delta_scale;
    t = 0.f;
    tmax = 2e3f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float inv = (float)(1.0 / (double)d[i]);
        const float t1 = (-0.5f - o[i]) * inv, t2 = ((float)g.size[i] - 0.5f - o[i]) * inv;
        if (d[i] != 0.f) {
            t = fmaxf(t, fminf(t1, t2));
            tmax = fminf(tmax, fmaxf(t1, t2));
        }
    }
    return delta_scale;
}

// atomicMax(float) of cuda_util.cuh:41-50 (a CAS loop around fmaxf).  For value >= 0 the integer order of the bit
// patterns is the float order against any non-NaN content, so one RED does it; other values take the loop.
__device__ __forceinline__ void gt_atomic_max(float *addr, float value) {
    if (value >= 0.f) {
        atomicMax(reinterpret_cast<int *>(addr), __float_as_int(value));
    } else {
        unsigned *a = reinterpret_cast<unsigned *>(addr);
        unsigned old = *a, assumed;
        do {
            assumed = old;
            old = atomicCAS(a, assumed, __float_as_uint(fmaxf(value, __uint_as_float(assumed))));
        } while (old != assumed);
    }
}

// sample voxel + interpolation offsets at t (misc_kernel.cu:233-239)
__device__ __forceinline__ void gt_sample(const GtXf &g, const float *o, const float *d, float t, int *l, float *pos) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float p = fmaf(t, d[j], o[j]);
        p = fminf(fmaxf(p, 0.f), (float)g.size[j] - 1.f);
        l[j] = min((int)p, g.size[j] - 2);
        pos[j] = p - (float)l[j];
    }
}

__global__ void __launch_bounds__(GT_THREADS)
dilate_kernel(const uint8_t *__restrict__ in, int sx, int sy, int sz, uint8_t *__restrict__ out) {
    const int64_t n = (int64_t)sx * sy * sz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % sz);
        const int64_t xy = i / sz;
        const int y = (int)(xy % sy), x = (int)(xy / sy);
        const int xl = max(x - 1, 0), xr = min(x + 1, sx - 1), yl = max(y - 1, 0), yr = min(y + 1, sy - 1);
        const int zl = max(z - 1, 0), zr = min(z + 1, sz - 1);
        unsigned any = 0;
        for (int a = xl; a <= xr; ++a)
            for (int b = yl; b <= yr; ++b) {
                const uint8_t *row = in + ((int64_t)a * sy + b) * sz;
                any |= __ldg(row + zl) | __ldg(row + z) | __ldg(row + zr);
            }
        out[i] = any ? 1 : 0;
    }
}

// SPARSE = false: grid_trace_ray on a dense (X,Y,Z) sigma volume, vertices receive the max WEIGHT;
// SPARSE = true:  sprase_grid_trace_ray through links, vertices receive the max TRANSMITTANCE in front of the sample.
template <bool SPARSE>
__global__ void __launch_bounds__(GT_THREADS)
weight_render_kernel(const float *__restrict__ data, const int32_t *__restrict__ links, const GtXf g, const GtCam cam,
                     float step_size, float stop_thresh, int last_sample_opaque, float *__restrict__ grid_weight) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)cam.width * cam.height) return;
    const int iy = (int)(tid / cam.width), ix = (int)(tid % cam.width);
    float o[3], d[3], t, tmax;
    gt_cam_ray(cam, ix, iy, o, d);
    const float world_step = gt_ray_bounds(g, o, d, t, tmax) * step_size;
    if (t > tmax) return;
    float log_light = 0.f;
    const int64_t s0 = (int64_t)g.size[1] * g.size[2];
    const int s1 = g.size[2];
    while (t <= tmax) {
        int l[3];
        float pos[3];
        gt_sample(g, o, d, t, l, pos);
        const int64_t idx = l[0] * s0 + (int64_t)l[1] * s1 + l[2];
        float v[8];
        if (SPARSE) {
            const int32_t *lp = links + idx;
            const int32_t k[8] = {__ldg(lp), __ldg(lp + 1), __ldg(lp + s1), __ldg(lp + s1 + 1),
                                  __ldg(lp + s0), __ldg(lp + s0 + 1), __ldg(lp + s0 + s1), __ldg(lp + s0 + s1 + 1)};
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = k[c] >= 0 ? __ldg(data + k[c]) : 0.f;
        } else {
            const float *p = data + idx;
            v[0] = __ldg(p); v[1] = __ldg(p + 1); v[2] = __ldg(p + s1); v[3] = __ldg(p + s1 + 1);
            v[4] = __ldg(p + s0); v[5] = __ldg(p + s0 + 1); v[6] = __ldg(p + s0 + s1); v[7] = __ldg(p + s0 + s1 + 1);
        }
        float sigma = trilerp8(v, pos);
        if (!SPARSE && last_sample_opaque && t + step_size > tmax) {
            sigma += 1e9f;
            log_light = 0.f;
        }
        if (sigma > 1e-8f) {
            const float log_att = -world_step * sigma;
            const float w = SPARSE ? __expf(log_light) : __expf(log_light) * (1.f - __expf(log_att));
            float *q = grid_weight + idx;
            gt_atomic_max(q, w);
            gt_atomic_max(q + 1, w);
            gt_atomic_max(q + s1, w);
            gt_atomic_max(q + s1 + 1, w);
            gt_atomic_max(q + s0, w);
            gt_atomic_max(q + s0 + 1, w);
            gt_atomic_max(q + s0 + s1, w);
            gt_atomic_max(q + s0 + s1 + 1, w);
            log_light += log_att;
            if (__expf(log_light) < stop_thresh) break;
        }
        t += step_size;
    }
}

__global__ void __launch_bounds__(GT_THREADS)
mask_render_kernel(const int32_t *__restrict__ links, const GtXf g, const float *__restrict__ origins,
                   const float *__restrict__ dirs, int64_t Q, float near_clip, float *__restrict__ grid_mask) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray_id >= Q) return;
    float o[3], d[3], t, tmax;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        o[i] = origins[ray_id * 3 + i];
        d[i] = dirs[ray_id * 3 + i];
    }
    gt_ray_bounds(g, o, d, t, tmax);
    if (t < near_clip) t = near_clip;
    if (t > tmax) return;
    const float step_size = 0.1f;   // fixed in the reference (:409)
    const int64_t s0 = (int64_t)g.size[1] * g.size[2];
    const int s1 = g.size[2];
    int64_t last = -1;
    while (t <= tmax) {
        int l[3];
        float pos[3];
        gt_sample(g, o, d, t, l, pos);
        const int64_t idx = l[0] * s0 + (int64_t)l[1] * s1 + l[2];
        if (idx != last) {   // ~10 consecutive samples share a voxel; the mark is idempotent
            last = idx;
            const int32_t *lp = links + idx;
            const int32_t k[8] = {__ldg(lp), __ldg(lp + 1), __ldg(lp + s1), __ldg(lp + s1 + 1),
                                  __ldg(lp + s0), __ldg(lp + s0 + 1), __ldg(lp + s0 + s1), __ldg(lp + s0 + s1 + 1)};
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (k[c] >= 0) gt_atomic_max(grid_mask + k[c], 1.f);
        }
        t += step_size;
    }
}

__global__ void __launch_bounds__(GT_THREADS)
visibility_surf_kernel(const int32_t *__restrict__ links, const float *__restrict__ surface,
                       const float *__restrict__ level_set, int level_set_num, const GtXf g, const GtCam cam,
                       float *__restrict__ visibility) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (int64_t)cam.width * cam.height) return;
    const int iy = (int)(tid / cam.width), ix = (int)(tid % cam.width);
    float o[3], d[3], t, tmax;
    gt_cam_ray(cam, ix, iy, o, d);
    gt_ray_bounds(g, o, d, t, tmax);
    if (t > tmax) return;
    int nv[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) nv[j] = min(max((int)fmaf(t, d[j], o[j]), 0), g.size[j] - 2);
    const int64_t offx = (int64_t)g.size[1] * g.size[2];
    const int offy = g.size[2];
    const double dd[3] = {(double)d[0], (double)d[1], (double)d[2]};
    while (t <= tmax) {
        const int v[3] = {nv[0], nv[1], nv[2]};
        const int32_t *lp = links + (offx * v[0] + (int64_t)offy * v[1] + v[2]);
        const int32_t k[8] = {__ldg(lp), __ldg(lp + 1), __ldg(lp + offy), __ldg(lp + offy + 1),
                              __ldg(lp + offx), __ldg(lp + offx + 1), __ldg(lp + offx + offy), __ldg(lp + offx + offy + 1)};
        bool all = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (k[c] >= 0) atomicAdd(visibility + k[c], 1.f);
            all &= (k[c] >= 0);
        }
        // one step of the reference DDA (:584-618)
        float tc = 0.f, tf[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int close_plane = d[j] > 0.f ? v[j] : v[j] + 1, far_plane = d[j] > 0.f ? v[j] + 1 : v[j];
            tc = fmaxf(tc, ((float)close_plane - o[j]) / d[j]);
            tf[j] = ((float)far_plane - o[j]) / d[j];
        }
        const float t_far = fminf(fminf(tf[0], tf[1]), tf[2]);
        t = t_far;
        const int a = (t_far == tf[0]) ? 0 : ((t_far == tf[1]) ? 1 : 2);
        nv[a] += (d[a] > 0.f) ? 1 : -1;
        // the reference sets t = ray.tmax + 1 from a SingleRaySpec whose tmax was never initialised (:600-609): leaving
        // the voxel range ends the march here
        if ((nv[a] < 0) || (nv[a] >= g.size[a] - 1)) t = tmax + 1.f;
        if (!all) continue;
        float sf[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) sf[c] = __ldg(surface + k[c]);
        float smin = sf[0], smax = sf[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) {
            smin = fminf(smin, sf[c]);
            smax = fmaxf(smax, sf[c]);
        }
        const float nof[3] = {fmaf(tc, d[0], o[0]), fmaf(tc, d[1], o[1]), fmaf(tc, d[2], o[2])};
        const double nno[3] = {(double)nof[0] - v[0], (double)nof[1] - v[1], (double)nof[2] - v[2]};
        bool fs_ready = false;
        double fs[4];
        for (int i = 0; i < level_set_num; ++i) {
            const float lv = __ldg(level_set + i);
            if ((lv < smin) || (lv > smax)) continue;
            if (!fs_ready) {
                double s[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) s[c] = (double)sf[c];
                field_to_cubic(s, nno, dd, fs);
                fs_ready = true;
            }
            double st[3] = {-1, -1, -1};
            solve_cubic(fs[0] - (double)lv, fs[1], fs[2], fs[3], st);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (st[j] <= 0) continue;
                const float stf = (float)st[j];
                const float p0 = fmaf(stf, d[0], nof[0]) - (float)v[0], p1 = fmaf(stf, d[1], nof[1]) - (float)v[1],
                            p2 = fmaf(stf, d[2], nof[2]) - (float)v[2];
                if ((p0 < 0) | (p0 > 1) | (p1 < 0) | (p1 > 1) | (p2 < 0) | (p2 > 1)) continue;
                return;   // first intersection: everything behind it is occluded for this pixel
            }
        }
    }
}

int gt_xf(const int32_t size[3], const float offset[3], const float scaling[3], GtXf &g, const char *who) {
    ASURF_REQUIRE(size && offset && scaling, ASURF_E_INVALID, "%s: null size / offset / scaling", who);
    ASURF_REQUIRE(size[0] >= 2 && size[1] >= 2 && size[2] >= 2, ASURF_E_INVALID, "%s: grid smaller than 2^3", who);
    for (int i = 0; i < 3; ++i) {
        g.size[i] = size[i];
        g.offset[i] = offset[i];
        g.scaling[i] = scaling[i];
    }
    return 0;
}

int gt_cam(const float *c2w_host, float fx, float fy, float cx, float cy, int32_t width, int32_t height, GtCam &cam,
           const char *who) {
    ASURF_REQUIRE(c2w_host, ASURF_E_INVALID, "%s: null camera matrix", who);
    ASURF_REQUIRE(width >= 0 && height >= 0, ASURF_E_INVALID, "%s: negative image size", who);
    for (int i = 0; i < 12; ++i) cam.c2w[i] = c2w_host[i];
    cam.fx = fx; cam.fy = fy; cam.cx = cx; cam.cy = cy;
    cam.width = width; cam.height = height;
    return 0;
}

inline int gt_grid(int64_t n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + GT_THREADS - 1) / GT_THREADS;
    const int64_t cap = (int64_t)sms * 32;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace asurf

using namespace asurf;

extern "C" int asurf_dilate(const uint8_t *grid, const int32_t size[3], uint8_t *out, void *stream) {
    ASURF_REQUIRE(grid && size && out, ASURF_E_INVALID, "dilate: null argument");
    const int64_t n = (int64_t)size[0] * size[1] * size[2];
    if (n <= 0) return 0;
    dilate_kernel<<<gt_grid(n), GT_THREADS, 0, (cudaStream_t)stream>>>(grid, size[0], size[1], size[2], out);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "dilate launch");
}

extern "C" int asurf_grid_weight_render(const float *data, const int32_t size[3], const float offset[3], const float scaling[3],
                                        const float *c2w_host, float fx, float fy, float cx, float cy, int32_t width,
                                        int32_t height, float step_size, float stop_thresh, int32_t last_sample_opaque,
                                        float *grid_weight_out, void *stream) {
    ASURF_REQUIRE(data && grid_weight_out, ASURF_E_INVALID, "grid_weight_render: null tensor");
    GtXf g;
    GtCam cam;
    int rc = gt_xf(size, offset, scaling, g, "grid_weight_render");
    if (!rc) rc = gt_cam(c2w_host, fx, fy, cx, cy, width, height, cam, "grid_weight_render");
    if (rc) return rc;
    const int64_t Q = (int64_t)width * height;
    if (Q == 0) return 0;
    weight_render_kernel<false><<<div_up(Q, GT_THREADS), GT_THREADS, 0, (cudaStream_t)stream>>>(
        data, nullptr, g, cam, step_size, stop_thresh, last_sample_opaque, grid_weight_out);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "grid_weight_render launch");
}

extern "C" int asurf_sparse_grid_weight_render(const int32_t *links, const float *density, const int32_t size[3],
                                               const float offset[3], const float scaling[3], const float *c2w_host, float fx,
                                               float fy, float cx, float cy, int32_t width, int32_t height, float step_size,
                                               float stop_thresh, float *grid_weight_out, void *stream) {
    ASURF_REQUIRE(links && density && grid_weight_out, ASURF_E_INVALID, "sparse_grid_weight_render: null tensor");
    GtXf g;
    GtCam cam;
    int rc = gt_xf(size, offset, scaling, g, "sparse_grid_weight_render");
    if (!rc) rc = gt_cam(c2w_host, fx, fy, cx, cy, width, height, cam, "sparse_grid_weight_render");
    if (rc) return rc;
    const int64_t Q = (int64_t)width * height;
    if (Q == 0) return 0;
    weight_render_kernel<true><<<div_up(Q, GT_THREADS), GT_THREADS, 0, (cudaStream_t)stream>>>(
        density, links, g, cam, step_size, stop_thresh, 0, grid_weight_out);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "sparse_grid_weight_render launch");
}

extern "C" int asurf_sparse_grid_mask_render(const int32_t *links, const int32_t size[3], const float offset[3],
                                             const float scaling[3], const float *origins, const float *dirs, int64_t n_rays,
                                             float near_clip, float *grid_mask, void *stream) {
    ASURF_REQUIRE(links && grid_mask, ASURF_E_INVALID, "sparse_grid_mask_render: null tensor");
    ASURF_REQUIRE(n_rays >= 0, ASURF_E_INVALID, "sparse_grid_mask_render: negative ray count");
    if (n_rays == 0) return 0;
    ASURF_REQUIRE(origins && dirs, ASURF_E_INVALID, "sparse_grid_mask_render: null ray tensor");
    GtXf g;
    int rc = gt_xf(size, offset, scaling, g, "sparse_grid_mask_render");
    if (rc) return rc;
    mask_render_kernel<<<div_up(n_rays, GT_THREADS), GT_THREADS, 0, (cudaStream_t)stream>>>(links, g, origins, dirs, n_rays,
                                                                                            near_clip, grid_mask);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "sparse_grid_mask_render launch");
}

extern "C" int asurf_sparse_grid_visibility_render_surf(const int32_t *links, const float *surface, const float *level_set,
                                                        int32_t level_set_num, const int32_t size[3], const float offset[3],
                                                        const float scaling[3], const float *c2w_host, float fx, float fy,
                                                        float cx, float cy, int32_t width, int32_t height,
                                                        float *visibility_out, void *stream) {
    ASURF_REQUIRE(links && surface && level_set && visibility_out, ASURF_E_INVALID,
                  "sparse_grid_visbility_render_surf: null tensor");
    GtXf g;
    GtCam cam;
    int rc = gt_xf(size, offset, scaling, g, "sparse_grid_visbility_render_surf");
    if (!rc) rc = gt_cam(c2w_host, fx, fy, cx, cy, width, height, cam, "sparse_grid_visbility_render_surf");
    if (rc) return rc;
    const int64_t Q = (int64_t)width * height;
    if (Q == 0) return 0;
    visibility_surf_kernel<<<div_up(Q, GT_THREADS), GT_THREADS, 0, (cudaStream_t)stream>>>(links, surface, level_set,
                                                                                          level_set_num, g, cam, visibility_out);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "sparse_grid_visbility_render_surf launch");
}
