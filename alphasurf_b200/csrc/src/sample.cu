// alphasurf_b200: point queries of the sparse grid -- what resample / resample_surface and the point
// extraction scripts call.
//
// Replaces, from /root/reference/svox2/csrc/svox2_kernel.cu:
//   sample_grid (:384-420), sample_grid_sh_surf (:423-462), sample_grid_raw_alpha (:464-489)
//       kernels sample_grid_{sh,density,alpha,surface}_kernel (:11-148): trilinear gather of one (N,C) tensor at world-space
//       points, missing corners reading a constant (0, empty_raw, default_surf)
//   sample_grid_backward (:491-539), kernels :151-246: the transposed scatter
//   cubic_extract_iso_pts (:542-582, kernel :248-376): zero crossings of the trilinear level function along a lattice of
//       axis-parallel lines through each listed cell
//
// One kernel serves the four forward variants (tensor, width and fill value are arguments) and one the two backward ones:
// a thread per (point, channel) with the channel fastest, so the 8 corner rows are read / updated as contiguous runs.  The
// reference creates two streams per call and synchronises the host on them; here everything is enqueued on the caller's
// stream.  The iso-point extraction runs one thread per (cell, line) instead of one per cell (3 n^2 cubic solves in a loop).
#include "common.cuh"
#include "surf_math.cuh"

namespace asurf {
namespace {

constexpr int SP_THREADS = 256;

struct SpXf {
    float offset[3], scaling[3];
    int size[3];
};

// world point -> voxel + interpolation weights (svox2_kernel.cu:20-29)
__device__ __forceinline__ void sp_locate(const SpXf &g, const float *__restrict__ points, int64_t pid, int *l, float *w) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float p = fmaf(points[pid * 3 + i], g.scaling[i], g.offset[i]);
        p = fminf(fmaxf(p, 0.f), (float)g.size[i] - 1.f);
        l[i] = min((int)p, g.size[i] - 2);
        w[i] = p - (float)l[i];
    }
}

__device__ __forceinline__ void sp_links(const SpXf &g, const int32_t *__restrict__ links, const int *l, int32_t *k) {
    const int offy = g.size[2];
    const int64_t offx = (int64_t)g.size[1] * g.size[2];
    const int32_t *lp = links + (offx * l[0] + (int64_t)offy * l[1] + l[2]);
    k[0] = __ldg(lp); k[1] = __ldg(lp + 1); k[2] = __ldg(lp + offy); k[3] = __ldg(lp + offy + 1);
    k[4] = __ldg(lp + offx); k[5] = __ldg(lp + offx + 1); k[6] = __ldg(lp + offx + offy); k[7] = __ldg(lp + offx + offy + 1);
}

__global__ void __launch_bounds__(SP_THREADS)
sample_kernel(const int32_t *__restrict__ links, const SpXf g, const float *__restrict__ data, int n_cols, float missing,
              const float *__restrict__ points, int64_t n_points, float *__restrict__ out) {
    const int64_t n = n_points * n_cols;
    for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pid = tid / n_cols;
        const int idx = (int)(tid - pid * n_cols);
        int l[3];
        float w[3];
        int32_t k[8];
        sp_locate(g, points, pid, l, w);
        sp_links(g, links, l, k);
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = (k[c] >= 0) ? __ldg(data + (int64_t)k[c] * n_cols + idx) : missing;
        out[tid] = trilerp8(v, w);
    }
}

__global__ void __launch_bounds__(SP_THREADS)
sample_backward_kernel(const int32_t *__restrict__ links, const SpXf g, const float *__restrict__ points, int64_t n_points,
                       const float *__restrict__ grad_out, int n_cols, float *__restrict__ grad_data) {
    const int64_t n = n_points * n_cols;
    for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pid = tid / n_cols;
        const int idx = (int)(tid - pid * n_cols);
        int l[3];
        float w[3];
        int32_t k[8];
        sp_locate(g, points, pid, l, w);
        sp_links(g, links, l, k);
        const float go = grad_out[tid];
        const float xb = w[0], yb = w[1], zb = w[2];
        const float xa = 1.f - w[0], ya = 1.f - w[1], za = 1.f - w[2];
        // products in the reference's order (:178-196)
        const float xago = xa * go, xbgo = xb * go;
        const float t00 = ya * xago, t01 = yb * xago, t10 = ya * xbgo, t11 = yb * xbgo;
        const float c[8] = {t00 * za, t00 * zb, t01 * za, t01 * zb, t10 * za, t10 * zb, t11 * za, t11 * zb};
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (k[q] >= 0) atomicAdd(grad_data + (int64_t)k[q] * n_cols + idx, c[q]);
    }
}

// one thread per (cell, i, j, axis): the line through lattice point (i, j) of the face normal to `axis`
__global__ void __launch_bounds__(SP_THREADS)
iso_pts_kernel(const int32_t *__restrict__ links, int sx, int sy, int sz, const float *__restrict__ level,
               const float *__restrict__ maskv, const int32_t *__restrict__ cell_ids, int64_t n_cells, int n_sample,
               float density_thresh, float *__restrict__ out) {
    const int per_cell = 3 * n_sample * n_sample;
    const int64_t n = n_cells * per_cell;
    for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = tid / per_cell;
        const int slot = (int)(tid - cell * per_cell);   // i * n_sample * 3 + j * 3 + dir_id
        const int dir_id = slot % 3, j = (slot / 3) % n_sample, i = slot / (3 * n_sample);
        const int xyz = __ldg(cell_ids + cell);
        const int z = xyz % sz, xy = xyz / sz, y = xy % sy, x = xy / sy;
        if ((x >= sx - 1) || (y >= sy - 1) || (z >= sz - 1)) continue;
        const int offy = sz;
        const int64_t offx = (int64_t)sy * sz;
        const int32_t *lp = links + (offx * x + (int64_t)offy * y + z);
        const int32_t k[8] = {__ldg(lp), __ldg(lp + 1), __ldg(lp + offy), __ldg(lp + offy + 1),
                              __ldg(lp + offx), __ldg(lp + offx + 1), __ldg(lp + offx + offy), __ldg(lp + offx + offy + 1)};
        bool ok = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) ok &= (k[c] >= 0);
        if (!ok) continue;
        double s[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) s[c] = (double)__ldg(level + k[c]);
        const float step = 1.f / (float)(n_sample - 1);
        const float pos1 = (float)i * step, pos2 = (float)j * step;
        double dirs[3] = {0., 0., 0.}, origin[3] = {0., 0., 0.};
        if (dir_id == 0) { dirs[0] = 1.; origin[1] = pos1; origin[2] = pos2; }
        else if (dir_id == 1) { dirs[1] = 1.; origin[0] = pos1; origin[2] = pos2; }
        else { dirs[2] = 1.; origin[0] = pos1; origin[1] = pos2; }
        double fs[4], st[3] = {-1, -1, -1};
        field_to_cubic(s, origin, dirs, fs);
        solve_cubic(fs[0], fs[1], fs[2], fs[3], st);
        float mv[8];
        bool mv_loaded = false;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (!((st[r] >= 0.) && (st[r] <= 1.))) continue;
            const float pt[3] = {(float)(origin[0] + dirs[0] * st[r]), (float)(origin[1] + dirs[1] * st[r]),
                                 (float)(origin[2] + dirs[2] * st[r])};
            if (!mv_loaded) {
#pragma unroll
                for (int c = 0; c < 8; ++c) mv[c] = __ldg(maskv + k[c]);
                mv_loaded = true;
            }
            if (trilerp8(mv, pt) >= density_thresh) {   // first root that passes the mask (:362-369)
                float *o = out + tid * 3;
                o[0] = pt[0] + (float)x;
                o[1] = pt[1] + (float)y;
                o[2] = pt[2] + (float)z;
                break;
            }
        }
    }
}

int sp_xf(const int32_t size[3], const float offset[3], const float scaling[3], SpXf &g, const char *who) {
    ASURF_REQUIRE(size && offset && scaling, ASURF_E_INVALID, "%s: null size / offset / scaling", who);
    ASURF_REQUIRE(size[0] >= 2 && size[1] >= 2 && size[2] >= 2, ASURF_E_INVALID, "%s: grid smaller than 2^3", who);
    for (int i = 0; i < 3; ++i) {
        g.size[i] = size[i];
        g.offset[i] = offset[i];
        g.scaling[i] = scaling[i];
    }
    return 0;
}

inline int sp_grid(int64_t n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + SP_THREADS - 1) / SP_THREADS;
    const int64_t cap = (int64_t)sms * 32;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace asurf

using namespace asurf;

extern "C" int asurf_sample_grid(const int32_t *links, const int32_t size[3], const float offset[3], const float scaling[3],
                                 const float *data, int32_t n_cols, float missing, const float *points, int64_t n_points,
                                 float *out, void *stream) {
    ASURF_REQUIRE(links && data, ASURF_E_INVALID, "sample_grid: null grid tensor");
    ASURF_REQUIRE(n_cols > 0 && n_points >= 0, ASURF_E_INVALID, "sample_grid: bad shape");
    if (n_points == 0) return 0;
    ASURF_REQUIRE(points && out, ASURF_E_INVALID, "sample_grid: null points / output");
    SpXf g;
    int rc = sp_xf(size, offset, scaling, g, "sample_grid");
    if (rc) return rc;
    sample_kernel<<<sp_grid(n_points * n_cols), SP_THREADS, 0, (cudaStream_t)stream>>>(links, g, data, n_cols, missing, points,
                                                                                       n_points, out);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "sample_grid launch");
}

extern "C" int asurf_sample_grid_backward(const int32_t *links, const int32_t size[3], const float offset[3],
                                          const float scaling[3], const float *points, int64_t n_points, const float *grad_out,
                                          int32_t n_cols, float *grad_data, void *stream) {
    ASURF_REQUIRE(links, ASURF_E_INVALID, "sample_grid_backward: null links");
    ASURF_REQUIRE(n_cols > 0 && n_points >= 0, ASURF_E_INVALID, "sample_grid_backward: bad shape");
    if (n_points == 0) return 0;
    ASURF_REQUIRE(points && grad_out && grad_data, ASURF_E_INVALID, "sample_grid_backward: null tensor");
    SpXf g;
    int rc = sp_xf(size, offset, scaling, g, "sample_grid_backward");
    if (rc) return rc;
    sample_backward_kernel<<<sp_grid(n_points * n_cols), SP_THREADS, 0, (cudaStream_t)stream>>>(links, g, points, n_points,
                                                                                                grad_out, n_cols, grad_data);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "sample_grid_backward launch");
}

extern "C" int asurf_cubic_extract_iso_pts(const int32_t *links, const int32_t size[3], const float *level_data,
                                           const float *mask_data, const int32_t *cell_ids, int64_t n_cells, int32_t n_sample,
                                           float density_thresh, float *out, void *stream) {
    ASURF_REQUIRE(links && size && level_data && mask_data, ASURF_E_INVALID, "cubic_extract_iso_pts: null tensor");
    ASURF_REQUIRE(n_sample >= 2 && n_cells >= 0, ASURF_E_INVALID, "cubic_extract_iso_pts: n_sample must be >= 2");
    if (n_cells == 0) return 0;
    ASURF_REQUIRE(cell_ids && out, ASURF_E_INVALID, "cubic_extract_iso_pts: null cell list / output");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = n_cells * 3 * n_sample * n_sample;
    ASURF_CUDA(cudaMemsetAsync(out, 0, (size_t)n * 3 * sizeof(float), st));
    iso_pts_kernel<<<sp_grid(n), SP_THREADS, 0, st>>>(links, size[0], size[1], size[2], level_data, mask_data, cell_ids, n_cells,
                                                      n_sample, density_thresh, out);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "cubic_extract_iso_pts launch");
}
