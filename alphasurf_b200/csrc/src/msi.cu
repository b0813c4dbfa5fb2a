// alphasurf_b200: multi-sphere-image (MSI) background of the svox2 renderers, and its TV regulariser.
//
// Replaces, from /root/reference/svox2/csrc:
//   render_background_forward / _backward            render_lerp_kernel_surf_trav.cu:2914-3137 (same code in
//                                                    render_lerp_kernel_cuvol.cu:540-760)
//   render_background_kernel / _backward_kernel      :3370-3455
//   ray_find_bounds_bg, ConcentricSpheresIntersector include/render_util.cuh:619-649, :703-745
//   trilerp_bg_one / trilerp_backward_bg_one         include/render_util.cuh:207-283
//   msi_tv_grad_sparse                               loss_kernel.cu:979-1064, host :1429-1463
//
// The background is a stack of `nlayers` equirectangular images on spheres of radius 1 .. infinity around the grid; a
// foreground render leaves, per ray, the log-transmittance behind the grid (and, in the backward pass, what is left of the
// running sum `accum`), and the background pass continues the same compositing through the layers.  One thread per ray
// as in the reference: a ray meets nlayers / step_size + 2 spheres and reads 4 texels x 2 layers x 4 channels per sphere,
// all from a table of a few MB -- L2 resident, latency-bound, a small fraction of the foreground pass.
#include "common.cuh"

namespace asurf {
namespace {

constexpr float MSI_C0 = 0.28209479177387814f;   // SH DC factor (render_util.cuh:373)

struct MsiP {
    const int32_t *links;   // (2 reso, reso)
    const float *data;      // (n, nlayers, 4): r, g, b, sigma
    int reso, nlayers;
    int size[3];
    float offset[3], scaling[3];
};

struct MsiRay {
    float o[3], d[3];
    float world_step;
};

// ray_find_bounds_bg: world -> grid -> the unit sphere around the grid
__device__ __forceinline__ void msi_ray_setup(const MsiP &m, MsiRay &r) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.o[i] = fmaf(r.o[i], m.scaling[i], m.offset[i]);
        r.d[i] *= m.scaling[i];
    }
    const float delta_scale = rnorm3df(r.d[0], r.d[1], r.d[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) r.d[i] *= delta_scale;
    r.world_step = delta_scale;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float ss = 2.f / (float)m.size[i];
        r.o[i] = fmaf(r.o[i] + 0.5f, ss, -1.f);
        r.d[i] = r.d[i] * ss;
    }
    const float inorm = rnorm3df(r.d[0], r.d[1], r.d[2]);
    r.world_step *= inorm;
#pragma unroll
    for (int i = 0; i < 3; ++i) r.d[i] *= inorm;
}

struct MsiSpheres {   // ConcentricSpheresIntersector: far intersection with the sphere of radius r around the origin
    float q2a, qb, f;
    __device__ __forceinline__ MsiSpheres(const float *o, const float *d) {
        q2a = 2 * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        qb = 2 * (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]);
        f = qb * qb - 2 * q2a * (o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    }
    __device__ __forceinline__ bool intersect(float r, float *out) const {
        const float det = f + 2 * q2a * r * r;
        if (det < 0) return false;
        *out = (-qb + sqrtf(det)) / q2a;
        return true;
    }
};

__device__ __forceinline__ float msi_inner_radius(const float *o, const float *d) {   // _dist_ray_to_origin + 1e-3, >= 1
    const float c0 = o[1] * d[2] - o[2] * d[1], c1 = o[2] * d[0] - o[0] * d[2], c2 = o[0] * d[1] - o[1] * d[0];
    return fmaxf(norm3df(c0, c1, c2) + 1e-3f, 1.f);
}

// position on the sphere of radius r -> texel (l) and interpolation offsets (pos); false when the ray misses the sphere
__device__ __forceinline__ bool msi_sample_pos(const MsiP &m, const MsiRay &r, const MsiSpheres &csi, float radius,
                                               float inner_radius, int *l, float *pos, float &invr_mid) {
    float t;
    if (radius < inner_radius || !csi.intersect(radius, &t)) return false;
#pragma unroll
    for (int j = 0; j < 3; ++j) pos[j] = fmaf(t, r.d[j], r.o[j]);
    invr_mid = rnorm3df(pos[0], pos[1], pos[2]);
#pragma unroll
    for (int j = 0; j < 3; ++j) pos[j] *= invr_mid;
    // _unitvec2equirect (double constants as in the reference, :555-563)
    const float lat = asinf(pos[1]);
    const float lon = atan2f(pos[0], pos[2]);
    pos[0] = (float)(m.reso * 2 * (0.5 + lon * 0.5 * 0.318309886183790671538));
    pos[1] = (float)(m.reso * (0.5 - lat * 0.318309886183790671538));
    pos[2] = fminf(fmaxf((1.f - invr_mid) * m.nlayers - 0.5f, 0.f), (float)(m.nlayers - 1));
#pragma unroll
    for (int j = 0; j < 3; ++j) l[j] = (int)pos[j];
    l[0] = min(l[0], m.reso * 2 - 1);
    l[1] = min(l[1], m.reso - 1);
    l[2] = min(l[2], m.nlayers - 2);
#pragma unroll
    for (int j = 0; j < 3; ++j) pos[j] -= (float)l[j];
    return true;
}

__device__ __forceinline__ float lerpf(float a, float b, float w) { return fmaf(w, b - a, a); }

// trilerp_bg_one: bilinear over the (wrapping) equirect texels, linear over two layers
__device__ __forceinline__ float msi_trilerp(const MsiP &m, const int *l, const float *pos, int idx) {
    const int ny = l[1] < (m.reso - 1) ? (l[1] + 1) : 0;
    const int nx = l[0] < (2 * m.reso - 1) ? (l[0] + 1) : 0;
    const int u[4] = {m.reso * l[0] + l[1], m.reso * l[0] + ny, m.reso * nx + l[1], m.reso * nx + ny};
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int link = __ldg(m.links + u[c]);
        if (link >= 0) {
            const float *dp = m.data + ((int64_t)link * m.nlayers + l[2]) * 4 + idx;
            v[c] = lerpf(__ldg(dp), __ldg(dp + 4), pos[2]);
        } else {
            v[c] = 0.f;
        }
    }
    const float ix0 = lerpf(v[0], v[1], pos[1]);
    const float ix1 = lerpf(v[2], v[3], pos[1]);
    return lerpf(ix0, ix1, pos[0]);
}

// trilerp_backward_bg_one
__device__ __forceinline__ void msi_trilerp_backward(const MsiP &m, float *__restrict__ grad, uint8_t *__restrict__ mask,
                                                     const int *l, const float *pos, float grad_out, int idx) {
    const float ay = 1.f - pos[1], az = 1.f - pos[2];
    const int ny = l[1] < (m.reso - 1) ? (l[1] + 1) : 0;
    const int nx = l[0] < (2 * m.reso - 1) ? (l[0] + 1) : 0;
    const int u[4] = {m.reso * l[0] + l[1], m.reso * l[0] + ny, m.reso * nx + l[1], m.reso * nx + ny};
    const float xo0 = (1.0f - pos[0]) * grad_out, xo1 = pos[0] * grad_out;
    const float w[4] = {ay * xo0, pos[1] * xo0, ay * xo1, pos[1] * xo1};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int link = __ldg(m.links + u[c]);
        if (link >= 0) {
            const int64_t row = (int64_t)link * m.nlayers + l[2];
            float *gp = grad + row * 4 + idx;
            atomicAdd(gp, w[c] * az);
            atomicAdd(gp + 4, w[c] * pos[2]);
            if (mask) { mask[row] = 1; mask[row + 1] = 1; }
        }
    }
}

struct MsiCam {          // CameraSpec (include/data_spec.hpp:124-137), c2w row-major 3x4; width == 0: rays come from tensors
    float c2w[12];
    float fx, fy, cx, cy;
    int width, height;
};

__global__ void __launch_bounds__(128)
msi_forward_kernel(const MsiP m, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                   const MsiCam cam, int64_t Q, const float *__restrict__ log_transmit_in, float *__restrict__ rgb_out) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray_id >= Q) return;
    float log_transmit = log_transmit_in[ray_id];
    if (log_transmit < -25.f) return;
    MsiRay r;
    if (cam.width > 0) {   // cam2world_ray (include/render_util.cuh:599-617), render_background_image_kernel :3389-3411
        const int ix = (int)(ray_id % cam.width), iy = (int)(ray_id / cam.width);
        float x = ((float)ix + 0.5f - cam.cx) / cam.fx;
        float y = ((float)iy + 0.5f - cam.cy) / cam.fy;
        float z = sqrtf((float)((double)(x * x + y * y) + 1.0));
        x /= z; y /= z; z = 1.0f / z;
        r.d[0] = cam.c2w[0] * x + cam.c2w[1] * y + cam.c2w[2] * z;
        r.d[1] = cam.c2w[4] * x + cam.c2w[5] * y + cam.c2w[6] * z;
        r.d[2] = cam.c2w[8] * x + cam.c2w[9] * y + cam.c2w[10] * z;
        r.o[0] = cam.c2w[3]; r.o[1] = cam.c2w[7]; r.o[2] = cam.c2w[11];
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            r.o[i] = origins[ray_id * 3 + i];
            r.d[i] = dirs[ray_id * 3 + i];
        }
    }
    msi_ray_setup(m, r);
    const MsiSpheres csi(r.o, r.d);
    const float inner_radius = msi_inner_radius(r.o, r.d);
    float invr_last = 1.f / inner_radius;
    const int n_steps = (int)(m.nlayers / opt.step_size) + 2;
    float outv[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < n_steps; ++i) {
        const float radius = (float)(n_steps / (n_steps - i - 0.5));   // between 1 and infinity (double arithmetic, :2933)
        int l[3];
        float pos[3], invr_mid;
        if (!msi_sample_pos(m, r, csi, radius, inner_radius, l, pos, invr_mid)) continue;
        const float sigma = msi_trilerp(m, l, pos, 3);
        if (sigma > 0.f) {
            const float pcnt = (invr_last - invr_mid) * r.world_step * sigma;
            const float weight = __expf(log_transmit) * (1.f - __expf(-pcnt));
            log_transmit -= pcnt;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float color = msi_trilerp(m, l, pos, c) * MSI_C0;
                outv[c] += weight * fmaxf(color + 0.5f, 0.f);
            }
            if (__expf(log_transmit) < opt.stop_thresh) break;
        }
        invr_last = invr_mid;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb_out[ray_id * 3 + c] += outv[c] + __expf(log_transmit) * opt.background_brightness;
}

__global__ void __launch_bounds__(128)
msi_backward_kernel(const MsiP m, const asurf_opt_t opt, const float *__restrict__ origins, const float *__restrict__ dirs,
                    int64_t Q, const float *__restrict__ grad_in, const float *__restrict__ color_cache, int grad_is_rgb,
                    float norm_factor, const float *__restrict__ log_transmit_in, const float *__restrict__ accum_in,
                    float beta_loss, float sparsity_loss, float *__restrict__ grad_bg, uint8_t *__restrict__ mask_bg) {
    const int64_t ray_id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray_id >= Q) return;
    float log_transmit = log_transmit_in[ray_id];
    if (log_transmit < -25.f) return;
    MsiRay r;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.o[i] = origins[ray_id * 3 + i];
        r.d[i] = dirs[ray_id * 3 + i];
    }
    msi_ray_setup(m, r);
    float go[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
        go[c] = grad_is_rgb ? (color_cache[ray_id * 3 + c] - grad_in[ray_id * 3 + c]) * norm_factor : grad_in[ray_id * 3 + c];
    // what the foreground backward left of its running sum; a ray the foreground pass never visited (no work voxel on it)
    // carries the initial value (:1795-1797) minus the beta term the foreground cancels at its end (:2901-2905)
    float accum = accum_in[ray_id];
    if (isnan(accum) || isinf(accum)) {   // +Inf: the ray missed the grid (the foreground returned before its loop: no beta term)
        const bool missed = isinf(accum);
        accum = fmaf(color_cache[ray_id * 3 + 0], go[0], fmaf(color_cache[ray_id * 3 + 1], go[1], color_cache[ray_id * 3 + 2] * go[2]));
        if (!missed) accum -= beta_loss;
    }
    const MsiSpheres csi(r.o, r.d);
    const int n_steps = (int)(m.nlayers / opt.step_size) + 2;
    const float inner_radius = msi_inner_radius(r.o, r.d);
    float invr_last = 1.f / inner_radius;
    for (int i = 0; i < n_steps; ++i) {
        const float radius = (float)(n_steps / (n_steps - i - 0.5));
        int l[3];
        float pos[3], invr_mid;
        if (!msi_sample_pos(m, r, csi, radius, inner_radius, l, pos, invr_mid)) continue;
        const float sigma = msi_trilerp(m, l, pos, 3);
        if (sigma > 0.f) {
            float total_color = 0.f;
            const float pcnt = r.world_step * (invr_last - invr_mid) * sigma;
            const float weight = __expf(log_transmit) * (1.f - __expf(-pcnt));
            log_transmit -= pcnt;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float color = msi_trilerp(m, l, pos, c) * MSI_C0 + 0.5f;
                total_color += fmaxf(color, 0.f) * go[c];
                if (color > 0.f) msi_trilerp_backward(m, grad_bg, nullptr, l, pos, MSI_C0 * weight * go[c], c);
            }
            accum -= weight * total_color;
            float curr_grad_sigma = r.world_step * (invr_last - invr_mid) * (total_color * __expf(log_transmit) - accum);
            if (sparsity_loss > 0.f) curr_grad_sigma += sparsity_loss * (4 * sigma / (1 + 2 * (sigma * sigma)));
            msi_trilerp_backward(m, grad_bg, mask_bg, l, pos, curr_grad_sigma, 3);
            if (__expf(log_transmit) < opt.stop_thresh) break;
        }
        invr_last = invr_mid;
    }
}

// msi_tv_grad_sparse_kernel: thread per (cell of the list, channel)
__global__ void __launch_bounds__(256)
msi_tv_kernel(const int32_t *__restrict__ links, const float *__restrict__ msi, int lx, int ly, int nlayers, int nch,
              const int32_t *__restrict__ cells, float scale, float scale_last, int64_t Q, uint8_t *__restrict__ mask,
              float *__restrict__ grad) {
    for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < Q; tid += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(tid % nch);
        const int msi_idx = __ldg(cells + tid / nch);
        const int z = msi_idx % nlayers;
        const int tmp = msi_idx / nlayers;
        const int y = tmp % ly, x = tmp / ly;
        const int nx = (x == lx - 1) ? 0 : x + 1, ny = (y == ly - 1) ? 0 : y + 1;
        const int l00 = __ldg(links + x * ly + y), l01 = __ldg(links + x * ly + ny), l10 = __ldg(links + nx * ly + y);
        const float v00 = l00 >= 0 ? __ldg(msi + ((int64_t)l00 * nlayers + z) * nch + ch) : 0.f;
        const float v_nxl = (l00 >= 0 && z + 1 < nlayers) ? __ldg(msi + ((int64_t)l00 * nlayers + z + 1) * nch + ch)
                                                          : ((ch == nch - 1) ? 0.f : v00);
        const float v01 = l01 >= 0 ? __ldg(msi + ((int64_t)l01 * nlayers + z) * nch + ch) : 0.f;
        const float v10 = l10 >= 0 ? __ldg(msi + ((int64_t)l10 * nlayers + z) * nch + ch) : 0.f;
        const float sc = (ch == nch - 1) ? scale_last : scale;
        float dx = v10 - v00, dy = v01 - v00, dz = v_nxl - v00;
        const float idelta = sc * rsqrtf(1e-9f + dx * dx + dy * dy + dz * dz);
        dx *= lx * (1.f / 256.f);
        dy *= ly * (1.f / 256.f);
        dz *= nlayers * (1.f / 256.f);
        const float sm = -(dx + dy + dz);
        if (l00 >= 0 && sm != 0.f) {
            atomicAdd(grad + ((int64_t)l00 * nlayers + z) * nch + ch, sm * idelta);
            if (mask) mask[(int64_t)l00 * nlayers + z] = 1;
        }
        if (z + 1 < nlayers && l00 >= 0 && dz != 0.f) {
            atomicAdd(grad + ((int64_t)l00 * nlayers + z + 1) * nch + ch, dz * idelta);
            if (mask) mask[(int64_t)l00 * nlayers + z + 1] = 1;
        }
        if (l01 >= 0 && dy != 0.f) {
            atomicAdd(grad + ((int64_t)l01 * nlayers + z) * nch + ch, dy * idelta);
            if (mask) mask[(int64_t)l01 * nlayers + z] = 1;
        }
        if (l10 >= 0 && dx != 0.f) {
            atomicAdd(grad + ((int64_t)l10 * nlayers + z) * nch + ch, dx * idelta);
            if (mask) mask[(int64_t)l10 * nlayers + z] = 1;
        }
    }
}

int make_msi(const asurf_grid_t *grid, MsiP &m, const char *who) {
    ASURF_REQUIRE(grid && grid->background_links && grid->background_data, ASURF_E_INVALID, "%s: the grid has no background", who);
    ASURF_REQUIRE(grid->background_reso > 0 && grid->background_nlayers > 1, ASURF_E_INVALID,
                  "%s: the background needs at least 2 layers", who);
    m.links = grid->background_links;
    m.data = grid->background_data;
    m.reso = grid->background_reso;
    m.nlayers = grid->background_nlayers;
    for (int i = 0; i < 3; ++i) {
        m.size[i] = grid->size[i];
        m.offset[i] = grid->offset[i];
        m.scaling[i] = grid->scaling[i];
    }
    return 0;
}

Workspace g_ws_bg;   // (2, Q) floats: log-transmittance and leftover accum of the last foreground pass
int64_t g_bg_q = 0;

}  // namespace

int bg_state_reserve(int64_t Q, float **lt, float **accum) {
    int rc = g_ws_bg.reserve((size_t)Q * 2 * sizeof(float));
    if (rc) return rc;
    *lt = (float *)g_ws_bg.ptr;
    *accum = *lt + Q;
    g_bg_q = Q;
    return 0;
}
void msi_release() { g_ws_bg.release(); }

}  // namespace asurf

using namespace asurf;

extern "C" int asurf_msi_forward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                 const float *log_transmit, float *rgb_out, void *stream) {
    ASURF_REQUIRE(rays && opt && log_transmit && rgb_out, ASURF_E_INVALID, "msi_forward: null argument");
    const int64_t Q = rays->n_rays;
    if (Q <= 0) return 0;
    MsiP m;
    int rc = make_msi(grid, m, "msi_forward");
    if (rc) return rc;
    MsiCam cam = {};
    msi_forward_kernel<<<(int)((Q + 127) / 128), 128, 0, (cudaStream_t)stream>>>(m, *opt, rays->origins, rays->dirs, cam, Q,
                                                                                 log_transmit, rgb_out);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "msi_forward launch");
}

extern "C" int asurf_msi_forward_image(const asurf_grid_t *grid, const float *c2w_host, float fx, float fy, float cx, float cy,
                                       int32_t width, int32_t height, const asurf_opt_t *opt, const float *log_transmit,
                                       float *rgb_out, void *stream) {
    ASURF_REQUIRE(c2w_host && opt && log_transmit && rgb_out && width > 0 && height > 0, ASURF_E_INVALID,
                  "msi_forward_image: bad argument");
    MsiP m;
    int rc = make_msi(grid, m, "msi_forward_image");
    if (rc) return rc;
    MsiCam cam;
    for (int i = 0; i < 12; ++i) cam.c2w[i] = c2w_host[i];
    cam.fx = fx; cam.fy = fy; cam.cx = cx; cam.cy = cy;
    cam.width = width; cam.height = height;
    const int64_t Q = (int64_t)width * height;
    msi_forward_kernel<<<(int)((Q + 127) / 128), 128, 0, (cudaStream_t)stream>>>(m, *opt, nullptr, nullptr, cam, Q, log_transmit,
                                                                                 rgb_out);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "msi_forward_image launch");
}

extern "C" int asurf_msi_backward(const asurf_grid_t *grid, const asurf_rays_t *rays, const asurf_opt_t *opt,
                                  const float *grad_in, const float *color_cache, int32_t grad_is_rgb, int64_t norm_rays,
                                  const float *log_transmit, const float *accum, float beta_loss, float sparsity_loss,
                                  const asurf_grads_t *grads, void *stream) {
    ASURF_REQUIRE(rays && opt && grad_in && log_transmit && accum && grads, ASURF_E_INVALID, "msi_backward: null argument");
    ASURF_REQUIRE(color_cache || !grad_is_rgb, ASURF_E_INVALID, "msi_backward: the fused form needs the rendered colours");
    ASURF_REQUIRE(grads->grad_background, ASURF_E_INVALID, "msi_backward: null background gradient buffer");
    const int64_t Q = rays->n_rays;
    if (Q <= 0) return 0;
    MsiP m;
    int rc = make_msi(grid, m, "msi_backward");
    if (rc) return rc;
    const int64_t qn = norm_rays > 0 ? norm_rays : Q;
    msi_backward_kernel<<<(int)((Q + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        m, *opt, rays->origins, rays->dirs, Q, grad_in, color_cache, grad_is_rgb, 2.f / (float)(3 * (int)qn), log_transmit, accum,
        beta_loss, sparsity_loss, grads->grad_background, grads->mask_background);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "msi_backward launch");
}

extern "C" int asurf_msi_tv_grad_sparse(const int32_t *links, int32_t links_x, int32_t links_y, const float *msi,
                                        int32_t nlayers, int32_t n_channels, const int32_t *rand_cells, int64_t n_cells,
                                        uint8_t *mask_out, float scale, float scale_last, float *grad_msi, void *stream) {
    ASURF_REQUIRE(links && msi && grad_msi, ASURF_E_INVALID, "msi_tv_grad_sparse: null tensor");
    ASURF_REQUIRE(links_x > 0 && links_y > 0 && nlayers > 0 && n_channels > 0, ASURF_E_INVALID, "msi_tv_grad_sparse: bad shape");
    if (n_cells <= 0) return 0;
    ASURF_REQUIRE(rand_cells, ASURF_E_INVALID, "msi_tv_grad_sparse: null cell list");
    const int64_t Q = n_cells * n_channels;
    const float nl = (float)(int)n_cells;
    const int64_t want = (Q + 255) / 256;
    const int blocks = (int)(want < 148 * 32 ? want : 148 * 32);
    msi_tv_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(links, msi, links_x, links_y, nlayers, n_channels, rand_cells,
                                                            scale / nl, scale_last / nl, Q, mask_out, grad_msi);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "msi_tv_grad_sparse launch");
}

extern "C" int asurf_debug_bg_state(float *log_transmit_out, float *accum_out, int64_t n_rays) {
    ASURF_REQUIRE(g_ws_bg.ptr && n_rays == g_bg_q, ASURF_E_INVALID, "debug_bg_state: no foreground pass of that size has run");
    const float *lt = (const float *)g_ws_bg.ptr;
    if (log_transmit_out) ASURF_CUDA(cudaMemcpy(log_transmit_out, lt, (size_t)n_rays * sizeof(float), cudaMemcpyDeviceToDevice));
    if (accum_out) ASURF_CUDA(cudaMemcpy(accum_out, lt + n_rays, (size_t)n_rays * sizeof(float), cudaMemcpyDeviceToDevice));
    return 0;
}
