// alphasurf_b200: the compiled `svox2.csrc` module -- a thin PyTorch C++ shim over the C ABI of include/asurf.h.
//
// Same module surface as the reference's pybind extension (/root/reference/svox2/csrc/svox2.cpp:139-291): function names,
// positional argument order, spec classes with the same read / write fields.  Each function checks its tensors the way the
// reference's CHECK_INPUT / TORCH_CHECK do (RuntimeError), fills the POD blocks of asurf.h from tensor.data_ptr() and calls
// libasurf.so on the current CUDA stream; a non-zero return code becomes a RuntimeError (NotImplementedError for
// ASURF_E_UNSUPPORTED).  It holds no arithmetic and no kernels: `g++` builds it (alphasurf_b200/build_shim.py), `nvcc` is not
// involved.  alphasurf_b200/svox2_csrc.py is the same layer written with ctypes; the GPU parity suite currently runs through
// that one, this file is the drop-in a maintainer installs as svox2/csrc*.so.
// Every exported function runs inside an NVTX range named "svox2.csrc.<function>" (tracing).
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <map>
#include <string>
#include <tuple>

#include <nvtx3/nvToolsExt.h>   // header-only; ranges show up in Nsight Systems / Compute timelines, cost nothing otherwise

#include "asurf.h"

namespace py = pybind11;
using torch::Tensor;

namespace {

// NVTX range around an exported function: REG(fn) registers Traced<&fn>::call, which has fn's signature
template <auto F>
struct Traced;
template <typename R, typename... A, R (*F)(A...)>
struct Traced<F> {
    static const char *name;
    static R call(A... a) {
        struct Range {
            explicit Range(const char *n) { nvtxRangePushA(n); }
            ~Range() { nvtxRangePop(); }
        } range(name);
        return F(std::forward<A>(a)...);
    }
};
template <typename R, typename... A, R (*F)(A...)>
const char *Traced<F>::name = "svox2.csrc";

constexpr int BASIS_TYPE_SH = 1;            // include/data_spec.hpp:11-15
constexpr int SURFACE_TYPE_NONE = 100;      // include/data_spec.hpp:17-23

// ---- spec classes (include/data_spec.hpp:39-201; fields registered at svox2.cpp:209-290) -----------------------------
struct SparseGridSpec {
    Tensor density_data, surface_data, level_set_data, sh_data, links, _offset, _scaling, basis_data, background_links,
        background_data;
    int basis_dim = 0;
    int basis_type = BASIS_TYPE_SH;
    int surface_type = SURFACE_TYPE_NONE;
    float fake_sample_std = 1.f;
    float truncated_vol_render_a = 1.f;
};
struct CameraSpec {
    Tensor c2w;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    int width = 0, height = 0;
    float ndc_coeffx = -1.f, ndc_coeffy = -1.f;
};
struct RaysSpec {
    Tensor origins, dirs, masks;
};
struct RayVoxIntersecSpec {
    Tensor voxel_ls, vox_start_i, vox_num;
};
struct RenderOptions {
    float background_brightness = 1.f, step_size = 0.5f, sigma_thresh = 1e-8f, stop_thresh = 1e-7f, near_clip = 0.f;
    bool use_spheric_clip = false, last_sample_opaque = false, surf_fake_sample = false;
    float surf_fake_sample_min_vox_len = 0.f;
    bool limited_fake_sample = false, no_surf_grad_from_sh = false;
    int alpha_activation_type = 0;
    bool fake_sample_l_dist = true, fake_sample_normalize_surf = false, only_outward_intersect = false,
         truncated_vol_render = false;
    float trunc_vol_weight_min = 0.f;
};
struct GridOutputGrads {
    Tensor grad_density_out, grad_sh_out, grad_surface_out, grad_fake_sample_std_out, grad_basis_out, grad_background_out,
        mask_out, mask_background_out;
};

// ---- checks and conversions ---------------------------------------------------------------------------------------------
void check_input(const Tensor &t, const char *name) {
    TORCH_CHECK(t.defined() && t.is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}
void check_cpu_input(const Tensor &t, const char *name) {
    TORCH_CHECK(t.defined() && !t.is_cuda(), name, " must be a CPU tensor");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}
void check_f32(const Tensor &t, const char *name) { TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32"); }

[[noreturn]] void not_on_path(const std::string &what) {
    PyErr_SetString(PyExc_NotImplementedError, what.c_str());
    throw py::error_already_set();
}

void check_rc(int rc, const char *what) {
    if (rc == 0) return;
    const std::string msg = std::string(what) + " failed (code " + std::to_string(rc) + "): " + asurf_last_error();
    if (rc == ASURF_E_UNSUPPORTED) not_on_path(msg);
    TORCH_CHECK(false, msg);
}

void check_grid(const SparseGridSpec &g) {
    check_input(g.density_data, "density_data");
    check_input(g.sh_data, "sh_data");
    check_input(g.links, "links");
    if (g.surface_type != SURFACE_TYPE_NONE) {
        check_input(g.surface_data, "surface_data");
        check_input(g.level_set_data, "level_set_data");
    }
    check_cpu_input(g._offset, "_offset");
    check_cpu_input(g._scaling, "_scaling");
    TORCH_CHECK(g.density_data.dim() == 2 && g.sh_data.dim() == 2 && g.links.dim() == 3,
                "density_data / sh_data must be 2-D and links 3-D");
    TORCH_CHECK(g.links.scalar_type() == torch::kInt32, "links must be int32");
    check_f32(g.density_data, "density_data");
    check_f32(g.sh_data, "sh_data");
    if (g.background_links.defined() && g.background_links.numel() > 0) {   // MSI background (data_spec.hpp:47-48)
        check_input(g.background_links, "background_links");
        check_input(g.background_data, "background_data");
        TORCH_CHECK(g.background_links.dim() == 2 && g.background_links.scalar_type() == torch::kInt32,
                    "background_links must be a 2-D int32 tensor");
        TORCH_CHECK(g.background_data.dim() == 3 && g.background_data.size(2) == 4 &&
                        g.background_data.scalar_type() == torch::kFloat32,
                    "background_data must be a float32 (n, nlayers, 4) tensor");
    }
    if (g.basis_type != BASIS_TYPE_SH) not_on_path("only the SH basis is on the B200 hot path");
}
void check_rays(const RaysSpec &r) {
    check_input(r.origins, "origins");
    check_input(r.dirs, "dirs");
    if (r.masks.defined()) check_input(r.masks, "masks");
    check_f32(r.origins, "origins");
    check_f32(r.dirs, "dirs");
}
void check_grads(const GridOutputGrads &g) {
    const std::pair<const Tensor *, const char *> ts[] = {{&g.grad_density_out, "grad_density_out"},
                                                          {&g.grad_sh_out, "grad_sh_out"},
                                                          {&g.grad_surface_out, "grad_surface_out"},
                                                          {&g.grad_fake_sample_std_out, "grad_fake_sample_std_out"}};
    for (const auto &p : ts)
        if (p.first->defined()) {
            check_input(*p.first, p.second);
            check_f32(*p.first, p.second);
        }
    if (g.mask_out.defined() && g.mask_out.numel() > 0) check_input(g.mask_out, "mask_out");
}

void *stream_of(const Tensor &t) { return (void *)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }
void size3(const Tensor &links, int32_t *sz) {
    for (int i = 0; i < 3; ++i) sz[i] = (int32_t)links.size(i);
}
void host3(const Tensor &t, float *out) {
    const Tensor f = t.to(torch::kCPU, torch::kFloat32).contiguous().reshape({-1});
    TORCH_CHECK(f.numel() >= 3, "expected 3 values");
    for (int i = 0; i < 3; ++i) out[i] = f.data_ptr<float>()[i];
}

// Occupancy pyramid of `links`, cached per tensor object and version (rebuilt on in-place edits and for any new tensor,
// also one that landed on a recycled address).
struct AccelEntry {
    c10::weak_intrusive_ptr<c10::TensorImpl> owner;
    uint32_t version;
    Tensor accel;
};
std::map<c10::TensorImpl *, AccelEntry> g_accel;

Tensor accel_for(const Tensor &links) {
    c10::TensorImpl *key = links.unsafeGetTensorImpl();
    const uint32_t ver = links._version();
    auto it = g_accel.find(key);
    if (it != g_accel.end()) {
        auto alive = it->second.owner.lock();
        if (alive && alive.get() == key && it->second.version == ver) return it->second.accel;
    }
    int32_t sz[3];
    size3(links, sz);
    Tensor acc = torch::empty({asurf_accel_words(sz)}, links.options().dtype(torch::kInt64));
    check_rc(asurf_accel_build(links.data_ptr<int32_t>(), sz, (uint64_t *)acc.data_ptr<int64_t>(), stream_of(links)),
             "asurf_accel_build");
    if (g_accel.size() > 8) g_accel.clear();
    g_accel.erase(key);
    g_accel.emplace(key, AccelEntry{c10::weak_intrusive_ptr<c10::TensorImpl>(links.getIntrusivePtr()), ver, acc});
    return acc;
}

struct GridArg {
    asurf_grid_t g;
    Tensor keep;   // the pyramid must outlive the call
};
GridArg grid_t(const SparseGridSpec &s, bool need_accel) {
    GridArg a;
    asurf_grid_t &g = a.g;
    g = asurf_grid_t();
    g.links = s.links.data_ptr<int32_t>();
    size3(s.links, g.size);
    g.density = s.density_data.data_ptr<float>();
    const bool has_surf = s.surface_type != SURFACE_TYPE_NONE && s.surface_data.defined();
    g.surface = has_surf ? s.surface_data.data_ptr<float>() : nullptr;
    g.level_set = has_surf ? s.level_set_data.data_ptr<float>() : nullptr;
    g.level_set_num = has_surf ? (int32_t)s.level_set_data.size(0) : 0;
    g.sh = s.sh_data.data_ptr<float>();
    g.basis_dim = s.basis_dim;
    g.sh_dim = (int32_t)s.sh_data.size(1);
    g.capacity = s.density_data.size(0);
    host3(s._offset, g.offset);
    host3(s._scaling, g.scaling);
    g.fake_sample_std = s.fake_sample_std;
    g.truncated_vol_render_a = s.truncated_vol_render_a;
    if (s.background_links.defined() && s.background_links.numel() > 0) {
        g.background_links = s.background_links.data_ptr<int32_t>();
        g.background_data = s.background_data.data_ptr<float>();
        g.background_reso = (int32_t)s.background_links.size(1);
        g.background_nlayers = (int32_t)s.background_data.size(1);
    }
    if (need_accel) {
        a.keep = accel_for(s.links);
        g.accel = (const uint64_t *)a.keep.data_ptr<int64_t>();
    }
    return a;
}
asurf_rays_t rays_t(const RaysSpec &r) {
    asurf_rays_t o;
    o.origins = r.origins.data_ptr<float>();
    o.dirs = r.dirs.data_ptr<float>();
    o.n_rays = r.origins.size(0);
    return o;
}
asurf_opt_t opt_t(const RenderOptions &r) {
    asurf_opt_t o;
    o.background_brightness = r.background_brightness;
    o.step_size = r.step_size;
    o.sigma_thresh = r.sigma_thresh;
    o.stop_thresh = r.stop_thresh;
    o.near_clip = r.near_clip;
    o.use_spheric_clip = r.use_spheric_clip;
    o.last_sample_opaque = r.last_sample_opaque;
    o.surf_fake_sample = r.surf_fake_sample;
    o.surf_fake_sample_min_vox_len = r.surf_fake_sample_min_vox_len;
    o.limited_fake_sample = r.limited_fake_sample;
    o.no_surf_grad_from_sh = r.no_surf_grad_from_sh;
    o.alpha_activation_type = r.alpha_activation_type;
    o.fake_sample_l_dist = r.fake_sample_l_dist;
    o.fake_sample_normalize_surf = r.fake_sample_normalize_surf;
    o.only_outward_intersect = r.only_outward_intersect;
    o.truncated_vol_render = r.truncated_vol_render;
    o.trunc_vol_weight_min = r.trunc_vol_weight_min;
    return o;
}
asurf_grads_t grads_t(const GridOutputGrads &g) {
    asurf_grads_t o = asurf_grads_t();
    if (g.grad_density_out.defined()) o.grad_density = g.grad_density_out.data_ptr<float>();
    if (g.grad_surface_out.defined()) o.grad_surface = g.grad_surface_out.data_ptr<float>();
    if (g.grad_sh_out.defined()) o.grad_sh = g.grad_sh_out.data_ptr<float>();
    if (g.grad_fake_sample_std_out.defined() && g.grad_fake_sample_std_out.numel() > 0)
        o.grad_fake_sample_std = g.grad_fake_sample_std_out.data_ptr<float>();
    if (g.mask_out.defined() && g.mask_out.numel() > 0) o.mask = (uint8_t *)g.mask_out.data_ptr();
    if (g.grad_background_out.defined() && g.grad_background_out.numel() > 0)
        o.grad_background = g.grad_background_out.data_ptr<float>();
    if (g.mask_background_out.defined() && g.mask_background_out.numel() > 0)
        o.mask_background = (uint8_t *)g.mask_background_out.data_ptr();
    return o;
}
struct CamArg {
    float c2w[12];
};
CamArg cam_t(const CameraSpec &cam) {
    check_input(cam.c2w, "c2w");
    if (cam.ndc_coeffx > 0.f) not_on_path("NDC cameras are outside the B200 hot path");
    const Tensor h = cam.c2w.to(torch::kCPU, torch::kFloat32).contiguous();
    TORCH_CHECK(h.dim() == 2 && h.size(0) >= 3 && h.size(1) == 4, "c2w must be (3,4) or (4,4)");
    CamArg a;
    for (int i = 0; i < 12; ++i) a.c2w[i] = h.data_ptr<float>()[i];
    return a;
}
uint8_t *mask_ptr(const Tensor &m) {   // an empty tensor means "no mask" (loss_kernel.cu:1368)
    check_input(m, "mask_out");
    return (m.dim() > 0 && m.numel() > 0) ? (uint8_t *)m.data_ptr() : nullptr;
}

int64_t g_norm_rays = 0;   // set_loss_norm_rays: global batch size of a ray-sharded run (0: this call's ray count)

// ---- surf_trav renderer (render_lerp_kernel_surf_trav.cu:3596-4081) -----------------------------------------------------
Tensor volume_render_surf_trav(SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt) {
    check_grid(grid);
    check_rays(rays);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    Tensor out = torch::empty_like(rays.origins);
    GridArg g = grid_t(grid, true);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    check_rc(asurf_surf_trav_forward(&g.g, &r, &o, out.data_ptr<float>(), nullptr, stream_of(out)), "volume_render_surf_trav");
    return out;
}

// march counters of SURVEY.md 8(d) for one forward march (bench.py's algorithmic bytes): surf_trav / cuvol flavour
py::dict stats_dict(const Tensor &st) {
    const Tensor h = st.cpu();
    const int64_t *v = h.data_ptr<int64_t>();
    py::dict d;
    d["n_steps"] = v[0];
    d["n_skips"] = v[1];
    d["n_linked"] = v[2];
    d["n_active"] = v[3];
    d["n_samples"] = v[4];
    return d;
}
py::dict render_stats(SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt) {
    check_grid(grid);
    check_rays(rays);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    Tensor out = torch::empty_like(rays.origins);
    Tensor st = torch::zeros({6}, rays.origins.options().dtype(torch::kInt64));
    GridArg g = grid_t(grid, true);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    check_rc(asurf_surf_trav_forward(&g.g, &r, &o, out.data_ptr<float>(), (asurf_stats_t *)st.data_ptr<int64_t>(), stream_of(out)),
             "render_stats");
    return stats_dict(st);
}
py::dict cuvol_render_stats(SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt) {
    check_grid(grid);
    check_rays(rays);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    Tensor st = torch::zeros({6}, rays.origins.options().dtype(torch::kInt64));
    GridArg g = grid_t(grid, false);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    check_rc(asurf_cuvol_stats(&g.g, &r, &o, (asurf_stats_t *)st.data_ptr<int64_t>(), stream_of(st)), "cuvol_render_stats");
    return stats_dict(st);
}

void volume_render_surf_trav_backward(SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt, Tensor grad_out,
                                      Tensor color_cache, GridOutputGrads &grads) {
    check_grid(grid);
    check_rays(rays);
    check_grads(grads);
    check_input(grad_out, "grad_out");
    check_input(color_cache, "color_cache");
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    GridArg g = grid_t(grid, true);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    asurf_grads_t gr = grads_t(grads);
    check_rc(asurf_surf_trav_backward(&g.g, &r, &o, grad_out.data_ptr<float>(), color_cache.data_ptr<float>(), &gr,
                                      stream_of(grad_out)),
             "volume_render_surf_trav_backward");
}

void volume_render_surf_trav_fused(SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt, Tensor rgb_gt, float beta_loss,
                                   float sparsity_loss, float fused_surf_norm_reg_scale, bool /*fused_surf_norm_reg_con_check*/,
                                   bool /*fused_surf_norm_reg_ignore_empty*/, float lambda_l2, float lambda_l1,
                                   float lambda_l_dist, float lambda_l_entropy, bool no_norm_weight_l_entropy,
                                   float lambda_l_dist_a, float lambda_l_entropy_a, float lambda_l_samp_dist, float lambda_l_di,
                                   float l_di_alpha_thresh, float surf_sparse_alpha_thresh, float lambda_inplace_surf_sparse,
                                   float lambda_inwards_norm_loss, float lambda_conv_mode_samp, int l_dist_max_sample,
                                   Tensor rgb_out, GridOutputGrads &grads) {
    check_input(rgb_gt, "rgb_gt");
    check_input(rgb_out, "rgb_out");
    check_grid(grid);
    check_rays(rays);
    check_grads(grads);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    asurf_fused_t f = asurf_fused_t();
    f.beta_loss = beta_loss;
    f.sparsity_loss = sparsity_loss;
    f.fused_surf_norm_reg_scale = fused_surf_norm_reg_scale;
    f.lambda_l2 = lambda_l2;
    f.lambda_l1 = lambda_l1;
    f.lambda_l_dist = lambda_l_dist;
    f.lambda_l_entropy = lambda_l_entropy;
    f.no_norm_weight_l_entropy = no_norm_weight_l_entropy;
    f.lambda_l_dist_a = lambda_l_dist_a;
    f.lambda_l_entropy_a = lambda_l_entropy_a;
    f.lambda_l_samp_dist = lambda_l_samp_dist;
    f.lambda_l_di = lambda_l_di;
    f.l_di_alpha_thresh = l_di_alpha_thresh;
    f.surf_sparse_alpha_thresh = surf_sparse_alpha_thresh;
    f.lambda_inplace_surf_sparse = lambda_inplace_surf_sparse;
    f.lambda_inwards_norm_loss = lambda_inwards_norm_loss;
    f.lambda_conv_mode_samp = lambda_conv_mode_samp;
    f.l_dist_max_sample = l_dist_max_sample;
    f.norm_rays = g_norm_rays;
    GridArg g = grid_t(grid, true);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    asurf_grads_t gr = grads_t(grads);
    check_rc(asurf_surf_trav_fused(&g.g, &r, &o, rgb_gt.data_ptr<float>(), &f, rgb_out.data_ptr<float>(), &gr, nullptr,
                                   stream_of(rgb_out)),
             "volume_render_surf_trav_fused");
}

std::tuple<Tensor, Tensor> surf_trav_scalar(const char *name, SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt,
                                            int mode, float param, int64_t width, int max_sample) {
    check_grid(grid);
    check_rays(rays);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    const int64_t Q = rays.origins.size(0);
    Tensor out = width == 1 ? torch::empty({Q}, rays.origins.options()) : torch::empty({Q, width}, rays.origins.options());
    Tensor out2 = mode == ASURF_SCALAR_EXTRACT_PTS ? torch::empty_like(out) : Tensor();
    GridArg g = grid_t(grid, true);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    check_rc(asurf_surf_trav_scalar(&g.g, &r, &o, mode, param, max_sample, out.data_ptr<float>(),
                                    out2.defined() ? out2.data_ptr<float>() : nullptr, stream_of(out)),
             name);
    return {out, out2};
}
Tensor volume_render_expected_term_surf_trav(SparseGridSpec &g, RaysSpec &r, RenderOptions &o) {
    return std::get<0>(surf_trav_scalar("volume_render_expected_term_surf_trav", g, r, o, ASURF_SCALAR_EXPECTED_TERM, 0.f, 1, 0));
}
Tensor volume_render_mode_term_surf_trav(SparseGridSpec &g, RaysSpec &r, RenderOptions &o, float weight_thresh) {
    return std::get<0>(surf_trav_scalar("volume_render_mode_term_surf_trav", g, r, o, ASURF_SCALAR_MODE_TERM, weight_thresh, 1, 0));
}
Tensor volume_render_sigma_thresh_surf_trav(SparseGridSpec &g, RaysSpec &r, RenderOptions &o, float sigma_thresh) {
    return std::get<0>(surf_trav_scalar("volume_render_sigma_thresh_surf_trav", g, r, o, ASURF_SCALAR_THRESH_DEPTH, sigma_thresh, 1, 0));
}
Tensor volume_render_alpha_surf_trav(SparseGridSpec &g, RaysSpec &r, RenderOptions &o, float thresh) {
    return std::get<0>(surf_trav_scalar("volume_render_alpha_surf_trav", g, r, o, ASURF_SCALAR_THRESH_ALPHA, thresh, 1, 0));
}
Tensor render_normal_surf_trav(SparseGridSpec &g, RaysSpec &r, RenderOptions &o) {
    return std::get<0>(surf_trav_scalar("render_normal_surf_trav", g, r, o, ASURF_SCALAR_NORMAL, 0.f, 3, 0));
}
std::tuple<Tensor, Tensor> extract_pts_surf_trav(SparseGridSpec &g, RaysSpec &r, RenderOptions &o, int max_sample,
                                                 float alpha_thresh) {
    if (max_sample <= 0) {
        Tensor z = torch::zeros({r.origins.size(0), 0}, r.origins.options());
        return {z, z.clone()};
    }
    return surf_trav_scalar("extract_pts_surf_trav", g, r, o, ASURF_SCALAR_EXTRACT_PTS, alpha_thresh, max_sample, max_sample);
}

// ---- cuvol renderer (render_lerp_kernel_cuvol.cu:1120-1442) ---------------------------------------------------------------
Tensor volume_render_cuvol(SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt) {
    check_grid(grid);
    check_rays(rays);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    Tensor out = torch::empty_like(rays.origins);
    GridArg g = grid_t(grid, false);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    check_rc(asurf_cuvol_forward(&g.g, &r, &o, out.data_ptr<float>(), nullptr, stream_of(out)), "volume_render_cuvol");
    return out;
}
Tensor volume_render_cuvol_image(SparseGridSpec &grid, CameraSpec &cam, RenderOptions &opt) {
    check_grid(grid);
    const CamArg c = cam_t(cam);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    Tensor out = torch::empty({cam.height, cam.width, 3}, grid.sh_data.options());
    GridArg g = grid_t(grid, false);
    asurf_opt_t o = opt_t(opt);
    check_rc(asurf_cuvol_image(&g.g, c.c2w, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, &o, out.data_ptr<float>(),
                               stream_of(out)),
             "volume_render_cuvol_image");
    return out;
}
void volume_render_cuvol_backward(SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt, Tensor grad_out, Tensor color_cache,
                                  GridOutputGrads &grads) {
    check_grid(grid);
    check_rays(rays);
    check_grads(grads);
    check_input(grad_out, "grad_out");
    check_input(color_cache, "color_cache");
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    GridArg g = grid_t(grid, false);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    asurf_grads_t gr = grads_t(grads);
    check_rc(asurf_cuvol_backward(&g.g, &r, &o, grad_out.data_ptr<float>(), color_cache.data_ptr<float>(), &gr,
                                  stream_of(grad_out)),
             "volume_render_cuvol_backward");
}
void volume_render_cuvol_fused(SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt, Tensor rgb_gt, float beta_loss,
                               float sparsity_loss, Tensor rgb_out, GridOutputGrads &grads) {
    check_input(rgb_gt, "rgb_gt");
    check_input(rgb_out, "rgb_out");
    check_grid(grid);
    check_rays(rays);
    check_grads(grads);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    GridArg g = grid_t(grid, false);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    asurf_grads_t gr = grads_t(grads);
    check_rc(asurf_cuvol_fused(&g.g, &r, &o, rgb_gt.data_ptr<float>(), beta_loss, sparsity_loss, g_norm_rays,
                               rgb_out.data_ptr<float>(), &gr, stream_of(rgb_out)),
             "volume_render_cuvol_fused");
}
std::tuple<Tensor, Tensor> cuvol_scalar(const char *name, SparseGridSpec &grid, RaysSpec &rays, RenderOptions &opt, int mode,
                                        float param, int max_sample) {
    check_grid(grid);
    check_rays(rays);
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    const int64_t Q = rays.origins.size(0);
    const bool med = mode == ASURF_CUVOL_MED_TERM;
    Tensor out = med ? torch::empty({Q, max_sample}, rays.origins.options()) : torch::empty({Q}, rays.origins.options());
    Tensor out2 = med ? torch::empty({Q, max_sample}, rays.origins.options()) : Tensor();
    GridArg g = grid_t(grid, false);
    asurf_rays_t r = rays_t(rays);
    asurf_opt_t o = opt_t(opt);
    check_rc(asurf_cuvol_scalar(&g.g, &r, &o, mode, param, max_sample, out.data_ptr<float>(),
                                med ? out2.data_ptr<float>() : nullptr, stream_of(out)),
             name);
    return {out, out2};
}
Tensor volume_render_expected_term(SparseGridSpec &g, RaysSpec &r, RenderOptions &o, float weight_thresh) {
    return std::get<0>(cuvol_scalar("volume_render_expected_term", g, r, o, ASURF_CUVOL_EXPECTED_TERM, weight_thresh, 0));
}
Tensor volume_render_mode_term(SparseGridSpec &g, RaysSpec &r, RenderOptions &o, float weight_thresh) {
    return std::get<0>(cuvol_scalar("volume_render_mode_term", g, r, o, ASURF_CUVOL_MODE_TERM, weight_thresh, 0));
}
std::tuple<Tensor, Tensor> volume_render_med_term(SparseGridSpec &g, RaysSpec &r, RenderOptions &o, int max_sample) {
    return cuvol_scalar("volume_render_med_term", g, r, o, ASURF_CUVOL_MED_TERM, 0.f, max_sample);
}
Tensor volume_render_sigma_thresh(SparseGridSpec &g, RaysSpec &r, RenderOptions &o, float sigma_thresh) {
    return std::get<0>(cuvol_scalar("volume_render_sigma_thresh", g, r, o, ASURF_CUVOL_SIGMA_THRESH, sigma_thresh, 0));
}

// ---- grid maintenance (misc_kernel.cu:1005-1175) ----------------------------------------------------------------------------
void accel_dist_prop(Tensor links) {
    check_input(links, "grid");
    TORCH_CHECK(!links.is_floating_point() && links.dim() == 3, "accel_dist_prop expects the 3-D integer links tensor");
    TORCH_CHECK(links.scalar_type() == torch::kInt32, "links must be int32");
    const c10::cuda::CUDAGuard guard(links.device());
    int32_t sz[3];
    size3(links, sz);
    check_rc(asurf_accel_dist_prop(links.data_ptr<int32_t>(), sz, stream_of(links)), "accel_dist_prop");
    links.add_(0);   // bump the version counter: cached pyramids of this tensor are rebuilt
}
Tensor dilate(Tensor grid) {
    check_input(grid, "grid");
    TORCH_CHECK(!grid.is_floating_point() && grid.dim() == 3, "dilate expects a 3-D non-floating tensor");
    TORCH_CHECK(grid.scalar_type() == torch::kBool, "dilate expects a bool tensor");
    const c10::cuda::CUDAGuard guard(grid.device());
    Tensor out = torch::empty_like(grid);
    int32_t sz[3];
    size3(grid, sz);
    check_rc(asurf_dilate((const uint8_t *)grid.data_ptr(), sz, (uint8_t *)out.data_ptr(), stream_of(grid)), "dilate");
    return out;
}
void grid_weight_render(Tensor data, CameraSpec &cam, float step_size, float stop_thresh, bool last_sample_opaque, Tensor offset,
                        Tensor scaling, Tensor grid_weight_out) {
    check_input(data, "data");
    check_input(offset, "offset");
    check_input(scaling, "scaling");
    check_input(grid_weight_out, "grid_weight_out");
    TORCH_CHECK(data.dim() == 3 && grid_weight_out.sizes() == data.sizes(),
                "grid_weight_render expects (X,Y,Z) data and an output of the same shape");
    const CamArg c = cam_t(cam);
    const c10::cuda::CUDAGuard guard(data.device());
    int32_t sz[3];
    size3(data, sz);
    float off[3], scl[3];
    host3(offset, off);
    host3(scaling, scl);
    check_rc(asurf_grid_weight_render(data.data_ptr<float>(), sz, off, scl, c.c2w, cam.fx, cam.fy, cam.cx, cam.cy, cam.width,
                                      cam.height, step_size, stop_thresh, last_sample_opaque, grid_weight_out.data_ptr<float>(),
                                      stream_of(data)),
             "grid_weight_render");
}
void sparse_grid_weight_render(SparseGridSpec &grid, CameraSpec &cam, float step_size, float stop_thresh, Tensor offset,
                               Tensor scaling, Tensor grid_weight_out) {
    check_grid(grid);
    check_input(offset, "offset");
    check_input(scaling, "scaling");
    check_input(grid_weight_out, "grid_weight_out");
    TORCH_CHECK(grid_weight_out.sizes() == grid.links.sizes(), "sparse_grid_weight_render expects an output shaped like links");
    const CamArg c = cam_t(cam);
    const c10::cuda::CUDAGuard guard(grid.links.device());
    int32_t sz[3];
    size3(grid.links, sz);
    float off[3], scl[3];
    host3(offset, off);
    host3(scaling, scl);
    check_rc(asurf_sparse_grid_weight_render(grid.links.data_ptr<int32_t>(), grid.density_data.data_ptr<float>(), sz, off, scl,
                                             c.c2w, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, step_size, stop_thresh,
                                             grid_weight_out.data_ptr<float>(), stream_of(grid.links)),
             "sparse_grid_weight_render");
}
void sparse_grid_mask_render(SparseGridSpec &grid, RaysSpec &rays, float near_clip, Tensor grid_mask) {
    check_grid(grid);
    check_rays(rays);
    check_input(grid_mask, "grid_mask");
    const c10::cuda::CUDAGuard guard(grid.links.device());
    int32_t sz[3];
    size3(grid.links, sz);
    float off[3], scl[3];
    host3(grid._offset, off);
    host3(grid._scaling, scl);
    check_rc(asurf_sparse_grid_mask_render(grid.links.data_ptr<int32_t>(), sz, off, scl, rays.origins.data_ptr<float>(),
                                           rays.dirs.data_ptr<float>(), rays.origins.size(0), near_clip,
                                           grid_mask.data_ptr<float>(), stream_of(grid.links)),
             "sparse_grid_mask_render");
}
void sparse_grid_visbility_render_surf(SparseGridSpec &grid, CameraSpec &cam, Tensor visibility_out) {
    check_grid(grid);
    check_input(visibility_out, "visibility_out");
    TORCH_CHECK(grid.surface_type != SURFACE_TYPE_NONE && grid.surface_data.defined(),
                "sparse_grid_visbility_render_surf needs a grid with surface data");
    const CamArg c = cam_t(cam);
    const c10::cuda::CUDAGuard guard(grid.links.device());
    int32_t sz[3];
    size3(grid.links, sz);
    float off[3], scl[3];
    host3(grid._offset, off);
    host3(grid._scaling, scl);
    check_rc(asurf_sparse_grid_visibility_render_surf(
                 grid.links.data_ptr<int32_t>(), grid.surface_data.data_ptr<float>(), grid.level_set_data.data_ptr<float>(),
                 (int32_t)grid.level_set_data.size(0), sz, off, scl, c.c2w, cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height,
                 visibility_out.data_ptr<float>(), stream_of(grid.links)),
             "sparse_grid_visbility_render_surf");
}

// ---- point queries (svox2_kernel.cu:384-582) -----------------------------------------------------------------------------------
void check_points(const Tensor &p) {
    check_input(p, "points");
    TORCH_CHECK(p.dim() == 2 && p.size(1) == 3, "points must be (P, 3)");
}
Tensor sample_one(const SparseGridSpec &grid, const Tensor &data, float missing, const Tensor &points, const char *name) {
    Tensor out = torch::empty({points.size(0), data.size(1)}, points.options());
    int32_t sz[3];
    size3(grid.links, sz);
    float off[3], scl[3];
    host3(grid._offset, off);
    host3(grid._scaling, scl);
    check_rc(asurf_sample_grid(grid.links.data_ptr<int32_t>(), sz, off, scl, data.data_ptr<float>(), (int32_t)data.size(1), missing,
                               points.data_ptr<float>(), points.size(0), out.data_ptr<float>(), stream_of(points)),
             name);
    return out;
}
std::tuple<Tensor, Tensor> sample_grid(SparseGridSpec &grid, Tensor points, bool want_colors) {
    check_grid(grid);
    check_points(points);
    const c10::cuda::CUDAGuard guard(points.device());
    Tensor dens = sample_one(grid, grid.density_data, 0.f, points, "sample_grid");
    Tensor sh = want_colors ? sample_one(grid, grid.sh_data, 0.f, points, "sample_grid")
                            : torch::empty({0, grid.sh_data.size(1)}, points.options());
    return {dens, sh};
}
std::tuple<Tensor, Tensor> sample_grid_sh_surf(SparseGridSpec &grid, Tensor points, bool want_colors, bool want_surfaces,
                                               float default_surf) {
    check_grid(grid);
    check_points(points);
    const c10::cuda::CUDAGuard guard(points.device());
    Tensor sh = want_colors ? sample_one(grid, grid.sh_data, 0.f, points, "sample_grid_sh_surf")
                            : torch::empty({0, grid.sh_data.size(1)}, points.options());
    Tensor surf = want_surfaces ? sample_one(grid, grid.surface_data, default_surf, points, "sample_grid_sh_surf")
                                : torch::empty({0, grid.surface_data.size(1)}, points.options());
    return {sh, surf};
}
Tensor sample_grid_raw_alpha(SparseGridSpec &grid, Tensor points, float empty_raw) {
    check_grid(grid);
    check_points(points);
    const c10::cuda::CUDAGuard guard(points.device());
    return sample_one(grid, grid.density_data, empty_raw, points, "sample_grid_raw_alpha");
}
void sample_grid_backward(SparseGridSpec &grid, Tensor points, Tensor grad_out_density, Tensor grad_out_sh, Tensor grad_density_out,
                          Tensor grad_sh_out, bool want_colors) {
    check_grid(grid);
    check_points(points);
    check_input(grad_out_density, "grad_out_density");
    check_input(grad_out_sh, "grad_out_sh");
    check_input(grad_density_out, "grad_density_out");
    check_input(grad_sh_out, "grad_sh_out");
    TORCH_CHECK(grad_out_density.dim() == 2 && grad_out_sh.dim() == 2, "sample_grid_backward expects 2-D output gradients");
    const c10::cuda::CUDAGuard guard(points.device());
    int32_t sz[3];
    size3(grid.links, sz);
    float off[3], scl[3];
    host3(grid._offset, off);
    host3(grid._scaling, scl);
    check_rc(asurf_sample_grid_backward(grid.links.data_ptr<int32_t>(), sz, off, scl, points.data_ptr<float>(), points.size(0),
                                        grad_out_density.data_ptr<float>(), (int32_t)grad_density_out.size(1),
                                        grad_density_out.data_ptr<float>(), stream_of(points)),
             "sample_grid_backward");
    if (want_colors)
        check_rc(asurf_sample_grid_backward(grid.links.data_ptr<int32_t>(), sz, off, scl, points.data_ptr<float>(), points.size(0),
                                            grad_out_sh.data_ptr<float>(), (int32_t)grad_sh_out.size(1),
                                            grad_sh_out.data_ptr<float>(), stream_of(points)),
                 "sample_grid_backward");
}
Tensor cubic_extract_iso_pts(Tensor links, Tensor level_data, Tensor mask_data, Tensor cell_ids, int n_sample,
                             float density_thresh) {
    check_input(level_data, "level_data");
    check_input(mask_data, "mask_data");
    check_input(links, "links");
    check_input(cell_ids, "cell_ids");
    TORCH_CHECK(links.scalar_type() == torch::kInt32 && cell_ids.scalar_type() == torch::kInt32, "links and cell_ids must be int32");
    const c10::cuda::CUDAGuard guard(level_data.device());
    Tensor out = torch::zeros({cell_ids.size(0), 3 * (int64_t)n_sample * n_sample, 3}, level_data.options());
    int32_t sz[3];
    size3(links, sz);
    check_rc(asurf_cubic_extract_iso_pts(links.data_ptr<int32_t>(), sz, level_data.data_ptr<float>(), mask_data.data_ptr<float>(),
                                         cell_ids.data_ptr<int32_t>(), cell_ids.size(0), n_sample, density_thresh,
                                         out.data_ptr<float>(), stream_of(level_data)),
             "cubic_extract_iso_pts");
    return out;
}

// ---- optimizer steps (optim_kernel.cu:154-267) -------------------------------------------------------------------------------
struct Indexer {
    int kind;
    const void *ptr;
    int64_t n;
};
Indexer indexer_of(const Tensor &ix) {   // dispatch of optim_kernel.cu:175-215
    check_input(ix, "indexer");
    if (ix.dim() == 0) return {0, nullptr, -1};
    if (ix.size(0) == 0) return {1, nullptr, 0};
    if (ix.scalar_type() == torch::kBool) return {1, ix.data_ptr(), ix.size(0)};
    TORCH_CHECK(ix.scalar_type() == torch::kInt64, "indexer must be a bool mask or an int64 index list");
    return {2, ix.data_ptr(), ix.size(0)};
}
void rmsprop_step(Tensor data, Tensor rms, Tensor grad, Tensor indexer, float beta, float lr, float epsilon, float minval,
                  float lr_last) {
    check_input(data, "data");
    check_input(rms, "rms");
    check_input(grad, "grad");
    const Indexer ix = indexer_of(indexer);
    if (ix.n == 0) return;
    const c10::cuda::CUDAGuard guard(data.device());
    check_rc(asurf_rmsprop_step(data.data_ptr<float>(), rms.data_ptr<float>(), grad.data_ptr<float>(), data.size(0),
                                (int32_t)data.size(1), ix.kind, ix.ptr, ix.n > 0 ? ix.n : 0, beta, lr, epsilon, minval, lr_last,
                                stream_of(data)),
             "rmsprop_step");
}
void sgd_step(Tensor data, Tensor grad, Tensor indexer, float lr, float lr_last) {
    check_input(data, "data");
    check_input(grad, "grad");
    const Indexer ix = indexer_of(indexer);
    if (ix.n == 0) return;
    const c10::cuda::CUDAGuard guard(data.device());
    check_rc(asurf_sgd_step(data.data_ptr<float>(), grad.data_ptr<float>(), data.size(0), (int32_t)data.size(1), ix.kind, ix.ptr,
                            ix.n > 0 ? ix.n : 0, lr, lr_last, stream_of(data)),
             "sgd_step");
}

// ---- grid-side regularisers (loss_kernel.cu:1214-1622) ----------------------------------------------------------------------
void check_loss_common(const Tensor &links, const Tensor &data, const Tensor *grad_data) {
    check_input(data, "data");
    check_input(links, "links");
    if (grad_data) {
        check_input(*grad_data, "grad_data");
        TORCH_CHECK(grad_data->is_floating_point() && grad_data->dim() == 2, "grad_data must be a 2-D floating point tensor");
    }
    TORCH_CHECK(data.is_floating_point() && !links.is_floating_point() && data.dim() == 2 && links.dim() == 3,
                "data must be a 2-D floating point tensor and links a 3-D integer tensor");
    check_f32(data, "data");
    TORCH_CHECK(links.scalar_type() == torch::kInt32, "links must be int32");
}
void check_cells(const Tensor &c) {
    check_input(c, "rand_cells");
    TORCH_CHECK(c.scalar_type() == torch::kInt32, "rand_cells must be int32");
}
Tensor tv(Tensor links, Tensor data, int start_dim, int end_dim, bool, float, bool ignore_edge, float, float) {
    check_loss_common(links, data, nullptr);
    const c10::cuda::CUDAGuard guard(data.device());
    Tensor out = torch::zeros({}, data.options());
    int32_t sz[3];
    size3(links, sz);
    check_rc(asurf_tv(links.data_ptr<int32_t>(), sz, data.data_ptr<float>(), (int32_t)data.size(1), start_dim, end_dim, ignore_edge,
                      out.data_ptr<float>(), stream_of(data)),
             "tv");
    return out;
}
void tv_grad(Tensor links, Tensor data, int start_dim, int end_dim, float scale, bool, float, bool ignore_edge, float, float,
             Tensor grad_data) {
    check_loss_common(links, data, &grad_data);
    const c10::cuda::CUDAGuard guard(data.device());
    int32_t sz[3];
    size3(links, sz);
    check_rc(asurf_tv_grad(links.data_ptr<int32_t>(), sz, data.data_ptr<float>(), (int32_t)data.size(1), start_dim, end_dim, scale,
                           ignore_edge, grad_data.data_ptr<float>(), stream_of(data)),
             "tv_grad");
}
void tv_grad_sparse(Tensor links, Tensor data, Tensor rand_cells, Tensor mask_out, int start_dim, int end_dim, float scale, bool,
                    float, bool ignore_edge, bool ignore_last_z, float, float, Tensor grad_data) {
    check_loss_common(links, data, &grad_data);
    check_cells(rand_cells);
    const c10::cuda::CUDAGuard guard(data.device());
    int32_t sz[3];
    size3(links, sz);
    check_rc(asurf_tv_grad_sparse(links.data_ptr<int32_t>(), sz, data.data_ptr<float>(), (int32_t)data.size(1),
                                  rand_cells.data_ptr<int32_t>(), rand_cells.size(0), mask_ptr(mask_out), start_dim, end_dim, scale,
                                  ignore_edge, ignore_last_z, grad_data.data_ptr<float>(), stream_of(data)),
             "tv_grad_sparse");
}
void surf_tv_grad_sparse(Tensor links, Tensor data, Tensor density_data, Tensor rand_cells, Tensor mask_out, int start_dim,
                         int end_dim, float scale, bool ignore_edge, float edge_value, bool ignore_last_z, float, float,
                         bool alpha_dependency, Tensor grad_data) {
    check_loss_common(links, data, &grad_data);
    check_input(density_data, "density_data");
    check_cells(rand_cells);
    const c10::cuda::CUDAGuard guard(data.device());
    int32_t sz[3];
    size3(links, sz);
    Tensor acc = accel_for(links);
    check_rc(asurf_surf_tv_grad_sparse(links.data_ptr<int32_t>(), sz, data.data_ptr<float>(), (int32_t)data.size(1),
                                       density_data.data_ptr<float>(), (int32_t)density_data.size(1),
                                       rand_cells.data_ptr<int32_t>(), rand_cells.size(0), mask_ptr(mask_out), start_dim, end_dim,
                                       scale, ignore_edge, edge_value, ignore_last_z, alpha_dependency,
                                       grad_data.data_ptr<float>(), (const uint64_t *)acc.data_ptr<int64_t>(), stream_of(data)),
             "surf_tv_grad_sparse");
}
void alpha_surf_sparsify_grad_sparse(Tensor links, Tensor alpha_data, Tensor surf_data, Tensor rand_cells, Tensor mask_out,
                                     float scale_alpha, float scale_surf, bool surf_sparse_decrease, float surf_sparse_thresh,
                                     float alpha_bound, float surf_bound, Tensor grad_alpha, Tensor grad_surf) {
    check_loss_common(links, alpha_data, &grad_alpha);
    check_input(surf_data, "surf_data");
    check_input(grad_surf, "grad_surf");
    check_cells(rand_cells);
    const c10::cuda::CUDAGuard guard(alpha_data.device());
    int32_t sz[3];
    size3(links, sz);
    check_rc(asurf_alpha_surf_sparsify_grad_sparse(
                 links.data_ptr<int32_t>(), sz, alpha_data.data_ptr<float>(), (int32_t)alpha_data.size(1),
                 surf_data.data_ptr<float>(), (int32_t)surf_data.size(1), rand_cells.data_ptr<int32_t>(), rand_cells.size(0),
                 mask_ptr(mask_out), scale_alpha, scale_surf, surf_sparse_decrease, surf_sparse_thresh, alpha_bound, surf_bound,
                 grad_alpha.data_ptr<float>(), grad_surf.data_ptr<float>(), stream_of(alpha_data)),
             "alpha_surf_sparsify_grad_sparse");
}
void surface_normal_grad_sparse(Tensor links, Tensor data, Tensor rand_cells, Tensor mask_out, float lv_set, int start_dim,
                                int end_dim, float scale, float /*eikonal_scale*/, float, float, bool con_check, bool ignore_empty,
                                bool use_l1, Tensor grad_data) {
    check_loss_common(links, data, &grad_data);
    check_cells(rand_cells);
    const c10::cuda::CUDAGuard guard(data.device());
    int32_t sz[3];
    size3(links, sz);
    Tensor acc = accel_for(links);
    check_rc(asurf_surface_normal_grad_sparse(links.data_ptr<int32_t>(), sz, data.data_ptr<float>(), rand_cells.data_ptr<int32_t>(),
                                              rand_cells.size(0), mask_ptr(mask_out), lv_set, start_dim, end_dim, scale, con_check,
                                              ignore_empty, use_l1, grad_data.data_ptr<float>(),
                                              (const uint64_t *)acc.data_ptr<int64_t>(), stream_of(data)),
             "surface_normal_grad_sparse");
}

// surf_sign_change_grad_sparse, loss_kernel.cu:1429-1466
void surf_sign_change_grad_sparse(Tensor links, Tensor data, Tensor rand_cells, Tensor mask_out, int start_dim, int end_dim,
                                  float scale, Tensor grad_data) {
    check_loss_common(links, data, &grad_data);
    check_cells(rand_cells);
    const c10::cuda::CUDAGuard guard(data.device());
    int32_t sz[3];
    size3(links, sz);
    check_rc(asurf_surf_sign_change_grad_sparse(links.data_ptr<int32_t>(), sz, data.data_ptr<float>(), (int32_t)data.size(1),
                                                rand_cells.data_ptr<int32_t>(), rand_cells.size(0), mask_ptr(mask_out), start_dim,
                                                end_dim, scale, grad_data.data_ptr<float>(), stream_of(data)),
             "surf_sign_change_grad_sparse");
}

// surface_normal_grad (dense), loss_kernel.cu:1289-1325 (the ndc coefficients are unused there too)
void surface_normal_grad(Tensor links, Tensor data, float lv_set, int start_dim, int end_dim, float scale, float ndc_coeffx,
                         float ndc_coeffy, Tensor grad_data) {
    check_loss_common(links, data, &grad_data);
    const c10::cuda::CUDAGuard guard(data.device());
    int32_t sz[3];
    size3(links, sz);
    check_rc(asurf_surface_normal_grad(links.data_ptr<int32_t>(), sz, data.data_ptr<float>(), (int32_t)data.size(1), lv_set,
                                       start_dim, end_dim, scale, grad_data.data_ptr<float>(), stream_of(data)),
             "surface_normal_grad");
}

// lumisphere_tv_grad_sparse, loss_kernel.cu:1661-1697
void lumisphere_tv_grad_sparse(SparseGridSpec &grid, Tensor rand_cells, Tensor basis_fn, Tensor basis_fn_u, float scale,
                               float ndc_coeffx, float ndc_coeffy, float dir_factor, GridOutputGrads &grads) {
    check_input(grid.sh_data, "grid.sh_data");
    check_input(grid.links, "grid.links");
    check_cells(rand_cells);
    check_input(basis_fn, "basis_fn");
    check_input(basis_fn_u, "basis_fn_u");
    check_f32(basis_fn, "basis_fn");
    check_f32(basis_fn_u, "basis_fn_u");
    TORCH_CHECK(basis_fn.dim() == 1, "basis_fn must be 1-D");
    TORCH_CHECK(grads.grad_sh_out.defined(), "grads.grad_sh_out is required");
    check_input(grads.grad_sh_out, "grads.grad_sh_out");
    TORCH_CHECK(basis_fn.numel() >= grid.basis_dim && basis_fn_u.numel() >= grid.basis_dim,
                "basis_fn / basis_fn_u hold fewer than basis_dim values");
    const c10::cuda::CUDAGuard guard(grid.sh_data.device());
    int32_t sz[3];
    size3(grid.links, sz);
    check_rc(asurf_lumisphere_tv_grad_sparse(grid.links.data_ptr<int32_t>(), sz, grid.sh_data.data_ptr<float>(),
                                             (int32_t)grid.sh_data.size(1), grid.basis_dim, rand_cells.data_ptr<int32_t>(),
                                             rand_cells.size(0), basis_fn.data_ptr<float>(), basis_fn_u.data_ptr<float>(), scale,
                                             dir_factor, grads.mask_out.defined() ? mask_ptr(grads.mask_out) : nullptr,
                                             grads.grad_sh_out.data_ptr<float>(), stream_of(grid.sh_data)),
             "lumisphere_tv_grad_sparse");
}

// msi_tv_grad_sparse, loss_kernel.cu:1624-1659
void msi_tv_grad_sparse(Tensor links, Tensor msi, Tensor rand_cells, Tensor mask_out, float scale, float scale_last,
                        Tensor grad_msi) {
    check_input(links, "links");
    check_input(msi, "msi");
    check_input(grad_msi, "grad_msi");
    check_cells(rand_cells);
    check_input(mask_out, "mask_out");
    TORCH_CHECK(msi.is_floating_point() && grad_msi.is_floating_point() && msi.dim() == 3 && links.dim() == 2,
                "msi must be a (n, nlayers, channels) floating point tensor and links 2-D");
    const c10::cuda::CUDAGuard guard(msi.device());
    uint8_t *mp = (mask_out.numel() > 0) ? (uint8_t *)mask_out.data_ptr() : nullptr;
    check_rc(asurf_msi_tv_grad_sparse(links.data_ptr<int32_t>(), (int32_t)links.size(0), (int32_t)links.size(1),
                                      msi.data_ptr<float>(), (int32_t)msi.size(1), (int32_t)msi.size(2),
                                      rand_cells.data_ptr<int32_t>(), rand_cells.size(0), mp, scale, scale_last,
                                      grad_msi.data_ptr<float>(), stream_of(msi)),
             "msi_tv_grad_sparse");
}

// ---- names of svox2.cpp that are outside the hot path: present (the reference's Python probes by name), never silent ----------
py::object off_path(const std::string &name) {
    return py::cpp_function(
        [name](py::args, py::kwargs) -> py::object {
            not_on_path("svox2.csrc." + name + " is outside the B200 hot path of alphasurf_b200 (SURVEY.md 8f)");
        },
        py::name(name.c_str()));
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "alphasurf_b200: svox2.csrc-compatible module over libasurf.so (include/asurf.h)";
#define REG(fn) (Traced<&fn>::name = "svox2.csrc." #fn, m.def(#fn, &Traced<&fn>::call))
    REG(sample_grid);
    REG(sample_grid_backward);
    REG(cubic_extract_iso_pts);
    REG(sample_grid_sh_surf);
    REG(sample_grid_raw_alpha);
    REG(volume_render_surf_trav);
    REG(volume_render_surf_trav_backward);
    REG(volume_render_surf_trav_fused);
    REG(volume_render_cuvol);
    REG(volume_render_cuvol_image);
    REG(volume_render_cuvol_backward);
    REG(volume_render_cuvol_fused);
    REG(volume_render_expected_term);
    REG(volume_render_mode_term);
    REG(volume_render_med_term);
    REG(volume_render_sigma_thresh);
    REG(volume_render_expected_term_surf_trav);
    REG(volume_render_mode_term_surf_trav);
    REG(volume_render_sigma_thresh_surf_trav);
    REG(volume_render_alpha_surf_trav);
    REG(extract_pts_surf_trav);
    REG(render_normal_surf_trav);
    REG(tv);
    REG(tv_grad);
    REG(surface_normal_grad_sparse);
    REG(alpha_surf_sparsify_grad_sparse);
    REG(tv_grad_sparse);
    REG(surf_tv_grad_sparse);
    REG(msi_tv_grad_sparse);
    REG(surf_sign_change_grad_sparse);
    REG(surface_normal_grad);
    REG(lumisphere_tv_grad_sparse);
    REG(dilate);
    REG(accel_dist_prop);
    REG(grid_weight_render);
    REG(sparse_grid_weight_render);
    REG(sparse_grid_visbility_render_surf);
    REG(sparse_grid_mask_render);
    REG(rmsprop_step);
    REG(sgd_step);
#undef REG
    // volume_render_surf_trav_image stays ABSENT: svox2.py:3660 probes for it and would switch the image path
    for (const char *name : {"test_cubic_root_grad", "volume_render_surface", "volume_render_surface_backward",
                             "volume_render_surface_fused", "volume_render_nvol", "volume_render_nvol_backward",
                             "volume_render_nvol_fused", "volume_render_svox1", "volume_render_svox1_backward",
                             "volume_render_svox1_fused"})
        m.attr(name) = off_path(name);
    m.def("set_loss_norm_rays", [](py::object n) { g_norm_rays = n.is_none() ? 0 : n.cast<int64_t>(); },
          "global ray count used to normalise the fused losses in a ray-sharded run (None: per call)");
    m.def("abi_version", []() { return asurf_abi_version(); });
    m.def("render_stats", &render_stats, "march counters of one surf_trav forward pass (SURVEY.md 8d)");
    m.def("cuvol_render_stats", &cuvol_render_stats, "march counters of one cuvol forward pass (SURVEY.md 8d)");
    m.def("accel_for", &accel_for, "occupancy pyramid of a links tensor (built on the current stream and cached)");

    py::class_<SparseGridSpec>(m, "SparseGridSpec")
        .def(py::init<>())
        .def_readwrite("density_data", &SparseGridSpec::density_data)
        .def_readwrite("surface_data", &SparseGridSpec::surface_data)
        .def_readwrite("level_set_data", &SparseGridSpec::level_set_data)
        .def_readwrite("sh_data", &SparseGridSpec::sh_data)
        .def_readwrite("links", &SparseGridSpec::links)
        .def_readwrite("_offset", &SparseGridSpec::_offset)
        .def_readwrite("_scaling", &SparseGridSpec::_scaling)
        .def_readwrite("basis_dim", &SparseGridSpec::basis_dim)
        .def_readwrite("basis_type", &SparseGridSpec::basis_type)
        .def_readwrite("surface_type", &SparseGridSpec::surface_type)
        .def_readwrite("basis_data", &SparseGridSpec::basis_data)
        .def_readwrite("background_links", &SparseGridSpec::background_links)
        .def_readwrite("background_data", &SparseGridSpec::background_data)
        .def_readwrite("fake_sample_std", &SparseGridSpec::fake_sample_std)
        .def_readwrite("truncated_vol_render_a", &SparseGridSpec::truncated_vol_render_a);
    py::class_<CameraSpec>(m, "CameraSpec")
        .def(py::init<>())
        .def_readwrite("c2w", &CameraSpec::c2w)
        .def_readwrite("fx", &CameraSpec::fx)
        .def_readwrite("fy", &CameraSpec::fy)
        .def_readwrite("cx", &CameraSpec::cx)
        .def_readwrite("cy", &CameraSpec::cy)
        .def_readwrite("width", &CameraSpec::width)
        .def_readwrite("height", &CameraSpec::height)
        .def_readwrite("ndc_coeffx", &CameraSpec::ndc_coeffx)
        .def_readwrite("ndc_coeffy", &CameraSpec::ndc_coeffy);
    py::class_<RaysSpec>(m, "RaysSpec")
        .def(py::init<>())
        .def_readwrite("origins", &RaysSpec::origins)
        .def_readwrite("dirs", &RaysSpec::dirs)
        .def_readwrite("masks", &RaysSpec::masks);
    py::class_<RayVoxIntersecSpec>(m, "RayVoxIntersecSpec")
        .def(py::init<>())
        .def_readwrite("voxel_ls", &RayVoxIntersecSpec::voxel_ls)
        .def_readwrite("vox_start_i", &RayVoxIntersecSpec::vox_start_i)
        .def_readwrite("vox_num", &RayVoxIntersecSpec::vox_num);
    py::class_<RenderOptions>(m, "RenderOptions")
        .def(py::init<>())
        .def_readwrite("background_brightness", &RenderOptions::background_brightness)
        .def_readwrite("step_size", &RenderOptions::step_size)
        .def_readwrite("sigma_thresh", &RenderOptions::sigma_thresh)
        .def_readwrite("stop_thresh", &RenderOptions::stop_thresh)
        .def_readwrite("near_clip", &RenderOptions::near_clip)
        .def_readwrite("use_spheric_clip", &RenderOptions::use_spheric_clip)
        .def_readwrite("last_sample_opaque", &RenderOptions::last_sample_opaque)
        .def_readwrite("surf_fake_sample", &RenderOptions::surf_fake_sample)
        .def_readwrite("surf_fake_sample_min_vox_len", &RenderOptions::surf_fake_sample_min_vox_len)
        .def_readwrite("limited_fake_sample", &RenderOptions::limited_fake_sample)
        .def_readwrite("no_surf_grad_from_sh", &RenderOptions::no_surf_grad_from_sh)
        .def_readwrite("alpha_activation_type", &RenderOptions::alpha_activation_type)
        .def_readwrite("fake_sample_l_dist", &RenderOptions::fake_sample_l_dist)
        .def_readwrite("fake_sample_normalize_surf", &RenderOptions::fake_sample_normalize_surf)
        .def_readwrite("only_outward_intersect", &RenderOptions::only_outward_intersect)
        .def_readwrite("truncated_vol_render", &RenderOptions::truncated_vol_render)
        .def_readwrite("trunc_vol_weight_min", &RenderOptions::trunc_vol_weight_min);
    py::class_<GridOutputGrads>(m, "GridOutputGrads")
        .def(py::init<>())
        .def_readwrite("grad_density_out", &GridOutputGrads::grad_density_out)
        .def_readwrite("grad_sh_out", &GridOutputGrads::grad_sh_out)
        .def_readwrite("grad_surface_out", &GridOutputGrads::grad_surface_out)
        .def_readwrite("grad_fake_sample_std_out", &GridOutputGrads::grad_fake_sample_std_out)
        .def_readwrite("grad_basis_out", &GridOutputGrads::grad_basis_out)
        .def_readwrite("grad_background_out", &GridOutputGrads::grad_background_out)
        .def_readwrite("mask_out", &GridOutputGrads::mask_out)
        .def_readwrite("mask_background_out", &GridOutputGrads::mask_background_out);
}
