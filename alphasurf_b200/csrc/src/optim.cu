// alphasurf_b200: in-place RMSprop / SGD steps on (N, C) grid tensors.
//
// Replaces rmsprop_step / sgd_step of /root/reference/svox2/csrc/optim_kernel.cu:154-267.
//   rms  = rms == 0 ? g^2 : lerp(g^2, rms, beta);  x = max(x - lr*g/(sqrt(rms)+eps), minval);  g = 0
//   (optim_kernel.cu:15-25); the last channel may use lr_last (:37-46).
// Three indexer modes as the reference dispatches them (:175-215): all rows, bool row mask, int64 row list.
//
// Pure streaming, HBM-bound: 12 B read + 12 B written per touched element.  The dense and masked kernels
// run a grid-stride loop with 64-bit element indices (the reference's `int tid` overflows at
// 512^3 x 27, cuda_util.cuh:14); when C is a multiple of 4 or rows are taken whole, loads are float4.
#include "common.cuh"

namespace asurf {
namespace {

__device__ __forceinline__ void rmsprop_once(float &x, float &rms, float &g, float beta, float lr, float eps,
                                             float minval) {
    const float g2 = g * g;
    rms = (rms == 0.f) ? g2 : fmaf(beta, rms - g2, g2);
    x = fmaxf(x - __fdiv_rn(lr * g, sqrtf(rms) + eps), minval);
    g = 0.f;
}

// Dense / masked: one thread per element, grid-stride; consecutive threads touch consecutive floats.
template <bool MASKED>
__global__ void __launch_bounds__(256) rmsprop_dense_kernel(float *__restrict__ data, float *__restrict__ rms,
                                                             float *__restrict__ grad, const uint8_t *__restrict__ mask,
                                                             int64_t n_elem, int n_cols, float beta, float lr,
                                                             float eps, float minval, float lr_last) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
        const int64_t row = i / n_cols;
        const int c = (int)(i - row * n_cols);
        if (MASKED && !mask[row]) continue;
        float x = data[i], r = rms[i], g = grad[i];
        rmsprop_once(x, r, g, beta, (c == n_cols - 1) ? lr_last : lr, eps, minval);
        data[i] = x;
        rms[i] = r;
        grad[i] = 0.f;
    }
}

// Single-column tensors (density, surface): four rows per thread -- float4 loads / stores of data, rms and grad and one
// 32-bit load of the four mask bytes.  The alpha-Surf regularisers touch every stored row each step, so nearly all
// groups take the full-vector path; a partially masked group stores its set rows one by one.
template <bool MASKED, bool RMS>
__global__ void __launch_bounds__(256) step_col1_vec4_kernel(float *__restrict__ data, float *__restrict__ rms,
                                                              float *__restrict__ grad, const uint8_t *__restrict__ mask,
                                                              int64_t n_rows, float beta, float lr, float eps,
                                                              float minval) {
    const int64_t n4 = n_rows >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < n4; i += stride) {
        unsigned m = 0x01010101u;
        if (MASKED) {
            m = reinterpret_cast<const unsigned *>(mask)[i];
            if (m == 0u) continue;
        }
        float4 x = reinterpret_cast<float4 *>(data)[i];
        float4 g = reinterpret_cast<float4 *>(grad)[i];
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (RMS) r = reinterpret_cast<float4 *>(rms)[i];
        float *xs = &x.x, *gs = &g.x, *rs = &r.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (RMS) rmsprop_once(xs[k], rs[k], gs[k], beta, lr, eps, minval);
            else { xs[k] = fmaf(-lr, gs[k], xs[k]); gs[k] = 0.f; }
        }
        const bool all = !MASKED || ((m & 0xffu) && (m & 0xff00u) && (m & 0xff0000u) && (m & 0xff000000u));
        if (all) {
            reinterpret_cast<float4 *>(data)[i] = x;
            if (RMS) reinterpret_cast<float4 *>(rms)[i] = r;
            reinterpret_cast<float4 *>(grad)[i] = g;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if ((m >> (8 * k)) & 0xffu) {
                    data[i * 4 + k] = xs[k];
                    if (RMS) rms[i * 4 + k] = rs[k];
                    grad[i * 4 + k] = 0.f;
                }
            }
        }
    }
    // tail rows
    for (int64_t i = n4 * 4 + t0; i < n_rows; i += stride) {
        if (MASKED && !mask[i]) continue;
        float x = data[i], g = grad[i];
        if (RMS) {
            float r = rms[i];
            rmsprop_once(x, r, g, beta, lr, eps, minval);
            rms[i] = r;
        } else {
            x = fmaf(-lr, g, x);
        }
        data[i] = x;
        grad[i] = 0.f;
    }
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

// Row-masked steps: training touches a tiny fraction of the rows (the voxels the batch's samples hit), so scanning all
// N*C elements for their row's mask byte wastes the launch.  Each warp takes 128 rows, reads their mask bytes as 32 words
// in one coalesced load, and then walks only the set rows with the lanes striding over the C channels (coalesced rows).
template <bool RMS>
__global__ void __launch_bounds__(256) masked_rows_kernel(float *__restrict__ data, float *__restrict__ rms,
                                                           float *__restrict__ grad, const uint8_t *__restrict__ mask,
                                                           int64_t n_rows, int n_cols, float beta, float lr, float eps,
                                                           float minval, float lr_last) {
    __shared__ uint8_t s_rows[256 / 32][128];
    const int lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool word_ok = ((uintptr_t)mask & 3u) == 0;
    for (int64_t base = warp0 * 128; base < n_rows; base += n_warps * 128) {
        // 128 rows per round: every lane fetches the mask bytes of 4 consecutive rows (one 32-bit load when it can)
        const int64_t row0 = base + lane * 4;
        unsigned w = 0u;
        if (word_ok && row0 + 3 < n_rows) {
            w = reinterpret_cast<const unsigned *>(mask)[row0 >> 2];
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (row0 + k < n_rows && mask[row0 + k]) w |= 1u << (8 * k);
        }
        unsigned bits = ((w & 0xffu) ? 1u : 0u) | ((w & 0xff00u) ? 2u : 0u) | ((w & 0xff0000u) ? 4u : 0u) |
                        ((w & 0xff000000u) ? 8u : 0u);
        if (!__ballot_sync(0xffffffffu, bits != 0u)) continue;
        // the set rows of the round, compacted: the touched rows cluster (they follow the surface), so a round often holds
        // dozens of them -- their (row, channel) elements are spread over the lanes instead of one row at a time
        int incl = __popc(bits);
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        const int n_set = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        {
            int at = incl - __popc(bits);
            for (unsigned b = bits; b; b &= b - 1) s_rows[warp_in_cta][at++] = (uint8_t)(lane * 4 + __ffs(b) - 1);
        }
        __syncwarp();
        // four elements per lane and pass: the touched rows cluster, so a few warps hold most of the work, and one dependent
        // load -> update -> store round trip per 32 elements made them the whole kernel
        const int n_elem = n_set * n_cols;
        constexpr int U = 4;
        for (int e0 = lane; e0 < n_elem; e0 += 32 * U) {
            int64_t off[U];
            float x[U], q[U], g[U], l[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + 32 * u;
                off[u] = -1;
                if (e < n_elem) {
                    const int k = e / n_cols, c = e - k * n_cols;
                    off[u] = (base + s_rows[warp_in_cta][k]) * n_cols + c;
                    l[u] = (c == n_cols - 1) ? lr_last : lr;
                    x[u] = data[off[u]];
                    g[u] = grad[off[u]];
                    if (RMS) q[u] = rms[off[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (off[u] < 0) continue;
                if (RMS) {
                    rmsprop_once(x[u], q[u], g[u], beta, l[u], eps, minval);
                    data[off[u]] = x[u];
                    rms[off[u]] = q[u];
                } else {
                    data[off[u]] = fmaf(-l[u], g[u], x[u]);
                }
                grad[off[u]] = 0.f;
            }
        }
    }
}

__global__ void __launch_bounds__(256) rmsprop_index_kernel(float *__restrict__ data, float *__restrict__ rms,
                                                             float *__restrict__ grad, const int64_t *__restrict__ idx,
                                                             int64_t n_index, int n_cols, float beta, float lr,
                                                             float eps, float minval, float lr_last) {
    const int64_t n_elem = n_index * n_cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
        const int64_t k = i / n_cols;
        const int c = (int)(i - k * n_cols);
        // the reference narrows the int64 index to int32 (optim_kernel.cu:86); rows never exceed 2^31
        const int64_t off = (int64_t)(int32_t)idx[k] * n_cols + c;
        float x = data[off], r = rms[off], g = grad[off];
        rmsprop_once(x, r, g, beta, (c == n_cols - 1) ? lr_last : lr, eps, minval);
        data[off] = x;
        rms[off] = r;
        grad[off] = 0.f;
    }
}

template <bool MASKED>
__global__ void __launch_bounds__(256) sgd_dense_kernel(float *__restrict__ data, float *__restrict__ grad,
                                                         const uint8_t *__restrict__ mask, int64_t n_elem, int n_cols,
                                                         float lr, float lr_last) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
        const int64_t row = i / n_cols;
        const int c = (int)(i - row * n_cols);
        if (MASKED && !mask[row]) continue;
        const float l = (c == n_cols - 1) ? lr_last : lr;
        data[i] = fmaf(-l, grad[i], data[i]);   // nvcc contracts `x -= lr * g` (optim_kernel.cu:100)
        grad[i] = 0.f;
    }
}

__global__ void __launch_bounds__(256) sgd_index_kernel(float *__restrict__ data, float *__restrict__ grad,
                                                         const int64_t *__restrict__ idx, int64_t n_index, int n_cols,
                                                         float lr, float lr_last) {
    const int64_t n_elem = n_index * n_cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
        const int64_t k = i / n_cols;
        const int c = (int)(i - k * n_cols);
        const int64_t off = (int64_t)(int32_t)idx[k] * n_cols + c;
        const float l = (c == n_cols - 1) ? lr_last : lr;
        data[off] = fmaf(-l, grad[off], data[off]);
        grad[off] = 0.f;
    }
}

int stream_grid(int64_t n_elem) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n_elem + 255) / 256;
    const int64_t cap = (int64_t)sms * 16;  // 16 CTAs x 256 threads = 2 full waves of resident threads per SM
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace asurf

using namespace asurf;

extern "C" int asurf_rmsprop_step(float *data, float *rms, float *grad, int64_t n_rows, int32_t n_cols,
                                  int32_t indexer_kind, const void *indexer, int64_t n_index, float beta, float lr,
                                  float eps, float minval, float lr_last, void *stream) {
    ASURF_REQUIRE(data && rms && grad, ASURF_E_INVALID, "rmsprop_step: null tensor");
    ASURF_REQUIRE(n_cols > 0 && n_rows >= 0, ASURF_E_INVALID, "rmsprop_step: bad shape");
    if (lr_last < 0.f) lr_last = lr;  // optim_kernel.cu:171
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_elem = n_rows * n_cols;
    const bool vec1 = n_cols == 1 && aligned16(data) && aligned16(rms) && aligned16(grad);   // lr of the only column: lr_last
    if (indexer_kind == 0) {
        if (n_elem == 0) return 0;
        if (vec1)
            step_col1_vec4_kernel<false, true><<<stream_grid(n_elem / 4 + 1), 256, 0, st>>>(data, rms, grad, nullptr, n_rows,
                                                                                             beta, lr_last, eps, minval);
        else
            rmsprop_dense_kernel<false><<<stream_grid(n_elem), 256, 0, st>>>(data, rms, grad, nullptr, n_elem, n_cols, beta,
                                                                             lr, eps, minval, lr_last);
    } else if (indexer_kind == 1) {
        if (n_index == 0 || n_elem == 0) return 0;  // size(0) == 0 -> skip (:189)
        // narrow tensors (density, surface: C = 1) gain nothing from a warp per row: one thread per element instead
        if (vec1 && (((uintptr_t)indexer & 3u) == 0))
            step_col1_vec4_kernel<true, true><<<stream_grid(n_elem / 4 + 1), 256, 0, st>>>(
                data, rms, grad, (const uint8_t *)indexer, n_rows, beta, lr_last, eps, minval);
        else if (n_cols < 8)
            rmsprop_dense_kernel<true><<<stream_grid(n_elem), 256, 0, st>>>(data, rms, grad, (const uint8_t *)indexer, n_elem,
                                                                            n_cols, beta, lr, eps, minval, lr_last);
        else
            masked_rows_kernel<true><<<stream_grid(n_rows / 4 + 1), 256, 0, st>>>(data, rms, grad, (const uint8_t *)indexer,
                                                                              n_rows, n_cols, beta, lr, eps, minval, lr_last);
    } else if (indexer_kind == 2) {
        if (n_index == 0) return 0;
        rmsprop_index_kernel<<<stream_grid(n_index * n_cols), 256, 0, st>>>(data, rms, grad, (const int64_t *)indexer,
                                                                           n_index, n_cols, beta, lr, eps, minval,
                                                                           lr_last);
    } else {
        ASURF_REQUIRE(false, ASURF_E_INVALID, "rmsprop_step: bad indexer kind %d", indexer_kind);
    }
    note_launches(1);
    return check_cuda(cudaGetLastError(), "rmsprop_step launch");
}

extern "C" int asurf_sgd_step(float *data, float *grad, int64_t n_rows, int32_t n_cols, int32_t indexer_kind,
                              const void *indexer, int64_t n_index, float lr, float lr_last, void *stream) {
    ASURF_REQUIRE(data && grad, ASURF_E_INVALID, "sgd_step: null tensor");
    ASURF_REQUIRE(n_cols > 0 && n_rows >= 0, ASURF_E_INVALID, "sgd_step: bad shape");
    if (lr_last < 0.f) lr_last = lr;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_elem = n_rows * n_cols;
    const bool vec1 = n_cols == 1 && aligned16(data) && aligned16(grad);
    if (indexer_kind == 0) {
        if (n_elem == 0) return 0;
        if (vec1)
            step_col1_vec4_kernel<false, false><<<stream_grid(n_elem / 4 + 1), 256, 0, st>>>(data, nullptr, grad, nullptr, n_rows,
                                                                                              0.f, lr_last, 0.f, 0.f);
        else
            sgd_dense_kernel<false><<<stream_grid(n_elem), 256, 0, st>>>(data, grad, nullptr, n_elem, n_cols, lr, lr_last);
    } else if (indexer_kind == 1) {
        if (n_index == 0 || n_elem == 0) return 0;
        if (vec1 && (((uintptr_t)indexer & 3u) == 0))
            step_col1_vec4_kernel<true, false><<<stream_grid(n_elem / 4 + 1), 256, 0, st>>>(
                data, nullptr, grad, (const uint8_t *)indexer, n_rows, 0.f, lr_last, 0.f, 0.f);
        else if (n_cols < 8)
            sgd_dense_kernel<true><<<stream_grid(n_elem), 256, 0, st>>>(data, grad, (const uint8_t *)indexer, n_elem, n_cols,
                                                                        lr, lr_last);
        else
            masked_rows_kernel<false><<<stream_grid(n_rows / 4 + 1), 256, 0, st>>>(data, nullptr, grad, (const uint8_t *)indexer,
                                                                               n_rows, n_cols, 0.f, lr, 0.f, 0.f, lr_last);
    } else if (indexer_kind == 2) {
        if (n_index == 0) return 0;
        sgd_index_kernel<<<stream_grid(n_index * n_cols), 256, 0, st>>>(data, grad, (const int64_t *)indexer, n_index,
                                                                       n_cols, lr, lr_last);
    } else {
        ASURF_REQUIRE(false, ASURF_E_INVALID, "sgd_step: bad indexer kind %d", indexer_kind);
    }
    note_launches(1);
    return check_cuda(cudaGetLastError(), "sgd_step launch");
}
