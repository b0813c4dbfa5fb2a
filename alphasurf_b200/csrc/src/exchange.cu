// alphasurf_b200: packing of the touched gradient rows for the multi-GPU gradient exchange (alphasurf_b200/dist.py).
//
// The ray-sharded data-parallel step all-reduces only the rows of (density.grad, surface.grad, sh.grad) that some rank
// touched.  `rows` (int64, ascending, identical on every rank) comes from the OR-ed touched masks; a bucket row is
// [density, surface, sh_0 .. sh_{D-1}].  One warp per row: the 2 + D floats are read / written as one coalesced segment.
// Pure data movement, HBM-bound: (2 + D) * 4 B read + written per row (twice for pack-and-clear).
#include "common.cuh"

namespace asurf {
namespace {

template <bool PACK>
__global__ void __launch_bounds__(256) rows_kernel(const int64_t *__restrict__ rows, int64_t n, float *__restrict__ density,
                                                   float *__restrict__ surface, float *__restrict__ sh, int D,
                                                   float *__restrict__ bucket, int clear) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int W = 2 + D;
    for (int64_t i = warp0; i < n; i += n_warps) {
        const int64_t r = rows[i];
        if (r < 0) {   // padding of a fixed-capacity row list (torch.nonzero_static): zeros travel, nothing is added back
            if (PACK)
                for (int c = lane; c < W; c += 32) bucket[i * W + c] = 0.f;
            continue;
        }
        for (int c = lane; c < W; c += 32) {
            float *src = (c == 0) ? (density + r) : ((c == 1) ? (surface + r) : (sh + r * D + (c - 2)));
            if (PACK) {
                bucket[i * W + c] = *src;
                if (clear) *src = 0.f;
            } else {
                *src += bucket[i * W + c];
            }
        }
    }
}

// Touched-row masks travel bit-packed: NCCL has no bitwise OR, so every rank packs its (N,) bool mask into N / 32 words,
// the words are all-gathered (N / 8 bytes per rank instead of an N-byte MAX all-reduce) and OR-ed while unpacking.
__global__ void __launch_bounds__(256) mask_pack_kernel(const uint8_t *__restrict__ mask, int64_t n, uint32_t *__restrict__ words) {
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + lane;
        const unsigned w = __ballot_sync(0xffffffffu, (i < n) && (mask[i] != 0));
        if (lane == 0) words[i0 >> 5] = w;
    }
}
__global__ void __launch_bounds__(256) mask_unpack_or_kernel(const uint32_t *__restrict__ words_all, int world, int64_t n_words,
                                                            int64_t n, uint8_t *__restrict__ mask) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w = 0;
        for (int r = 0; r < world; ++r) w |= __ldg(words_all + (int64_t)r * n_words + (i >> 5));
        mask[i] = (uint8_t)((w >> (i & 31)) & 1u);
    }
}

int launch(bool pack, const int64_t *rows, int64_t n, float *density, float *surface, float *sh, int D, float *bucket,
           int clear, cudaStream_t st) {
    if (n <= 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n * 32 + 255) / 256;
    const int blocks = (int)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
    if (pack) rows_kernel<true><<<blocks, 256, 0, st>>>(rows, n, density, surface, sh, D, bucket, clear);
    else rows_kernel<false><<<blocks, 256, 0, st>>>(rows, n, density, surface, sh, D, bucket, 0);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "rows pack/unpack launch");
}

}  // namespace
}  // namespace asurf

using namespace asurf;

extern "C" int asurf_rows_pack(const int64_t *rows, int64_t n_rows, float *grad_density, float *grad_surface, float *grad_sh,
                               int32_t sh_dim, float *bucket, int32_t clear_rows, void *stream) {
    ASURF_REQUIRE(n_rows <= 0 || (rows && grad_density && grad_surface && grad_sh && bucket), ASURF_E_INVALID,
                  "rows_pack: null pointer");
    return launch(true, rows, n_rows, grad_density, grad_surface, grad_sh, sh_dim, bucket, clear_rows, (cudaStream_t)stream);
}

extern "C" int asurf_rows_unpack_add(const int64_t *rows, int64_t n_rows, float *grad_density, float *grad_surface,
                                     float *grad_sh, int32_t sh_dim, const float *bucket, void *stream) {
    ASURF_REQUIRE(n_rows <= 0 || (rows && grad_density && grad_surface && grad_sh && bucket), ASURF_E_INVALID,
                  "rows_unpack_add: null pointer");
    return launch(false, rows, n_rows, grad_density, grad_surface, grad_sh, sh_dim, (float *)bucket, 0, (cudaStream_t)stream);
}

extern "C" int asurf_mask_pack(const uint8_t *mask, int64_t n, uint32_t *words, void *stream) {
    ASURF_REQUIRE(n <= 0 || (mask && words), ASURF_E_INVALID, "mask_pack: null pointer");
    if (n <= 0) return 0;
    const int64_t want = (n + 255) / 256;
    mask_pack_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, (cudaStream_t)stream>>>(mask, n, words);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "mask_pack launch");
}

extern "C" int asurf_mask_unpack_or(const uint32_t *words_all, int32_t world, int64_t n, uint8_t *mask, void *stream) {
    ASURF_REQUIRE(n <= 0 || (words_all && mask && world > 0), ASURF_E_INVALID, "mask_unpack_or: bad argument");
    if (n <= 0) return 0;
    const int64_t want = (n + 255) / 256;
    mask_unpack_or_kernel<<<(int)(want < 148 * 16 ? want : 148 * 16), 256, 0, (cudaStream_t)stream>>>(words_all, world,
                                                                                                    (n + 31) / 32, n, mask);
    note_launches(1);
    return check_cuda(cudaGetLastError(), "mask_unpack_or launch");
}
