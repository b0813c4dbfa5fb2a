// alphasurf_b200: grid maintenance that sits right in front of the render path.
//
// accel_dist_prop (/root/reference/svox2/csrc/misc_kernel.cu:1022-1058, kernels :113-182): every empty vertex of `links`
// (value < 0) receives -(1 + k), k = number of levels of the implicit octree above it that hold no stored vertex (stopping
// below the first occupied ancestor); the cuvol marcher turns the code into the side of an empty aligned block it may
// skip (compute_skip_dist, include/render_util.cuh:286-368).  The reference marks ancestors with one scattered byte store
// per stored vertex per level; here each level is reduced from the one below (8 children -> 1 byte), so the mark pass
// reads `links` once and writes each pyramid byte once.  Integer work, HBM-bound: 4 B read per vertex for the marks,
// 4 B read + 4 B written per empty vertex for the codes.
#include "common.cuh"

namespace asurf {
namespace {

struct PyrLevel {
    int sx, sy, sz;      // size of this level
    int64_t off;         // byte offset in the pyramid buffer
};
constexpr int MAX_LEVELS = 16;
struct Pyr {
    PyrLevel lv[MAX_LEVELS];
    int n;
};

inline int div2_ceil(int x) { return (x + 1) >> 1; }

// level 0 of the pyramid (half resolution) straight from links
__global__ void __launch_bounds__(256) mark_level0_kernel(const int32_t *__restrict__ links, int gx, int gy, int gz, PyrLevel L,
                                                          uint8_t *__restrict__ pyr) {
    const int64_t n = (int64_t)L.sx * L.sy * L.sz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % L.sz), y = (int)((i / L.sz) % L.sy), x = (int)(i / ((int64_t)L.sz * L.sy));
        bool any = false;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int cx = 2 * x + (c >> 2), cy = 2 * y + ((c >> 1) & 1), cz = 2 * z + (c & 1);
            if (cx < gx && cy < gy && cz < gz) any |= (links[((int64_t)cx * gy + cy) * gz + cz] >= 0);
        }
        pyr[L.off + i] = any ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) mark_up_kernel(PyrLevel lo, PyrLevel hi, uint8_t *__restrict__ pyr) {
    const int64_t n = (int64_t)hi.sx * hi.sy * hi.sz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % hi.sz), y = (int)((i / hi.sz) % hi.sy), x = (int)(i / ((int64_t)hi.sz * hi.sy));
        bool any = false;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int cx = 2 * x + (c >> 2), cy = 2 * y + ((c >> 1) & 1), cz = 2 * z + (c & 1);
            if (cx < lo.sx && cy < lo.sy && cz < lo.sz) any |= (pyr[lo.off + ((int64_t)cx * lo.sy + cy) * lo.sz + cz] != 0);
        }
        pyr[hi.off + i] = any ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) dist_code_kernel(int32_t *__restrict__ links, int gx, int gy, int gz, Pyr P,
                                                        const uint8_t *__restrict__ pyr) {
    const int64_t n = (int64_t)gx * gy * gz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (links[i] >= 0) continue;
        int z = (int)(i % gz), y = (int)((i / gz) % gy), x = (int)(i / ((int64_t)gz * gy));
        int result = -1;
        for (int l = 0; l < P.n; ++l) {
            x >>= 1; y >>= 1; z >>= 1;
            const PyrLevel &L = P.lv[l];
            if (pyr[L.off + ((int64_t)x * L.sy + y) * L.sz + z]) break;
            result -= 1;
        }
        links[i] = result;
    }
}

Workspace g_ws_pyr;

}  // namespace
void misc_release() { g_ws_pyr.release(); }
}  // namespace asurf

using namespace asurf;

extern "C" int asurf_accel_dist_prop(int32_t *links, const int32_t size[3], void *stream) {
    ASURF_REQUIRE(links && size, ASURF_E_INVALID, "accel_dist_prop: null argument");
    ASURF_REQUIRE(size[0] >= 1 && size[1] >= 1 && size[2] >= 1, ASURF_E_INVALID, "accel_dist_prop: bad grid size");
    cudaStream_t st = (cudaStream_t)stream;
    Pyr P;
    P.n = 0;
    int sx = size[0], sy = size[1], sz = size[2];
    int64_t bytes = 0;
    while (sx > 1 && sy > 1 && sz > 1 && P.n < MAX_LEVELS) {   // same level sequence as misc_kernel.cu:1035-1041
        sx = div2_ceil(sx); sy = div2_ceil(sy); sz = div2_ceil(sz);
        P.lv[P.n] = {sx, sy, sz, bytes};
        bytes += (int64_t)sx * sy * sz;
        ++P.n;
    }
    if (P.n == 0) {   // a grid with an axis of length 1 has no pyramid: every empty vertex gets -1
        Pyr none;
        none.n = 0;
        dist_code_kernel<<<148 * 8, 256, 0, st>>>(links, size[0], size[1], size[2], none, nullptr);
        note_launches(1);
        return check_cuda(cudaGetLastError(), "accel_dist_prop launch");
    }
    int rc = g_ws_pyr.reserve((size_t)bytes);
    if (rc) return rc;
    uint8_t *pyr = (uint8_t *)g_ws_pyr.ptr;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto blocks = [&](int64_t n) { const int64_t w = (n + 255) / 256; return (int)(w < (int64_t)sms * 16 ? (w > 0 ? w : 1) : (int64_t)sms * 16); };
    mark_level0_kernel<<<blocks((int64_t)P.lv[0].sx * P.lv[0].sy * P.lv[0].sz), 256, 0, st>>>(links, size[0], size[1], size[2],
                                                                                              P.lv[0], pyr);
    for (int l = 1; l < P.n; ++l)
        mark_up_kernel<<<blocks((int64_t)P.lv[l].sx * P.lv[l].sy * P.lv[l].sz), 256, 0, st>>>(P.lv[l - 1], P.lv[l], pyr);
    dist_code_kernel<<<blocks((int64_t)size[0] * size[1] * size[2]), 256, 0, st>>>(links, size[0], size[1], size[2], P, pyr);
    note_launches(P.n + 1);
    return check_cuda(cudaGetLastError(), "accel_dist_prop launch");
}
