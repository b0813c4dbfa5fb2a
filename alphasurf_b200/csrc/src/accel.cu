// alphasurf_b200: occupancy pyramid over the `links` table.
//
// The reference marches voxel by voxel and reads the 8 corner links of every visited voxel
// (/root/reference/svox2/csrc/render_lerp_kernel_surf_trav.cu:212-221, USE_ACC_SKIP=false at :31).
// A voxel only matters when all 8 links are >= 0, so we precompute that predicate once per `links`
// version as a bitmap (16.7 MB at 512^3 -- L2 resident on B200) plus two coarser levels, and the marcher
// tests one bit (and skips whole empty 4^3 / 16^3 / 64^3 blocks) instead of gathering 8 scattered int32s.
// Layout: see AccelLayout in common.cuh.
#include "common.cuh"

namespace asurf {
namespace {

// One thread per level-0 word (4x4x4 cells = 5x5x5 vertices).  Vertex reads along z are contiguous;
// neighbouring threads handle neighbouring z-blocks so a warp covers long z runs of the links table.
__global__ void __launch_bounds__(128) accel_level0_kernel(const int32_t *__restrict__ links, int sx, int sy, int sz,
                                                            AccelLayout lay, uint64_t *__restrict__ out) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= lay.count(0)) return;
    const int bz = (int)(w % lay.b[0][2]);
    const int by = (int)((w / lay.b[0][2]) % lay.b[0][1]);
    const int bx = (int)(w / ((int64_t)lay.b[0][2] * lay.b[0][1]));
    const int x0 = bx * 4, y0 = by * 4, z0 = bz * 4;
    // vertex validity bits: 5x5 columns of 5 z-bits
    uint32_t col[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            uint32_t bits = 0;
            const int x = x0 + i, y = y0 + j;
            if (x < sx && y < sy) {
                const int32_t *p = links + ((int64_t)x * sy + y) * sz;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int z = z0 + k;
                    if (z < sz && p[z] >= 0) bits |= 1u << k;
                }
            }
            col[i][j] = bits;
        }
    }
    uint64_t word = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = col[i][j] & col[i][j + 1] & col[i + 1][j] & col[i + 1][j + 1];
            const uint32_t cells = c & (c >> 1) & 0xFu;  // cell k active iff z-bits k and k+1 of all 4 columns
            word |= (uint64_t)cells << ((i << 4) | (j << 2));
        }
    }
    out[w] = word;
}

__global__ void __launch_bounds__(128) accel_coarsen_kernel(AccelLayout lay, int level, uint64_t *__restrict__ buf) {
    // builds level `level` (>= 1) from level-1
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= lay.count(level)) return;
    const int *bc = lay.b[level], *bf = lay.b[level - 1];
    const uint64_t *fine = buf + lay.off[level - 1];
    const int cz = (int)(w % bc[2]);
    const int cy = (int)((w / bc[2]) % bc[1]);
    const int cx = (int)(w / ((int64_t)bc[2] * bc[1]));
    uint64_t word = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 4; ++k) {
                const int bx = cx * 4 + i, by = cy * 4 + j, bz = cz * 4 + k;
                if (bx < bf[0] && by < bf[1] && bz < bf[2]) {
                    if (fine[((int64_t)bx * bf[1] + by) * bf[2] + bz] != 0) word |= 1ull << ((i << 4) | (j << 2) | k);
                }
            }
    buf[lay.off[level] + w] = word;
}

}  // namespace
}  // namespace asurf

using namespace asurf;

extern "C" int64_t asurf_accel_words(const int32_t size[3]) {
    AccelLayout lay(size);
    return lay.off[3];
}

extern "C" int asurf_accel_build(const int32_t *links, const int32_t size[3], uint64_t *accel_out, void *stream) {
    ASURF_REQUIRE(links && accel_out, ASURF_E_INVALID, "accel_build: null pointer");
    ASURF_REQUIRE(size[0] >= 2 && size[1] >= 2 && size[2] >= 2, ASURF_E_INVALID, "accel_build: grid smaller than 2^3");
    AccelLayout lay(size);
    cudaStream_t st = (cudaStream_t)stream;
    accel_level0_kernel<<<div_up(lay.count(0), 128), 128, 0, st>>>(links, size[0], size[1], size[2], lay, accel_out);
    accel_coarsen_kernel<<<div_up(lay.count(1), 128), 128, 0, st>>>(lay, 1, accel_out);
    accel_coarsen_kernel<<<div_up(lay.count(2), 128), 128, 0, st>>>(lay, 2, accel_out);
    return check_cuda(cudaGetLastError(), "accel_build launch");
}
