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

// Work bitmap, level 0: one warp per 4x4x4-cell block.  The 5x5x5 vertices of the block are staged once in shared
// memory (surface scalar + raw density), then each lane classifies two cells.
constexpr int WORK_WARPS = 8;
__global__ void __launch_bounds__(WORK_WARPS * 32)
work_level0_kernel(const int32_t *__restrict__ links, const float *__restrict__ density,
                   const float *__restrict__ surface, const float *__restrict__ level_set, int level_set_num,
                   float sigma_thresh, int every_voxel, int sx, int sy, int sz, AccelLayout lay,
                   const uint64_t *__restrict__ occ, uint64_t *__restrict__ out) {
    __shared__ float s_surf[WORK_WARPS][128];
    __shared__ float s_dens[WORK_WARPS][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_words = lay.count(0);
    for (int64_t w = (int64_t)blockIdx.x * WORK_WARPS + warp; w < n_words; w += (int64_t)gridDim.x * WORK_WARPS) {
        const uint64_t o = occ[w];
        if (o == 0) {
            if (lane == 0) out[w] = 0;
            continue;
        }
        const int bz = (int)(w % lay.b[0][2]);
        const int by = (int)((w / lay.b[0][2]) % lay.b[0][1]);
        const int bx = (int)(w / ((int64_t)lay.b[0][2] * lay.b[0][1]));
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int v = lane + 32 * r;
            if (v < 125) {
                const int x = bx * 4 + v / 25, y = by * 4 + (v / 5) % 5, z = bz * 4 + v % 5;
                float sv = 0.f, dv = 0.f;
                if (x < sx && y < sy && z < sz) {
                    const int32_t l = links[((int64_t)x * sy + y) * sz + z];
                    if (l >= 0) {
                        sv = surface[l];
                        dv = density[l];
                    }
                }
                s_surf[warp][v] = sv;
                s_dens[warp][v] = dv;
            }
        }
        __syncwarp();
        unsigned bits[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            bool work = false;
            if ((o >> c) & 1ull) {
                const int base = (c >> 4) * 25 + ((c >> 2) & 3) * 5 + (c & 3);
                float smin = INFINITY, smax = -INFINITY;
                bool gate = false;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int idx = base + (k >> 2) * 25 + ((k >> 1) & 1) * 5 + (k & 1);
                    const float sv = s_surf[warp][idx];
                    smin = fminf(smin, sv);
                    smax = fmaxf(smax, sv);
                    gate |= !(s_dens[warp][idx] < sigma_thresh);
                }
                bool has_surf = every_voxel != 0;
                for (int i = 0; i < level_set_num && !has_surf; ++i) {
                    const float lv = level_set[i];
                    has_surf = !((lv < smin) || (lv > smax));
                }
                work = gate && has_surf;
            }
            bits[h] = __ballot_sync(0xffffffffu, work);
        }
        if (lane == 0) out[w] = (uint64_t)bits[0] | ((uint64_t)bits[1] << 32);
    }
}

}  // namespace
}  // namespace asurf

using namespace asurf;

extern "C" int asurf_work_build(const asurf_grid_t *grid, const asurf_opt_t *opt, uint64_t *work_out, void *stream) {
    ASURF_REQUIRE(grid && opt && work_out, ASURF_E_INVALID, "work_build: null argument");
    ASURF_REQUIRE(grid->links && grid->density && grid->surface && grid->level_set && grid->accel, ASURF_E_INVALID,
                  "work_build: the grid needs links, density, surface, level sets and its occupancy pyramid");
    AccelLayout lay(grid->size);
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (lay.count(0) + WORK_WARPS - 1) / WORK_WARPS;
    const int blocks = (int)(want < (int64_t)sms * 32 ? want : (int64_t)sms * 32);
    const int every_voxel = (opt->surf_fake_sample && !opt->limited_fake_sample) ? 1 : 0;
    work_level0_kernel<<<blocks, WORK_WARPS * 32, 0, st>>>(grid->links, grid->density, grid->surface, grid->level_set,
                                                           grid->level_set_num, opt->sigma_thresh, every_voxel,
                                                           grid->size[0], grid->size[1], grid->size[2], lay, grid->accel,
                                                           work_out);
    accel_coarsen_kernel<<<div_up(lay.count(1), 128), 128, 0, st>>>(lay, 1, work_out);
    accel_coarsen_kernel<<<div_up(lay.count(2), 128), 128, 0, st>>>(lay, 2, work_out);
    note_launches(3);
    return check_cuda(cudaGetLastError(), "work_build launch");
}

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
    note_launches(3);
    return check_cuda(cudaGetLastError(), "accel_build launch");
}
