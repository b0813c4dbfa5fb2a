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

// Work bitmap, levels 0 and 1 in one pass: one CTA per non-empty 16^3-cell block (= one level-1 word), walking the list
// of such blocks that asurf_accel_build leaves behind the occupancy pyramid.  The block's 17^3 vertices (surface scalar +
// raw density through `links`) are staged once in shared memory -- 1.2x the block's own vertices, against 1.95x for a 5^3
// halo per 4^3 block -- with many independent gathers in flight per thread; then each warp classifies 8 of the 64 4^3
// sub-blocks.  Words of blocks without linked cells are cleared by a memset before the launch.
constexpr int WB_THREADS = 256;
constexpr int WB_V = 17;
__global__ void __launch_bounds__(WB_THREADS)
work_block16_kernel(const int32_t *__restrict__ links, const float *__restrict__ density,
                    const float *__restrict__ surface, const float *__restrict__ level_set, int level_set_num,
                    float sigma_thresh, int every_voxel, int sx, int sy, int sz, AccelLayout lay,
                    const uint64_t *__restrict__ occ, uint64_t *__restrict__ out) {
    __shared__ float s_surf[WB_V * WB_V * WB_V];
    __shared__ float s_dens[WB_V * WB_V * WB_V];
    __shared__ unsigned s_l1[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_active = (int64_t)occ[lay.off[3]];
    const uint32_t *active = (const uint32_t *)(occ + lay.off[3] + 1);
    constexpr int NV = WB_V * WB_V * WB_V;
    for (int64_t it = blockIdx.x; it < n_active; it += gridDim.x) {
        __syncthreads();   // the shared buffers of the previous block are free
        const int64_t w1 = active[it];
        const int cz = (int)(w1 % lay.b[1][2]);
        const int cy = (int)((w1 / lay.b[1][2]) % lay.b[1][1]);
        const int cx = (int)(w1 / ((int64_t)lay.b[1][2] * lay.b[1][1]));
        if (tid < 2) s_l1[tid] = 0u;
        const int x0 = cx * 16, y0 = cy * 16, z0 = cz * 16;
        // stage the vertices: batches of 4 independent link loads, then 8 independent gathers
        for (int vb = tid; vb < NV; vb += 4 * WB_THREADS) {
            int32_t l[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int v = vb + r * WB_THREADS;
                l[r] = -1;
                if (v < NV) {
                    const int x = x0 + v / (WB_V * WB_V), y = y0 + (v / WB_V) % WB_V, z = z0 + v % WB_V;
                    if (x < sx && y < sy && z < sz) l[r] = __ldg(links + (((int64_t)x * sy + y) * sz + z));
                }
            }
            float sv[4], dv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                sv[r] = 0.f;
                dv[r] = 0.f;
                if (l[r] >= 0) {
                    sv[r] = __ldg(surface + l[r]);
                    dv[r] = __ldg(density + l[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int v = vb + r * WB_THREADS;
                if (v < NV) {
                    s_surf[v] = sv[r];
                    s_dens[v] = dv[r];
                }
            }
        }
        __syncthreads();
        unsigned long long l1bits = 0;   // lane 0 of each warp collects the non-empty sub-blocks it produced
        for (int sb = warp * 8; sb < warp * 8 + 8; ++sb) {
            const int i = sb >> 4, j = (sb >> 2) & 3, k = sb & 3;
            const int bx = cx * 4 + i, by = cy * 4 + j, bz = cz * 4 + k;
            if (!(bx < lay.b[0][0] && by < lay.b[0][1] && bz < lay.b[0][2])) continue;
            const int64_t k0 = ((int64_t)bx * lay.b[0][1] + by) * lay.b[0][2] + bz;
            const uint64_t o = occ[k0];
            if (o == 0) continue;   // the word was cleared by the memset
            unsigned bits[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                bool work = false;
                if ((o >> c) & 1ull) {
                    const int vx = i * 4 + (c >> 4), vy = j * 4 + ((c >> 2) & 3), vz = k * 4 + (c & 3);
                    const int base = (vx * WB_V + vy) * WB_V + vz;
                    float smin = INFINITY, smax = -INFINITY;
                    bool gate = false;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int idx = base + (q >> 2) * WB_V * WB_V + ((q >> 1) & 1) * WB_V + (q & 1);
                        const float sv = s_surf[idx];
                        smin = fminf(smin, sv);
                        smax = fmaxf(smax, sv);
                        gate |= !(s_dens[idx] < sigma_thresh);
                    }
                    bool has_surf = every_voxel != 0;
                    for (int q = 0; q < level_set_num && !has_surf; ++q) {
                        const float lv = level_set[q];
                        has_surf = !((lv < smin) || (lv > smax));
                    }
                    work = gate && has_surf;
                }
                bits[h] = __ballot_sync(0xffffffffu, work);
            }
            const uint64_t word = (uint64_t)bits[0] | ((uint64_t)bits[1] << 32);
            if (lane == 0) {
                out[k0] = word;
                if (word != 0) l1bits |= 1ull << sb;
            }
        }
        if (lane == 0 && l1bits != 0) {
            atomicOr(&s_l1[0], (unsigned)(l1bits & 0xffffffffull));
            atomicOr(&s_l1[1], (unsigned)(l1bits >> 32));
        }
        __syncthreads();
        if (tid == 0) out[lay.off[1] + w1] = (uint64_t)s_l1[0] | ((uint64_t)s_l1[1] << 32);
    }
}

// number of vertices with a link >= 0 (kept behind the block list; the regularisers use it to recognise the list of all
// stored vertices)
__global__ void __launch_bounds__(256) count_stored_kernel(const int32_t *__restrict__ links, int64_t n,
                                                            unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c += (links[i] >= 0) ? 1ull : 0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// After the three pyramid levels the occupancy buffer holds the list of non-empty level-1 (16^3) blocks:
// word off[3] = their number, then their indices as uint32 (2 per word).  The work-pyramid build walks this list.
__global__ void __launch_bounds__(256) accel_list1_kernel(AccelLayout lay, uint64_t *__restrict__ buf) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = (w < lay.count(1)) && (buf[lay.off[1] + w] != 0);
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd((unsigned long long *)(buf + lay.off[3]), (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (on) ((uint32_t *)(buf + lay.off[3] + 1))[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)w;
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
    const int every_voxel = (opt->surf_fake_sample && !opt->limited_fake_sample) ? 1 : 0;
    // words of blocks without linked cells stay zero: clear levels 0 and 1, then visit the non-empty 16^3 blocks only
    ASURF_CUDA(cudaMemsetAsync(work_out, 0, (size_t)lay.off[2] * sizeof(uint64_t), st));
    const int64_t n1 = lay.count(1);
    const int ctas = (int)(n1 < (int64_t)sms * 5 ? n1 : (int64_t)sms * 5);
    work_block16_kernel<<<ctas, WB_THREADS, 0, st>>>(grid->links, grid->density, grid->surface, grid->level_set,
                                                     grid->level_set_num, opt->sigma_thresh, every_voxel, grid->size[0],
                                                     grid->size[1], grid->size[2], lay, grid->accel, work_out);
    accel_coarsen_kernel<<<div_up(lay.count(2), 128), 128, 0, st>>>(lay, 2, work_out);
    note_launches(2);
    return check_cuda(cudaGetLastError(), "work_build launch");
}

extern "C" int64_t asurf_accel_words(const int32_t size[3]) {
    AccelLayout lay(size);
    return lay.off[3] + 1 + (lay.count(1) + 1) / 2 + 1;   // pyramid + list of non-empty level-1 blocks + stored-vertex count
}

extern "C" int asurf_accel_build(const int32_t *links, const int32_t size[3], uint64_t *accel_out, void *stream) {
    ASURF_REQUIRE(links && accel_out, ASURF_E_INVALID, "accel_build: null pointer");
    ASURF_REQUIRE(size[0] >= 2 && size[1] >= 2 && size[2] >= 2, ASURF_E_INVALID, "accel_build: grid smaller than 2^3");
    AccelLayout lay(size);
    cudaStream_t st = (cudaStream_t)stream;
    accel_level0_kernel<<<div_up(lay.count(0), 128), 128, 0, st>>>(links, size[0], size[1], size[2], lay, accel_out);
    accel_coarsen_kernel<<<div_up(lay.count(1), 128), 128, 0, st>>>(lay, 1, accel_out);
    accel_coarsen_kernel<<<div_up(lay.count(2), 128), 128, 0, st>>>(lay, 2, accel_out);
    ASURF_CUDA(cudaMemsetAsync(accel_out + lay.off[3], 0, sizeof(uint64_t), st));
    accel_list1_kernel<<<div_up(lay.count(1), 256), 256, 0, st>>>(lay, accel_out);
    uint64_t *n_stored = accel_out + lay.off[3] + 1 + (lay.count(1) + 1) / 2;
    ASURF_CUDA(cudaMemsetAsync(n_stored, 0, sizeof(uint64_t), st));
    count_stored_kernel<<<148 * 8, 256, 0, st>>>(links, (int64_t)size[0] * size[1] * size[2], (unsigned long long *)n_stored);
    note_launches(5);
    return check_cuda(cudaGetLastError(), "accel_build launch");
}
