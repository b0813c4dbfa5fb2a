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


// List of the 16^3 VERTEX blocks (vertices [16b, 16b + 16) per axis) that hold a stored vertex, kept behind the stored-vertex
// count: word = their number, then their block coordinates packed as bx | by << 10 | bz << 20 (uint32, 2 per word).  The tiled regularisers (loss.cu) walk it: the
// non-empty level-1 list above only names blocks with a COMPLETE cell, which a TV term does not need.
__global__ void __launch_bounds__(256) vblock_list_kernel(const int32_t *__restrict__ links, int sx, int sy, int sz, int nby,
                                                           int nbz, uint64_t *__restrict__ out) {
    const int w = blockIdx.x;
    const int bz = w % nbz, by = (w / nbz) % nby, bx = w / (nbz * nby);
    const int x = bx * 16 + (threadIdx.x >> 4), y = by * 16 + (threadIdx.x & 15), z0 = bz * 16;
    int any = 0;
    if (x < sx && y < sy) {
        const int32_t *p = links + ((int64_t)x * sy + y) * sz;
        const int ze = (z0 + 16 < sz) ? z0 + 16 : sz;
        for (int z = z0; z < ze; ++z) any |= (__ldg(p + z) >= 0) ? 1 : 0;
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0 && any) {
        const unsigned long long at = atomicAdd((unsigned long long *)out, 1ull);
        ((uint32_t *)(out + 1))[at] = (uint32_t)bx | ((uint32_t)by << 10) | ((uint32_t)bz << 20);
    }
}

// Stored vertices per (x, y) column (warp per column), then their exclusive prefix sum in place (one CTA: the table has
// X * Y + 1 entries and is rebuilt only when `links` changes).
__global__ void __launch_bounds__(256) column_count_kernel(const int32_t *__restrict__ links, int64_t n_cols, int sz,
                                                            uint32_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    for (int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; col < n_cols;
         col += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int32_t *p = links + col * sz;
        unsigned c = 0;
        for (int z = lane; z < sz; z += 32) c += (__ldg(p + z) >= 0) ? 1u : 0u;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if (lane == 0) out[col] = c;
    }
}
__global__ void __launch_bounds__(1024) column_prefix_kernel(uint32_t *__restrict__ tab, int64_t n_cols) {
    // in: tab[i] = count of column i (i < n_cols); out: tab[i] = sum of the counts of the columns below i, tab[n_cols] = total
    __shared__ unsigned s_part[1024];
    const int t = threadIdx.x;
    const int64_t per = (n_cols + 1023) / 1024;
    const int64_t b = (int64_t)t * per, e = (b + per < n_cols) ? b + per : n_cols;
    unsigned sum = 0;
    for (int64_t i = b; i < e; ++i) sum += tab[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned run = 0;
        for (int i = 0; i < 1024; ++i) {
            const unsigned v = s_part[i];
            s_part[i] = run;
            run += v;
        }
        tab[n_cols] = run;
    }
    __syncthreads();
    unsigned run = s_part[t];
    for (int64_t i = b; i < e; ++i) {
        const unsigned v = tab[i];
        tab[i] = run;
        run += v;
    }
}

// ---- incremental maintenance of the work pyramid across calls -------------------------------------------------------------
// A voxel's work bit is a function of its 8 corner vertices through three per-vertex predicates only:
// (surface < lv), (surface > lv) [smin <= lv <= smax  <=>  not all corners > lv and not all < lv] and (density >= sigma_thresh).
// Training changes every stored value a little every step, but those predicates flip for a handful of vertices.  So the
// library keeps, per grid, the class byte of every data row; each call re-reads surface / density once, in row order
// (coalesced, 8 B per row instead of 1.2 x 12 B per vertex through `links`), and re-derives -- from the actual data -- only
// the <= 8 voxels around each vertex whose class changed.  The cache validates itself against the data on every call;
// links / options changes are caught by the key (pointers, options, generation of the occupancy buffer).
__global__ void __launch_bounds__(256)
inverse_links_kernel(const int32_t *__restrict__ links, int64_t n_vertices, int32_t *__restrict__ inv) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vertices; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t l = links[i];
        if (l >= 0) inv[l] = (int32_t)i;
    }
}

__device__ __forceinline__ uint8_t vertex_class(float s, float dn, float lv, float sigma_thresh) {
    return (uint8_t)((s < lv ? 1 : 0) | (s > lv ? 2 : 0) | (!(dn < sigma_thresh) ? 4 : 0));
}

// INIT: fill the class bytes.  Otherwise: compare, update, and list the rows whose class changed.
template <bool INIT>
__global__ void __launch_bounds__(256)
class_scan_kernel(const float *__restrict__ surface, const float *__restrict__ density, int64_t n_rows,
                  const float *__restrict__ level_set, float sigma_thresh, uint8_t *__restrict__ cls,
                  int32_t *__restrict__ changed, unsigned long long *__restrict__ n_changed) {
    const int lane = threadIdx.x & 31;
    const float lv = __ldg(level_set);   // read on the device every call: a changed level set is just more changed classes
    for (int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ll; r0 < n_rows; r0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = r0 + lane;
        bool diff = false;
        if (r < n_rows) {
            const uint8_t c = vertex_class(surface[r], density[r], lv, sigma_thresh);
            if (INIT) {
                cls[r] = c;
            } else if (c != cls[r]) {
                cls[r] = c;
                diff = true;
            }
        }
        if (!INIT) {
            const unsigned m = __ballot_sync(0xffffffffu, diff);
            if (m) {
                unsigned long long base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(n_changed, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (diff) changed[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)r;
            }
        }
    }
}

// 8 threads per changed row: thread k re-derives the voxel whose corner k is the changed vertex.
__global__ void __launch_bounds__(256)
work_patch_kernel(const int32_t *__restrict__ links, const float *__restrict__ density, const float *__restrict__ surface,
                  const float *__restrict__ level_set, float sigma_thresh, int sx, int sy, int sz, AccelLayout lay,
                  const uint64_t *__restrict__ occ,
                  const int32_t *__restrict__ inv, const int32_t *__restrict__ changed,
                  const unsigned long long *__restrict__ n_changed, uint64_t *__restrict__ work) {
    const int64_t n = (int64_t)*n_changed * 8;
    const float lv = __ldg(level_set);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t flat = inv[changed[i >> 3]];
        const int k = (int)(i & 7);
        const int vz = flat % sz, vy = (flat / sz) % sy, vx = flat / (sz * sy);
        const int x = vx - (k >> 2), y = vy - ((k >> 1) & 1), z = vz - (k & 1);
        if (x < 0 || y < 0 || z < 0 || x >= sx - 1 || y >= sy - 1 || z >= sz - 1) continue;
        const int64_t k0 = ((int64_t)(x >> 2) * lay.b[0][1] + (y >> 2)) * lay.b[0][2] + (z >> 2);
        const int bit = ((x & 3) << 4) | ((y & 3) << 2) | (z & 3);
        if (!((occ[k0] >> bit) & 1ull)) continue;   // not all 8 links: never a work voxel
        float smin = INFINITY, smax = -INFINITY;
        bool gate = false;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int32_t l = links[((int64_t)(x + (q >> 2)) * sy + (y + ((q >> 1) & 1))) * sz + (z + (q & 1))];
            const float sv = surface[l];
            smin = fminf(smin, sv);
            smax = fmaxf(smax, sv);
            gate |= !(density[l] < sigma_thresh);
        }
        const bool w = gate && !((lv < smin) || (lv > smax));
        if (w) atomicOr((unsigned long long *)(work + k0), 1ull << bit);
        else atomicAnd((unsigned long long *)(work + k0), ~(1ull << bit));
    }
}

// The update scan, four rows per thread (128-bit loads of surface / density, one 32-bit load of the class bytes); changed
// rows are rare (a few thousand per step), so they are appended one by one.
__global__ void __launch_bounds__(256)
class_scan4_kernel(const float *__restrict__ surface, const float *__restrict__ density, int64_t n_rows,
                   const float *__restrict__ level_set, float sigma_thresh, uint8_t *__restrict__ cls,
                   int32_t *__restrict__ changed, unsigned long long *__restrict__ n_changed) {
    const float lv = __ldg(level_set);
    const int64_t n4 = n_rows >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
        const float4 s = __ldg((const float4 *)surface + q), dn = __ldg((const float4 *)density + q);
        const uint32_t old = ((const uint32_t *)cls)[q];
        const uint32_t now = (uint32_t)vertex_class(s.x, dn.x, lv, sigma_thresh) | ((uint32_t)vertex_class(s.y, dn.y, lv, sigma_thresh) << 8) |
                             ((uint32_t)vertex_class(s.z, dn.z, lv, sigma_thresh) << 16) |
                             ((uint32_t)vertex_class(s.w, dn.w, lv, sigma_thresh) << 24);
        if (now != old) {
            ((uint32_t *)cls)[q] = now;
            const uint32_t diff = now ^ old;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((diff >> (8 * k)) & 0xffu) changed[atomicAdd(n_changed, 1ull)] = (int32_t)(q * 4 + k);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n_rows & 3)) {   // tail rows
        const int64_t r = (n4 << 2) + threadIdx.x;
        const uint8_t c = vertex_class(surface[r], density[r], lv, sigma_thresh);
        if (c != cls[r]) {
            cls[r] = c;
            changed[atomicAdd(n_changed, 1ull)] = (int32_t)r;
        }
    }
}

// After work_patch_kernel: the level-1 / level-2 bits above the voxels that were re-derived, recomputed from the (now final)
// words of the level below.  Replaces two full coarsening passes by two passes over the changed list.
__global__ void __launch_bounds__(256)
work_fix_level_kernel(int level, int sx, int sy, int sz, AccelLayout lay, const int32_t *__restrict__ inv,
                      const int32_t *__restrict__ changed, const unsigned long long *__restrict__ n_changed,
                      uint64_t *__restrict__ work) {
    const int64_t n = (int64_t)*n_changed * 8;
    const int sh = 2 * level;   // voxel coordinate -> word coordinate of level-1 below: >> sh
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t flat = inv[changed[i >> 3]];
        const int k = (int)(i & 7);
        const int vz = flat % sz, vy = (flat / sz) % sy, vx = flat / (sz * sy);
        const int x = vx - (k >> 2), y = vy - ((k >> 1) & 1), z = vz - (k & 1);
        if (x < 0 || y < 0 || z < 0 || x >= sx - 1 || y >= sy - 1 || z >= sz - 1) continue;
        // word of the level below that holds the voxel, and its bit in the word of this level
        const int cx = x >> sh, cy = y >> sh, cz = z >> sh;
        const uint64_t below = work[lay.off[level - 1] + ((int64_t)cx * lay.b[level - 1][1] + cy) * lay.b[level - 1][2] + cz];
        const int64_t kw = lay.off[level] + ((int64_t)(cx >> 2) * lay.b[level][1] + (cy >> 2)) * lay.b[level][2] + (cz >> 2);
        const int bit = ((cx & 3) << 4) | ((cy & 3) << 2) | (cz & 3);
        if (below != 0) atomicOr((unsigned long long *)(work + kw), 1ull << bit);
        else atomicAnd((unsigned long long *)(work + kw), ~(1ull << bit));
    }
}

struct WorkCache {
    Workspace work, cls, inv, changed, ctr;
    const void *links = nullptr, *surface = nullptr, *density = nullptr, *accel = nullptr;
    int32_t size[3] = {0, 0, 0};
    int64_t capacity = 0;
    float sigma_thresh = 0.f;
    unsigned long long accel_gen = 0;
    bool valid = false;
} g_wcs[2];   // [0] the training renderer's options, [1] the gate-free predicate of the evaluation renders

unsigned long long g_accel_gen = 0;          // bumped by every asurf_accel_build
const void *g_accel_last[16] = {nullptr};    // occupancy buffers built by this library and their generation
unsigned long long g_accel_last_gen[16] = {0};

unsigned long long accel_generation(const void *accel) {
    for (int i = 0; i < 16; ++i)
        if (g_accel_last[i] == accel) return g_accel_last_gen[i];
    return 0;   // unknown buffer: never cached
}

}  // namespace

// Work pyramid for the render call: incremental when the grid is the one of the previous call, full build otherwise.
int work_pyramid_for_call(const asurf_grid_t *grid, const asurf_opt_t *opt, cudaStream_t st, const uint64_t **work_out,
                          int slot) {
    WorkCache &g_wc = g_wcs[slot ? 1 : 0];
    AccelLayout lay(grid->size);
    const bool every_voxel = opt->surf_fake_sample && !opt->limited_fake_sample;
    const unsigned long long gen = accel_generation(grid->accel);
    const bool cacheable = gen != 0 && grid->level_set_num == 1 && !every_voxel && grid->capacity > 0 &&
                           (int64_t)grid->size[0] * grid->size[1] * grid->size[2] < ((int64_t)1 << 31);
    int rc = g_wc.work.reserve((size_t)lay.off[3] * sizeof(uint64_t));
    if (rc) return rc;
    uint64_t *work = (uint64_t *)g_wc.work.ptr;
    *work_out = work;
    const bool hit = cacheable && g_wc.valid && g_wc.links == grid->links && g_wc.surface == grid->surface &&
                     g_wc.density == grid->density && g_wc.accel == grid->accel && g_wc.accel_gen == gen &&
                     g_wc.capacity == grid->capacity && g_wc.size[0] == grid->size[0] && g_wc.size[1] == grid->size[1] &&
                     g_wc.size[2] == grid->size[2] && g_wc.sigma_thresh == opt->sigma_thresh;
    const int64_t N = grid->capacity;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int scan_blocks = (int)((N + 255) / 256 < (int64_t)sms * 16 ? (N + 255) / 256 : (int64_t)sms * 16);
    if (hit) {
        unsigned long long *ctr = (unsigned long long *)g_wc.ctr.ptr;
        ASURF_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st));
        const bool vec4 = ((((uintptr_t)grid->surface) | ((uintptr_t)grid->density)) & 15) == 0;
        if (vec4)
            class_scan4_kernel<<<sms * 8, 256, 0, st>>>(grid->surface, grid->density, N, grid->level_set, opt->sigma_thresh,
                                                        (uint8_t *)g_wc.cls.ptr, (int32_t *)g_wc.changed.ptr, ctr);
        else
            class_scan_kernel<false><<<scan_blocks, 256, 0, st>>>(grid->surface, grid->density, N, grid->level_set,
                                                                  opt->sigma_thresh, (uint8_t *)g_wc.cls.ptr,
                                                                  (int32_t *)g_wc.changed.ptr, ctr);
        work_patch_kernel<<<sms * 4, 256, 0, st>>>(grid->links, grid->density, grid->surface, grid->level_set, opt->sigma_thresh,
                                                   grid->size[0], grid->size[1], grid->size[2], lay, grid->accel,
                                                   (const int32_t *)g_wc.inv.ptr, (const int32_t *)g_wc.changed.ptr, ctr, work);
        // levels 1 and 2: only the bits above the re-derived voxels can have changed
        for (int level = 1; level <= 2; ++level)
            work_fix_level_kernel<<<sms * 2, 256, 0, st>>>(level, grid->size[0], grid->size[1], grid->size[2], lay,
                                                          (const int32_t *)g_wc.inv.ptr, (const int32_t *)g_wc.changed.ptr, ctr,
                                                          work);
        note_launches(4);
        return check_cuda(cudaGetLastError(), "work pyramid update");
    }
    g_wc.valid = false;
    rc = asurf_work_build(grid, opt, work, st);
    if (rc || !cacheable) return rc;
    // start a cache for this grid: class bytes, row -> vertex map
    rc = g_wc.cls.reserve((size_t)N);
    if (!rc) rc = g_wc.inv.reserve((size_t)N * sizeof(int32_t));
    if (!rc) rc = g_wc.changed.reserve((size_t)N * sizeof(int32_t));
    if (!rc) rc = g_wc.ctr.reserve(sizeof(unsigned long long));
    if (rc) return rc;
    const int64_t nv = (int64_t)grid->size[0] * grid->size[1] * grid->size[2];
    inverse_links_kernel<<<sms * 16, 256, 0, st>>>(grid->links, nv, (int32_t *)g_wc.inv.ptr);
    class_scan_kernel<true><<<scan_blocks, 256, 0, st>>>(grid->surface, grid->density, N, grid->level_set, opt->sigma_thresh,
                                                         (uint8_t *)g_wc.cls.ptr, nullptr, nullptr);
    note_launches(2);
    g_wc.links = grid->links; g_wc.surface = grid->surface; g_wc.density = grid->density; g_wc.accel = grid->accel;
    g_wc.accel_gen = gen; g_wc.capacity = N; g_wc.sigma_thresh = opt->sigma_thresh;
    for (int i = 0; i < 3; ++i) g_wc.size[i] = grid->size[i];
    g_wc.valid = true;
    return check_cuda(cudaGetLastError(), "work pyramid cache init");
}

int work_cache_copy(uint64_t *out, int64_t words, cudaStream_t st) {
    WorkCache &g_wc = g_wcs[0];
    ASURF_REQUIRE(g_wc.work.ptr && (size_t)words * sizeof(uint64_t) <= g_wc.work.bytes, ASURF_E_INVALID,
                  "work cache: nothing cached / size mismatch");
    return check_cuda(cudaMemcpyAsync(out, g_wc.work.ptr, (size_t)words * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st),
                      "work cache copy");
}

void work_cache_release() {
    for (WorkCache &g_wc : g_wcs) {
        g_wc.work.release(); g_wc.cls.release(); g_wc.inv.release(); g_wc.changed.release(); g_wc.ctr.release();
        g_wc.valid = false;
    }
}

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
    // pyramid + list of non-empty level-1 blocks + stored-vertex count + list of the vertex blocks with a stored vertex
    // + per-column prefix counts of the stored vertices
    return accel_total_words(size);
}

extern "C" int asurf_accel_build(const int32_t *links, const int32_t size[3], uint64_t *accel_out, void *stream) {
    ASURF_REQUIRE(links && accel_out, ASURF_E_INVALID, "accel_build: null pointer");
    ASURF_REQUIRE(size[0] >= 2 && size[1] >= 2 && size[2] >= 2, ASURF_E_INVALID, "accel_build: grid smaller than 2^3");
    {   // a new generation for this buffer: work pyramids cached for an older content of the address are dropped
        ++g_accel_gen;
        int slot = (int)(g_accel_gen % 16);
        for (int i = 0; i < 16; ++i)
            if (g_accel_last[i] == (const void *)accel_out) slot = i;
        g_accel_last[slot] = (const void *)accel_out;
        g_accel_last_gen[slot] = g_accel_gen;
    }
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
    uint64_t *vbl = accel_out + accel_vblock_offset(size);
    ASURF_CUDA(cudaMemsetAsync(vbl, 0, sizeof(uint64_t), st));
    const int nby = (size[1] + 15) / 16, nbz = (size[2] + 15) / 16;
    vblock_list_kernel<<<(unsigned)accel_vblock_count(size), 256, 0, st>>>(links, size[0], size[1], size[2], nby, nbz, vbl);
    uint32_t *colp = (uint32_t *)(accel_out + accel_colprefix_offset(size));
    const int64_t n_columns = (int64_t)size[0] * size[1];
    column_count_kernel<<<148 * 8, 256, 0, st>>>(links, n_columns, size[2], colp);
    column_prefix_kernel<<<1, 1024, 0, st>>>(colp, n_columns);
    note_launches(8);
    return check_cuda(cudaGetLastError(), "accel_build launch");
}

extern "C" int asurf_debug_work_cache_copy(uint64_t *out, int64_t words, void *stream) {
    return work_cache_copy(out, words, (cudaStream_t)stream);
}
extern "C" int32_t asurf_debug_work_cache_valid(void) { return g_wcs[0].valid ? 1 : 0; }
