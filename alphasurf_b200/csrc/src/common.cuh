// alphasurf_b200: shared host/device declarations for the sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "asurf.h"

namespace asurf {

// ---- error plumbing -------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

#define ASURF_CUDA(expr)                                   \
    do {                                                   \
        int _rc = ::asurf::check_cuda((expr), #expr);      \
        if (_rc != 0) return _rc;                          \
    } while (0)

#define ASURF_REQUIRE(cond, code, ...)                     \
    do {                                                   \
        if (!(cond)) {                                     \
            ::asurf::set_error(__VA_ARGS__);               \
            return (code);                                 \
        }                                                  \
    } while (0)

// ---- growable device workspace (one per purpose, owned by the library) -----------------------------------
struct Workspace {
    void *ptr = nullptr;
    size_t bytes = 0;
    int device = -1;
    int reserve(size_t need);   // grows geometrically; contents are NOT preserved
    void release();
};

// ---- launch accounting (asurf_launch_count): every kernel this library enqueues is counted ------------------
void note_launches(int n);
unsigned long long launches_read(int reset);

// work pyramid of a render call (accel.cu): incremental update of a cached pyramid when the grid is unchanged
int work_pyramid_for_call(const asurf_grid_t *grid, const asurf_opt_t *opt, cudaStream_t st, const uint64_t **work_out,
                          int slot = 0);
void work_cache_release();
void msi_release();
// per-ray state a foreground pass leaves for the MSI background pass (msi.cu): two (Q,) float arrays in a library workspace
int bg_state_reserve(int64_t Q, float **log_transmit, float **accum);
static inline bool grid_has_background(const asurf_grid_t *g) {
    return g->background_links != nullptr && g->background_data != nullptr && g->background_nlayers > 0;
}
void loss_release();    // per-file workspaces, freed by asurf_release
void cuvol_release();
void misc_release();
int work_cache_copy(uint64_t *out, int64_t words, cudaStream_t st);

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// The asurf_debug_set_* switches (algorithm variants the tests compare) are inert unless the process was started with
// ASURF_DEBUG_HOOKS=1: a production process cannot be reconfigured through the ABI by accident.
static inline bool debug_hooks_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("ASURF_DEBUG_HOOKS");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// ---- occupancy pyramid layout (accel.cu) -----------------------------------------------------------------
// A "cell" (x,y,z) is the voxel spanned by vertices (x..x+1, y..y+1, z..z+1); it is ACTIVE iff its 8 corner
// links are all >= 0.  Level 0: one 64-bit word per 4x4x4 block of cells, bit = (x&3)<<4 | (y&3)<<2 | (z&3).
// Level k+1: one word per 4x4x4 block of level-k words, bit set iff that child word != 0.
// So "word == 0" at level k means an empty block of 4^(k+1) cells per side (4, 16, 64).
// All levels live in one buffer: [level 0 | level 1 | level 2].
struct AccelLayout {
    int b[3][3];      // b[level][axis]
    int64_t off[4];   // word offset of each level; off[3] = total
    __host__ __device__ AccelLayout() {}
    __host__ __device__ explicit AccelLayout(const int32_t size[3]) {
        off[0] = 0;
        for (int l = 0; l < 3; ++l) {
            for (int i = 0; i < 3; ++i) {
                const int prev = (l == 0) ? (size[i] - 1 > 0 ? size[i] - 1 : 1) : b[l - 1][i];
                b[l][i] = (prev + 3) >> 2;
            }
            off[l + 1] = off[l] + (int64_t)b[l][0] * b[l][1] * b[l][2];
        }
    }
    __host__ __device__ int64_t count(int l) const { return off[l + 1] - off[l]; }
};

// Tail of the occupancy buffer (asurf_accel_build): [pyramid | n level-1 blocks, their list | n stored vertices |
// n vertex blocks with a stored vertex, their list].  Word offsets of the last two parts:
static inline int64_t accel_stored_offset(const int32_t size[3]) {
    AccelLayout lay(size);
    return lay.off[3] + 1 + (lay.count(1) + 1) / 2;
}
static inline int64_t accel_vblock_offset(const int32_t size[3]) { return accel_stored_offset(size) + 1; }
static inline int64_t accel_vblock_count(const int32_t size[3]) {
    return (int64_t)((size[0] + 15) / 16) * ((size[1] + 15) / 16) * ((size[2] + 15) / 16);
}
// ... followed by the exclusive prefix count of stored vertices per (x, y) column: X * Y + 1 uint32 (the number of stored
// vertices with a flat id below any bound is then one table read plus a partial column)
static inline int64_t accel_colprefix_offset(const int32_t size[3]) {
    return accel_vblock_offset(size) + 1 + (accel_vblock_count(size) + 1) / 2;
}
static inline int64_t accel_total_words(const int32_t size[3]) {
    return accel_colprefix_offset(size) + ((int64_t)size[0] * size[1] + 2) / 2;
}

}  // namespace asurf
