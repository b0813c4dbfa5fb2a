// alphasurf_b200: C-ABI plumbing (error reporting, workspaces).  See include/asurf.h.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace asurf {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

static unsigned long long g_launches = 0;
void note_launches(int n) { g_launches += (unsigned long long)n; }
unsigned long long launches_read(int reset) {
    const unsigned long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int Workspace::reserve(size_t need) {
    int dev = 0;
    ASURF_CUDA(cudaGetDevice(&dev));
    if (ptr != nullptr && dev == device && bytes >= need) return 0;
    if (ptr != nullptr) {
        cudaDeviceSynchronize();
        cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
    size_t want = need + need / 2 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = need;
        e = cudaMalloc(&ptr, want);
    }
    if (e != cudaSuccess) {
        ptr = nullptr;
        set_error("workspace allocation of %zu bytes failed: %s", need, cudaGetErrorString(e));
        cudaGetLastError();
        return ASURF_E_NOMEM;
    }
    bytes = want;
    device = dev;
    return 0;
}

void Workspace::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    device = -1;
}

}  // namespace asurf

extern "C" const char *asurf_last_error(void) { return asurf::g_err; }
extern "C" int asurf_abi_version(void) { return ASURF_ABI_VERSION; }
extern "C" uint64_t asurf_launch_count(int32_t reset) { return asurf::launches_read(reset); }
