// Host-side plumbing of libmvster_b200: thread-local error string, launch counter, device guard, version.
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace mvster {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int fail(mvster_status st, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return (int)st;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MVSTER_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return MVSTER_ERR_CUDA;
}

DeviceGuard::DeviceGuard(const void* ptr) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        status = fail(MVSTER_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return;
    }
    cudaPointerAttributes attr;
    e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        status = fail(MVSTER_ERR_NO_DEVICE, "output pointer %p is not CUDA device memory", ptr);
        return;
    }
    dev = attr.device;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) {
        e = cudaSetDevice(dev);
        if (e != cudaSuccess) status = check_cuda(e, "cudaSetDevice");
    }
}

DeviceGuard::~DeviceGuard() {
    if (prev >= 0 && dev >= 0 && prev != dev) cudaSetDevice(prev);
}

}  // namespace mvster

extern "C" int mvster_version(void) { return MVSTER_ABI_VERSION; }
extern "C" const char* mvster_last_error(void) { return mvster::g_err; }
extern "C" uint64_t mvster_launch_count(void) { return mvster::g_launches.load(std::memory_order_relaxed); }
