// Host-side plumbing of libmvster_b200: thread-local error string, launch counter, device guard, version.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>

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

// Scratch buffers of kernels that need one (K1 forward: per-tile bounding boxes).  One grow-only allocation per
// (device, stream): launches on one stream reuse it in stream order, launches on different streams never share it.
struct Workspace {
    int dev;
    cudaStream_t stream;
    void* ptr;
    size_t bytes;
};
static std::mutex g_ws_mutex;
static std::vector<Workspace> g_ws;

void* epi_workspace(cudaStream_t stream, size_t bytes, int* status) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    Workspace* w = nullptr;
    for (auto& e : g_ws)
        if (e.dev == dev && e.stream == stream) { w = &e; break; }
    if (w && w->bytes >= bytes) return w->ptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) cudaGetLastError();
    if (cap != cudaStreamCaptureStatusNone) {
        *status = fail(MVSTER_ERR_CUDA, "epi_fwd: scratch of %zu bytes must be allocated outside CUDA graph capture; "
                       "run the same call once eagerly on this stream first", bytes);
        return nullptr;
    }
    const size_t want = bytes + bytes / 4 + 4096;
    void* ptr = nullptr;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) { *status = check_cuda(e, "epi_fwd: cudaMalloc(scratch)"); return nullptr; }
    if (w) {
        // the old buffer may still be read by kernels in flight on this stream
        cudaStreamSynchronize(stream);
        cudaFree(w->ptr);
        w->ptr = ptr; w->bytes = want;
    } else {
        g_ws.push_back(Workspace{dev, stream, ptr, want});
    }
    return ptr;
}

}  // namespace mvster

extern "C" int mvster_version(void) { return MVSTER_ABI_VERSION; }
extern "C" const char* mvster_last_error(void) { return mvster::g_err; }
extern "C" uint64_t mvster_launch_count(void) { return mvster::g_launches.load(std::memory_order_relaxed); }
