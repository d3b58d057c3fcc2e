// Shared device/host helpers for libmvster_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mvster_b200.h"

namespace mvster {

// ---- host-side error plumbing (thread-local message, no exceptions across the C boundary) ------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int fail(mvster_status st, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// RAII device guard keyed on the device that owns `ptr`
struct DeviceGuard {
    int prev = -1;
    int dev = -1;
    int status = MVSTER_OK;
    explicit DeviceGuard(const void* ptr);
    ~DeviceGuard();
};

// Opt a kernel into more than 48 KB of dynamic shared memory.  The attribute is per (function, device): a process
// that drives several GPUs (nn.DataParallel, reference test_mvs4.py:393) must set it once on each of them.  The
// largest size set per device is remembered and the limit is raised again when a later call needs more (kernels whose
// shared-memory size depends on run-time arguments, e.g. the Sinkhorn history).  Safe to call from concurrent host threads.
template <typename Kernel>
static inline int ensure_dynamic_smem_bytes(Kernel kernel, int bytes, int (&largest)[64], const char* what) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (__atomic_load_n(&largest[dev], __ATOMIC_ACQUIRE) >= bytes) return MVSTER_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return check_cuda(e, what);
    int seen = __atomic_load_n(&largest[dev], __ATOMIC_ACQUIRE);
    while (seen < bytes && !__atomic_compare_exchange_n(&largest[dev], &seen, bytes, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
    }
    return MVSTER_OK;
}

#define MVSTER_CHECK_LAUNCH(what)                                   \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return mvster::check_cuda(e__, what); \
    } while (0)

// ---- device helpers ------------------------------------------------------------------------------------------
struct alignas(32) F8 {
    float v[8];
};

// 256-bit read-only global load (LDG.E.256, new on sm_100): one instruction fetches a lane's 8 fp32 channels.
__device__ __forceinline__ F8 ldg256(const float* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
                   "=f"(r.v[7])
                 : "l"(p));
    return r;
}

// 8 bf16 channels (16 bytes) -> 8 floats
__device__ __forceinline__ F8 ldg_bf16x8(const __nv_bfloat16* p) {
    uint4 q;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p));
    F8 r;
    r.v[0] = __uint_as_float(q.x << 16);
    r.v[1] = __uint_as_float(q.x & 0xffff0000u);
    r.v[2] = __uint_as_float(q.y << 16);
    r.v[3] = __uint_as_float(q.y & 0xffff0000u);
    r.v[4] = __uint_as_float(q.z << 16);
    r.v[5] = __uint_as_float(q.z & 0xffff0000u);
    r.v[6] = __uint_as_float(q.w << 16);
    r.v[7] = __uint_as_float(q.w & 0xffff0000u);
    return r;
}

template <typename T>
__device__ __forceinline__ F8 load8(const T* p);
template <>
__device__ __forceinline__ F8 load8<float>(const float* p) {
    return ldg256(p);
}
template <>
__device__ __forceinline__ F8 load8<__nv_bfloat16>(const __nv_bfloat16* p) {
    return ldg_bf16x8(p);
}

// streaming (read-once) scalar load that does not allocate in L1
__device__ __forceinline__ float ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// streaming store (evict-first): outputs are consumed by the next kernel, never re-read here
__device__ __forceinline__ void stg_stream(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// vector reduction into global memory (no return value): RED.E.ADD.F32x4 on sm_90+
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, new on sm_100): one issue slot for two lanes of math --------
typedef unsigned long long f32x2;  // two floats in one 64-bit register pair (lo = even channel, hi = odd channel)

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// 8 channels of one texel as 4 packed pairs
struct P8 {
    f32x2 q[4];
};

// 256-bit read-only load straight into four 64-bit register pairs (LDG.E.256)
__device__ __forceinline__ P8 ldg256_pairs(const void* p) {
    P8 r;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=l"(r.q[0]), "=l"(r.q[1]), "=l"(r.q[2]), "=l"(r.q[3])
                 : "l"(p));
    return r;
}
// 8 bf16 channels (16 bytes) widened to 4 packed fp32 pairs
__device__ __forceinline__ P8 ldg_bf16x8_pairs(const void* p) {
    uint4 q;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p));
    P8 r;
    r.q[0] = pack2(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u));
    r.q[1] = pack2(__uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
    r.q[2] = pack2(__uint_as_float(q.z << 16), __uint_as_float(q.z & 0xffff0000u));
    r.q[3] = pack2(__uint_as_float(q.w << 16), __uint_as_float(q.w & 0xffff0000u));
    return r;
}
template <typename T>
__device__ __forceinline__ P8 load_pairs(const void* p);
template <>
__device__ __forceinline__ P8 load_pairs<float>(const void* p) {
    return ldg256_pairs(p);
}
template <>
__device__ __forceinline__ P8 load_pairs<__nv_bfloat16>(const void* p) {
    return ldg_bf16x8_pairs(p);
}

// 1/x to <= 1 ulp without the IEEE-division slow path: MUFU.RCP + one Newton step
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    const float e = fmaf(-x, r, 1.0f);
    return fmaf(r, e, r);
}

// Per-(batch, view) homography [R | t], 12 floats, uniform across a CTA.
struct Homography {
    float r00, r01, r02, t0, r10, r11, r12, t1, r20, r21, r22, t2;
};
__device__ __forceinline__ Homography load_homography(const float* rt) {
    Homography h;
    const float4* q = reinterpret_cast<const float4*>(rt);
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    h.r00 = a.x; h.r01 = a.y; h.r02 = a.z; h.t0 = a.w;
    h.r10 = b.x; h.r11 = b.y; h.r12 = b.z; h.t1 = b.w;
    h.r20 = c.x; h.r21 = c.y; h.r22 = c.z; h.t2 = c.w;
    return h;
}

// Bilinear footprint of one sample: 4 clamped tap offsets (in texels) and 4 weights; out-of-image taps have
// weight 0 (grid_sample padding_mode='zeros', align_corners=True; reference models/mvs4net_utils.py:59).
struct Taps {
    int o00, o01, o10, o11;   // texel offsets  y*Ws + x
    float w00, w01, w10, w11; // (row y0: x0, x0+1) (row y0+1: x0, x0+1)
    bool any;
};

// Sample position of one hypothesis, clamped to [-1, Ws] x [-1, Hs].  A clamped coordinate has both of its taps
// outside the image (or a zero weight on the one inside), so the sample contributes nothing - exactly like the
// unclamped out-of-range sample under padding_mode='zeros' - while staying next to the image for the bounding box.
// NaN collapses to -1 (the CPU reference samples nothing for NaN coordinates either).
// p = R*[x,y,1]*d + t ; z==0 -> 1e-9 ; p.xy / z                                         (reference :42-48)
__device__ __forceinline__ void sample_pos(float ax, float ay, float az, const Homography& h, float d, float wlim,
                                           float hlim, float& sx, float& sy) {
    const float px = fmaf(ax, d, h.t0);
    const float py = fmaf(ay, d, h.t1);
    float pz = fmaf(az, d, h.t2);
    pz = (pz == 0.0f) ? 1e-9f : pz;
    const float rz = fast_rcp(pz);
    sx = fminf(fmaxf(px * rz, -1.0f), wlim);
    sy = fminf(fmaxf(py * rz, -1.0f), hlim);
}

// Taps of one sample for the direct-gather kernels: clamped texel offsets and bounds-zeroed bilinear weights
// (grid_sample padding_mode='zeros', align_corners=True; reference models/mvs4net_utils.py:59).
__device__ __forceinline__ Taps make_taps(float ax, float ay, float az, const Homography& h, float d, int Hs,
                                          int Ws) {
    float sx, sy;
    sample_pos(ax, ay, az, h, d, (float)Ws, (float)Hs, sx, sy);
    const float x0f = floorf(sx), y0f = floorf(sy);
    const float fx = sx - x0f, fy = sy - y0f;
    const int x0 = (int)x0f, y0 = (int)y0f;  // in [-1, Ws] x [-1, Hs] after the clamp
    const bool vx0 = (unsigned)x0 < (unsigned)Ws, vx1 = (unsigned)(x0 + 1) < (unsigned)Ws;
    const bool vy0 = (unsigned)y0 < (unsigned)Hs, vy1 = (unsigned)(y0 + 1) < (unsigned)Hs;
    const int xc0 = min(max(x0, 0), Ws - 1), xc1 = min(x0 + 1, Ws - 1);
    const int yc0 = min(max(y0, 0), Hs - 1), yc1 = min(y0 + 1, Hs - 1);
    const float gx = vx0 ? 1.0f - fx : 0.0f, hx = vx1 ? fx : 0.0f;
    const float gy = vy0 ? 1.0f - fy : 0.0f, hy = vy1 ? fy : 0.0f;
    Taps t;
    t.o00 = yc0 * Ws + xc0;
    t.o01 = yc0 * Ws + xc1;
    t.o10 = yc1 * Ws + xc0;
    t.o11 = yc1 * Ws + xc1;
    t.w00 = gx * gy;
    t.w01 = hx * gy;
    t.w10 = gx * hy;
    t.w11 = hx * hy;
    t.any = (t.w00 != 0.0f) || (t.w01 != 0.0f) || (t.w10 != 0.0f) || (t.w11 != 0.0f);
    return t;
}

}  // namespace mvster
