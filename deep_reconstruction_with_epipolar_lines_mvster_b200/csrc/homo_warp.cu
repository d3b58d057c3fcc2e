// Compatibility kernel for the reference's homo_warping(src_fea, src_proj, ref_proj, depth_values)
// (models/mvs4net_utils.py:21-67): materialises the [B, C, D, H, W] warped volume.  The fused path (epi_fwd.cu)
// never calls this; it exists so that code written against the reference function keeps working.
#include "common.cuh"

namespace mvster {

template <typename T>
__global__ void __launch_bounds__(256) homo_warp_kernel(const T* __restrict__ src, const float* __restrict__ rt,
                                                        const float* __restrict__ hypo, float* __restrict__ warped,
                                                        int C, int D, int H, int W, int Hs, int Ws) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z / D, d = blockIdx.z % D;
    if (x >= W || y >= H) return;
    const Homography h = load_homography(rt + (size_t)b * 12);
    const float fx = (float)x, fy = (float)y;
    const float ax = fmaf(h.r00, fx, fmaf(h.r01, fy, h.r02));
    const float ay = fmaf(h.r10, fx, fmaf(h.r11, fy, h.r12));
    const float az = fmaf(h.r20, fx, fmaf(h.r21, fy, h.r22));
    const size_t plane = (size_t)H * W, pix = (size_t)y * W + x;
    const Taps t = make_taps(ax, ay, az, h, hypo[((size_t)b * D + d) * plane + pix], Hs, Ws);
    const T* sp = src + (size_t)b * Hs * Ws * C;
    for (int c0 = 0; c0 < C; c0 += 8) {
        float wv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t.any) {
            const F8 a = load8<T>(sp + (size_t)t.o00 * C + c0), bq = load8<T>(sp + (size_t)t.o01 * C + c0);
            const F8 cq = load8<T>(sp + (size_t)t.o10 * C + c0), dq = load8<T>(sp + (size_t)t.o11 * C + c0);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                wv[c] = fmaf(t.w00, a.v[c], fmaf(t.w01, bq.v[c], fmaf(t.w10, cq.v[c], t.w11 * dq.v[c])));
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) stg_stream(warped + (((size_t)b * C + c0 + c) * D + d) * plane + pix, wv[c]);
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_homo_warp(const void* src, const float* rt, const float* hypo, float* warped, int B, int C,
                                int D, int H, int W, int Hs, int Ws, int dtype, void* stream) {
    if (!src || !rt || !hypo || !warped) return fail(MVSTER_ERR_BAD_ARG, "homo_warp: null pointer");
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0)
        return fail(MVSTER_ERR_BAD_ARG, "homo_warp: non-positive dimension");
    if (C % 8) return fail(MVSTER_ERR_UNSUPPORTED, "homo_warp: C=%d is not a multiple of 8", C);
    if ((size_t)B * D > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "homo_warp: B*D > 65535");
    if ((double)B * Hs * Ws * C >= 2147483648.0) return fail(MVSTER_ERR_UNSUPPORTED, "homo_warp: tensor too large");
    const uintptr_t align = (dtype == MVSTER_BF16) ? 16 : 32;
    if (((uintptr_t)src) % align) return fail(MVSTER_ERR_ALIGN, "homo_warp: src not %d-byte aligned", (int)align);
    if (((uintptr_t)rt) % 16) return fail(MVSTER_ERR_ALIGN, "homo_warp: rt not 16-byte aligned");
    DeviceGuard guard(warped);
    if (guard.status != MVSTER_OK) return guard.status;
    dim3 grid((W + 31) / 32, (H + 7) / 8, B * D);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == MVSTER_F32)
        homo_warp_kernel<float><<<grid, 256, 0, s>>>((const float*)src, rt, hypo, warped, C, D, H, W, Hs, Ws);
    else if (dtype == MVSTER_BF16)
        homo_warp_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)src, rt, hypo, warped, C, D, H, W,
                                                            Hs, Ws);
    else
        return fail(MVSTER_ERR_BAD_ARG, "homo_warp: unknown dtype %d", dtype);
    count_launch();
    MVSTER_CHECK_LAUNCH("homo_warp launch");
    return MVSTER_OK;
}
