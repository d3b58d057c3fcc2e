// K2a: depth / confidence tail of stagenet.forward (models/mvs4net_utils.py:1109-1156) as one streaming kernel,
// and the hypothesis schedule (models/mvs4net_utils.py:79-94).  All HBM-bound elementwise work over [B,D,H,W]:
// one thread per pixel, D planes read with coalesced 128-byte warp requests, everything else in registers.
#include "tail_common.cuh"

namespace mvster {

struct TailParams {
    const float* logits;
    const float* hypo;
    float* attn;
    float* depth;
    float* conf;
    float* inv_min;
    float* inv_max;
    float split_itv;
    int mode;
    int D;
    size_t plane;  // H*W
    size_t total;  // B*H*W
};

// D compile-time (registers) when DT > 0, else a three-pass loop over p.D planes (re-reads hit L1/L2).
template <int DT>
__global__ void __launch_bounds__(256) tail_kernel(const TailParams p) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.total) return;
    const size_t b = i / p.plane, r = i - b * p.plane;
    const int D = DT > 0 ? DT : p.D;
    const float* lg = p.logits + b * D * p.plane + r;
    const float* hy = p.hypo + b * D * p.plane + r;
    float* at = p.attn + b * D * p.plane + r;

    if constexpr (DT > 0) {
        float l[DT > 0 ? DT : 1], h[DT > 0 ? DT : 1], a[DT > 0 ? DT : 1];
#pragma unroll
        for (int d = 0; d < DT; ++d) l[d] = ldg_stream(lg + d * p.plane);
#pragma unroll
        for (int d = 0; d < DT; ++d) h[d] = ldg_stream(hy + d * p.plane);
        const TailOut r = tail_pixel<(DT > 0 ? DT : 2)>(l, h, p.mode, p.split_itv, a);
#pragma unroll
        for (int d = 0; d < DT; ++d) stg_stream(at + d * p.plane, a[d]);
        p.depth[i] = r.depth;
        if (p.conf != nullptr) p.conf[i] = r.conf;
        if (p.inv_min != nullptr) {
            p.inv_min[i] = r.inv_min;
            p.inv_max[i] = r.inv_max;
        }
        return;
    } else {
        float depth, h1 = 0.f, h2 = 0.f, lmax, lsum = 0.f;
        lmax = lg[0];
        for (int d = 0; d < D; ++d) { const float v = lg[d * p.plane]; lmax = fmaxf(lmax, v); lsum += v; }
        float es = 0.f;
        for (int d = 0; d < D; ++d) es += expf(lg[d * p.plane] - lmax);
        float best = -1.f, reg = 0.f;
        depth = 0.f;
        for (int d = 0; d < D; ++d) {
            const float a = expf(lg[d * p.plane] - lmax) / es;
            const float h = hy[d * p.plane];
            stg_stream(at + d * p.plane, a);
            if (a > best) { best = a; depth = h; }
            reg = fmaf(a, h, reg);
            if (d == 1) h1 = h;
            if (d == 2) h2 = h;
        }
        if (p.mode == MVSTER_DEPTH_REGRESS) depth = reg;
        p.depth[i] = depth;
        // photometric confidence on the raw logits: max / sum (reference :1109-1113,1138)
        if (p.conf != nullptr) p.conf[i] = lmax / lsum;
        if (p.inv_min != nullptr) {
            // last_depth_itv = 1/hypo[:,2] - 1/hypo[:,1]  (reference :1152)
            const float itv = 1.0f / h2 - 1.0f / h1;
            const float inv = 1.0f / depth;
            p.inv_min[i] = inv + p.split_itv * itv;
            p.inv_max[i] = inv - p.split_itv * itv;
        }
    }
}

struct TailBwdParams {
    const float* attn;
    const float* hypo;
    const float* depth;
    const float* g_attn;
    const float* g_depth;
    float* g_logits;
    int mode;
    int D;
    size_t plane;
    size_t total;
};

// d logits = attn * (g - sum_d attn*g) with g = g_attn (+ g_depth * hypo when depth = sum attn*hypo)
__global__ void __launch_bounds__(256) tail_bwd_kernel(const TailBwdParams p) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.total) return;
    const size_t b = i / p.plane, r = i - b * p.plane;
    const size_t base = b * p.D * p.plane + r;
    const float gd = (p.mode == MVSTER_DEPTH_REGRESS && p.g_depth != nullptr) ? p.g_depth[i] : 0.f;
    float dot = 0.f;
    for (int d = 0; d < p.D; ++d) {
        const size_t o = base + d * p.plane;
        float g = p.g_attn != nullptr ? p.g_attn[o] : 0.f;
        if (gd != 0.f) g = fmaf(gd, p.hypo[o], g);
        dot = fmaf(p.attn[o], g, dot);
    }
    for (int d = 0; d < p.D; ++d) {
        const size_t o = base + d * p.plane;
        float g = p.g_attn != nullptr ? p.g_attn[o] : 0.f;
        if (gd != 0.f) g = fmaf(gd, p.hypo[o], g);
        p.g_logits[o] = p.attn[o] * (g - dot);
    }
}

// ---- hypothesis schedule --------------------------------------------------------------------------------------
// stage 1 (reference :79-85): inverse depth uniform between 1/d[:, -1] (index 0 = far) and 1/d[:, 0]
__global__ void __launch_bounds__(256) init_inverse_range_kernel(const float* __restrict__ depth_values, int nvals,
                                                                 float* __restrict__ hypo, int D, size_t plane,
                                                                 size_t total /* B*D*plane */) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t bd = i / plane;
    const int d = (int)(bd % D);
    const size_t b = bd / D;
    const float inv_min = 1.0f / depth_values[b * nvals];
    const float inv_max = 1.0f / depth_values[b * nvals + nvals - 1];
    const float itv = (float)d / (float)(D - 1);
    const float inv = inv_max + (inv_min - inv_max) * itv;
    hypo[i] = 1.0f / inv;
}

// stages 2..4 (reference :87-94): per-pixel lerp in inverse depth at (H/2, W/2), upsampled x2 with
// F.interpolate(trilinear, align_corners=True) - the depth axis keeps its size, so it is a bilinear upsample of
// every hypothesis plane - then the reciprocal.  The lerp commutes with the bilinear weights only approximately in
// fp32, so the order of the reference is kept: lerp at the four low-res taps first, then interpolate.
__global__ void __launch_bounds__(256) schedule_inverse_range_kernel(const float* __restrict__ inv_min,
                                                                     const float* __restrict__ inv_max,
                                                                     float* __restrict__ hypo, int D, int H, int W,
                                                                     int Hl, int Wl, float sy, float sx,
                                                                     size_t total /* B*H*W */) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t plane = (size_t)H * W;
    const size_t b = i / plane, r = i - b * plane;
    const int y = (int)(r / W), x = (int)(r - (size_t)y * W);
    const float fy = sy * (float)y, fx = sx * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hl - 1), x1 = x0 + (x0 < Wl - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const size_t lb = b * (size_t)Hl * Wl;
    const size_t o00 = lb + (size_t)y0 * Wl + x0, o01 = lb + (size_t)y0 * Wl + x1;
    const size_t o10 = lb + (size_t)y1 * Wl + x0, o11 = lb + (size_t)y1 * Wl + x1;
    const float mx00 = inv_max[o00], mx01 = inv_max[o01], mx10 = inv_max[o10], mx11 = inv_max[o11];
    const float df00 = inv_min[o00] - mx00, df01 = inv_min[o01] - mx01;
    const float df10 = inv_min[o10] - mx10, df11 = inv_min[o11] - mx11;
    for (int d = 0; d < D; ++d) {
        const float itv = (float)d / (float)(D - 1);
        const float v00 = mx00 + df00 * itv, v01 = mx01 + df01 * itv;
        const float v10 = mx10 + df10 * itv, v11 = mx11 + df11 * itv;
        const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
        stg_stream(hypo + (b * D + d) * plane + r, 1.0f / v);
    }
}

// Same arithmetic, four consecutive pixels of a row per thread (W % 4 == 0): the eight low-resolution values a
// pixel quad needs sit in at most three columns, and every hypothesis plane is written with one 16-byte store.
// 3-D grid (quad column, row, batch) and the D interpolation fractions d / (D - 1) as kernel parameters: the first
// version derived (b, y, x) from a flat 64-bit index and divided d by D - 1 per thread and hypothesis - 805
// instructions per thread, issue-bound at 30 % of the HBM write rate (profiles/r02_step_kernels_ncu.md).
constexpr int kSchedMaxD = 32;
struct ScheduleItv {
    float v[kSchedMaxD];  // (float)d / (float)(D - 1), the reference's itv (models/mvs4net_utils.py:91)
};
__global__ void __launch_bounds__(128) schedule_inverse_range_x4_kernel(const float* __restrict__ inv_min,
                                                                        const float* __restrict__ inv_max,
                                                                        float* __restrict__ hypo,
                                                                        const __grid_constant__ ScheduleItv itv, int D,
                                                                        int H, int W, int Hl, int Wl, float sy, float sx) {
    const int W4 = W >> 2;
    const int q = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, b = blockIdx.z;
    if (q >= W4 || y >= H) return;
    const int xq = q * 4;
    const float fy = sy * (float)y;
    const int y0 = (int)fy, y1 = y0 + (y0 < Hl - 1);
    const float ly = fy - (float)y0, hy = 1.0f - ly;
    const int lb = b * Hl * Wl;  // B * Hl * Wl < 2^31 (checked on the host)
    const float* mx0 = inv_max + lb + y0 * Wl;
    const float* mx1 = inv_max + lb + y1 * Wl;
    const float* mn0 = inv_min + lb + y0 * Wl;
    const float* mn1 = inv_min + lb + y1 * Wl;
    float m00[4], m01[4], m10[4], m11[4], d00[4], d01[4], d10[4], d11[4], lx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float fx = sx * (float)(xq + k);
        const int x0 = (int)fx, x1 = x0 + (x0 < Wl - 1);
        lx[k] = fx - (float)x0;
        m00[k] = __ldg(mx0 + x0); m01[k] = __ldg(mx0 + x1); m10[k] = __ldg(mx1 + x0); m11[k] = __ldg(mx1 + x1);
        d00[k] = __ldg(mn0 + x0) - m00[k]; d01[k] = __ldg(mn0 + x1) - m01[k];
        d10[k] = __ldg(mn1 + x0) - m10[k]; d11[k] = __ldg(mn1 + x1) - m11[k];
    }
    const size_t plane = (size_t)H * W;
    float* out = hypo + (size_t)b * D * plane + (size_t)y * W + xq;
    for (int d = 0; d < D; ++d) {
        const float t = itv.v[d];
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float hx = 1.0f - lx[k];
            const float v00 = m00[k] + d00[k] * t, v01 = m01[k] + d01[k] * t;
            const float v10 = m10[k] + d10[k] * t, v11 = m11[k] + d11[k] * t;
            // __frcp_rn is the correctly rounded reciprocal: the same value as the IEEE division 1.0f / v
            o[k] = __frcp_rn(hy * (hx * v00 + lx[k] * v01) + ly * (hx * v10 + lx[k] * v11));
        }
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(out + (size_t)d * plane), "f"(o[0]), "f"(o[1]),
                     "f"(o[2]), "f"(o[3]) : "memory");
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_tail(const float* logits, const float* hypo, float split_itv, int depth_mode, float* attn,
                           float* depth, float* conf, float* inv_min, float* inv_max, int B, int D, int H, int W,
                           void* stream) {
    if (!logits || !hypo || !attn || !depth) return fail(MVSTER_ERR_BAD_ARG, "tail: null pointer");
    if ((inv_min == nullptr) != (inv_max == nullptr))
        return fail(MVSTER_ERR_BAD_ARG, "tail: inv_min and inv_max must both be given or both be NULL");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "tail: non-positive dimension");
    if (inv_min != nullptr && D < 3)
        return fail(MVSTER_ERR_BAD_ARG, "tail: inverse-depth outputs need D >= 3 (reference indexes hypo[:,2])");
    if (depth_mode != MVSTER_DEPTH_ARGMAX && depth_mode != MVSTER_DEPTH_REGRESS)
        return fail(MVSTER_ERR_BAD_ARG, "tail: unknown depth_mode %d", depth_mode);
    DeviceGuard guard(depth);
    if (guard.status != MVSTER_OK) return guard.status;
    TailParams p{logits, hypo, attn, depth, conf, inv_min, inv_max, split_itv, depth_mode, D,
                 (size_t)H * W, (size_t)B * H * W};
    const unsigned blocks = (unsigned)((p.total + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (D == 4) tail_kernel<4><<<blocks, 256, 0, s>>>(p);
    else if (D == 8) tail_kernel<8><<<blocks, 256, 0, s>>>(p);
    else tail_kernel<0><<<blocks, 256, 0, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("tail launch");
    return MVSTER_OK;
}

extern "C" int mvster_tail_bwd(const float* attn, const float* hypo, const float* depth, const float* g_attn,
                               const float* g_depth, int depth_mode, float* g_logits, int B, int D, int H, int W,
                               void* stream) {
    (void)depth;
    if (!attn || !hypo || !g_logits) return fail(MVSTER_ERR_BAD_ARG, "tail_bwd: null pointer");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "tail_bwd: non-positive dimension");
    DeviceGuard guard(g_logits);
    if (guard.status != MVSTER_OK) return guard.status;
    TailBwdParams p{attn, hypo, depth, g_attn, g_depth, g_logits, depth_mode, D, (size_t)H * W, (size_t)B * H * W};
    tail_bwd_kernel<<<(unsigned)((p.total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("tail_bwd launch");
    return MVSTER_OK;
}

extern "C" int mvster_init_inverse_range(const float* depth_values, int nvals, float* hypo, int B, int D, int H,
                                         int W, void* stream) {
    if (!depth_values || !hypo) return fail(MVSTER_ERR_BAD_ARG, "init_inverse_range: null pointer");
    if (B <= 0 || D < 2 || H <= 0 || W <= 0 || nvals < 1)
        return fail(MVSTER_ERR_BAD_ARG, "init_inverse_range: bad dimension (need D >= 2)");
    DeviceGuard guard(hypo);
    if (guard.status != MVSTER_OK) return guard.status;
    const size_t plane = (size_t)H * W, total = (size_t)B * D * plane;
    init_inverse_range_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        depth_values, nvals, hypo, D, plane, total);
    count_launch();
    MVSTER_CHECK_LAUNCH("init_inverse_range launch");
    return MVSTER_OK;
}

extern "C" int mvster_schedule_inverse_range(const float* inv_min, const float* inv_max, float* hypo, int B, int D,
                                             int H, int W, void* stream) {
    if (!inv_min || !inv_max || !hypo) return fail(MVSTER_ERR_BAD_ARG, "schedule_inverse_range: null pointer");
    if (B <= 0 || D < 2 || H < 2 || W < 2)
        return fail(MVSTER_ERR_BAD_ARG, "schedule_inverse_range: bad dimension (need D >= 2, H, W >= 2)");
    DeviceGuard guard(hypo);
    if (guard.status != MVSTER_OK) return guard.status;
    const int Hl = H / 2, Wl = W / 2;  // reference :90 (H//2, W//2)
    // align_corners=True source-index scale, computed like ATen's area_pixel_compute_scale
    const float sy = H > 1 ? (float)(Hl - 1) / (float)(H - 1) : 0.f;
    const float sx = W > 1 ? (float)(Wl - 1) / (float)(W - 1) : 0.f;
    const size_t total = (size_t)B * H * W;
    if (W % 4 == 0 && ((uintptr_t)hypo) % 16 == 0 && D <= kSchedMaxD && B <= 65535 && (H + 3) / 4 <= 65535 &&
        (double)B * Hl * Wl < 2147483648.0) {
        ScheduleItv itv{};
        for (int d = 0; d < D; ++d) itv.v[d] = (float)d / (float)(D - 1);
        dim3 grid((W / 4 + 31) / 32, (H + 3) / 4, B);
        schedule_inverse_range_x4_kernel<<<grid, dim3(32, 4), 0, (cudaStream_t)stream>>>(inv_min, inv_max, hypo, itv, D, H, W,
                                                                                       Hl, Wl, sy, sx);
    } else {
        schedule_inverse_range_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            inv_min, inv_max, hypo, D, H, W, Hl, Wl, sy, sx, total);
    }
    count_launch();
    MVSTER_CHECK_LAUNCH("schedule_inverse_range launch");
    return MVSTER_OK;
}
