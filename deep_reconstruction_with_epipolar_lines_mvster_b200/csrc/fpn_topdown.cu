// FPN4 top-down step fused with its output convolution ("FPN4 emitting NHWC features directly").
//
// Reference (models/mvs4net_utils.py:488-495), per pyramid level:
//     intra = F.interpolate(intra, scale_factor=2, mode="bilinear", align_corners=True) + inner(lateral)   # 1x1, bias
//     out   = out_conv(intra)                                                                              # 3x3, no bias
// At the finest level `intra` is a 64-channel full-resolution tensor (1.2 GB for 5 views of 832x1152) that cuDNN /
// ATen write and re-read four times (upsample, 1x1 conv, add, 3x3 conv: 12 ms of a 30 ms forward on B200).  Here a
// CTA builds the 64-channel `intra` tile (32x8 pixels + 1-pixel halo) in shared memory - bilinear taps from the
// coarser level, the 1x1 lateral convolution from the 8/16-channel encoder map - and immediately applies the 3x3
// output convolution from shared memory, writing the K1-ready NHWC feature map.  `intra` reaches HBM only where the
// next (finer) level needs it (intra_out != NULL).  All weights (3x3x64xCO + CLx64 + 64 floats, <= 23 KB) travel as
// kernel parameters and are consumed through uniform constant loads.
//
// The output convolution of a level with 16 output channels (36.8 KB of weights) is done in two 8-channel slices; the
// second launch reads the `intra` tile that the first one stored (intra_in != NULL) instead of recomputing it.
#include <string.h>

#include "common.cuh"

namespace mvster {

constexpr int kTdTW = 32, kTdTH = 8;                 // output tile
constexpr int kTdHW = kTdTW + 2, kTdHH = kTdTH + 2;  // tile + halo
constexpr int kTdRS = 36;                            // shared-memory row stride in floats (even: 8-byte aligned pairs)
constexpr int kTdSmem = (64 * kTdHH * kTdRS + 16 * 64 + 64) * 4;  // tile (92160 bytes) + lateral weights and bias

template <int CL, int CO>
struct TopDownParams {
    float w_out[9 * 64 * CO];  // [ky][kx][c64][co]
    float w_in[CL * 64];       // [cl][c64]
    float b_in[64];
    const float* prev;      // [B,64,H/2,W/2] planar: intra of the coarser level
    const float* lat;       // [B,CL,H,W]     planar: encoder map of this level
    const float* intra_in;  // nullable [B,64,H,W]: load the tile instead of computing it
    float* intra_out;       // nullable [B,64,H,W]
    void* feat;             // NHWC [B,H,W,co_total], fp32 or bf16 (the kernel's OutT)
    int B, H, W, co_total, co_off;
    float sy, sx;           // align_corners=True source scale (Hl-1)/(H-1), (Wl-1)/(W-1)
};

template <int CL, int CO, int HALF>
__device__ __forceinline__ void topdown_conv_half(const TopDownParams<CL, CO>& p, const float* tb0, float (&acc)[2][CO]) {
    constexpr int CS = kTdHH * kTdRS;
    const float* tb = tb0 + (HALF * 32) * CS;
#pragma unroll 2
    for (int ci = 0; ci < 32; ++ci) {
        const float* tc = tb + ci * CS;
        float in[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float2 a = *reinterpret_cast<const float2*>(tc + r * kTdRS);
            const float2 c = *reinterpret_cast<const float2*>(tc + r * kTdRS + 2);
            in[r][0] = a.x; in[r][1] = a.y; in[r][2] = c.x; in[r][3] = c.y;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int co = 0; co < CO; ++co) {
                    const float wv = p.w_out[((ky * 3 + kx) * 64 + HALF * 32 + ci) * CO + co];
                    acc[0][co] = fmaf(wv, in[ky][kx], acc[0][co]);
                    acc[1][co] = fmaf(wv, in[ky][kx + 1], acc[1][co]);
                }
    }
}

constexpr int kTdThreads = 256;

__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x = lo (low half), .y = hi
    return *reinterpret_cast<const unsigned*>(&v);
}

template <int CL, int CO, typename OutT>
__global__ void __launch_bounds__(kTdThreads, 2) fpn_topdown_kernel(const __grid_constant__ TopDownParams<CL, CO> p) {
    extern __shared__ float tile[];  // [64][kTdHH][kTdRS]
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int tx0 = blockIdx.x * kTdTW, ty0 = blockIdx.y * kTdTH;
    const int H = p.H, W = p.W, Hl = H / 2, Wl = W / 2;
    const size_t plane = (size_t)H * W, lplane = (size_t)Hl * Wl;
    constexpr int CS = kTdHH * kTdRS;  // channel stride of the tile

    // ---- phase 1: the 64-channel intra tile (+ halo) -> shared memory ------------------------------------------------
#if !defined(MVSTER_TD_SKIP) || MVSTER_TD_SKIP != 1
    // work item = (32 consecutive halo pixels, 8 channels): 11 pixel chunks x 8 channel groups = 88 warp-items, exactly
    // 11 per warp (a pixel-major split leaves two thirds of the CTA idle at the barrier in its second round).  The
    // channel group is warp-uniform but not CTA-uniform, so the lateral 1x1 weights are read as shared-memory
    // broadcasts instead of uniform constant loads.
    float* w_s = tile + 64 * CS;         // [CL][64] lateral weights, then [64] bias
    for (int i = tid; i < CL * 64 + 64; i += kTdThreads) w_s[i] = i < CL * 64 ? p.w_in[i] : p.b_in[i - CL * 64];
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NCHUNK = (kTdHH * kTdHW + 31) / 32;  // 11
#pragma unroll 1
    for (int wi = warp; wi < 8 * NCHUNK; wi += kTdThreads / 32) {
        const int cg = wi / NCHUNK, hp = (wi - cg * NCHUNK) * 32 + lane;
        if (hp >= kTdHH * kTdHW) continue;
        const int ry = hp / kTdHW, rx = hp - ry * kTdHW;
        const int gy = ty0 - 1 + ry, gx = tx0 - 1 + rx;
        float* ts = tile + (cg * 8) * CS + ry * kTdRS + rx;
        const bool inside = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
        if (!inside) {  // zero padding of the 3x3 output convolution
#pragma unroll
            for (int c = 0; c < 8; ++c) ts[c * CS] = 0.0f;
            continue;
        }
        const size_t go = (size_t)gy * W + gx;
        if (p.intra_in != nullptr) {
            const float* ip = p.intra_in + ((size_t)b * 64 + cg * 8) * plane + go;
#pragma unroll
            for (int c = 0; c < 8; ++c) ts[c * CS] = __ldg(ip + (size_t)c * plane);
            continue;
        }
        // bilinear x2, align_corners=True, with ATen's arithmetic (upsample_bilinear2d: source = scale * dst)
        const float fy = p.sy * (float)gy, fx = p.sx * (float)gx;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < Hl - 1), x1 = x0 + (x0 < Wl - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float hy = 1.0f - ly, hx = 1.0f - lx;
        const float* pp = p.prev + ((size_t)b * 64 + cg * 8) * lplane;
        const size_t o00 = (size_t)y0 * Wl + x0, o01 = (size_t)y0 * Wl + x1;
        const size_t o10 = (size_t)y1 * Wl + x0, o11 = (size_t)y1 * Wl + x1;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {   // 32 independent taps in flight
            const float* q = pp + (size_t)c * lplane;
            v[c] = hy * (hx * __ldg(q + o00) + lx * __ldg(q + o01)) + ly * (hx * __ldg(q + o10) + lx * __ldg(q + o11));
        }
        float t[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) t[c] = 0.0f;
        const float* lp = p.lat + (size_t)b * CL * plane + go;
#pragma unroll
        for (int k = 0; k < CL; ++k) {
            const float lk = __ldg(lp + (size_t)k * plane);
            const float4 wa = *reinterpret_cast<const float4*>(w_s + k * 64 + cg * 8);
            const float4 wb = *reinterpret_cast<const float4*>(w_s + k * 64 + cg * 8 + 4);
            t[0] = fmaf(wa.x, lk, t[0]); t[1] = fmaf(wa.y, lk, t[1]); t[2] = fmaf(wa.z, lk, t[2]); t[3] = fmaf(wa.w, lk, t[3]);
            t[4] = fmaf(wb.x, lk, t[4]); t[5] = fmaf(wb.y, lk, t[5]); t[6] = fmaf(wb.z, lk, t[6]); t[7] = fmaf(wb.w, lk, t[7]);
        }
        const bool interior = p.intra_out != nullptr && ry >= 1 && ry <= kTdTH && rx >= 1 && rx <= kTdTW;
        float* iop = interior ? p.intra_out + ((size_t)b * 64 + cg * 8) * plane + go : nullptr;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float val = v[c] + (t[c] + w_s[CL * 64 + cg * 8 + c]);
            ts[c * CS] = val;
            if (iop != nullptr) iop[(size_t)c * plane] = val;
        }
    }
#endif
    __syncthreads();
#if defined(MVSTER_TD_SKIP) && MVSTER_TD_SKIP == 2
    if (tile[threadIdx.x] != 12345.0f) return;
#endif

    // ---- phase 2: 3x3 output convolution from shared memory; a thread owns 1x2 pixels x CO channels for HALF of the
    // 64 input channels (threads 0-127: channels 0-31, threads 128-255: channels 32-63), halves summed through smem ----
    const int half = tid >> 7, t7 = tid & 127;
    const int tx = t7 & 15, ty = t7 >> 4;
    float acc[2][CO];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[c][co] = 0.0f;
    // the branch is warp-uniform; inside each arm the weight offsets are compile-time constants (uniform loads)
    if (half == 0) topdown_conv_half<CL, CO, 0>(p, tile + ty * kTdRS + 2 * tx, acc);
    else topdown_conv_half<CL, CO, 1>(p, tile + ty * kTdRS + 2 * tx, acc);
    __syncthreads();  // everyone is done reading the tile: reuse it for the cross-half reduction
    float* red = tile + t7;  // [2*CO][128]
    if (half == 1) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int co = 0; co < CO; ++co) red[(c * CO + co) * 128] = acc[c][co];
    }
    __syncthreads();
    if (half == 1) return;
    const int gy = ty0 + ty, gx = tx0 + 2 * tx;
    if (gy >= H) return;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (gx + c >= W) break;
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[c][co] += red[(c * CO + co) * 128];
        const size_t fo = ((size_t)b * plane + (size_t)gy * W + gx + c) * p.co_total + p.co_off;
        if constexpr (sizeof(OutT) == 4) {
            float* fp = static_cast<float*>(p.feat) + fo;
#pragma unroll
            for (int q = 0; q < CO; q += 4)
                *reinterpret_cast<float4*>(fp + q) = make_float4(acc[c][q], acc[c][q + 1], acc[c][q + 2], acc[c][q + 3]);
        } else {  // bf16 (round to nearest even, as torch's .to(bfloat16)): 8 channels = one 16-byte store
            __nv_bfloat16* fp = static_cast<__nv_bfloat16*>(p.feat) + fo;
#pragma unroll
            for (int q = 0; q < CO; q += 8) {
                uint4 v;
                v.x = pack_bf16x2(acc[c][q], acc[c][q + 1]);
                v.y = pack_bf16x2(acc[c][q + 2], acc[c][q + 3]);
                v.z = pack_bf16x2(acc[c][q + 4], acc[c][q + 5]);
                v.w = pack_bf16x2(acc[c][q + 6], acc[c][q + 7]);
                *reinterpret_cast<uint4*>(fp + q) = v;
            }
        }
    }
}

template <int CL, int CO, typename OutT>
static int launch_topdown(const float* prev, const float* lat, const float* intra_in, float* intra_out, void* feat,
                          const float* w_out, const float* w_in, const float* b_in, int B, int H, int W, int co_total,
                          int co_off, cudaStream_t s) {
    static thread_local TopDownParams<CL, CO> p;
    static_assert(sizeof(TopDownParams<CL, CO>) <= 32000, "weights must fit the kernel-parameter space");
    memcpy(p.w_out, w_out, sizeof(p.w_out));
    memcpy(p.w_in, w_in, sizeof(p.w_in));
    memcpy(p.b_in, b_in, sizeof(p.b_in));
    p.prev = prev; p.lat = lat; p.intra_in = intra_in; p.intra_out = intra_out; p.feat = feat;
    p.B = B; p.H = H; p.W = W; p.co_total = co_total; p.co_off = co_off;
    const int Hl = H / 2, Wl = W / 2;
    p.sy = H > 1 ? (float)(Hl - 1) / (float)(H - 1) : 0.f;  // ATen area_pixel_compute_scale, align_corners=True
    p.sx = W > 1 ? (float)(Wl - 1) / (float)(W - 1) : 0.f;
    static int attr_done[64] = {};  // largest size set per device
    const int st = ensure_dynamic_smem_bytes(fpn_topdown_kernel<CL, CO, OutT>, kTdSmem, attr_done, "fpn_topdown: cudaFuncSetAttribute");
    if (st != MVSTER_OK) return st;
    dim3 grid((W + kTdTW - 1) / kTdTW, (H + kTdTH - 1) / kTdTH, B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "fpn_topdown: grid too large");
    fpn_topdown_kernel<CL, CO, OutT><<<grid, kTdThreads, kTdSmem, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("fpn_topdown launch");
    return MVSTER_OK;
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_fpn_topdown_ex(const float* prev, const float* lat, const float* intra_in, float* intra_out,
                                     void* feat, int feat_dtype, const float* w_out_host, const float* w_in_host,
                                     const float* b_in_host, int B, int Clat, int Cout, int Cout_total, int co_off,
                                     int H, int W, void* stream) {
    if (!feat || !w_out_host || !w_in_host || !b_in_host) return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown: null pointer");
    if (feat_dtype != MVSTER_F32 && feat_dtype != MVSTER_BF16) return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown: feat_dtype must be MVSTER_F32 or MVSTER_BF16");
    if (!intra_in && (!prev || !lat))
        return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown: need prev + lat (compute the tile) or intra_in (reload it)");
    if (B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown: H, W must be positive and even");
    const int cal = feat_dtype == MVSTER_BF16 ? 8 : 4;  // 16-byte stores
    if (co_off < 0 || co_off + Cout > Cout_total || (co_off % cal) || (Cout_total % cal))
        return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown: bad output channel slice");
    if (((uintptr_t)feat) % 16) return fail(MVSTER_ERR_ALIGN, "fpn_topdown: feat must be 16-byte aligned");
    DeviceGuard guard(feat);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    const bool bf = feat_dtype == MVSTER_BF16;
#define MVSTER_TD_ARGS prev, lat, intra_in, intra_out, feat, w_out_host, w_in_host, b_in_host, B, H, W, Cout_total, co_off, s
    if (Clat == 8 && Cout == 8)
        return bf ? launch_topdown<8, 8, __nv_bfloat16>(MVSTER_TD_ARGS) : launch_topdown<8, 8, float>(MVSTER_TD_ARGS);
    if (Clat == 16 && Cout == 8)
        return bf ? launch_topdown<16, 8, __nv_bfloat16>(MVSTER_TD_ARGS) : launch_topdown<16, 8, float>(MVSTER_TD_ARGS);
#undef MVSTER_TD_ARGS
    return fail(MVSTER_ERR_UNSUPPORTED, "fpn_topdown: no kernel for Clat=%d Cout=%d (built: (8,8), (16,8))", Clat, Cout);
}

extern "C" int mvster_fpn_topdown(const float* prev, const float* lat, const float* intra_in, float* intra_out,
                                  float* feat, const float* w_out_host, const float* w_in_host, const float* b_in_host,
                                  int B, int Clat, int Cout, int Cout_total, int co_off, int H, int W, void* stream) {
    return mvster_fpn_topdown_ex(prev, lat, intra_in, intra_out, feat, MVSTER_F32, w_out_host, w_in_host, b_in_host, B,
                                 Clat, Cout, Cout_total, co_off, H, W, stream);
}
