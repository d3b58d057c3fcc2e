// K2a': the last layers of the cost regulariser fused with the stagenet tail.
//
// Reference (eval mode): reg2d.forward, models/mvs4net_utils.py:923-926
//     x      = conv0 + relu(bn(conv11(low)))        conv11 = ConvTranspose3d(16, 8, (1,3,3), stride (1,2,2),
//                                                             padding (0,1,1), output_padding (0,1,1), bias=False)
//     logits = prob(x).squeeze(1)                   prob   = Conv3d(8, 1, 1) with bias
// followed by the tail of stagenet.forward, :1109-1156 (softmax over D, arg-max depth, confidence, inverse range).
// Unfused, the full-resolution [B,8,D,H,W] tensor is written by conv11, re-read and re-written by BN/ReLU and by the
// skip add, read by prob, and the [B,D,H,W] logits make one more round trip into the tail kernel - about 3 GB of HBM
// traffic per launch at the DTU stage-4 shape (B=8).  Here every thread owns one 2x2 block of output pixels for ALL D
// hypotheses: it reads the 2x2 low-resolution neighbourhood the transposed convolution needs (16 channels), the skip
// tensor, the hypotheses, and writes only the tail outputs.
//
// A stride-2 transposed 3x3 convolution touches 1 / 2 / 2 / 4 taps for the four pixel parities of a 2x2 block:
//     out[2j  ,2i  ] = W11 in[j,i]
//     out[2j  ,2i+1] = W10 in[j,i+1] + W12 in[j,i]
//     out[2j+1,2i  ] = W01 in[j+1,i] + W21 in[j,i]
//     out[2j+1,2i+1] = W00 in[j+1,i+1] + W02 in[j+1,i] + W20 in[j,i+1] + W22 in[j,i]        (Wkykx: [16 x 8])
// i.e. 9 x 16 x 8 = 1152 FMAs per block and hypothesis.  The BN-folded weights travel as a __grid_constant__ kernel
// parameter (4.7 KB): they are fetched with uniform constant loads (LDCU) into uniform registers that the FFMAs take
// as operands - no per-thread load instruction, no shared memory, no vector register for them.  The kernel is FP32-FMA bound (9.2 GFMA per launch at stage 4), not HBM bound.
#include "tail_common.cuh"

namespace mvster {

struct RegTailParams {
    float w[3][3][16][8];  // [ky][kx][ci][co], BatchNorm scale folded in
    float shift[8];        // BatchNorm shift (beta - mean * scale)
    float pw[8];           // prob weights
    float pb;              // prob bias
    float split_itv;
    const float* low;      // [B,16,D,Hh,Wh]
    const float* skip;     // [B, 8,D,H ,W ]   (conv0 output)
    const float* hypo;     // [B,D,H,W]
    float* attn;           // [B,D,H,W]
    float* depth;          // [B,H,W]
    float* conf;
    float* inv_min;
    float* inv_max;
    int mode, B, Hh, Wh;
};

__device__ __forceinline__ float2 ldg_stream2(const float* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream2(float* p, float a, float b) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

#ifndef MVSTER_REGTAIL_UNROLL
#define MVSTER_REGTAIL_UNROLL 2
#endif
#ifndef MVSTER_REGTAIL_MINB
#define MVSTER_REGTAIL_MINB 4
#endif
constexpr int kCiUnroll = MVSTER_REGTAIL_UNROLL;

template <int D>
__global__ void __launch_bounds__(128, MVSTER_REGTAIL_MINB) regtail_kernel(const __grid_constant__ RegTailParams p) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);
    const int j = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (i >= p.Wh || j >= p.Hh) return;
    const int W = 2 * p.Wh;
    const size_t lplane = (size_t)p.Hh * p.Wh, plane = 4 * lplane;
    const bool has_r = (i + 1 < p.Wh), has_d = (j + 1 < p.Hh);
    const size_t l00 = (size_t)j * p.Wh + i;
    const size_t l01 = has_r ? l00 + 1 : l00, l10 = has_d ? l00 + p.Wh : l00;
    const size_t l11 = has_d ? l01 + p.Wh : l01;
    const size_t o0 = (size_t)(2 * j) * W + 2 * i;  // top-left pixel of the block; the row below is o0 + W

    // The hypothesis loop stays rolled (one copy of the 1152 constant-operand FFMAs, bounded register pressure); the
    // four logits of each hypothesis are parked in a thread-private shared-memory column until the tail needs all D.
    __shared__ float lg_s[D * 4 * 128];
#pragma unroll 1
    for (int d = 0; d < D; ++d) {
        float a[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int co = 0; co < 8; ++co) a[k][co] = 0.0f;
        const float* lp = p.low + ((size_t)b * 16 * D + d) * lplane;
        // partially unrolled on purpose: fully unrolled, ptxas hoists all 64 loads of a hypothesis above the FFMAs and
        // needs 254 registers; the weights are fetched with uniform loads (LDCU) either way
#pragma unroll kCiUnroll
        for (int ci = 0; ci < 16; ++ci) {
            const float* q = lp + (size_t)ci * D * lplane;
            const float v00 = __ldg(q + l00);
            const float v01 = has_r ? __ldg(q + l01) : 0.0f;  // zero padding right of / below the low-res map
            const float v10 = has_d ? __ldg(q + l10) : 0.0f;
            const float v11 = (has_r && has_d) ? __ldg(q + l11) : 0.0f;
#pragma unroll
            for (int co = 0; co < 8; ++co) {
                a[0][co] = fmaf(p.w[1][1][ci][co], v00, a[0][co]);
                a[1][co] = fmaf(p.w[1][0][ci][co], v01, fmaf(p.w[1][2][ci][co], v00, a[1][co]));
                a[2][co] = fmaf(p.w[0][1][ci][co], v10, fmaf(p.w[2][1][ci][co], v00, a[2][co]));
                a[3][co] = fmaf(p.w[0][0][ci][co], v11,
                                fmaf(p.w[0][2][ci][co], v10, fmaf(p.w[2][0][ci][co], v01, fmaf(p.w[2][2][ci][co], v00, a[3][co]))));
            }
        }
        // x = skip + relu(deconv + shift); logit = prob_w . x + prob_b
        float l0 = p.pb, l1 = p.pb, l2 = p.pb, l3 = p.pb;
        const float* sp = p.skip + ((size_t)b * 8 * D + d) * plane + o0;
#pragma unroll
        for (int co = 0; co < 8; ++co) {
            const float2 s0 = ldg_stream2(sp + (size_t)co * D * plane);
            const float2 s1 = ldg_stream2(sp + (size_t)co * D * plane + W);
            l0 = fmaf(p.pw[co], s0.x + fmaxf(a[0][co] + p.shift[co], 0.0f), l0);
            l1 = fmaf(p.pw[co], s0.y + fmaxf(a[1][co] + p.shift[co], 0.0f), l1);
            l2 = fmaf(p.pw[co], s1.x + fmaxf(a[2][co] + p.shift[co], 0.0f), l2);
            l3 = fmaf(p.pw[co], s1.y + fmaxf(a[3][co] + p.shift[co], 0.0f), l3);
        }
        float* ls = lg_s + d * 4 * 128 + threadIdx.x;
        ls[0] = l0; ls[128] = l1; ls[256] = l2; ls[384] = l3;
    }
    // ---- tail, one output row (two pixels) at a time: same arithmetic as tail_kernel ---------------------------
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        const size_t orow = o0 + (size_t)r * W;
        const float* hp = p.hypo + (size_t)b * D * plane + orow;
        float l0[D], l1[D], h0[D], h1[D], a0[D], a1[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float2 t = ldg_stream2(hp + (size_t)d * plane);
            h0[d] = t.x; h1[d] = t.y;
            l0[d] = lg_s[(d * 4 + 2 * r) * 128 + threadIdx.x];
            l1[d] = lg_s[(d * 4 + 2 * r + 1) * 128 + threadIdx.x];
        }
        const TailOut r0 = tail_pixel<D>(l0, h0, p.mode, p.split_itv, a0);
        const TailOut r1 = tail_pixel<D>(l1, h1, p.mode, p.split_itv, a1);
        float* ap = p.attn + (size_t)b * D * plane + orow;
#pragma unroll
        for (int d = 0; d < D; ++d) stg_stream2(ap + (size_t)d * plane, a0[d], a1[d]);
        const size_t ob = (size_t)b * plane + orow;
        stg_stream2(p.depth + ob, r0.depth, r1.depth);
        if (p.conf != nullptr) stg_stream2(p.conf + ob, r0.conf, r1.conf);
        if (p.inv_min != nullptr) {
            stg_stream2(p.inv_min + ob, r0.inv_min, r1.inv_min);
            stg_stream2(p.inv_max + ob, r0.inv_max, r1.inv_max);
        }
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_regtail(const float* low, const float* skip, const float* w_host, const float* params_host,
                              const float* hypo, float split_itv, int depth_mode, float* attn, float* depth,
                              float* conf, float* inv_min, float* inv_max, int B, int D, int H, int W, void* stream) {
    if (!low || !skip || !w_host || !params_host || !hypo || !attn || !depth)
        return fail(MVSTER_ERR_BAD_ARG, "regtail: null pointer");
    if ((inv_min == nullptr) != (inv_max == nullptr))
        return fail(MVSTER_ERR_BAD_ARG, "regtail: inv_min and inv_max must both be given or both be NULL");
    if (B <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "regtail: non-positive dimension");
    if ((H & 1) || (W & 1)) return fail(MVSTER_ERR_BAD_ARG, "regtail: H and W must be even (stride-2 transposed conv)");
    if (D != 4 && D != 8) return fail(MVSTER_ERR_UNSUPPORTED, "regtail: D=%d not in {4,8}", D);
    if (depth_mode != MVSTER_DEPTH_ARGMAX && depth_mode != MVSTER_DEPTH_REGRESS)
        return fail(MVSTER_ERR_BAD_ARG, "regtail: unknown depth_mode %d", depth_mode);
    if (B > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "regtail: B too large");
    const uintptr_t al = (uintptr_t)skip | (uintptr_t)hypo | (uintptr_t)attn | (uintptr_t)depth | (uintptr_t)conf |
                         (uintptr_t)inv_min | (uintptr_t)inv_max;
    if (al % 8) return fail(MVSTER_ERR_ALIGN, "regtail: full-resolution tensors must be 8-byte aligned");
    DeviceGuard guard(depth);
    if (guard.status != MVSTER_OK) return guard.status;
    static thread_local RegTailParams p;
    memcpy(p.w, w_host, sizeof(p.w));
    memcpy(p.shift, params_host, 8 * sizeof(float));
    memcpy(p.pw, params_host + 8, 8 * sizeof(float));
    p.pb = params_host[16];
    p.split_itv = split_itv;
    p.low = low; p.skip = skip; p.hypo = hypo; p.attn = attn; p.depth = depth; p.conf = conf;
    p.inv_min = inv_min; p.inv_max = inv_max;
    p.mode = depth_mode; p.B = B; p.Hh = H / 2; p.Wh = W / 2;
    dim3 grid((p.Wh + 31) / 32, (p.Hh + 3) / 4, B);
    if (grid.y > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "regtail: grid too large");
    cudaStream_t s = (cudaStream_t)stream;
    if (D == 4) regtail_kernel<4><<<grid, 128, 0, s>>>(p);
    else regtail_kernel<8><<<grid, 128, 0, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("regtail launch");
    return MVSTER_OK;
}
