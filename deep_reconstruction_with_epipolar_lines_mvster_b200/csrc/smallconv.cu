// Direct fp32 convolutions for the high-resolution, few-channel layers of the cost regulariser reg2d
// (models/mvs4net_utils.py:889-912, eval mode, BatchNorm folded into the weights, ReLU fused) - SURVEY.md §8f rank 1.
//
// cuDNN's fp32 kernels are built for wide layers; at 4..32 channels and millions of pixels they run at a few percent
// of the machine (measured on B200: conv0 + BatchNorm + ReLU of stage 4 take ~5 ms per scene for 1.1 GFMA).  These
// layers have so few weights (288 .. 6912 floats) that the whole filter bank fits in the kernel-parameter constant
// bank: a thread keeps a small tile of output pixels x ALL output channels in registers, walks the input channels,
// and every FFMA takes its weight from a uniform register filled by a uniform constant load (no per-thread weight
// loads, no shared memory).  Activations are NCDHW fp32 as the reference's; x is the fastest axis, so a warp reads
// and writes whole 128/256-byte row segments per channel plane.
//
//   mode 0  Conv3d(k = (KD,3,3), stride 1, padding (KD/2,1,1))            thread tile 2x2 output pixels
//   mode 1  Conv3d(k = (1,3,3), stride (1,2,2), padding (0,1,1))          thread tile 1x2 output pixels
//   mode 2  ConvTranspose3d(k = (1,3,3), stride (1,2,2), padding (0,1,1), output_padding (0,1,1)), optional skip
//           added AFTER the ReLU (reg2d: x = conv2 + conv9(x))            thread tile 2x2 outputs of one input pixel
#include <string.h>

#include "common.cuh"

namespace mvster {

#ifndef MVSTER_MID_UNROLL
#define MVSTER_MID_UNROLL 2  // input channels per loop body of midconv_s1_kernel
#endif
constexpr int kMidUnroll = MVSTER_MID_UNROLL;
#ifndef MVSTER_MID_MINB
#define MVSTER_MID_MINB 3  // CTAs per SM of the midconv kernels
#endif
constexpr int kScUnroll = 2;  // input channels per loop body (bounded registers; weights come through LDCU)

template <int KD, int CIN, int COUT>
struct SmallConvParams {
    float w[KD * 9 * CIN * COUT];  // [kd][ky][kx][ci][co]
    float bias[COUT];
    const float* x;
    const float* skip;
    float* y;
    int B, D, H, W;  // INPUT height / width
    int relu;
    int co_total, co_off;  // the COUT channels computed here are channels co_off .. co_off+COUT-1 of a co_total-channel y
};

__device__ __forceinline__ float ldz(const float* p, bool ok) { return ok ? __ldg(p) : 0.0f; }
__device__ __forceinline__ float2 ldz2(const float* p, bool ok) {
    return ok ? __ldg(reinterpret_cast<const float2*>(p)) : make_float2(0.f, 0.f);
}
__device__ __forceinline__ float4 ldz4(const float* p, bool ok) {
    return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- mode 0: stride 1 ------------------------------------------------------------------------------------------
template <int KD, int CIN, int COUT>
__global__ void __launch_bounds__(128, 3) smallconv_s1_kernel(const __grid_constant__ SmallConvParams<KD, CIN, COUT> p) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), j = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.z / p.D, d = blockIdx.z % p.D;
    const int H = p.H, W = p.W;
    if (2 * i >= W || 2 * j >= H) return;
    const int x0 = 2 * i, y0 = 2 * j;
    const size_t plane = (size_t)H * W;
    float acc[2][2][COUT];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int co = 0; co < COUT; ++co) acc[a][c][co] = 0.0f;
    const bool vl = x0 > 0, vr = x0 + 2 < W;
    bool vy[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) vy[r] = (unsigned)(y0 - 1 + r) < (unsigned)H;
#pragma unroll 1
    for (int kd = 0; kd < KD; ++kd) {
        const int dz = d + kd - KD / 2;
        if ((unsigned)dz >= (unsigned)p.D) continue;
        const float* xp = p.x + (((size_t)b * CIN) * p.D + dz) * plane + (size_t)(y0 - 1) * W + x0;
        const float* wk = p.w + (size_t)kd * 9 * CIN * COUT;
#pragma unroll kScUnroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float* q = xp + (size_t)ci * p.D * plane;
            float in[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float* qr = q + (size_t)r * W;
                const float2 m = ldz2(qr, vy[r]);
                in[r][0] = ldz(qr - 1, vy[r] && vl);
                in[r][1] = m.x;
                in[r][2] = m.y;
                in[r][3] = ldz(qr + 2, vy[r] && vr);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int co = 0; co < COUT; ++co) {
                        const float wv = wk[((ky * 3 + kx) * CIN + ci) * COUT + co];
                        acc[0][0][co] = fmaf(wv, in[ky][kx], acc[0][0][co]);
                        acc[0][1][co] = fmaf(wv, in[ky][kx + 1], acc[0][1][co]);
                        acc[1][0][co] = fmaf(wv, in[ky + 1][kx], acc[1][0][co]);
                        acc[1][1][co] = fmaf(wv, in[ky + 1][kx + 1], acc[1][1][co]);
                    }
        }
    }
    float* yp = p.y + (((size_t)b * p.co_total + p.co_off) * p.D + d) * plane + (size_t)y0 * W + x0;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        float v[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                v[a][c] = acc[a][c][co] + p.bias[co];
                if (p.relu) v[a][c] = fmaxf(v[a][c], 0.0f);
            }
        float* yc = yp + (size_t)co * p.D * plane;
        *reinterpret_cast<float2*>(yc) = make_float2(v[0][0], v[0][1]);
        *reinterpret_cast<float2*>(yc + W) = make_float2(v[1][0], v[1][1]);
    }
}

// ---- mode 1: stride 2 (KD = 1) ---------------------------------------------------------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(128, 3) smallconv_s2_kernel(const __grid_constant__ SmallConvParams<1, CIN, COUT> p) {
    const int Ho = p.H / 2, Wo = p.W / 2;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.z / p.D, d = blockIdx.z % p.D;
    if (2 * i >= Wo || oy >= Ho) return;
    const int H = p.H, W = p.W;
    const int xi = 4 * i;  // first input column of the aligned float4; outputs 2i, 2i+1 read columns xi-1 .. xi+3
    const size_t plane = (size_t)H * W, oplane = (size_t)Ho * Wo;
    float acc[2][COUT];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[c][co] = 0.0f;
    bool vy[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) vy[r] = (unsigned)(2 * oy - 1 + r) < (unsigned)H;
    const bool vl = xi > 0;
    const float* xp = p.x + (((size_t)b * CIN) * p.D + d) * plane + (size_t)(2 * oy - 1) * W + xi;
#pragma unroll kScUnroll
    for (int ci = 0; ci < CIN; ++ci) {
        const float* q = xp + (size_t)ci * p.D * plane;
        float in[3][5];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float* qr = q + (size_t)r * W;
            const float4 m = ldz4(qr, vy[r]);
            in[r][0] = ldz(qr - 1, vy[r] && vl);
            in[r][1] = m.x; in[r][2] = m.y; in[r][3] = m.z; in[r][4] = m.w;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    const float wv = p.w[((ky * 3 + kx) * CIN + ci) * COUT + co];
                    acc[0][co] = fmaf(wv, in[ky][kx], acc[0][co]);
                    acc[1][co] = fmaf(wv, in[ky][kx + 2], acc[1][co]);
                }
    }
    float* yp = p.y + (((size_t)b * p.co_total + p.co_off) * p.D + d) * oplane + (size_t)oy * Wo + 2 * i;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        float v0 = acc[0][co] + p.bias[co], v1 = acc[1][co] + p.bias[co];
        if (p.relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
        *reinterpret_cast<float2*>(yp + (size_t)co * p.D * oplane) = make_float2(v0, v1);
    }
}

// ---- mode 2: transposed stride 2 (KD = 1), skip added after the ReLU ----------------------------------------------
// out[2j,2i] = W11 in[j,i];  out[2j,2i+1] = W10 in[j,i+1] + W12 in[j,i];  out[2j+1,2i] = W01 in[j+1,i] + W21 in[j,i];
// out[2j+1,2i+1] = W00 in[j+1,i+1] + W02 in[j+1,i] + W20 in[j,i+1] + W22 in[j,i]
template <int CIN, int COUT>
__global__ void __launch_bounds__(128, 3) smallconv_t2_kernel(const __grid_constant__ SmallConvParams<1, CIN, COUT> p) {
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), j = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.z / p.D, d = blockIdx.z % p.D;
    if (i >= p.W || j >= p.H) return;
    const int Wo = 2 * p.W;
    const size_t plane = (size_t)p.H * p.W, oplane = 4 * plane;
    const bool has_r = i + 1 < p.W, has_d = j + 1 < p.H;
    float acc[4][COUT];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[k][co] = 0.0f;
    const float* xp = p.x + (((size_t)b * CIN) * p.D + d) * plane + (size_t)j * p.W + i;
#pragma unroll kScUnroll
    for (int ci = 0; ci < CIN; ++ci) {
        const float* q = xp + (size_t)ci * p.D * plane;
        const float v00 = __ldg(q);
        const float v01 = ldz(q + 1, has_r);
        const float v10 = ldz(q + p.W, has_d);
        const float v11 = ldz(q + p.W + 1, has_r && has_d);
#define SCW(ky, kx) p.w[(((ky) * 3 + (kx)) * CIN + ci) * COUT + co]
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            acc[0][co] = fmaf(SCW(1, 1), v00, acc[0][co]);
            acc[1][co] = fmaf(SCW(1, 0), v01, fmaf(SCW(1, 2), v00, acc[1][co]));
            acc[2][co] = fmaf(SCW(0, 1), v10, fmaf(SCW(2, 1), v00, acc[2][co]));
            acc[3][co] = fmaf(SCW(0, 0), v11, fmaf(SCW(0, 2), v10, fmaf(SCW(2, 0), v01, fmaf(SCW(2, 2), v00, acc[3][co]))));
        }
#undef SCW
    }
    const size_t o = (((size_t)b * p.co_total + p.co_off) * p.D + d) * oplane + (size_t)(2 * j) * Wo + 2 * i;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = acc[k][co] + p.bias[co];
            if (p.relu) v[k] = fmaxf(v[k], 0.0f);
        }
        const size_t oc = o + (size_t)co * p.D * oplane;
        if (p.skip != nullptr) {
            const float2 s0 = __ldg(reinterpret_cast<const float2*>(p.skip + oc));
            const float2 s1 = __ldg(reinterpret_cast<const float2*>(p.skip + oc + Wo));
            v[0] += s0.x; v[1] += s0.y; v[2] += s1.x; v[3] += s1.y;
        }
        *reinterpret_cast<float2*>(p.y + oc) = make_float2(v[0], v[1]);
        *reinterpret_cast<float2*>(p.y + oc + Wo) = make_float2(v[2], v[3]);
    }
}

// ---- 5x5, stride 2, padding 2 (FPN4 down-sampling layers, models/mvs4net_utils.py:438,444; D = 1) -------------------
template <int CIN, int COUT>
struct Conv5Params {
    float w[25 * CIN * COUT];  // [ky][kx][ci][co]
    float bias[COUT];
    const float* x;
    float* y;
    int B, H, W, relu, co_total, co_off;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(128, 3) conv5s2_kernel(const __grid_constant__ Conv5Params<CIN, COUT> p) {
    const int Ho = p.H / 2, Wo = p.W / 2;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (2 * i >= Wo || oy >= Ho) return;
    const int H = p.H, W = p.W;
    const int xi = 4 * i;  // outputs 2i, 2i+1 read input columns xi-2 .. xi+4
    const size_t plane = (size_t)H * W, oplane = (size_t)Ho * Wo;
    float acc[2][COUT];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[c][co] = 0.0f;
    bool vy[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) vy[r] = (unsigned)(2 * oy - 2 + r) < (unsigned)H;
    const bool vl = xi > 0, vr = xi + 4 < W;
    const float* xp = p.x + ((size_t)b * CIN) * plane + (size_t)(2 * oy - 2) * W + xi;
#pragma unroll kScUnroll
    for (int ci = 0; ci < CIN; ++ci) {
        const float* q = xp + (size_t)ci * plane;
        float in[5][7];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const float* qr = q + (size_t)r * W;
            const float2 l = ldz2(qr - 2, vy[r] && vl);
            const float4 m = ldz4(qr, vy[r]);
            in[r][0] = l.x; in[r][1] = l.y; in[r][2] = m.x; in[r][3] = m.y; in[r][4] = m.z; in[r][5] = m.w;
            in[r][6] = ldz(qr + 4, vy[r] && vr);
        }
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
            for (int kx = 0; kx < 5; ++kx)
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    const float wv = p.w[((ky * 5 + kx) * CIN + ci) * COUT + co];
                    acc[0][co] = fmaf(wv, in[ky][kx], acc[0][co]);
                    acc[1][co] = fmaf(wv, in[ky][kx + 2], acc[1][co]);
                }
    }
    float* yp = p.y + ((size_t)b * p.co_total + p.co_off) * oplane + (size_t)oy * Wo + 2 * i;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
        float v0 = acc[0][co] + p.bias[co], v1 = acc[1][co] + p.bias[co];
        if (p.relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
        *reinterpret_cast<float2*>(yp + (size_t)co * oplane) = make_float2(v0, v1);
    }
}

// ---- 32/64-channel stride-1 layers (reg2d conv4 / conv6, FPN4 conv3.1 / conv3.2) ---------------------------------------
// Their filter banks (110 .. 442 KB) do not fit the kernel-parameter space, and slicing them into <= 32 KB launches
// leaves each launch with a fraction of a wave at 1/4 .. 1/8 resolution.  Here the folded weights stay in global memory
// as [kd][ky][kx][ci][co]: a thread owns 2x2 output pixels x 16 output channels, blockIdx.z enumerates (batch, depth,
// 16-channel slice) so that ONE launch covers every output channel.  The CTA stages the 9 x CIN x 16 weights of the
// current depth plane in shared memory (<= 36.9 KB) and reads the 16 weights of a (tap, ci) as four warp-uniform
// LDS.128 broadcasts feeding 64 FFMAs.  (First version: CTA-uniform LDG.128 straight from L1/L2 - the compiler hoisted
// 36 loads per input channel to hide their latency, 254 registers, 8 warps per SM, 30 % of FP32 peak.)
template <int KD, int CIN>
struct MidConvParams {
    const float* x;
    const float* w;     // dev [KD][3][3][CIN][co_total]
    const float* bias;  // dev [co_total]
    float* y;
    int B, D, H, W, relu, co_total;
};

template <int KD, int CIN>
__global__ void __launch_bounds__(128, MVSTER_MID_MINB) midconv_s1_kernel(const MidConvParams<KD, CIN> p) {
    constexpr int COT = 16;
    __shared__ float4 wsm[9 * CIN * (COT / 4)];  // [ky][kx][ci][16] of the current kd plane (<= 36.9 KB)
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), j = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int nsl = p.co_total / COT;
    const int sl = blockIdx.z % nsl, bd = blockIdx.z / nsl;
    const int b = bd / p.D, d = bd % p.D;
    const int H = p.H, W = p.W;
    const bool active = 2 * i < W && 2 * j < H;  // no early exit: every thread stages weights and meets the barriers
    const int x0 = min(2 * i, W - 2), y0 = min(2 * j, H - 2);
    const size_t plane = (size_t)H * W;
    float acc[2][2][COT];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int co = 0; co < COT; ++co) acc[a][c][co] = 0.0f;
    const bool vl = x0 > 0, vr = x0 + 2 < W;
    bool vy[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) vy[r] = (unsigned)(y0 - 1 + r) < (unsigned)H;
#pragma unroll 1
    for (int kd = 0; kd < KD; ++kd) {
        const int dz = d + kd - KD / 2;  // CTA-uniform
        if ((unsigned)dz >= (unsigned)p.D) continue;
        if (kd > 0) __syncthreads();  // everyone is done with the previous plane
        {
            const float* wk = p.w + ((size_t)kd * 9 * CIN) * p.co_total + sl * COT;
            for (int idx = threadIdx.x; idx < 9 * CIN * (COT / 4); idx += 128)
                wsm[idx] = __ldg(reinterpret_cast<const float4*>(wk + (size_t)(idx >> 2) * p.co_total) + (idx & 3));
        }
        __syncthreads();
        const float* xp = p.x + (((size_t)b * CIN) * p.D + dz) * plane + (size_t)(y0 - 1) * W + x0;
#pragma unroll kMidUnroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float* q = xp + (size_t)ci * p.D * plane;
            float in[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float* qr = q + (size_t)r * W;
                const float2 m = ldz2(qr, vy[r]);
                in[r][0] = ldz(qr - 1, vy[r] && vl);
                in[r][1] = m.x;
                in[r][2] = m.y;
                in[r][3] = ldz(qr + 2, vy[r] && vr);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4* wq = wsm + ((ky * 3 + kx) * CIN + ci) * (COT / 4);  // warp-uniform: LDS broadcast
#pragma unroll
                    for (int k = 0; k < COT / 4; ++k) {
                        const float4 w4 = wq[k];
                        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int co = 4 * k + e;
                            acc[0][0][co] = fmaf(wv[e], in[ky][kx], acc[0][0][co]);
                            acc[0][1][co] = fmaf(wv[e], in[ky][kx + 1], acc[0][1][co]);
                            acc[1][0][co] = fmaf(wv[e], in[ky + 1][kx], acc[1][0][co]);
                            acc[1][1][co] = fmaf(wv[e], in[ky + 1][kx + 1], acc[1][1][co]);
                        }
                    }
                }
        }
    }
    if (!active) return;
    float* yp = p.y + (((size_t)b * p.co_total + sl * COT) * p.D + d) * plane + (size_t)y0 * W + x0;
#pragma unroll
    for (int co = 0; co < COT; ++co) {
        const float bv = __ldg(p.bias + sl * COT + co);
        float v[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                v[a][c] = acc[a][c][co] + bv;
                if (p.relu) v[a][c] = fmaxf(v[a][c], 0.0f);
            }
        float* yc = yp + (size_t)co * p.D * plane;
        *reinterpret_cast<float2*>(yc) = make_float2(v[0][0], v[0][1]);
        *reinterpret_cast<float2*>(yc + W) = make_float2(v[1][0], v[1][1]);
    }
}

template <int KD, int CIN>
static int launch_mid(const float* x, const float* w, const float* bias, float* y, int B, int Cout, int D, int H, int W,
                      int relu, cudaStream_t s) {
    MidConvParams<KD, CIN> p{x, w, bias, y, B, D, H, W, relu, Cout};
    const long long z = (long long)B * D * (Cout / 16);
    if (z > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_mid: B*D*Cout/16 too large");
    dim3 grid((W / 2 + 31) / 32, (H / 2 + 3) / 4, (unsigned)z);
    midconv_s1_kernel<KD, CIN><<<grid, 128, 0, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("conv3d_mid launch");
    return MVSTER_OK;
}

// ---- 5x5 stride-2 layers with device-resident weights (FPN4 conv1.0 / conv2.0 / conv3.0) --------------------------------
// ncu on conv5s2_kernel: with a 25 KB filter bank in the kernel-parameter space the constant cache thrashes (issue
// 15-35 %).  Same recipe as midconv_s1_kernel instead: weights [ky][kx][ci][co] in global memory read as CTA-uniform
// shared-memory broadcasts (staged per chunk of 16 input channels, 25.6 KB), a thread owns 2x2 output pixels x 16
// output channels (1600 FFMAs per input channel for 21 input and 100 weight loads), blockIdx.z enumerates (batch,
// 16-channel slice).
template <int CIN>
struct MidConv5Params {
    const float* x;
    const float* w;     // dev [5][5][CIN][co_total]
    const float* bias;  // dev [co_total]
    float* y;
    int B, H, W, relu, co_total;
};

template <int CIN>
__global__ void __launch_bounds__(128, MVSTER_MID_MINB) midconv5s2_kernel(const MidConv5Params<CIN> p) {
    constexpr int COT = 16;
    constexpr int CC = CIN < 16 ? CIN : 16;  // input channels per staged weight chunk (25 * 16 * 16 floats = 25.6 KB)
    __shared__ float4 wsm[25 * CC * (COT / 4)];  // [ky][kx][cc][16]
    const int Ho = p.H / 2, Wo = p.W / 2;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31), j = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int nsl = p.co_total / COT;
    const int sl = blockIdx.z % nsl, b = blockIdx.z / nsl;
    const bool active = 2 * i < Wo && 2 * j < Ho;  // no early exit: every thread stages weights and meets the barriers
    const int H = p.H, W = p.W;
    const int xi = min(4 * i, W - 4), yi = min(4 * j, H - 4);  // outputs read input rows yi-2..yi+4, columns xi-2..xi+4
    const size_t plane = (size_t)H * W, oplane = (size_t)Ho * Wo;
    float acc[2][2][COT];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int co = 0; co < COT; ++co) acc[a][c][co] = 0.0f;
    bool vy[7];
#pragma unroll
    for (int r = 0; r < 7; ++r) vy[r] = (unsigned)(yi - 2 + r) < (unsigned)H;
    const bool vl = xi > 0, vr = xi + 4 < W;
    const float* xp = p.x + ((size_t)b * CIN) * plane + (size_t)(yi - 2) * W + xi;
    const float* wk = p.w + sl * COT;
#pragma unroll 1
    for (int c0 = 0; c0 < CIN; c0 += CC) {
        if (c0 > 0) __syncthreads();
        for (int idx = threadIdx.x; idx < 25 * CC * (COT / 4); idx += 128) {
            const int tap = (idx >> 2) / CC, cc = (idx >> 2) % CC;
            wsm[idx] = __ldg(reinterpret_cast<const float4*>(wk + ((size_t)tap * CIN + c0 + cc) * p.co_total) + (idx & 3));
        }
        __syncthreads();
#pragma unroll 1
        for (int cc = 0; cc < CC; ++cc) {
            const float* q = xp + (size_t)(c0 + cc) * plane;
            float in[7][7];
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                const float* qr = q + (size_t)r * W;
                const float2 l = ldz2(qr - 2, vy[r] && vl);
                const float4 m = ldz4(qr, vy[r]);
                in[r][0] = l.x; in[r][1] = l.y; in[r][2] = m.x; in[r][3] = m.y; in[r][4] = m.z; in[r][5] = m.w;
                in[r][6] = ldz(qr + 4, vy[r] && vr);
            }
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const float4* wq = wsm + ((ky * 5 + kx) * CC + cc) * (COT / 4);  // warp-uniform: LDS broadcast
#pragma unroll
                    for (int k = 0; k < COT / 4; ++k) {
                        const float4 w4 = wq[k];
                        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int co = 4 * k + e;
                            acc[0][0][co] = fmaf(wv[e], in[ky][kx], acc[0][0][co]);
                            acc[0][1][co] = fmaf(wv[e], in[ky][kx + 2], acc[0][1][co]);
                            acc[1][0][co] = fmaf(wv[e], in[ky + 2][kx], acc[1][0][co]);
                            acc[1][1][co] = fmaf(wv[e], in[ky + 2][kx + 2], acc[1][1][co]);
                        }
                    }
                }
        }
    }
    if (!active) return;
    float* yp = p.y + ((size_t)b * p.co_total + sl * COT) * oplane + (size_t)(2 * j) * Wo + 2 * i;
#pragma unroll
    for (int co = 0; co < COT; ++co) {
        const float bv = __ldg(p.bias + sl * COT + co);
        float v[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                v[a][c] = acc[a][c][co] + bv;
                if (p.relu) v[a][c] = fmaxf(v[a][c], 0.0f);
            }
        float* yc = yp + (size_t)co * oplane;
        *reinterpret_cast<float2*>(yc) = make_float2(v[0][0], v[0][1]);
        *reinterpret_cast<float2*>(yc + Wo) = make_float2(v[1][0], v[1][1]);
    }
}

template <int CIN>
static int launch_mid5(const float* x, const float* w, const float* bias, float* y, int B, int Cout, int H, int W, int relu,
                       cudaStream_t s) {
    MidConv5Params<CIN> p{x, w, bias, y, B, H, W, relu, Cout};
    const long long z = (long long)B * (Cout / 16);
    if (z > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_mid5: B*Cout/16 too large");
    dim3 grid((W / 4 + 31) / 32, (H / 4 + 3) / 4, (unsigned)z);
    midconv5s2_kernel<CIN><<<grid, 128, 0, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("conv2d_mid5 launch");
    return MVSTER_OK;
}

template <int CIN, int COUT>
static int launch_conv5(const float* x, const float* w_host, const float* bias_host, float* y, int B, int H, int W,
                        int relu, int co_total, int co_off, cudaStream_t s) {
    static thread_local Conv5Params<CIN, COUT> p;
    static_assert(sizeof(Conv5Params<CIN, COUT>) <= 32000, "filter bank must fit the kernel-parameter space");
    memcpy(p.w, w_host, sizeof(p.w));
    memcpy(p.bias, bias_host, sizeof(p.bias));
    p.x = x; p.y = y; p.B = B; p.H = H; p.W = W; p.relu = relu; p.co_total = co_total; p.co_off = co_off;
    if (B > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_small: B too large");
    dim3 grid((W / 4 + 31) / 32, (H / 2 + 3) / 4, B);
    conv5s2_kernel<CIN, COUT><<<grid, 128, 0, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("conv2d_small launch");
    return MVSTER_OK;
}

template <int KD, int CIN, int COUT, int MODE>
static int launch_small(const float* x, const float* w_host, const float* bias_host, const float* skip, float* y, int B,
                        int D, int H, int W, int relu, cudaStream_t s, int co_total = COUT, int co_off = 0) {
    static thread_local SmallConvParams<KD, CIN, COUT> p;
    static_assert(sizeof(SmallConvParams<KD, CIN, COUT>) <= 32000, "filter bank must fit the kernel-parameter space");
    static_assert(MODE == 0 || KD == 1, "strided / transposed layers of reg2d are (1,3,3)");
    memcpy(p.w, w_host, sizeof(p.w));
    memcpy(p.bias, bias_host, sizeof(p.bias));
    p.x = x; p.skip = skip; p.y = y; p.B = B; p.D = D; p.H = H; p.W = W; p.relu = relu;
    p.co_total = co_total; p.co_off = co_off;
    if ((long long)B * D > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_small: B*D too large");
    if constexpr (MODE == 0) {
        dim3 grid((W / 2 + 31) / 32, (H / 2 + 3) / 4, B * D);
        smallconv_s1_kernel<KD, CIN, COUT><<<grid, 128, 0, s>>>(p);
    } else if constexpr (MODE == 1) {
        dim3 grid((W / 4 + 31) / 32, (H / 2 + 3) / 4, B * D);
        smallconv_s2_kernel<CIN, COUT><<<grid, 128, 0, s>>>(p);
    } else {
        dim3 grid((W + 31) / 32, (H + 3) / 4, B * D);
        smallconv_t2_kernel<CIN, COUT><<<grid, 128, 0, s>>>(p);
    }
    count_launch();
    MVSTER_CHECK_LAUNCH("conv3d_small launch");
    return MVSTER_OK;
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_conv3d_small(const float* x, const float* w_host, const float* bias_host, const float* skip,
                                   float* y, int B, int Cin, int Cout, int D, int H, int W, int kd, int mode, int relu,
                                   void* stream) {
    if (!x || !w_host || !bias_host || !y) return fail(MVSTER_ERR_BAD_ARG, "conv3d_small: null pointer");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "conv3d_small: non-positive dimension");
    if (mode < 0 || mode > 2) return fail(MVSTER_ERR_BAD_ARG, "conv3d_small: mode %d not in {0,1,2}", mode);
    if (skip != nullptr && mode != 2) return fail(MVSTER_ERR_BAD_ARG, "conv3d_small: skip is only defined for mode 2");
    if (mode == 0 && ((H & 1) || (W & 1))) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_small: stride-1 needs even H, W");
    if (mode == 1 && ((H & 1) || (W & 3))) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_small: stride-2 needs H%%2==0, W%%4==0");
    if ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)skip)) % 16)
        return fail(MVSTER_ERR_ALIGN, "conv3d_small: tensors must be 16-byte aligned");
    DeviceGuard guard(y);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
#define SC_CASE(KD_, CI_, CO_, MODE_)                                     \
    if (kd == KD_ && Cin == CI_ && Cout == CO_ && mode == MODE_)          \
        return launch_small<KD_, CI_, CO_, MODE_>(x, w_host, bias_host, skip, y, B, D, H, W, relu, s);
    SC_CASE(1, 4, 8, 0)    // reg2d.conv0, G = 4
    SC_CASE(1, 8, 8, 0)    // reg2d.conv0, G = 8
    SC_CASE(1, 8, 16, 1)   // reg2d.conv1
    SC_CASE(3, 16, 16, 0)  // reg2d.conv2
    SC_CASE(1, 16, 32, 1)  // reg2d.conv3
    SC_CASE(1, 32, 16, 2)  // reg2d.conv9 (+ conv2 skip)
    SC_CASE(1, 16, 8, 2)   // reg2d.conv11 (unfused form of mvster_regtail's first half)
#undef SC_CASE
    return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_small: no kernel for Cin=%d Cout=%d kd=%d mode=%d", Cin, Cout, kd, mode);
}

// One slice of `Cout` output channels (co_off .. co_off+Cout-1 of a Cout_total-channel y) of a strided / transposed
// reg2d layer whose whole filter bank exceeds the kernel-parameter space (conv5: 32->64 stride 2, conv7: 64->32
// transposed + skip); w_host [1,3,3,Cin,Cout] and bias_host [Cout] hold the slice.
extern "C" int mvster_conv3d_small_slice(const float* x, const float* w_host, const float* bias_host, const float* skip,
                                         float* y, int B, int Cin, int Cout, int Cout_total, int co_off, int D, int H,
                                         int W, int mode, int relu, void* stream) {
    if (!x || !w_host || !bias_host || !y) return fail(MVSTER_ERR_BAD_ARG, "conv3d_small_slice: null pointer");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || Cout <= 0 || co_off < 0 || co_off + Cout > Cout_total)
        return fail(MVSTER_ERR_BAD_ARG, "conv3d_small_slice: bad dimension / channel slice");
    if (mode != 1 && mode != 2) return fail(MVSTER_ERR_BAD_ARG, "conv3d_small_slice: mode %d not in {1,2}", mode);
    if (skip != nullptr && mode != 2) return fail(MVSTER_ERR_BAD_ARG, "conv3d_small_slice: skip is only defined for mode 2");
    if (mode == 1 && ((H & 1) || (W & 3)))
        return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_small_slice: stride-2 needs H%%2==0, W%%4==0");
    if ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)skip)) % 16)
        return fail(MVSTER_ERR_ALIGN, "conv3d_small_slice: tensors must be 16-byte aligned");
    DeviceGuard guard(y);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (Cin == 32 && Cout == 16 && mode == 1)  // reg2d.conv5
        return launch_small<1, 32, 16, 1>(x, w_host, bias_host, skip, y, B, D, H, W, relu, s, Cout_total, co_off);
    if (Cin == 64 && Cout == 8 && mode == 2)   // reg2d.conv7 (+ conv4 skip)
        return launch_small<1, 64, 8, 2>(x, w_host, bias_host, skip, y, B, D, H, W, relu, s, Cout_total, co_off);
    return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_small_slice: no kernel for Cin=%d Cout=%d mode=%d", Cin, Cout, mode);
}

// 32/64-channel stride-1 layers with the folded weights in device memory (see midconv_s1_kernel).
extern "C" int mvster_conv3d_mid(const float* x, const float* w_dev, const float* bias_dev, float* y, int B, int Cin,
                                 int Cout, int D, int H, int W, int kd, int relu, void* stream) {
    if (!x || !w_dev || !bias_dev || !y) return fail(MVSTER_ERR_BAD_ARG, "conv3d_mid: null pointer");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "conv3d_mid: non-positive dimension");
    if ((H & 1) || (W & 1)) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_mid: needs even H, W");
    if (Cout % 16) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_mid: Cout=%d is not a multiple of 16", Cout);
    if ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)w_dev)) % 16)
        return fail(MVSTER_ERR_ALIGN, "conv3d_mid: tensors must be 16-byte aligned");
    DeviceGuard guard(y);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (kd == 1 && Cin == 32) return launch_mid<1, 32>(x, w_dev, bias_dev, y, B, Cout, D, H, W, relu, s);
    if (kd == 1 && Cin == 64) return launch_mid<1, 64>(x, w_dev, bias_dev, y, B, Cout, D, H, W, relu, s);
    if (kd == 3 && Cin == 32) return launch_mid<3, 32>(x, w_dev, bias_dev, y, B, Cout, D, H, W, relu, s);
    if (kd == 3 && Cin == 64) return launch_mid<3, 64>(x, w_dev, bias_dev, y, B, Cout, D, H, W, relu, s);
    return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_mid: no kernel for Cin=%d kd=%d", Cin, kd);
}

// 5x5 stride-2 padding-2 layers of FPN4 with the folded weights in device memory (see midconv5s2_kernel).
extern "C" int mvster_conv2d_mid5(const float* x, const float* w_dev, const float* bias_dev, float* y, int B, int Cin,
                                  int Cout, int H, int W, int relu, void* stream) {
    if (!x || !w_dev || !bias_dev || !y) return fail(MVSTER_ERR_BAD_ARG, "conv2d_mid5: null pointer");
    if (B <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "conv2d_mid5: non-positive dimension");
    if ((H & 3) || (W & 3)) return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_mid5: needs H%%4==0, W%%4==0");
    if (Cout % 16) return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_mid5: Cout=%d is not a multiple of 16", Cout);
    if ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)w_dev)) % 16)
        return fail(MVSTER_ERR_ALIGN, "conv2d_mid5: tensors must be 16-byte aligned");
    DeviceGuard guard(y);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    if (Cin == 8) return launch_mid5<8>(x, w_dev, bias_dev, y, B, Cout, H, W, relu, s);
    if (Cin == 16) return launch_mid5<16>(x, w_dev, bias_dev, y, B, Cout, H, W, relu, s);
    if (Cin == 32) return launch_mid5<32>(x, w_dev, bias_dev, y, B, Cout, H, W, relu, s);
    return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_mid5: no kernel for Cin=%d", Cin);
}

// 2-D layers of FPN4 (models/mvs4net_utils.py:431-449, eval mode, BatchNorm folded): NCHW planar fp32 in and out.
// A layer with more output channels than one filter-bank slice holds is launched once per slice of `Cout` channels
// (co_off .. co_off+Cout-1 of a Cout_total-channel output).
extern "C" int mvster_conv2d_small(const float* x, const float* w_host, const float* bias_host, float* y, int B, int Cin,
                                   int Cout, int Cout_total, int co_off, int H, int W, int ksize, int stride, int relu,
                                   void* stream) {
    if (!x || !w_host || !bias_host || !y) return fail(MVSTER_ERR_BAD_ARG, "conv2d_small: null pointer");
    if (B <= 0 || H <= 0 || W <= 0 || Cout <= 0 || co_off < 0 || co_off + Cout > Cout_total)
        return fail(MVSTER_ERR_BAD_ARG, "conv2d_small: bad dimension / channel slice");
    if (!((ksize == 3 && stride == 1) || (ksize == 5 && stride == 2)))
        return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_small: only 3x3 stride 1 and 5x5 stride 2 are built");
    if (stride == 1 && ((H & 1) || (W & 1))) return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_small: 3x3 needs even H, W");
    if (stride == 2 && ((H & 1) || (W & 3))) return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_small: 5x5/2 needs H%%2==0, W%%4==0");
    if ((((uintptr_t)x) | ((uintptr_t)y)) % 16) return fail(MVSTER_ERR_ALIGN, "conv2d_small: tensors must be 16-byte aligned");
    DeviceGuard guard(y);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
#define C3_CASE(CI_, CO_)                                    \
    if (ksize == 3 && Cin == CI_ && Cout == CO_)             \
        return launch_small<1, CI_, CO_, 0>(x, w_host, bias_host, nullptr, y, B, 1, H, W, relu, s, Cout_total, co_off);
#define C5_CASE(CI_, CO_)                                    \
    if (ksize == 5 && Cin == CI_ && Cout == CO_)             \
        return launch_conv5<CI_, CO_>(x, w_host, bias_host, y, B, H, W, relu, Cout_total, co_off, s);
    C3_CASE(3, 8)     // FPN4.conv0.0
    C3_CASE(8, 8)     // FPN4.conv0.1
    C3_CASE(16, 16)   // FPN4.conv1.1 / conv1.2
    C3_CASE(32, 16)   // FPN4.conv2.1 / conv2.2, two slices of 16 output channels
    C5_CASE(8, 16)    // FPN4.conv1.0
    C5_CASE(16, 16)   // FPN4.conv2.0, two slices of 16 output channels
#undef C3_CASE
#undef C5_CASE
    return fail(MVSTER_ERR_UNSUPPORTED, "conv2d_small: no kernel for Cin=%d Cout=%d k=%d", Cin, Cout, ksize);
}
