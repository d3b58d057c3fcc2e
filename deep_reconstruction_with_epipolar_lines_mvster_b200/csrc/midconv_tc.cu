// 32/64-channel stride-1 convolutions (reg2d conv4 / conv6: Conv3d (3,3,3); FPN4 conv3.1 / conv3.2 / out2: 3x3 with
// D = 1; models/mvs4net_utils.py:456-468,896-903, eval mode, BatchNorm folded) as an implicit GEMM on the tensor cores
// with fp32-grade accuracy: every operand is split into two TF32 numbers (x = hi + lo, hi = rna_tf32(x),
// lo = rna_tf32(x - hi)) and a product is accumulated as  lo*hi + hi*lo + hi*hi  in fp32 ("3xTF32"; the dropped lo*lo
// term is 2^-22 relative).  Measured errors against float64 are in tests/test_gpu_network.py.
//
// Why here and not in the few-channel layers: these layers are real contractions (K = 9 x Cin x kd = 288 .. 1728,
// N = 32 / 64 output channels).  What it buys, measured on B200: mma.sync (the legacy HMMA path, the only one that takes
// register fragments) issues 0.465 m16n8k8 TF32 instructions per clock and SM = 476 FMA/clk/SM, 159 after the
// three-way split - against 128 for the FP32 pipe (scripts/micro/hmma_rate.cu, ffma_rate.cu).  This kernel is 8 %
// faster than the FP32 SIMT kernel on the same layers (1.84 vs 1.99 ms per 832x1152 scene); a real step needs tcgen05
// (2048 TF32 FMA/clk/SM) with the operands split in shared memory - not built.
//
// GEMM view per CTA: M = 128 output pixels (a 32 x 4 tile of one (batch, depth) plane; warp w owns row w = two m16
// tiles), N = Cout, K walks (kd, 8-channel chunk, tap).  Per chunk the CTA stages the 8 x 6 x 34 input window (with
// the zero padding of the convolution) and the 9 x 8 x Cout weights in shared memory, both already split into hi / lo;
// the strides (232 floats per input channel plane, Cout + 8 per weight row) make every mma.m16n8k8 fragment load a
// conflict-free LDS.32: the four k-lanes of a quad land 8 banks apart, the eight row / column lanes are consecutive.
// Weights are split once per layer on the device (mvster_tf32_split) and cached by the caller.
#include "common.cuh"

namespace mvster {

constexpr int kTcTW = 32, kTcTH = 4;             // output tile
constexpr int kTcXR = kTcTH + 2, kTcXC = kTcTW + 2;  // input window rows / columns
constexpr int kTcXS = 232;                       // floats per input-channel plane in shared memory (204 used; 232 % 32 == 8)
constexpr int kTcKC = 8;                         // input channels per chunk = one k8 step per tap

struct MidTcParams {
    const float* x;
    const float* w_hi;  // dev [KD][3][3][CIN][CO], TF32 values
    const float* w_lo;
    const float* bias;  // dev [CO]
    float* y;
    int B, D, H, W, relu;
};

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// d = a * b + c (d and c may be different registers)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const float (&c)[4]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}

template <int CO>
struct MidTcCfg {
    static constexpr int CS = CO + 8;                                    // weight row stride (CS % 32 == 8)
    static constexpr int XF = kTcKC * kTcXS;                             // floats of one input array (hi or lo)
    static constexpr int WF = 9 * kTcKC * CS;                            // floats of one weight array
    static constexpr int SMEM = (2 * XF + 2 * WF) * 4;
    static constexpr int MINB = CO == 64 ? 3 : 4;
};

template <int KD, int CIN, int CO>
__global__ void __launch_bounds__(128, MidTcCfg<CO>::MINB) midconv_tc_kernel(const MidTcParams p) {
    using K = MidTcCfg<CO>;
    constexpr int NT = CO / 8, CS = K::CS;
    extern __shared__ __align__(16) float smem_tc[];
    float* Xh = smem_tc;
    float* Xl = Xh + K::XF;
    float* Wh = Xl + K::XF;
    float* Wl = Wh + K::WF;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tx0 = blockIdx.x * kTcTW, ty0 = blockIdx.y * kTcTH;
    const int b = blockIdx.z / p.D, d = blockIdx.z % p.D;
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;

    float acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.0f;

    bool first = true;
#pragma unroll 1
    for (int kd = 0; kd < KD; ++kd) {
        const int dz = d + kd - KD / 2;  // CTA-uniform
        if ((unsigned)dz >= (unsigned)p.D) continue;
#pragma unroll 1
        for (int c0 = 0; c0 < CIN; c0 += kTcKC) {
            if (!first) __syncthreads();  // every warp is done with the previous chunk
            first = false;
            // ---- stage the input window of 8 channels, split into hi / lo --------------------------------------------
            const float* xb = p.x + (((size_t)b * CIN + c0) * p.D + dz) * plane;
            for (int idx = tid; idx < kTcKC * kTcXR * kTcXC; idx += 128) {
                const int ci = idx / (kTcXR * kTcXC), rem = idx - ci * (kTcXR * kTcXC);
                const int r = rem / kTcXC, c = rem - r * kTcXC;
                const int gy = ty0 - 1 + r, gx = tx0 - 1 + c;
                float v = 0.0f;
                if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W)
                    v = __ldg(xb + (size_t)ci * p.D * plane + (size_t)gy * W + gx);
                const uint32_t hi = to_tf32(v);
                const uint32_t lo = to_tf32(v - __uint_as_float(hi));
                Xh[ci * kTcXS + rem] = __uint_as_float(hi);
                Xl[ci * kTcXS + rem] = __uint_as_float(lo);
            }
            // ---- stage the 9 x 8 x CO weights of this chunk (already split) ------------------------------------------
            {
                constexpr int Q = CO / 4;  // float4 per weight row
                for (int idx = tid; idx < 9 * kTcKC * Q; idx += 128) {
                    const int row = idx / Q, q = idx - row * Q;       // row = tap * 8 + ci
                    const int tap = row >> 3, ci = row & 7;
                    const size_t src = ((size_t)(kd * 9 + tap) * CIN + c0 + ci) * CO + 4 * q;
                    *reinterpret_cast<float4*>(Wh + row * CS + 4 * q) = __ldg(reinterpret_cast<const float4*>(p.w_hi + src));
                    *reinterpret_cast<float4*>(Wl + row * CS + 4 * q) = __ldg(reinterpret_cast<const float4*>(p.w_lo + src));
                }
            }
            __syncthreads();
            // ---- 9 taps x (2 m-tiles x NT n-tiles x 3 products) ------------------------------------------------------
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    uint32_t ah[2][4], al[2][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const int base = (warp + ky) * kTcXC + mt * 16 + kx + g;
                        ah[mt][0] = __float_as_uint(Xh[t * kTcXS + base]);
                        ah[mt][1] = __float_as_uint(Xh[t * kTcXS + base + 8]);
                        ah[mt][2] = __float_as_uint(Xh[(t + 4) * kTcXS + base]);
                        ah[mt][3] = __float_as_uint(Xh[(t + 4) * kTcXS + base + 8]);
                        al[mt][0] = __float_as_uint(Xl[t * kTcXS + base]);
                        al[mt][1] = __float_as_uint(Xl[t * kTcXS + base + 8]);
                        al[mt][2] = __float_as_uint(Xl[(t + 4) * kTcXS + base]);
                        al[mt][3] = __float_as_uint(Xl[(t + 4) * kTcXS + base + 8]);
                    }
                    const float* wh = Wh + ((ky * 3 + kx) * kTcKC + t) * CS + g;
                    const float* wl = Wl + ((ky * 3 + kx) * kTcKC + t) * CS + g;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const uint32_t bh0 = __float_as_uint(wh[nt * 8]), bh1 = __float_as_uint(wh[4 * CS + nt * 8]);
                        const uint32_t bl0 = __float_as_uint(wl[nt * 8]), bl1 = __float_as_uint(wl[4 * CS + nt * 8]);
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            // The tensor core adds into its fp32 accumulator with truncation: a chain of K/8 x 3 = 648
                            // such adds (conv6) leaves a bias of ~1e-5 of the output range.  The three products of one
                            // k8 step are therefore summed from zero (small terms first) and added to the running sum
                            // with a rounded FADD: the long chain is round-to-nearest, like the FP32 kernel's.
                            const float zero[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                            float part[4];
                            mma_tf32(part, al[mt], bh0, bh1, zero);
                            mma_tf32(part, ah[mt], bl0, bl1, part);
                            mma_tf32(part, ah[mt], bh0, bh1, part);
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[mt][nt][e] += part[e];
                        }
                    }
                }
            }
        }
    }

    // ---- epilogue: bias, ReLU, planar store (c0/c1: pixel g, channels 2t / 2t+1; c2/c3: pixel g + 8) ----------------
    const int gy = ty0 + warp;
    if (gy >= H) return;
    float* yb = p.y + (((size_t)b * CO) * p.D + d) * plane + (size_t)gy * W;
    const size_t cstride = (size_t)p.D * plane;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int co = nt * 8 + 2 * t;
        const float b0 = __ldg(p.bias + co), b1 = __ldg(p.bias + co + 1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int gx = tx0 + mt * 16 + g + 8 * half;
                if (gx >= W) continue;
                float v0 = acc[mt][nt][2 * half] + b0, v1 = acc[mt][nt][2 * half + 1] + b1;
                if (p.relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
                yb[(size_t)co * cstride + gx] = v0;
                yb[(size_t)(co + 1) * cstride + gx] = v1;
            }
        }
    }
}

__global__ void tf32_split_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = w[i];
    const uint32_t h = to_tf32(v);
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(to_tf32(v - __uint_as_float(h)));
}

template <int KD, int CIN, int CO>
static int launch_mid_tc(const MidTcParams& p, cudaStream_t s) {
    using K = MidTcCfg<CO>;
    static int attr_done[64] = {};
    const int st = ensure_dynamic_smem_bytes(midconv_tc_kernel<KD, CIN, CO>, K::SMEM, attr_done, "conv3d_mid_tc: cudaFuncSetAttribute");
    if (st != MVSTER_OK) return st;
    dim3 grid((p.W + kTcTW - 1) / kTcTW, (p.H + kTcTH - 1) / kTcTH, p.B * p.D);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_mid_tc: grid too large");
    midconv_tc_kernel<KD, CIN, CO><<<grid, 128, K::SMEM, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("conv3d_mid_tc launch");
    return MVSTER_OK;
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_tf32_split(const float* w, float* hi, float* lo, long long n, void* stream) {
    if (!w || !hi || !lo || n <= 0) return fail(MVSTER_ERR_BAD_ARG, "tf32_split: null pointer or empty tensor");
    DeviceGuard guard(hi);
    if (guard.status != MVSTER_OK) return guard.status;
    tf32_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, hi, lo, (size_t)n);
    count_launch();
    MVSTER_CHECK_LAUNCH("tf32_split launch");
    return MVSTER_OK;
}

extern "C" int mvster_conv3d_mid_tc(const float* x, const float* w_hi, const float* w_lo, const float* bias_dev, float* y,
                                    int B, int Cin, int Cout, int D, int H, int W, int kd, int relu, void* stream) {
    if (!x || !w_hi || !w_lo || !bias_dev || !y) return fail(MVSTER_ERR_BAD_ARG, "conv3d_mid_tc: null pointer");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "conv3d_mid_tc: non-positive dimension");
    if ((((uintptr_t)w_hi) | ((uintptr_t)w_lo)) % 16) return fail(MVSTER_ERR_ALIGN, "conv3d_mid_tc: weights must be 16-byte aligned");
    DeviceGuard guard(y);
    if (guard.status != MVSTER_OK) return guard.status;
    MidTcParams p{x, w_hi, w_lo, bias_dev, y, B, D, H, W, relu};
    cudaStream_t s = (cudaStream_t)stream;
#define MVSTER_TC_CASE(KD_, CI_, CO_) \
    if (kd == KD_ && Cin == CI_ && Cout == CO_) return launch_mid_tc<KD_, CI_, CO_>(p, s);
    MVSTER_TC_CASE(1, 32, 32) MVSTER_TC_CASE(1, 64, 64) MVSTER_TC_CASE(1, 64, 32)
    MVSTER_TC_CASE(3, 32, 32) MVSTER_TC_CASE(3, 64, 64)
#undef MVSTER_TC_CASE
    return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_mid_tc: no kernel for Cin=%d Cout=%d kd=%d", Cin, Cout, kd);
}
