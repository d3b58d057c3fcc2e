// Weight gradient of the regulariser's 3-D convolutions (training mode; the autograd of nn.Conv3d / nn.ConvTranspose3d
// inside ConvBnReLU3D / Deconv3d, models/mvs4net_utils.py:123-130, 884-926):
//     dW[a][b][kd][ky][kx] = sum_{n,d,y,x} A[n,a,d,y,x] * B[n,b, d+kd-pd, s*y+ky-1, s*x+kx-1]        (B zero outside)
//   Conv3d          (kernel (KD,3,3), padding (KD/2,1,1), stride (1,s,s)): A = grad_output [Cout], B = input [Cin]
//   ConvTranspose3d (kernel (1,3,3), padding (0,1,1), output_padding (0,1,1), stride (1,2,2)):
//                                                                          A = input [Cin], B = grad_output [Cout], s = 2
// i.e. exactly PyTorch's weight layouts [Cout,Cin,KD,3,3] / [Cin,Cout,1,3,3].
//
// Why hand-written: these layers have 4..64 channels and up to 2.6 M positions; cuDNN's wgrad_alg1_nd_float_engine
// spends 26 ms of a 96 ms training step on them (profiles/r02b_train_step_torch_profiler.txt) for ~20 GFMA of work.
// Here a lane owns one B channel x one depth tap x 9 spatial taps x 8 A channels = 72 accumulators and walks down a
// column of positions (lanes = consecutive x: every load is a coalesced row segment; at stride 1 the 3x3 window slides,
// three new B values per position), 72 FMAs per 11-17 loads.  A warp reduces its 72 sums with shuffles and writes one
// partial per (position chunk, weight); a second kernel sums the partials in a fixed order: bit-reproducible, no atomics.
#include "common.cuh"

namespace mvster {

constexpr int kWgRows = 64;    // positions rows per warp task
constexpr int kWgWarps = 4;    // warps per CTA: four (B channel, A block, depth tap) tasks of the same position chunk, so
                               // that small planes (one row chunk) still fill their CTAs and the warps share rows in L1
constexpr int kWgABlk = 8;     // A channels per lane

struct WgradParams {
    const float* A;   // [N, CA, D, HA, WA]
    const float* B;   // [N, CB, D, HB, WB]
    float* partial;   // [nparts][CA*CB*KD*9]
    int N, CA, CB, KD, D, HA, WA, HB, WB;
    int nychunk, nxstrip;
};

template <int S>
__global__ void __launch_bounds__(kWgWarps * 32) wgrad3d_kernel(const WgradParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blockIdx.x: (x strip, row chunk); blockIdx.y * kWgWarps + warp: (b, a block, kd); blockIdx.z: n * D + d
    const int xs = blockIdx.x % p.nxstrip;
    const int yc = blockIdx.x / p.nxstrip;
    const int nab = p.CA / kWgABlk;
    int t = blockIdx.y * kWgWarps + warp;
    if (t >= p.CB * nab * p.KD) return;
    const int kd = t % p.KD; t /= p.KD;
    const int ab = t % nab;
    const int b = t / nab;
    const int n = blockIdx.z / p.D, d = blockIdx.z % p.D;
    const int db = d + kd - p.KD / 2;

    float acc[9][kWgABlk];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int a = 0; a < kWgABlk; ++a) acc[k][a] = 0.0f;

    const int x = xs * 32 + lane;
    const int y0 = yc * kWgRows, y1 = min(y0 + kWgRows, p.HA);
    if ((unsigned)db < (unsigned)p.D) {  // warp-uniform: a depth tap outside the volume contributes nothing
        const size_t planeA = (size_t)p.HA * p.WA, planeB = (size_t)p.HB * p.WB;
        const float* Ap = p.A + (((size_t)n * p.CA + ab * kWgABlk) * p.D + d) * planeA;
        const float* Bp = p.B + (((size_t)n * p.CB + b) * p.D + db) * planeB;
        const size_t strideA = (size_t)p.D * planeA;  // channel stride of A
        const bool xin = x < p.WA;
        const int bx = S * x - 1;
        const bool c0 = xin && (unsigned)bx < (unsigned)p.WB, c1 = xin && (unsigned)(bx + 1) < (unsigned)p.WB,
                   c2 = xin && (unsigned)(bx + 2) < (unsigned)p.WB;
        auto load_row = [&](int by, float& v0, float& v1, float& v2) {
            const bool rin = (unsigned)by < (unsigned)p.HB;
            const float* r = Bp + (size_t)(rin ? by : 0) * p.WB + bx;
            v0 = (rin && c0) ? __ldg(r) : 0.0f;
            v1 = (rin && c1) ? __ldg(r + 1) : 0.0f;
            v2 = (rin && c2) ? __ldg(r + 2) : 0.0f;
        };
        // U rows per trip with all their loads issued first: at 12 warps per SM (146 registers) one row's 72 FMAs
        // did not cover the L2 latency of the next row's loads (first version: 9 ms for the ~27 GFMA of a step)
        constexpr int U = 4;
        if (S == 1) {
            // sliding window: rows y-1 .. y+U of B in w[0 .. 3(U+2)); the first two are carried between trips
            float w[3 * (U + 2)];
            load_row(y0 - 1, w[0], w[1], w[2]);
            load_row(y0, w[3], w[4], w[5]);
            for (int y = y0; y < y1; y += U) {
                float av[U][kWgABlk];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    load_row(y + u + 1, w[3 * (u + 2)], w[3 * (u + 2) + 1], w[3 * (u + 2) + 2]);
                    const bool live = xin && (y + u < y1);
#pragma unroll
                    for (int a = 0; a < kWgABlk; ++a) av[u][a] = live ? __ldg(Ap + a * strideA + (size_t)(y + u) * p.WA + x) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < 9; ++k)
#pragma unroll
                        for (int a = 0; a < kWgABlk; ++a) acc[k][a] = fmaf(w[3 * u + k], av[u][a], acc[k][a]);
#pragma unroll
                for (int k = 0; k < 6; ++k) w[k] = w[3 * U + k];
            }
        } else {
            // stride 2: position y uses B rows 2y-1, 2y, 2y+1, so consecutive positions share one row: rows
            // 2y-1 .. 2y+2U-1 of a trip in w[0 .. 3(2U+1)), the last one carried into the next trip
            float w[3 * (2 * U + 1)];
            load_row(S * y0 - 1, w[0], w[1], w[2]);
            for (int y = y0; y < y1; y += U) {
                float av[U][kWgABlk];
#pragma unroll
                for (int r = 1; r <= 2 * U; ++r) load_row(S * y - 1 + r, w[3 * r], w[3 * r + 1], w[3 * r + 2]);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool live = xin && (y + u < y1);
#pragma unroll
                    for (int a = 0; a < kWgABlk; ++a) av[u][a] = live ? __ldg(Ap + a * strideA + (size_t)(y + u) * p.WA + x) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < 9; ++k)
#pragma unroll
                        for (int a = 0; a < kWgABlk; ++a) acc[k][a] = fmaf(w[6 * u + k], av[u][a], acc[k][a]);
#pragma unroll
                for (int k = 0; k < 3; ++k) w[k] = w[6 * U + k];
            }
        }
    }
    // warp reduction of the 72 sums; lane 0 writes the warp's partial
    const int part = (blockIdx.z * p.nychunk + yc) * p.nxstrip + xs;
    const size_t nw = (size_t)p.CA * p.CB * p.KD * 9;
    float* out = p.partial + (size_t)part * nw;
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int a = 0; a < kWgABlk; ++a) {
            float v = acc[k][a];
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
            if (lane == 0) out[(((size_t)(ab * kWgABlk + a) * p.CB + b) * p.KD + kd) * 9 + k] = v;
        }
}

// dW[i] = sum over parts (fixed order), double accumulation
__global__ void wgrad3d_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int nparts, size_t nw) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nw) return;
    double s = 0.0;
    for (int q = 0; q < nparts; ++q) s += (double)partial[(size_t)q * nw + i];
    dw[i] = (float)s;
}

// ---- prob, the regulariser's last layer (nn.Conv3d(8, 1, 1) with bias, models/mvs4net_utils.py:914): weight and bias
// gradient dW[c] = sum x[n,c,s] g[n,0,s], db = sum g as one streaming pass on a (chunk, sample) grid (cuDNN's wgrad
// took 3.4 ms per step for these 8 + 1 numbers); partials in double, summed in a fixed order.
constexpr int kPwThreads = 256, kPwVec = 4, kPwChunk = kPwThreads * kPwVec * 4;  // 4096 positions per CTA

template <int C>
__global__ void __launch_bounds__(kPwThreads) pointwise_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                                     double* __restrict__ partial, long long S) {
    __shared__ double red[kPwThreads / 32][C + 1];
    const long long n = blockIdx.y, s0 = (long long)blockIdx.x * kPwChunk;
    const long long nvec = min((long long)kPwChunk, S - s0) >> 2;
    const float4* gp = reinterpret_cast<const float4*>(g + n * S + s0);
    float acc[C + 1];
#pragma unroll
    for (int c = 0; c <= C; ++c) acc[c] = 0.0f;
#pragma unroll
    for (int k = 0; k < kPwVec; ++k) {
        const long long i = (long long)k * kPwThreads + threadIdx.x;
        if (i < nvec) {
            const float4 gv = __ldg(gp + i);
            acc[C] += (gv.x + gv.y) + (gv.z + gv.w);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (n * C + c) * S + s0) + i);
                acc[c] = fmaf(xv.x, gv.x, fmaf(xv.y, gv.y, fmaf(xv.z, gv.z, fmaf(xv.w, gv.w, acc[c]))));
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c <= C; ++c) {
        double v = (double)acc[c];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x <= C) {
        double t = 0.0;
        for (int w = 0; w < kPwThreads / 32; ++w) t += red[w][threadIdx.x];
        partial[((size_t)n * gridDim.x + blockIdx.x) * (C + 1) + threadIdx.x] = t;
    }
}

__global__ void pointwise_wgrad_reduce_kernel(const double* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                              int nparts, int C) {
    const int c = threadIdx.x;
    if (c > C) return;
    double s = 0.0;
    for (int q = 0; q < nparts; ++q) s += partial[(size_t)q * (C + 1) + c];
    if (c < C) dw[c] = (float)s;
    else if (db) db[0] = (float)s;
}

static void wgrad_geometry(int N, int D, int HA, int WA, int* nychunk, int* nxstrip, long long* nparts) {
    *nychunk = (HA + kWgRows - 1) / kWgRows;
    *nxstrip = (WA + 31) / 32;
    *nparts = (long long)N * D * *nychunk * *nxstrip;
}

}  // namespace mvster

using namespace mvster;

extern "C" long long mvster_conv3d_wgrad_workspace_bytes(int N, int CA, int CB, int KD, int D, int HA, int WA) {
    if (N <= 0 || CA <= 0 || CB <= 0 || KD <= 0 || D <= 0 || HA <= 0 || WA <= 0) return 0;
    int nyc, nxs;
    long long nparts;
    wgrad_geometry(N, D, HA, WA, &nyc, &nxs, &nparts);
    return nparts * (long long)CA * CB * KD * 9 * 4;
}

extern "C" int mvster_conv3d_wgrad(const float* A, const float* B, float* dw, int N, int CA, int CB, int KD, int D,
                                   int HA, int WA, int HB, int WB, int stride, void* workspace, void* stream) {
    if (!A || !B || !dw || !workspace) return fail(MVSTER_ERR_BAD_ARG, "conv3d_wgrad: null pointer");
    if (N <= 0 || CA <= 0 || CB <= 0 || D <= 0 || HA <= 0 || WA <= 0 || HB <= 0 || WB <= 0)
        return fail(MVSTER_ERR_BAD_ARG, "conv3d_wgrad: non-positive dimension");
    if (KD != 1 && KD != 3) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_wgrad: depth kernel %d not in {1,3}", KD);
    if (stride != 1 && stride != 2) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_wgrad: stride %d not in {1,2}", stride);
    if (CA % kWgABlk) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_wgrad: CA=%d must be a multiple of %d", CA, kWgABlk);
    if (stride == 1 ? (HB != HA || WB != WA) : ((HB + 1) / 2 != HA || (WB + 1) / 2 != WA))
        return fail(MVSTER_ERR_BAD_ARG, "conv3d_wgrad: B %dx%d does not match A %dx%d at stride %d", HB, WB, HA, WA, stride);
    int nyc, nxs;
    long long nparts;
    wgrad_geometry(N, D, HA, WA, &nyc, &nxs, &nparts);
    const long long tasks = (long long)CB * (CA / kWgABlk) * KD;
    if ((tasks + kWgWarps - 1) / kWgWarps > 65535 || (long long)N * D > 65535 || nparts > 2147483647LL)
        return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_wgrad: grid too large");
    DeviceGuard guard(dw);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    WgradParams p{A, B, static_cast<float*>(workspace), N, CA, CB, KD, D, HA, WA, HB, WB, nyc, nxs};
    dim3 grid(nxs * nyc, (unsigned)((tasks + kWgWarps - 1) / kWgWarps), N * D);
    if (stride == 1) wgrad3d_kernel<1><<<grid, kWgWarps * 32, 0, s>>>(p);
    else wgrad3d_kernel<2><<<grid, kWgWarps * 32, 0, s>>>(p);
    count_launch();
    const size_t nw = (size_t)CA * CB * KD * 9;
    wgrad3d_reduce_kernel<<<(unsigned)((nw + 127) / 128), 128, 0, s>>>(p.partial, dw, (int)nparts, nw);
    count_launch();
    MVSTER_CHECK_LAUNCH("conv3d_wgrad launch");
    return MVSTER_OK;
}

extern "C" long long mvster_conv1x1_wgrad_workspace_bytes(int N, int C, long long S) {
    if (N <= 0 || C <= 0 || S <= 0) return 0;
    return (long long)N * ((S + kPwChunk - 1) / kPwChunk) * (C + 1) * 8;
}

extern "C" int mvster_conv1x1_wgrad(const float* x, const float* g, float* dw, float* db, int N, int C, long long S,
                                    void* workspace, void* stream) {
    if (!x || !g || !dw || !workspace) return fail(MVSTER_ERR_BAD_ARG, "conv1x1_wgrad: null pointer");
    if (N <= 0 || S <= 0) return fail(MVSTER_ERR_BAD_ARG, "conv1x1_wgrad: non-positive dimension");
    if (C != 8) return fail(MVSTER_ERR_UNSUPPORTED, "conv1x1_wgrad: C=%d (built: 8 input channels, 1 output channel)", C);
    if (S % 4 || ((uintptr_t)x) % 16 || ((uintptr_t)g) % 16)
        return fail(MVSTER_ERR_ALIGN, "conv1x1_wgrad: 16-byte aligned tensors with a plane size %% 4 == 0");
    const long long nchunk = (S + kPwChunk - 1) / kPwChunk;
    if (N > 65535 || nchunk * N > 2147483647LL) return fail(MVSTER_ERR_UNSUPPORTED, "conv1x1_wgrad: grid too large");
    DeviceGuard guard(dw);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    double* partial = static_cast<double*>(workspace);
    pointwise_wgrad_kernel<8><<<dim3((unsigned)nchunk, N), kPwThreads, 0, s>>>(x, g, partial, S);
    count_launch();
    pointwise_wgrad_reduce_kernel<<<1, 32, 0, s>>>(partial, dw, db, (int)(nchunk * N), C);
    count_launch();
    MVSTER_CHECK_LAUNCH("conv1x1_wgrad launch");
    return MVSTER_OK;
}
