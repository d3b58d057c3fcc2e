// Training-mode BatchNorm (+ ReLU) over planar [N, C, S] fp32 activations: what nn.BatchNorm3d / nn.BatchNorm2d in
// train() followed by ReLU compute inside the reference's ConvBnReLU3D / Deconv3d / Conv2d blocks
// (models/mvs4net_utils.py:123-130, 231-258, 884-926), forward and backward.
//
// Why hand-written: the regulariser's activations have 8..64 channels and up to 1.3 M elements per (sample, channel)
// plane.  cuDNN's planar kernels (bn_fw_tr_1C11_kernel_NCHW, bn_bw_1C11_kernel_new) parallelise over channels and took
// 5.3 + 15.5 ms of a 116 ms training step at 512x640, B=2 (profiles/r01_train_step_torch_profiler.txt); the work is
// two streaming passes per direction, i.e. bandwidth:
//   forward : stats pass (read x) -> finalize (C threads) -> apply pass (read x, write y = relu(a x + b))
//   backward: stats pass (read x, y, dy: s1 = sum g, s2 = sum g xhat, g = dy [y > 0]) -> finalize (dgamma = s2, dbeta = s1)
//             -> apply pass (read x, y, dy, write dx = a (g - s1/M - xhat s2/M))
// Both passes run on a (chunk of the plane, plane = n * C + c) grid - thousands of CTAs whatever C is - with 16-byte
// accesses and no index arithmetic beyond one multiply; partial sums are kept per (plane, chunk) in double and summed
// in a fixed order by the finalize kernel: bit-reproducible, no atomics.
// Statistics follow torch: biased variance for the normalisation, unbiased for running_var, momentum update of the
// running statistics, invstd = 1 / sqrt(var + eps).
#include "common.cuh"

namespace mvster {

constexpr int kBnThreads = 256;
constexpr int kBnVecPerThread = 8;                                // float4 per thread and chunk
constexpr int kBnChunk = kBnThreads * kBnVecPerThread * 4;        // elements of a plane per CTA (8192)

__device__ __forceinline__ double bn_block_sum(double v, double* red) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kBnThreads / 32; ++w) t += red[w];
    return t;  // valid in thread 0
}

// partial[(plane * nchunk + chunk) * 2 + {0,1}] = (sum x, sum x^2) of the chunk
__global__ void __launch_bounds__(kBnThreads) bn_fwd_stats_kernel(const float* __restrict__ x, double* __restrict__ partial,
                                                                  long long S) {
    __shared__ double red[kBnThreads / 32];
    const long long plane = blockIdx.y;
    const long long s0 = (long long)blockIdx.x * kBnChunk;
    const float4* xp = reinterpret_cast<const float4*>(x + plane * S + s0);
    const long long nvec = min((long long)kBnChunk, S - s0) >> 2;
    float s = 0.0f, q = 0.0f;
#pragma unroll
    for (int k = 0; k < kBnVecPerThread; ++k) {
        const long long i = (long long)k * kBnThreads + threadIdx.x;
        if (i < nvec) {
            const float4 v = __ldg(xp + i);
            s += (v.x + v.y) + (v.z + v.w);
            q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
        }
    }
    const double ts = bn_block_sum((double)s, red);
    const double tq = bn_block_sum((double)q, red);
    if (threadIdx.x == 0) {
        double* o = partial + (plane * gridDim.x + blockIdx.x) * 2;
        o[0] = ts;
        o[1] = tq;
    }
}

// one thread per channel: fixed-order sums over (n, chunk); mean / invstd, affine coefficients, running statistics
__global__ void bn_fwd_finalize_kernel(const double* __restrict__ partial, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ mean_out,
                                       float* __restrict__ invstd_out, float* __restrict__ coef /* [C][2]: a, b */,
                                       float* running_mean, float* running_var, float momentum, float eps, int N, int C,
                                       int nchunk, double count) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int n = 0; n < N; ++n) {
        const double* p = partial + ((size_t)(n * C + c) * nchunk) * 2;
        for (int k = 0; k < nchunk; ++k) { s += p[2 * k]; q += p[2 * k + 1]; }
    }
    const double mean = s / count;
    const double var = fmax(q / count - mean * mean, 0.0);  // biased
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    const float a = (gamma ? gamma[c] : 1.0f) * invstd;
    coef[2 * c] = a;
    coef[2 * c + 1] = (beta ? beta[c] : 0.0f) - (float)mean * a;
    if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

template <bool RELU>
__global__ void __launch_bounds__(kBnThreads) bn_fwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ coef,
                                                                  float* __restrict__ y, long long S, int C) {
    const long long plane = blockIdx.y;
    const int c = (int)(plane % C);
    const float a = coef[2 * c], b = coef[2 * c + 1];
    const long long s0 = (long long)blockIdx.x * kBnChunk;
    const float4* xp = reinterpret_cast<const float4*>(x + plane * S + s0);
    float4* yp = reinterpret_cast<float4*>(y + plane * S + s0);
    const long long nvec = min((long long)kBnChunk, S - s0) >> 2;
#pragma unroll
    for (int k = 0; k < kBnVecPerThread; ++k) {
        const long long i = (long long)k * kBnThreads + threadIdx.x;
        if (i < nvec) {
            const float4 v = __ldg(xp + i);
            float4 o = make_float4(fmaf(v.x, a, b), fmaf(v.y, a, b), fmaf(v.z, a, b), fmaf(v.w, a, b));
            if (RELU) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
            yp[i] = o;
        }
    }
}

// partial = (sum g, sum g * xhat), g = dy masked by the ReLU
template <bool RELU>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                  const float* __restrict__ dy, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, double* __restrict__ partial,
                                                                  long long S, int C) {
    __shared__ double red[kBnThreads / 32];
    const long long plane = blockIdx.y;
    const int c = (int)(plane % C);
    const float mu = mean[c], is = invstd[c];
    const long long s0 = (long long)blockIdx.x * kBnChunk;
    const float4* xp = reinterpret_cast<const float4*>(x + plane * S + s0);
    const float4* yp = reinterpret_cast<const float4*>(y + plane * S + s0);
    const float4* gp = reinterpret_cast<const float4*>(dy + plane * S + s0);
    const long long nvec = min((long long)kBnChunk, S - s0) >> 2;
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < kBnVecPerThread; ++k) {
        const long long i = (long long)k * kBnThreads + threadIdx.x;
        if (i < nvec) {
            const float4 v = __ldg(xp + i);
            float4 g = __ldg(gp + i);
            if (RELU) {
                const float4 o = __ldg(yp + i);
                g.x = o.x > 0.0f ? g.x : 0.0f; g.y = o.y > 0.0f ? g.y : 0.0f;
                g.z = o.z > 0.0f ? g.z : 0.0f; g.w = o.w > 0.0f ? g.w : 0.0f;
            }
            s1 += (g.x + g.y) + (g.z + g.w);
            s2 = fmaf(g.x, (v.x - mu) * is, fmaf(g.y, (v.y - mu) * is, fmaf(g.z, (v.z - mu) * is, fmaf(g.w, (v.w - mu) * is, s2))));
        }
    }
    const double t1 = bn_block_sum((double)s1, red);
    const double t2 = bn_block_sum((double)s2, red);
    if (threadIdx.x == 0) {
        double* o = partial + (plane * gridDim.x + blockIdx.x) * 2;
        o[0] = t1;
        o[1] = t2;
    }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ partial, const float* __restrict__ gamma,
                                       const float* __restrict__ invstd, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ coef /* [C][3]: a, s1/M, s2/M */,
                                       int N, int C, int nchunk, double count) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s1 = 0.0, s2 = 0.0;
    for (int n = 0; n < N; ++n) {
        const double* p = partial + ((size_t)(n * C + c) * nchunk) * 2;
        for (int k = 0; k < nchunk; ++k) { s1 += p[2 * k]; s2 += p[2 * k + 1]; }
    }
    if (dgamma) dgamma[c] = (float)s2;
    if (dbeta) dbeta[c] = (float)s1;
    coef[3 * c] = (gamma ? gamma[c] : 1.0f) * invstd[c];
    coef[3 * c + 1] = (float)(s1 / count);
    coef[3 * c + 2] = (float)(s2 / count);
}

template <bool RELU>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                  const float* __restrict__ dy, const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, const float* __restrict__ coef,
                                                                  float* __restrict__ dx, long long S, int C) {
    const long long plane = blockIdx.y;
    const int c = (int)(plane % C);
    const float mu = mean[c], is = invstd[c];
    const float a = coef[3 * c], k1 = coef[3 * c + 1], k2 = coef[3 * c + 2];
    const long long s0 = (long long)blockIdx.x * kBnChunk;
    const float4* xp = reinterpret_cast<const float4*>(x + plane * S + s0);
    const float4* yp = reinterpret_cast<const float4*>(y + plane * S + s0);
    const float4* gp = reinterpret_cast<const float4*>(dy + plane * S + s0);
    float4* dp = reinterpret_cast<float4*>(dx + plane * S + s0);
    const long long nvec = min((long long)kBnChunk, S - s0) >> 2;
#pragma unroll
    for (int k = 0; k < kBnVecPerThread; ++k) {
        const long long i = (long long)k * kBnThreads + threadIdx.x;
        if (i < nvec) {
            const float4 v = __ldg(xp + i);
            float4 g = __ldg(gp + i);
            if (RELU) {
                const float4 o = __ldg(yp + i);
                g.x = o.x > 0.0f ? g.x : 0.0f; g.y = o.y > 0.0f ? g.y : 0.0f;
                g.z = o.z > 0.0f ? g.z : 0.0f; g.w = o.w > 0.0f ? g.w : 0.0f;
            }
            float4 o;
            o.x = a * (g.x - k1 - (v.x - mu) * is * k2);
            o.y = a * (g.y - k1 - (v.y - mu) * is * k2);
            o.z = a * (g.z - k1 - (v.z - mu) * is * k2);
            o.w = a * (g.w - k1 - (v.w - mu) * is * k2);
            dp[i] = o;
        }
    }
}

static int bn_check(const char* who, const void* a, const void* b, int N, int C, long long S) {
    if (N <= 0 || C <= 0 || S <= 0) return fail(MVSTER_ERR_BAD_ARG, "%s: non-positive dimension", who);
    if (S % 4) return fail(MVSTER_ERR_UNSUPPORTED, "%s: plane size must be a multiple of 4 (16-byte accesses)", who);
    if (((uintptr_t)a) % 16 || ((uintptr_t)b) % 16) return fail(MVSTER_ERR_ALIGN, "%s: tensors must be 16-byte aligned", who);
    if ((long long)N * C > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "%s: more than 65535 (sample, channel) planes", who);
    if ((S + kBnChunk - 1) / kBnChunk > 2147483647LL) return fail(MVSTER_ERR_UNSUPPORTED, "%s: plane too large", who);
    return MVSTER_OK;
}

}  // namespace mvster

using namespace mvster;

extern "C" long long mvster_bn_train_workspace_bytes(int N, int C, long long S) {
    if (N <= 0 || C <= 0 || S <= 0) return 0;
    const long long nchunk = (S + kBnChunk - 1) / kBnChunk;
    return (long long)N * C * nchunk * 2 * 8 + (long long)C * 3 * 4 + 64;
}

extern "C" int mvster_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                   float* invstd, float* running_mean, float* running_var, float momentum, float eps,
                                   int relu, int N, int C, long long S, void* workspace, void* stream) {
    if (!x || !y || !mean || !invstd || !workspace) return fail(MVSTER_ERR_BAD_ARG, "bn_train_fwd: null pointer");
    int st = bn_check("bn_train_fwd", x, y, N, C, S);
    if (st != MVSTER_OK) return st;
    if (((uintptr_t)workspace) % 8) return fail(MVSTER_ERR_ALIGN, "bn_train_fwd: workspace must be 8-byte aligned");
    DeviceGuard guard(y);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    const int nchunk = (int)((S + kBnChunk - 1) / kBnChunk);
    double* partial = static_cast<double*>(workspace);
    float* coef = reinterpret_cast<float*>(partial + (size_t)N * C * nchunk * 2);
    dim3 grid(nchunk, N * C);
    bn_fwd_stats_kernel<<<grid, kBnThreads, 0, s>>>(x, partial, S);
    count_launch();
    bn_fwd_finalize_kernel<<<(C + 63) / 64, 64, 0, s>>>(partial, gamma, beta, mean, invstd, coef, running_mean, running_var,
                                                         momentum, eps, N, C, nchunk, (double)N * (double)S);
    count_launch();
    if (relu) bn_fwd_apply_kernel<true><<<grid, kBnThreads, 0, s>>>(x, coef, y, S, C);
    else bn_fwd_apply_kernel<false><<<grid, kBnThreads, 0, s>>>(x, coef, y, S, C);
    count_launch();
    MVSTER_CHECK_LAUNCH("bn_train_fwd launch");
    return MVSTER_OK;
}

extern "C" int mvster_bn_train_bwd(const float* x, const float* y, const float* dy, const float* gamma,
                                   const float* mean, const float* invstd, float* dx, float* dgamma, float* dbeta,
                                   int relu, int N, int C, long long S, void* workspace, void* stream) {
    if (!x || !dy || !mean || !invstd || !dx || !workspace || (relu && !y))
        return fail(MVSTER_ERR_BAD_ARG, "bn_train_bwd: null pointer");
    int st = bn_check("bn_train_bwd", x, dx, N, C, S);
    if (st != MVSTER_OK) return st;
    if (((uintptr_t)dy) % 16 || (relu && ((uintptr_t)y) % 16)) return fail(MVSTER_ERR_ALIGN, "bn_train_bwd: tensors must be 16-byte aligned");
    if (((uintptr_t)workspace) % 8) return fail(MVSTER_ERR_ALIGN, "bn_train_bwd: workspace must be 8-byte aligned");
    DeviceGuard guard(dx);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    const int nchunk = (int)((S + kBnChunk - 1) / kBnChunk);
    double* partial = static_cast<double*>(workspace);
    float* coef = reinterpret_cast<float*>(partial + (size_t)N * C * nchunk * 2);
    dim3 grid(nchunk, N * C);
    const float* yy = relu ? y : x;
    if (relu) bn_bwd_stats_kernel<true><<<grid, kBnThreads, 0, s>>>(x, yy, dy, mean, invstd, partial, S, C);
    else bn_bwd_stats_kernel<false><<<grid, kBnThreads, 0, s>>>(x, yy, dy, mean, invstd, partial, S, C);
    count_launch();
    bn_bwd_finalize_kernel<<<(C + 63) / 64, 64, 0, s>>>(partial, gamma, invstd, dgamma, dbeta, coef, N, C, nchunk,
                                                         (double)N * (double)S);
    count_launch();
    if (relu) bn_bwd_apply_kernel<true><<<grid, kBnThreads, 0, s>>>(x, yy, dy, mean, invstd, coef, dx, S, C);
    else bn_bwd_apply_kernel<false><<<grid, kBnThreads, 0, s>>>(x, yy, dy, mean, invstd, coef, dx, S, C);
    count_launch();
    MVSTER_CHECK_LAUNCH("bn_train_bwd launch");
    return MVSTER_OK;
}
