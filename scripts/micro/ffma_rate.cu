// FP32 FMA issue rate on sm_100a: scalar FFMA with three register operands, FFMA with a constant-bank weight, and
// packed FFMA2 (fma.rn.f32x2).  Prints FMA/clk/SM for each; 16 independent accumulators per thread, 8 warps per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096, ACC = 16;

__global__ void __launch_bounds__(1024) k_scalar(float* out, float a, float b) {
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float w0 = a + threadIdx.x, w1 = b + threadIdx.x;   // per-thread registers: 3-register FFMA
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) acc[i] = fmaf(acc[i], w0, w1);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(1024) k_const(float* out, float a, float b) {
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float x = 1.0f + threadIdx.x * 1e-6f;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) acc[i] = fmaf(x, a, acc[i]);   // a: kernel parameter -> constant bank operand
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}

__global__ void __launch_bounds__(1024) k_packed(float* out, float a, float b) {
    unsigned long long acc[ACC / 2];
#pragma unroll
    for (int i = 0; i < ACC / 2; ++i) {
        float lo = threadIdx.x * 0.001f + i, hi = lo + 0.5f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(lo), "f"(hi));
    }
    unsigned long long w0, w1;
    {
        float x = a + threadIdx.x, y = b + threadIdx.x;
        asm("mov.b64 %0, {%1, %1};" : "=l"(w0) : "f"(x));
        asm("mov.b64 %0, {%1, %1};" : "=l"(w1) : "f"(y));
    }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < ACC / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(w0), "l"(w1));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC / 2; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the convolution pattern: acc[co] += w[co] * in with w read from shared memory (broadcast) - scalar and packed
__global__ void __launch_bounds__(256) k_conv_scalar(float* out, const float* wg) {
    __shared__ float4 ws[64];
    if (threadIdx.x < 64) ws[threadIdx.x] = reinterpret_cast<const float4*>(wg)[threadIdx.x];
    __syncthreads();
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[p][i] = 0.f;
    float in[4] = {threadIdx.x * 1e-3f, 1.f, 2.f, 3.f};
    for (int it = 0; it < ITER / 4; ++it) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 w4 = ws[(it * 4 + t * 4 + k) & 63];
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc[p][4 * k + e] = fmaf(wv[e], in[p], acc[p][4 * k + e]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[p][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_conv_packed(float* out, const float* wg) {
    __shared__ float4 ws[64];
    if (threadIdx.x < 64) ws[threadIdx.x] = reinterpret_cast<const float4*>(wg)[threadIdx.x];
    __syncthreads();
    unsigned long long acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[p][i] = 0ull;
    unsigned long long in2[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float v = threadIdx.x * 1e-3f + p;
        asm("mov.b64 %0, {%1, %1};" : "=l"(in2[p]) : "f"(v));
    }
    for (int it = 0; it < ITER / 4; ++it) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const ulonglong2 w2 = reinterpret_cast<const ulonglong2*>(ws)[(it * 4 + t * 4 + k) & 63];   // {w0,w1},{w2,w3}
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[p][2 * k]) : "l"(w2.x), "l"(in2[p]));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[p][2 * k + 1]) : "l"(w2.y), "l"(in2[p]));
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[p][i]));
            s += lo + hi;
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount;
    float *out, *wg;
    cudaMalloc(&out, sizeof(float) * sms * 2 * 1024);
    cudaMalloc(&wg, 1024);
    cudaMemset(wg, 0, 1024);
    const double clk = clk_khz * 1e3;
    {
        const float ms = timeit([&] { k_scalar<<<sms * 2, 1024>>>(out, 1.0f, 0.5f); });
        printf("scalar FFMA, 3 registers     : %.1f FMA/clk/SM\n", (double)sms * 2 * 1024 * ITER * ACC / (ms * 1e-3) / clk / sms);
    }
    {
        const float ms = timeit([&] { k_const<<<sms * 2, 1024>>>(out, 1.0f, 0.5f); });
        printf("scalar FFMA, constant operand: %.1f FMA/clk/SM\n", (double)sms * 2 * 1024 * ITER * ACC / (ms * 1e-3) / clk / sms);
    }
    {
        const float ms = timeit([&] { k_packed<<<sms * 2, 1024>>>(out, 1.0f, 0.5f); });
        printf("packed FFMA2                 : %.1f FMA/clk/SM\n", (double)sms * 2 * 1024 * ITER * ACC * 2 / (ms * 1e-3) / clk / sms);
    }
    {
        const float ms = timeit([&] { k_conv_scalar<<<sms * 6, 256>>>(out, wg); });
        printf("conv pattern, scalar (LDS.128 weights, 4 px x 16 co): %.1f FMA/clk/SM\n", (double)sms * 6 * 256 * (ITER / 4) * 4 * 64 / (ms * 1e-3) / clk / sms);
    }
    {
        const float ms = timeit([&] { k_conv_packed<<<sms * 6, 256>>>(out, wg); });
        printf("conv pattern, packed FFMA2                          : %.1f FMA/clk/SM\n", (double)sms * 6 * 256 * (ITER / 4) * 4 * 64 / (ms * 1e-3) / clk / sms);
    }
    printf("(clock %d MHz from the device attribute, %d SMs)\n", clk_khz / 1000, sms);
    return 0;
}
