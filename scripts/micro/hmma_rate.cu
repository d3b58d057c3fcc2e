// Legacy tensor-core path on sm_100a: issue rate of mma.sync.m16n8k8 TF32 (fp32 accumulate), independent accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int NACC>
__global__ void __launch_bounds__(256) k_hmma(float* out, uint32_t a0, uint32_t b0) {
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
    uint32_t a[4] = {a0 + threadIdx.x, a0, a0 + 1, a0 + 2}, b[2] = {b0, b0 + threadIdx.x};
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
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
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const double clk = clk_khz * 1e3;
    for (int wps = 1; wps <= 4; wps *= 2) {   // warps per scheduler
        const int ctas = sms * wps * 4 * 32 / 256;
        const float ms = timeit([&] { k_hmma<8><<<ctas, 256>>>(out, 0x3f800000u, 0x3f000000u); });
        const double mmas = (double)ctas * 8 * ITER * 8;   // warp-level instructions
        printf("%d warp(s) per scheduler: %.3f HMMA.1688.TF32 per clk per SM = %.0f FMA/clk/SM (x1/3 for the 3xTF32 split: %.0f)\n", wps,
               mmas / (ms * 1e-3) / clk / sms, mmas * 1024 / (ms * 1e-3) / clk / sms, mmas * 1024 / 3 / (ms * 1e-3) / clk / sms);
    }
    return 0;
}
