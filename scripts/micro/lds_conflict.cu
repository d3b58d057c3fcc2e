// Micro-benchmark: how does sm_100 split an LDS.128 into wavefronts?  Each pattern is one kernel (one warp, many
// repetitions) so that ncu's l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld / inst gives wavefronts per LDS.128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_conflict lds_conflict.cu
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum ./lds_conflict
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(const int* __restrict__ offs, float* out, int reps) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    const int o = offs[threadIdx.x];  // in 16-byte units; negative = lane inactive
    float acc = 0.f;
    if (o >= 0) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(sm) + o * 16;
        for (int r = 0; r < reps; ++r) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
            acc += v.x + v.y + v.z + v.w;
        }
    }
    out[threadIdx.x] = acc;
}

int main() {
    int* d_offs; float* d_out;
    cudaMalloc(&d_offs, 32 * 4); cudaMalloc(&d_out, 32 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    auto sw = [](int texel, int chunk) { int g = 2 * texel + chunk; return g ^ ((g >> 3) & 1); };  // 32B swizzle, 16B units
    struct Pat { const char* name; int o[32]; } pats[32];
    int np = 0;
    auto add = [&](const char* n, auto f) { pats[np].name = n; for (int l = 0; l < 32; ++l) pats[np].o[l] = f(l); ++np; };
    add("P0 consecutive texels, swizzled chunk0 (expect 4)", [&](int l) { return sw(l, 0); });
    add("P1 all lanes same address (broadcast)", [&](int l) { return 0; });
    add("P2 lane l -> 16B unit 8*l (all same bank group, 32 distinct)", [&](int l) { return 8 * l; });
    add("P3 quarter q reads units q*8+[0..7] rotated: same banks across quarters", [&](int l) { return (l / 8) * 64 + (l % 8); });
    add("P4 lanes 0-7 conflict-free; lanes 8-15 same as 0-7 +8 units; 16-31 inactive", [&](int l) { return l < 16 ? (l % 8) + (l / 8) * 8 : -1; });
    add("P5 lanes 0-3 unit l, lanes 4-7 unit l-4+8 (2-way inside a quarter), rest inactive", [&](int l) { return l < 4 ? l : (l < 8 ? l - 4 + 8 : -1); });
    add("P6 lanes 0-3 of every quarter active, conflict-free within 16 lanes", [&](int l) { return (l % 8) < 4 ? (l / 8) * 4 + (l % 8) : -1; });
    add("P7 only lanes 0,8,16,24 active, distinct banks", [&](int l) { return (l % 8) == 0 ? l / 8 : -1; });
    add("P8 only lanes 0,8,16,24 active, same bank different address", [&](int l) { return (l % 8) == 0 ? (l / 8) * 8 : -1; });
    add("P9 texels with stride 9/8 (lane 7 skips one): swizzled chunk0", [&](int l) { return sw(l + (l % 8 == 7 ? 1 : 0) + (l / 8), 0); });
    add("P10 lanes 0-7: texel l chunk0 but lane 7 in next row (+48 texels)", [&](int l) { int q = l / 8, i = l % 8; return sw(q * 8 + i + (i == 7 ? 48 : 0), 0); });
    add("P11 lanes 0-7: two lanes share a texel, one skipped twice", [&](int l) { int q = l / 8, i = l % 8; int t[8] = {0, 1, 1, 2, 3, 5, 6, 8}; return sw(q * 16 + t[i], 0); });
    add("P12 even lanes unit l/2, odd lanes unit 8+l/2 within 16 lanes: 2 addr per bank group per quarter?", [&](int l) { return (l & 1) ? 8 + (l % 16) / 2 : (l % 16) / 2; });
    add("P13 one duplicate per quarter {0,1,2,3,3,4,5,6}, swizzled chunk0", [&](int l) { int q = l / 8, i = l % 8; int t[8] = {0, 1, 2, 3, 3, 4, 5, 6}; return sw(q * 8 + t[i], 0); });
    add("P14 duplicate at start {0,0,1,2,3,4,5,6}", [&](int l) { int q = l / 8, i = l % 8; int t[8] = {0, 0, 1, 2, 3, 4, 5, 6}; return sw(q * 8 + t[i], 0); });
    add("P15 two duplicates {0,0,1,2,3,3,4,5}", [&](int l) { int q = l / 8, i = l % 8; int t[8] = {0, 0, 1, 2, 3, 3, 4, 5}; return sw(q * 8 + t[i], 0); });
    add("P16 warp-continuous texels with scale 0.9: t = floor(0.9*l)", [&](int l) { return sw((int)(0.9f * l), 0); });
    add("P17 warp-continuous scale 0.9 offset 3: t = 3 + floor(0.9*l+0.5)", [&](int l) { return sw(3 + (int)(0.9f * l + 0.5f), 0); });
    add("P18 quarter duplicates unswizzled 16B units {0,1,2,3,3,4,5,6}", [&](int l) { int q = l / 8, i = l % 8; int t[8] = {0, 1, 2, 3, 3, 4, 5, 6}; return q * 8 + t[i]; });
    add("P19 pairs of lanes share a unit: {0,0,1,1,2,2,3,3}", [&](int l) { return l / 2; });
    for (int p = 0; p < np; ++p) {
        cudaMemcpy(d_offs, pats[p].o, 128, cudaMemcpyHostToDevice);
        probe<<<1, 32, 65536>>>(d_offs, d_out, 1000);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%s : %s\n", pats[p].name, cudaGetErrorString(e));
    }
    return 0;
}
