// Layout helper: NCHW fp32 -> NHWC fp32 / bf16.  The reference FPN emits NCHW-contiguous feature maps
// (models/mvs4net_utils.py:504-507); the fused kernels read NHWC so that one bilinear tap is one contiguous
// channel vector.  A caller that already holds channels_last tensors never reaches this kernel.
#include "common.cuh"

namespace mvster {

template <typename OutT>
__device__ __forceinline__ OutT cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// one CTA transposes a [C, 32-pixel] slab through shared memory: coalesced 128-byte reads along pixels for every
// channel, contiguous 32*C-element writes in the pixel-major output
template <typename OutT>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, OutT* __restrict__ out, int C,
                                                           size_t plane) {
    extern __shared__ float tile[];  // [C][33]
    const size_t pix0 = (size_t)blockIdx.x * 32;
    const int b = blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int npix = (int)min((size_t)32, plane - pix0);
    const float* src = in + (size_t)b * C * plane + pix0;
    for (int c = ty; c < C; c += 8) tile[c * 33 + tx] = (tx < npix) ? ldg_stream(src + (size_t)c * plane + tx) : 0.f;
    __syncthreads();
    OutT* dst = out + ((size_t)b * plane + pix0) * C;
    const int n = npix * C;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int pixel = i / C, c = i - pixel * C;
        dst[i] = cvt_out<OutT>(tile[c * 33 + pixel]);
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_nchw_to_nhwc(const float* in, void* out, int B, int C, int H, int W, int out_dtype,
                                   void* stream) {
    if (!in || !out) return fail(MVSTER_ERR_BAD_ARG, "nchw_to_nhwc: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "nchw_to_nhwc: non-positive dimension");
    if (B > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "nchw_to_nhwc: B > 65535");
    if (C > 1024) return fail(MVSTER_ERR_UNSUPPORTED, "nchw_to_nhwc: C > 1024");
    DeviceGuard guard(out);
    if (guard.status != MVSTER_OK) return guard.status;
    const size_t plane = (size_t)H * W;
    dim3 grid((unsigned)((plane + 31) / 32), B);
    const size_t smem = (size_t)C * 33 * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == MVSTER_F32) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(nchw_to_nhwc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nchw_to_nhwc_kernel<float><<<grid, 256, smem, s>>>(in, (float*)out, C, plane);
    } else if (out_dtype == MVSTER_BF16) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(nchw_to_nhwc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem);
        nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(in, (__nv_bfloat16*)out, C, plane);
    } else {
        return fail(MVSTER_ERR_BAD_ARG, "nchw_to_nhwc: unknown dtype %d", out_dtype);
    }
    count_launch();
    MVSTER_CHECK_LAUNCH("nchw_to_nhwc launch");
    return MVSTER_OK;
}
