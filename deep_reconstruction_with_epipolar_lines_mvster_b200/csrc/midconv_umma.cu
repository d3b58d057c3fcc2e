// 32/64-channel stride-1 convolutions (reg2d conv4 / conv6: Conv3d (3,3,3); FPN4 conv3.1 / conv3.2 / out2: 3x3 with
// D = 1; models/mvs4net_utils.py:456-468,896-903, eval mode, BatchNorm folded) as an implicit GEMM on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32, operands in shared memory, accumulators in tensor memory.
//
// fp32-grade accuracy from TF32 hardware: x = hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi); a product is
// accumulated as lo*hi + hi*lo + hi*hi (3xTF32, the dropped lo*lo term is 2^-22 relative).  The tensor core adds into
// its fp32 accumulator with truncation, so the K chain is spread round-robin over four accumulators in tensor memory
// (each sees a quarter of the adds) and they are summed with rounded FADDs in the epilogue.
//
// GEMM view per CTA: M = 128 consecutive pixels of one image row of one (batch, depth) plane, N = Cout, K walks
// (kd, 8-channel chunk, tap).  Implicit im2col without copies: the staged input window is pixel-major with four
// channels per 16-byte row, [hi|lo][k-half][3 rows][136 pixels][4 channels], which IS the canonical K-major no-swizzle
// UMMA layout with 8-row core matrices 128 bytes apart - so the A operand of tap (ky, kx) is the same buffer with the
// descriptor's start address moved by (ky * 136 + kx) * 16 bytes.  Weights are packed once per layer
// (mvster_umma_pack_weights) into per-chunk blocks [hi|lo][tap][k-half][co][4 ci] and arrive by one bulk copy per chunk.
// Two stages: while the tensor core works on chunk k (27 MMAs, asynchronous, tracked by tcgen05.commit -> mbarrier),
// the 128 threads stage and split chunk k+1.
#include "epi_tma.cuh"

namespace mvster {

constexpr int kUmCols = 136;  // staged pixels per row (128 + 2 halo, padded)
constexpr int kUmAcc = 4;     // accumulators in tensor memory (round-robin over chunks)

struct UmmaConvParams {
    const float* x;
    const float* wt;    // packed: [KD][CIN/8][hi|lo][tap 9][k-half 2][CO][4]
    const float* bias;  // [CO]
    float* y;
    int B, D, H, W, relu;
};

template <int CO>
struct UmmaCfg {
    static constexpr int XF = 2 * 2 * 3 * kUmCols * 4;  // floats of one input stage (hi and lo)
    static constexpr int WF = 2 * 9 * 2 * CO * 4;       // floats of one weight stage (hi and lo)
    static constexpr int BAR_OFF = 2 * (XF + WF) * 4;
    static constexpr int SMEM = BAR_OFF + 64;
    static constexpr int TCOLS = kUmAcc * CO;           // 256 / 128: a power of two >= 32
};

__device__ __forceinline__ uint32_t rna_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// bounded wait: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_units, uint32_t sbo_units) {
    // K-major, no swizzle: 8-row core matrices of 16-byte rows; LBO = distance between the two 16-byte k-halves,
    // SBO = distance between 8-row groups, both in 16-byte units; descriptor version 1 (Blackwell)
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo_units & 0x3FFFu) << 16) |
           ((uint64_t)(sbo_units & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

template <int KD, int CIN, int CO>
__global__ void __launch_bounds__(128) midconv_umma_kernel(const UmmaConvParams p) {
    using K = UmmaCfg<CO>;
    extern __shared__ __align__(128) unsigned char smem_um[];
    float* Xs = reinterpret_cast<float*>(smem_um);            // [stage][hi|lo][k-half][row][col][4]
    float* Ws = Xs + 2 * K::XF;                               // [stage][hi|lo][tap][k-half][co][4]
    const uint32_t sbase = smem_u32(smem_um);
    const uint32_t bars = sbase + K::BAR_OFF;                 // wfull[2], mdone[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_um + K::BAR_OFF + 32);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * 128, y = blockIdx.y;
    const int b = blockIdx.z / p.D, d = blockIdx.z % p.D;
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(K::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CO >> 3) << 17) | ((128u >> 4) << 24);
    constexpr int NCH = CIN / 8;

    int k = 0;  // chunks processed so far
#pragma unroll 1
    for (int kd = 0; kd < KD; ++kd) {
        const int dz = d + kd - KD / 2;  // CTA-uniform
        if ((unsigned)dz >= (unsigned)p.D) continue;
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch, ++k) {
            const int s = k & 1, use = k >> 1;
            float* Xst = Xs + s * K::XF;
            float* Wst = Ws + s * K::WF;
            // ---- this chunk's input window -> registers first (all loads in flight at once; the tensor core is still
            //      busy with the previous chunk): item = (k-half, row, pixel), four channels each -------------------
            constexpr int NIT = (2 * 3 * 130 + 127) / 128;
            const float* xb = p.x + (((size_t)b * CIN + ch * 8) * p.D + dz) * plane;
            float v[NIT][4];
#pragma unroll
            for (int j = 0; j < NIT; ++j) {
                const int it = tid + j * 128;
                const int kh = it / 390, rem = it - kh * 390;
                const int r = rem / 130, c = rem - r * 130;
                const int gy = y - 1 + r, gx = x0 - 1 + c;
                const bool ok = it < 780 && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                const float* q = xb + (size_t)(kh * 4) * p.D * plane + (size_t)gy * W + gx;
#pragma unroll
                for (int e = 0; e < 4; ++e) v[j][e] = ok ? __ldg(q + (size_t)e * p.D * plane) : 0.0f;
            }
            // the tensor core is done with this stage's previous contents (chunk k - 2)
            if (k >= 2) mbar_wait_bounded(bars + 16u + 8u * s, (uint32_t)(use - 1) & 1u);
            if (tid == 0) {
                const uint32_t wf = bars + 8u * s;
                mbar_expect_tx(wf, (uint32_t)K::WF * 4u);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(Wst)), "l"(p.wt + ((size_t)kd * NCH + ch) * K::WF), "r"((uint32_t)K::WF * 4u), "r"(wf)
                             : "memory");
            }
            // ---- split and store: 16 bytes hi + 16 bytes lo per item ---------------------------------------------------
#pragma unroll
            for (int j = 0; j < NIT; ++j) {
                const int it = tid + j * 128;
                if (it < 780) {
                    const int kh = it / 390, rem = it - kh * 390;
                    const int r = rem / 130, c = rem - r * 130;
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        hi[e] = rna_tf32(v[j][e]);
                        lo[e] = rna_tf32(v[j][e] - __uint_as_float(hi[e]));
                    }
                    const int o = ((kh * 3 + r) * kUmCols + c) * 4;
                    *reinterpret_cast<uint4*>(Xst + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(Xst + K::XF / 2 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            __syncthreads();
            if (tid == 0) {
                mbar_wait_bounded(bars + 8u * s, (uint32_t)use & 1u);      // this chunk's weights have landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t xa = smem_u32(Xst), wa = smem_u32(Wst);
                const uint32_t acc = tmem + (uint32_t)((k % kUmAcc) * CO);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap - ky * 3;
                    const uint32_t a_hi = xa + (uint32_t)((ky * kUmCols + kx) * 16);
                    const uint32_t a_lo = a_hi + (uint32_t)(K::XF / 2) * 4u;
                    const uint32_t b_hi = wa + (uint32_t)(tap * 2 * CO * 16);
                    const uint32_t b_lo = b_hi + (uint32_t)(K::WF / 2) * 4u;
                    const uint64_t dah = umma_desc(a_hi, 3 * kUmCols, 8), dal = umma_desc(a_lo, 3 * kUmCols, 8);
                    const uint64_t dbh = umma_desc(b_hi, CO, 8), dbl = umma_desc(b_lo, CO, 8);
                    umma_tf32(acc, dal, dbh, IDESC, (k >= kUmAcc || tap > 0) ? 1u : 0u);  // small terms first
                    umma_tf32(acc, dah, dbl, IDESC, 1u);
                    umma_tf32(acc, dah, dbh, IDESC, 1u);
                }
                // arrives on mdone[s] when every MMA issued so far has completed (implies fence::before_thread_sync)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 16u + 8u * s) : "memory");
            }
        }
    }

    // ---- epilogue: wait for the last commit (it covers every MMA), sum the accumulators, bias, ReLU, planar store -----
    if (k > 0) mbar_wait_bounded(bars + 16u + 8u * ((k - 1) & 1), (uint32_t)((k - 1) >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int nacc = k < kUmAcc ? k : kUmAcc;
    const int gx = x0 + warp * 32 + lane;
    float* yb = p.y + (((size_t)b * CO) * p.D + d) * plane + (size_t)y * W + gx;
    const size_t cstride = (size_t)p.D * plane;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes of tensor memory
#pragma unroll 1
    for (int j = 0; j < CO; j += 8) {
        float sum[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) sum[e] = 0.0f;
        for (int a = 0; a < nacc; ++a) {   // warp-uniform trip count
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(trow + (uint32_t)(a * CO + j)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 8; ++e) sum[e] += __uint_as_float(r[e]);
        }
        if (gx < W) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = sum[e] + __ldg(p.bias + j + e);
                if (p.relu) v = fmaxf(v, 0.0f);
                yb[(size_t)(j + e) * cstride] = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(K::TCOLS) : "memory");
}

// w [KD][3][3][CIN][CO] fp32 -> packed [KD][CIN/8][hi|lo][tap][k-half][CO][4]
__global__ void umma_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int KD, int CIN, int CO) {
    const size_t n = (size_t)KD * 9 * CIN * CO;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // decode the destination index of the hi half
    const int e = (int)(i & 3);
    size_t r = i >> 2;
    const int co = (int)(r % CO); r /= CO;
    const int kh = (int)(r & 1); r >>= 1;
    const int tap = (int)(r % 9); r /= 9;
    const int ch = (int)(r % (CIN / 8));
    const int kd = (int)(r / (CIN / 8));
    const int ci = ch * 8 + kh * 4 + e;
    const float v = w[((size_t)(kd * 9 + tap) * CIN + ci) * CO + co];
    const uint32_t hi = rna_tf32(v);
    const uint32_t lo = rna_tf32(v - __uint_as_float(hi));
    const size_t half = (size_t)9 * 2 * CO * 4;
    const size_t blk = ((size_t)kd * (CIN / 8) + ch) * 2 * half;
    const size_t off = (((size_t)tap * 2 + kh) * CO + co) * 4 + e;
    out[blk + off] = __uint_as_float(hi);
    out[blk + half + off] = __uint_as_float(lo);
}

template <int KD, int CIN, int CO>
static int launch_umma(const UmmaConvParams& p, cudaStream_t s) {
    using K = UmmaCfg<CO>;
    static int attr_done[64] = {};
    const int st = ensure_dynamic_smem_bytes(midconv_umma_kernel<KD, CIN, CO>, K::SMEM, attr_done, "conv3d_mid_umma: cudaFuncSetAttribute");
    if (st != MVSTER_OK) return st;
    dim3 grid((p.W + 127) / 128, p.H, p.B * p.D);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_mid_umma: grid too large");
    midconv_umma_kernel<KD, CIN, CO><<<grid, 128, K::SMEM, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("conv3d_mid_umma launch");
    return MVSTER_OK;
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_umma_pack_weights(const float* w, float* packed, int kd, int Cin, int Cout, void* stream) {
    if (!w || !packed) return fail(MVSTER_ERR_BAD_ARG, "umma_pack_weights: null pointer");
    if ((kd != 1 && kd != 3) || Cin <= 0 || Cin % 8 || Cout <= 0 || Cout % 8)
        return fail(MVSTER_ERR_BAD_ARG, "umma_pack_weights: kd in {1,3}, Cin and Cout multiples of 8");
    DeviceGuard guard(packed);
    if (guard.status != MVSTER_OK) return guard.status;
    const size_t n = (size_t)kd * 9 * Cin * Cout;
    umma_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, packed, kd, Cin, Cout);
    count_launch();
    MVSTER_CHECK_LAUNCH("umma_pack_weights launch");
    return MVSTER_OK;
}

extern "C" int mvster_conv3d_mid_umma(const float* x, const float* w_packed, const float* bias_dev, float* y, int B, int Cin,
                                      int Cout, int D, int H, int W, int kd, int relu, void* stream) {
    if (!x || !w_packed || !bias_dev || !y) return fail(MVSTER_ERR_BAD_ARG, "conv3d_mid_umma: null pointer");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "conv3d_mid_umma: non-positive dimension");
    if (((uintptr_t)w_packed) % 16) return fail(MVSTER_ERR_ALIGN, "conv3d_mid_umma: packed weights must be 16-byte aligned");
    DeviceGuard guard(y);
    if (guard.status != MVSTER_OK) return guard.status;
    UmmaConvParams p{x, w_packed, bias_dev, y, B, D, H, W, relu};
    cudaStream_t s = (cudaStream_t)stream;
#define MVSTER_UM_CASE(KD_, CI_, CO_) \
    if (kd == KD_ && Cin == CI_ && Cout == CO_) return launch_umma<KD_, CI_, CO_>(p, s);
    MVSTER_UM_CASE(1, 32, 32) MVSTER_UM_CASE(1, 64, 64) MVSTER_UM_CASE(1, 64, 32)
    MVSTER_UM_CASE(3, 32, 32) MVSTER_UM_CASE(3, 64, 64)
#undef MVSTER_UM_CASE
    return fail(MVSTER_ERR_UNSUPPORTED, "conv3d_mid_umma: no kernel for Cin=%d Cout=%d kd=%d", Cin, Cout, kd);
}
