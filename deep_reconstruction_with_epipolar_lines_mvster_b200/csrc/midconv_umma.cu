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
// channels per 16-byte row, [hi|lo][k-half][3 rows][132 pixels][4 channels], which IS the canonical K-major no-swizzle
// UMMA layout with 8-row core matrices 128 bytes apart - so the A operand of tap (ky, kx) is the same buffer with the
// descriptor's start address moved by (ky * 132 + kx) * 16 bytes.  Weights are packed once per layer
// (mvster_umma_pack_weights) into blocks [hi|lo][tap][k-half][co][4 ci] of one kernel row (Cout = 64) or all nine taps
// (Cout = 32) of a chunk; they arrive by bulk copies through a ring of 5 / 3 blocks - small enough for two CTAs per SM,
// so that one CTA's prologue and epilogue hide behind the other's MMAs.
// Warp-specialised pipeline: four warps stage and split the input windows two stages deep (the loads of chunk k+1 are
// in flight while chunk k is stored), one thread keeps a four-stage ring of weight blocks filled and issues the 27
// asynchronous MMAs of a chunk as soon as both operands are there; tcgen05.commit -> mbarrier hands stages back.
#include "epi_tma.cuh"

namespace mvster {

constexpr int kUmCols = 132;  // staged pixels per row (128 + 2 halo, padded)
constexpr int kUmAcc = 4;     // accumulators in tensor memory (round-robin over chunks)

struct UmmaConvParams {
    const float* x;
    const float* wt;    // packed: [KD][CIN/8][block][hi|lo][tap in block][k-half 2][CO][4]
    const float* bias;  // [CO]
    float* y;
    int B, D, H, W, relu;
};


template <int CO>
struct UmmaCfg {
    static constexpr int XF = 2 * 2 * 3 * kUmCols * 4;  // floats of one input stage (hi and lo)
    static constexpr int TPU = CO == 64 ? 3 : 9;        // taps per weight block: one kernel row (Cout = 64) or all nine (measured)
    static constexpr int UPC = 9 / TPU;                 // weight blocks per chunk
    static constexpr int WF = 2 * TPU * 2 * CO * 4;     // floats of one weight block, hi and lo
    static constexpr int WST = CO == 64 ? 5 : 3;        // ring of weight blocks: 110 KB / 105 KB per CTA, two CTAs per SM
    static constexpr int BAR_OFF = (2 * XF + WST * WF) * 4;
    static constexpr int SMEM = BAR_OFF + 192;
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

__device__ __forceinline__ void mbar_arrive_one(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// 160 threads: warps 0-3 stage the input windows (and run the epilogue), lane 0 of warp 4 feeds the weight ring and
// issues the MMAs.  Nothing but mbarriers between the two roles inside the chunk loop:
//   xfull[s]  (128 arrivals)  stagers  -> issuer   input stage s holds chunk k
//   xempty[s] (tcgen05.commit) issuer  -> stagers  every MMA up to chunk k has completed: input stage s and the weight
//                                                  stage of chunk k are free again
//   wfull[w]  (bulk-copy bytes)         -> issuer   weight stage w holds its chunk
// Issued from warp-uniform code: every lane of the issuing warp executes the statement, elect.sync picks the one lane
// whose predicate lets the instruction through.  (Inside an `if (lane == 0)` region the compiler wraps every UTCHMMA in
// an ELECT / BRA.U.ANY loop and rebuilds the descriptors with ~20 uniform-datapath instructions per MMA - the single
// issuing thread then cannot keep the tensor core fed.)
__device__ __forceinline__ void umma_tf32_elect(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t desc_hi,
                                                uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, pe;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, {%6, %6, %6, %6}, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar)
        : "memory");
}

template <int KD, int CIN, int CO>
__global__ void __launch_bounds__(160) midconv_umma_kernel(const UmmaConvParams p) {
    using K = UmmaCfg<CO>;
    extern __shared__ __align__(128) unsigned char smem_um[];
    float* Xs = reinterpret_cast<float*>(smem_um);            // [2][hi|lo][k-half][row][col][4]
    float* Ws = Xs + 2 * K::XF;                               // [kUmWst][hi|lo][tap][k-half][co][4]
    const uint32_t sbase = smem_u32(smem_um);
    const uint32_t xfull = sbase + K::BAR_OFF, xempty = xfull + 16, wfull = xfull + 32, wempty = wfull + 64;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_um + K::BAR_OFF + 160);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * 128, y = blockIdx.y;
    const int b = blockIdx.z / p.D, d = blockIdx.z % p.D;
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;
    constexpr int NCH = CIN / 8, kUmWst = K::WST;
    const int kd_lo = max(0, KD / 2 - d), kd_hi = min(KD - 1, KD / 2 + p.D - 1 - d);   // depth taps inside the volume
    const int total = (kd_hi - kd_lo + 1) * NCH;                                        // chunks of this tile (>= NCH)

    if (tid == 0) {
        mbar_init(xfull, 128); mbar_init(xfull + 8, 128);
        mbar_init(xempty, 1); mbar_init(xempty + 8, 1);
#pragma unroll
        for (int i = 0; i < kUmWst; ++i) { mbar_init(wfull + 8u * i, 1); mbar_init(wempty + 8u * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(K::TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        // ---- issuer warp: warp-uniform control flow, single-lane side effects ---------------------------------------------
        constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CO >> 3) << 17) | ((128u >> 4) << 24);
        constexpr int TPU = K::TPU, UPC = K::UPC;
        // K-major, no swizzle: LBO (16-byte units, bits 16-29) = distance of the two k-halves, SBO (bits 32-45) = 8 units
        // between 8-row groups, descriptor version 1 in bits 46-47.  Start addresses are 16-byte aligned and below
        // 256 KB, so a tap's descriptor is the base's low word plus a constant.
        constexpr uint32_t DESC_HI = 8u | (1u << 14);
        constexpr uint32_t A_LBO = (uint32_t)(3 * kUmCols) << 16, B_LBO = (uint32_t)CO << 16;
        const int units = UPC * total;   // weight blocks
        auto fetch_weights = [&](int u) {   // block u -> ring slot u % kUmWst (one lane)
            const uint32_t wf = wfull + 8u * (u % kUmWst);
            const int c = u / UPC, g = u - UPC * c;
            const int kd = kd_lo + c / NCH, ch = c % NCH;
            mbar_expect_tx(wf, (uint32_t)K::WF * 4u);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(Ws + (u % kUmWst) * K::WF)), "l"(p.wt + (((size_t)kd * NCH + ch) * UPC + g) * K::WF),
                           "r"((uint32_t)K::WF * 4u), "r"(wf)
                         : "memory");
        };
        if (lane == 0)
            for (int u = 0; u < kUmWst && u < units; ++u) fetch_weights(u);
        __syncwarp();
#pragma unroll 1
        for (int u = 0; u < units; ++u) {
            const int k = u / UPC, g = u - UPC * k, s = k & 1, ws = u % kUmWst;
            if (g == 0) mbar_wait_bounded(xfull + 8u * s, (uint32_t)(k >> 1) & 1u);
            mbar_wait_bounded(wfull + 8u * ws, (uint32_t)(u / kUmWst) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // with one kernel row per block the row offset is the only run-time part of the A address
            const uint32_t xa = smem_u32(Xs + s * K::XF) + (UPC == 3 ? (uint32_t)(g * kUmCols * 16) : 0u);
            const uint32_t a_base = (xa >> 4) | A_LBO, b_base = (smem_u32(Ws + ws * K::WF) >> 4) | B_LBO;
            const uint32_t acc = tmem + (uint32_t)((k % kUmAcc) * CO);
#pragma unroll
            for (int tp = 0; tp < TPU; ++tp) {
                const int ky = UPC == 3 ? 0 : tp / 3, kx = tp % 3;
                const uint32_t a_hi = a_base + (uint32_t)(ky * kUmCols + kx), a_lo = a_hi + (uint32_t)(K::XF / 2) / 4u;
                const uint32_t b_hi = b_base + (uint32_t)(tp * 2 * CO), b_lo = b_hi + (uint32_t)(K::WF / 2) / 4u;
                umma_tf32_elect(acc, a_lo, b_hi, DESC_HI, IDESC, (k >= kUmAcc || g > 0 || tp > 0) ? 1u : 0u);  // small terms first
                umma_tf32_elect(acc, a_hi, b_lo, DESC_HI, IDESC, 1u);
                umma_tf32_elect(acc, a_hi, b_hi, DESC_HI, IDESC, 1u);
            }
            // commits arrive when every MMA issued so far has completed (they imply tcgen05.fence::before_thread_sync)
            umma_commit_elect(wempty + 8u * ws);
            if (g == UPC - 1) umma_commit_elect(xempty + 8u * s);
            // block u-2 has been consumed (the issuer stays two blocks ahead of what it waits for): refill its slot
            if (u >= 2 && u - 2 + kUmWst < units) {
                mbar_wait_bounded(wempty + 8u * ((u - 2) % kUmWst), (uint32_t)((u - 2) / kUmWst) & 1u);
                if (lane == 0) fetch_weights(u - 2 + kUmWst);
                __syncwarp();
            }
        }
    } else {
        // ---- stagers: the window of chunk k+1 is requested before chunk k is split and stored -------------------------
        constexpr int NIT = (2 * 3 * 130 + 127) / 128;
        auto load_chunk = [&](int c, float (&v)[NIT][4]) {
            const int kd = kd_lo + c / NCH, ch = c % NCH;
            const int dz = d + kd - KD / 2;
            const float* xb = p.x + (((size_t)b * CIN + ch * 8) * p.D + dz) * plane;
#pragma unroll
            for (int j = 0; j < NIT; ++j) {
                const int it = tid + j * 128;
                const int kh = it / 390, rem = it - kh * 390;
                const int r = rem / 130, c2 = rem - r * 130;
                const int gy = y - 1 + r, gx = x0 - 1 + c2;
                const bool ok = it < 780 && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                const float* q = xb + (size_t)(kh * 4) * p.D * plane + (size_t)gy * W + gx;
#pragma unroll
                for (int e = 0; e < 4; ++e) v[j][e] = ok ? __ldg(q + (size_t)e * p.D * plane) : 0.0f;
            }
        };
        float cur[NIT][4], nxt[NIT][4];
        load_chunk(0, cur);
#pragma unroll 1
        for (int k = 0; k < total; ++k) {
            const int s = k & 1;
            if (k + 1 < total) load_chunk(k + 1, nxt);
            if (k >= 2) mbar_wait_bounded(xempty + 8u * s, (uint32_t)((k - 2) >> 1) & 1u);   // chunk k-2 has left stage s
            float* Xst = Xs + s * K::XF;
#pragma unroll
            for (int j = 0; j < NIT; ++j) {
                const int it = tid + j * 128;
                if (it < 780) {
                    const int kh = it / 390, rem = it - kh * 390;
                    const int r = rem / 130, c2 = rem - r * 130;
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        hi[e] = rna_tf32(cur[j][e]);
                        lo[e] = rna_tf32(cur[j][e] - __uint_as_float(hi[e]));
                    }
                    const int o = ((kh * 3 + r) * kUmCols + c2) * 4;
                    *reinterpret_cast<uint4*>(Xst + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(Xst + K::XF / 2 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            mbar_arrive_one(xfull + 8u * s);
#pragma unroll
            for (int j = 0; j < NIT; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) cur[j][e] = nxt[j][e];
        }

        // ---- epilogue: the last commit covers every MMA; sum the accumulators, bias, ReLU, planar store -------------
        mbar_wait_bounded(xempty + 8u * ((total - 1) & 1), (uint32_t)((total - 1) >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int nacc = total < kUmAcc ? total : kUmAcc;
        const int gx = x0 + warp * 32 + lane;
        float* yb = p.y + (((size_t)b * CO) * p.D + d) * plane + (size_t)y * W + gx;
        const size_t cstride = (size_t)p.D * plane;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes of tensor memory
#pragma unroll 1
        for (int j = 0; j < CO; j += 8) {
            // all accumulators of these 8 channels are requested before the one wait (unused accumulators of a short
            // K loop hold garbage and are masked out)
            uint32_t r[kUmAcc][8];
#pragma unroll
            for (int a = 0; a < kUmAcc; ++a)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(r[a][0]), "=r"(r[a][1]), "=r"(r[a][2]), "=r"(r[a][3]), "=r"(r[a][4]), "=r"(r[a][5]),
                               "=r"(r[a][6]), "=r"(r[a][7])
                             : "r"(trow + (uint32_t)(a * CO + j)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float sum[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                sum[e] = __uint_as_float(r[0][e]);
#pragma unroll
                for (int a = 1; a < kUmAcc; ++a) sum[e] += a < nacc ? __uint_as_float(r[a][e]) : 0.0f;
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
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(K::TCOLS) : "memory");
}

// w [KD][3][3][CIN][CO] fp32 -> packed [KD][CIN/8][block][hi|lo][tap in block][k-half][CO][4], TPU taps per block
__global__ void umma_pack_kernel(const float* __restrict__ w, float* __restrict__ out, int KD, int CIN, int CO, int TPU) {
    const size_t n = (size_t)KD * 9 * CIN * CO;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // decode the destination index of the hi half: [kd][ch][tap][kh][co][e]
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
    const int blkid = tap / TPU, tin = tap - blkid * TPU;
    const size_t half = (size_t)TPU * 2 * CO * 4;
    const size_t blk = (((size_t)kd * (CIN / 8) + ch) * (9 / TPU) + blkid) * 2 * half;
    const size_t off = (((size_t)tin * 2 + kh) * CO + co) * 4 + e;
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
    midconv_umma_kernel<KD, CIN, CO><<<grid, 160, K::SMEM, s>>>(p);
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
    umma_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, packed, kd, Cin, Cout, Cout == 64 ? 3 : 9);
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
