// K1 forward, fine stages (texels of 16 / 32 / 64 bytes: C = 8 / 16, fp32 or bf16): the staged "box" kernels.
//
// A CTA owns a 32 x TILE_H pixel tile.  Per source view the texels its samples can touch are staged by ONE
// cp.async.bulk.tensor.4d of box {C, BW, BH, 1} into shared memory (TMA zero-fill outside the image IS grid_sample's
// padding_mode='zeros', so the gather has no bounds logic), and every lane gathers its four bilinear taps with LDS.128
// from the hardware-swizzled box (8 neighbouring texels hit 8 different bank groups).
//
// Measured facts that shaped this file (profiles/r02_k1_fwd_*.md):
//   * first version (r01): 147 warp instructions per (pixel, view, hypothesis) sample, one CTA barrier per view.
//     Knock-out builds showed where the time goes: without the gather LDS -12 %, without the blend math 0 %, without
//     both -27 %; the remaining 73 % was everything AROUND the gather (sample arithmetic, bounding boxes, barriers,
//     start-up latency of short-lived CTAs at 16 warps per SM).
//   * sample arithmetic is therefore done in packed fp32x2 across PAIRS of hypotheses (p = R [x y 1]^T d + t, Newton
//     step of the reciprocal, floor via FADD2.RM with 1.5 * 2^23 - box origin - which yields the floor and the
//     box-relative integer index at once -, fractions, the four bilinear weights): 12.5 instead of 26 issue slots per
//     sample; the tap address is two integer multiply-adds of the raw float bits.
//   * no clamps, no zero test and no bounds logic on the staged path: a view whose bounding box is larger than the
//     staging buffer - which includes every non-finite position - takes the exact direct-gather path for that tile.
//
// Two kernels share the per-view device functions below:
//   epi_fwd_box_kernel  (this file): one tile per CTA, exact per-view bounding box computed in the kernel (phase A).
//   epi_fwd_pipe_kernel (epi_fwd_pipe.cuh): persistent CTAs, boxes from a tiny pre-pass, TMA ring with full / empty
//                        mbarriers fed one tile ahead - no CTA-wide barrier in the steady state.
#pragma once

#include "epi_tma.cuh"

#ifndef MVSTER_BOX_WARPS
#define MVSTER_BOX_WARPS 4
#endif
#ifndef MVSTER_BOX_MINB
#define MVSTER_BOX_MINB 5   // 96 registers: fits since the column slots halved the live tap registers (A/B: 0.719 -> 0.674 ms)
#endif
#ifndef MVSTER_BOX_BW
#define MVSTER_BOX_BW 48
#endif
#ifndef MVSTER_BOX_BHX
#define MVSTER_BOX_BHX 3
#endif
#ifndef MVSTER_BOX_NBUF
#define MVSTER_BOX_NBUF 4
#endif
#ifndef MVSTER_BOX_WIDE_WARPS
#define MVSTER_BOX_WIDE_WARPS 8
#endif
#ifndef MVSTER_BOX_WIDE_MINB
#define MVSTER_BOX_WIDE_MINB 3
#endif
#ifndef MVSTER_BOX_SLOTS
#define MVSTER_BOX_SLOTS 1  // texel-column register slots shared by the hypotheses of a pixel (gather_view_slots); 0 = gather every tap
#endif
#ifndef MVSTER_BOX_KO
#define MVSTER_BOX_KO 0    // development only, wrong results: knock-out bits (1 no gather LDS, 2 no TMA, 8 no stores, 16 no blend math)
#endif

namespace mvster {

// Geometry shared by the staged kernels.  BWT / BHX: box width in texels / extra box rows beyond the tile height.
template <int C, int D, int ES, int BWT = MVSTER_BOX_BW, int BHX = MVSTER_BOX_BHX, int NBUFS = MVSTER_BOX_NBUF>
struct BoxCfg {
    static constexpr int TB = C * ES;                                    // texel bytes: 16, 32 or 64
    static constexpr int LD = (TB == 64 && D % 4 == 0) ? 2 : 1;          // lanes per pixel (each owns D/LD hypotheses)
    static constexpr int DL = D / LD;
    static constexpr int PPW = 32 / LD;                                  // pixels per warp
    static constexpr int WX = LD;                                        // warps side by side: the tile is 32 wide
    static constexpr int WARPS = LD == 2 ? MVSTER_BOX_WIDE_WARPS : MVSTER_BOX_WARPS;
    static constexpr int MINB = LD == 2 ? MVSTER_BOX_WIDE_MINB : MVSTER_BOX_MINB;
    static constexpr int TILE_W = 32, TILE_H = WARPS / WX;
    static constexpr int BW = BWT, BH = TILE_H + BHX;                    // staging box in texels
    static constexpr int ROW_BYTES = BW * TB, BUF_BYTES = ROW_BYTES * BH;
    static constexpr int NBUF = NBUFS;
    static constexpr int NCHUNK = C / 8;
    static constexpr int NSUB = TB / 16;                                 // 16-byte chunks per texel
    static constexpr uint32_t SWZ = (uint32_t)(NSUB - 1) << 4;           // swizzled 16-byte chunk index bits
    // A row offset adds (ROW >> 7) to the swizzle source bits [8:7].  With those bits of ROW in {0, 2} (64B mode) or
    // {0, 1} (32B mode) the addition is an XOR: the lower tap row reads chunk j at the upper row's address of chunk
    // j ^ ROW_SWZ - a compile-time renaming, no extra instruction.
    static constexpr int ROW_SWZ = (ROW_BYTES >> 7) & (NSUB - 1);
    static constexpr int SWZ_PERIOD = TB == 64 ? 512 : (TB == 32 ? 256 : 128);  // bytes after which the pattern repeats
    static constexpr int BAR_OFF = NBUF * BUF_BYTES;
    static constexpr int BBOX_OFF = BAR_OFF + 64;
    static constexpr int RT_OFF = BBOX_OFF + MVSTER_MAX_SRC_VIEWS * 16;
    static constexpr int SMEM = RT_OFF + MVSTER_MAX_SRC_VIEWS * 48 + 1024;  // + slack for the 1024-byte alignment
    static_assert(DL % 2 == 0, "hypotheses are processed in packed pairs");
    static_assert(BUF_BYTES % SWZ_PERIOD == 0 && (TB == 16 || ROW_BYTES % 128 == 0), "buffers must not change the swizzle phase");
    static_assert(TB != 64 || (ROW_SWZ & 1) == 0, "64B swizzle: the row pitch must be a multiple of 256 bytes");
    static_assert(NBUF >= 1 && NBUF <= 8, "NBUF");
};

__device__ __forceinline__ f32x2 add2_rm(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float min_nan(float a, float b) {
    float d;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float max_nan(float a, float b) {
    float d;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}

// Per-view constants of the sample-position arithmetic for one pixel; all of them carry the sign that makes the
// packed Newton step free of negations: sx = (-px) * (-1/pz).
struct PixelView {
    f32x2 naxy;       // -(R [x y 1]^T).xy
    float az;         //  (R [x y 1]^T).z
    float nt0, nt1, t2;
};

// Repack one [R|t] (12 floats, row-major 3x4) for pixel_view(): r00 r10 r01 r11 | r02 r12 -t0 -t1 | r20 r21 r22 t2
__device__ __forceinline__ float repack_rt(const float* rt12, int k) {
    const int srck = k < 8 ? (k & 1) * 4 + (k >> 1) : k;
    const float val = __ldg(rt12 + srck);
    return (k == 6 || k == 7) ? -val : val;
}

__device__ __forceinline__ PixelView pixel_view(const float* rt_s, float fxp, float fyp) {
    const float4* q = reinterpret_cast<const float4*>(rt_s);
    const float4 a = q[0], b = q[1], c = q[2];
    PixelView pv;
    // same nesting as the reference restatement: r0*x + (r1*y + r2)
    const f32x2 axy = fma2(pack2(a.x, a.y), pack2(fxp, fxp), fma2(pack2(a.z, a.w), pack2(fyp, fyp), pack2(b.x, b.y)));
    pv.naxy = mul2(axy, pack2(-1.0f, -1.0f));
    pv.az = fmaf(c.x, fxp, fmaf(c.y, fyp, c.z));
    pv.nt0 = b.z; pv.nt1 = b.w; pv.t2 = c.w;
    return pv;
}

// Positions of two hypotheses (h0, h1) of one pixel in one view: sx = (ax d + t0) / (az d + t2), same for sy, with the
// reciprocal as MUFU.RCP + one Newton step (<= 1 ulp).  No clamp, no zero test: the bounding-box step rejects what
// this produces for degenerate input.  Every caller of one kernel uses this same function, so they see identical bits.
__device__ __forceinline__ void positions2(const PixelView& pv, f32x2 hh, f32x2& sx, f32x2& sy) {
    float nax, nay;
    unpack2(pv.naxy, nax, nay);
    const f32x2 npx = fma2(pack2(nax, nax), hh, pack2(pv.nt0, pv.nt0));
    const f32x2 npy = fma2(pack2(nay, nay), hh, pack2(pv.nt1, pv.nt1));
    const f32x2 pz = fma2(pack2(pv.az, pv.az), hh, pack2(pv.t2, pv.t2));
    float z0, z1;
    unpack2(pz, z0, z1);
    const f32x2 nr = pack2(rcp_approx(-z0), rcp_approx(-z1));            // -1/pz (approx)
    const f32x2 e = fma2(pz, nr, pack2(1.0f, 1.0f));                     // 1 - pz * r
    const f32x2 nrz = fma2(nr, e, nr);                                   // -(r + r e)
    sx = mul2(npx, nrz);
    sy = mul2(npy, nrz);
}

constexpr float kFloorMagic = 12582912.0f;        // 1.5 * 2^23: (s + magic) rounded down has floor(s) in its low bits
constexpr int kFloorMagicBits = 0x4B400000;

// Box origin (texels) and "the view is staged" from a bounding box of magic-offset float bits {min x, min y, max x,
// max y}: inside the linear range of the floor trick, and the footprint (+1 for the right / bottom tap) fits.
template <typename K>
__device__ __forceinline__ bool box_fits(const int4 bb, int& bx, int& by) {
    bx = bb.x - kFloorMagicBits;
    by = bb.y - kFloorMagicBits;
    return (bb.x > kFloorMagicBits - (1 << 20)) && (bb.y > kFloorMagicBits - (1 << 20)) &&
           (bb.z < kFloorMagicBits + (1 << 20)) && (bb.w < kFloorMagicBits + (1 << 20)) &&
           (bb.z - bb.x + 2 <= K::BW) && (bb.w - bb.y + 2 <= K::BH);
}

// 8 channels (chunk c8 of the texel) of the upper (ROWSEL = 0) or lower (ROWSEL = 1) tap row.  aL = swizzled address of
// the texel's logical 16-byte chunk 0 in the UPPER row.
template <typename K, typename T, int ROWSEL>
__device__ __forceinline__ void lds_tap(uint32_t aL, int c8, P8& t) {
    constexpr int RX = ROWSEL ? K::ROW_SWZ : 0;
    constexpr int OFF = ROWSEL ? K::ROW_BYTES : 0;
    if constexpr (sizeof(T) == 4) {
        const uint32_t j0 = (uint32_t)((2 * c8) ^ RX) << 4, j1 = (uint32_t)((2 * c8 + 1) ^ RX) << 4;
        lds_pairs_at<OFF>(aL ^ j0, t.q[0], t.q[1]);
        lds_pairs_at<OFF>(aL ^ j1, t.q[2], t.q[3]);
    } else {
        const uint32_t j0 = (uint32_t)(c8 ^ RX) << 4;
        uint32_t x, y, z, w;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+%5];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(aL ^ j0), "n"(OFF));
        t.q[0] = pack2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
        t.q[1] = pack2(__uint_as_float(y << 16), __uint_as_float(y & 0xffff0000u));
        t.q[2] = pack2(__uint_as_float(z << 16), __uint_as_float(z & 0xffff0000u));
        t.q[3] = pack2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    }
}

// ---- one view, staged path: cor[g][d] = group-wise correlation of the reference with the bilinear samples taken
//      from the box at shared address `buf` whose origin is texel (bx, by) ------------------------------------------------
template <typename K, int CPG, typename T>
__device__ __forceinline__ void gather_view(const PixelView& pv, const f32x2 (&hh)[K::DL / 2], const f32x2* rf, uint32_t buf,
                                            int bx, int by, float (&cor)[K::NCHUNK * (8 / CPG)][K::DL]) {
    constexpr int DL = K::DL, NCHUNK = K::NCHUNK, TB = K::TB, ROW = K::ROW_BYTES, GPC = 8 / CPG;
    constexpr uint32_t SWZ = K::SWZ;
    const float kx = kFloorMagic - (float)bx, ky = kFloorMagic - (float)by;  // exact
    // address of texel (rx, ry) = buf + ry * ROW + rx * TB with rx = bits(tx) - magic bits: fold the constants
    const uint32_t bufk = buf - (uint32_t)kFloorMagicBits * (uint32_t)(ROW + TB);
#pragma unroll
    for (int k = 0; k < DL / 2; ++k) {
        f32x2 sx, sy;
        positions2(pv, hh[k], sx, sy);
        const f32x2 tx = add2_rm(sx, pack2(kx, kx)), ty = add2_rm(sy, pack2(ky, ky));
        const f32x2 flx = add2(tx, pack2(-kx, -kx)), fly = add2(ty, pack2(-ky, -ky));   // floor(s), exact
        const f32x2 m1 = pack2(-1.0f, -1.0f), one = pack2(1.0f, 1.0f);
        const f32x2 fx = fma2(flx, m1, sx), fy = fma2(fly, m1, sy);                     // s - floor(s), exact
        const f32x2 gx = fma2(fx, m1, one), gy = fma2(fy, m1, one);
        const f32x2 w00 = mul2(gx, gy), w01 = mul2(fx, gy), w10 = mul2(gx, fy), w11 = mul2(fx, fy);
        float wa[2][4];
        unpack2(w00, wa[0][0], wa[1][0]); unpack2(w01, wa[0][1], wa[1][1]);
        unpack2(w10, wa[0][2], wa[1][2]); unpack2(w11, wa[0][3], wa[1][3]);
        float txs[2], tys[2];
        unpack2(tx, txs[0], txs[1]);
        unpack2(ty, tys[0], tys[1]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int d = 2 * k + j;
            const uint32_t base = (uint32_t)__float_as_int(tys[j]) * (uint32_t)ROW +
                                  ((uint32_t)__float_as_int(txs[j]) * (uint32_t)TB + bufk);
            // hardware swizzle: 16-byte chunk index ^= address bits [8:7] (64B mode) / [7] (32B mode)
            const uint32_t baseR = base + TB;
            const uint32_t aL = base ^ ((base >> 3) & SWZ), aR = baseR ^ ((baseR >> 3) & SWZ);
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                P8 t00, t01, t10, t11;
                if constexpr ((MVSTER_BOX_KO & 1) != 0) {  // taps made of the addresses: no shared-memory traffic
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        t00.q[q] = pack2(__uint_as_float(aL), __uint_as_float(aR + q));
                        t01.q[q] = pack2(__uint_as_float(aR), __uint_as_float(aL + q));
                        t10.q[q] = t01.q[q]; t11.q[q] = t00.q[q];
                    }
                } else {
                    lds_tap<K, T, 0>(aL, c, t00);
                    lds_tap<K, T, 0>(aR, c, t01);
                    lds_tap<K, T, 1>(aL, c, t10);
                    lds_tap<K, T, 1>(aR, c, t11);
                }
                float cg[GPC];
                if constexpr ((MVSTER_BOX_KO & 16) != 0) {  // keep the loads alive, drop the 20 packed + 4 scalar math ops
                    const f32x2 xo = t00.q[0] ^ t01.q[1] ^ t10.q[2] ^ t11.q[3] ^ t00.q[1] ^ t01.q[2] ^ t10.q[3] ^ t11.q[0] ^
                                     t00.q[2] ^ t01.q[3] ^ t10.q[0] ^ t11.q[1] ^ t00.q[3] ^ t01.q[0] ^ t10.q[1] ^ t11.q[2];
                    float lo, hi;
                    unpack2(xo, lo, hi);
#pragma unroll
                    for (int g = 0; g < GPC; ++g) cg[g] = (g & 1) ? hi * wa[j][g & 3] : lo * wa[j][g & 3];
                } else {
                    blend_correlate<CPG>(t00, t01, t10, t11, wa[j][0], wa[j][1], wa[j][2], wa[j][3], rf + c * 4, cg);
                }
#pragma unroll
                for (int g = 0; g < GPC; ++g) cor[c * GPC + g][d] = cg[g];
            }
        }
    }
}

// Predicated variant: lanes with pred == false keep the previous contents of t (no shared-memory access, and a
// quarter-warp without an active lane costs no wavefront).
template <int OFF>
__device__ __forceinline__ void lds_pairs_pred(uint32_t addr, f32x2& a, f32x2& b, bool pred) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@p ld.shared.v2.b64 {%0,%1}, [%2+%3];\n\t}"
        : "+l"(a), "+l"(b)
        : "r"(addr), "n"(OFF), "r"((int)pred));
}
template <typename K, typename T, int ROWSEL>
__device__ __forceinline__ void lds_tap_pred(uint32_t aL, int c8, P8& t, bool pred) {
    constexpr int RX = ROWSEL ? K::ROW_SWZ : 0;
    constexpr int OFF = ROWSEL ? K::ROW_BYTES : 0;
    if constexpr (sizeof(T) == 4) {
        const uint32_t j0 = (uint32_t)((2 * c8) ^ RX) << 4, j1 = (uint32_t)((2 * c8 + 1) ^ RX) << 4;
        lds_pairs_pred<OFF>(aL ^ j0, t.q[0], t.q[1], pred);
        lds_pairs_pred<OFF>(aL ^ j1, t.q[2], t.q[3], pred);
    } else {
        if (pred) lds_tap<K, T, ROWSEL>(aL, c8, t);
    }
}

// ---- one view, staged path with texel-column slots -------------------------------------------------------------------
// The D hypotheses of a pixel walk along the epipolar line in sub-texel to ~1-texel steps (profiles/r02: 7.7 distinct
// texels behind the 16 taps of a stage-4 pixel and view), and the shared-memory data pipe is the busiest unit of this
// kernel.  A lane therefore keeps two texel COLUMNS (upper and lower tap row each) in registers and addresses them by
// the parity of the column index relative to the first hypothesis' cell: a sample whose cell is k columns away from
// sample 0 uses the slot of parity k for its left taps and the other one for its right taps; only a slot whose
// column changed is loaded again (predicated LDS - no branch, no data movement between registers; the bilinear x
// weights are swapped instead).  Because the parity is taken relative to the lane's own first cell, neighbouring lanes
// reload the same slot at the same hypothesis, so whole quarter-warps skip the access together.  A change of the tap
// row (rare: epipolar lines are close to horizontal or the step is sub-texel) reloads both slots.
template <typename K, int CPG, typename T>
__device__ __forceinline__ void gather_view_slots(const PixelView& pv, const f32x2 (&hh)[K::DL / 2], const f32x2* rf,
                                                  uint32_t buf, int bx, int by, float (&cor)[K::NCHUNK * (8 / CPG)][K::DL]) {
    constexpr int DL = K::DL, NCHUNK = K::NCHUNK, TB = K::TB, ROW = K::ROW_BYTES, GPC = 8 / CPG;
    constexpr uint32_t SWZ = K::SWZ;
    const float kx = kFloorMagic - (float)bx, ky = kFloorMagic - (float)by;  // exact
    const uint32_t bufk = buf - (uint32_t)kFloorMagicBits * (uint32_t)(ROW + TB);
    float fxs[DL], fys[DL];
    int txi[DL], tyi[DL];
#pragma unroll
    for (int k = 0; k < DL / 2; ++k) {
        f32x2 sx, sy;
        positions2(pv, hh[k], sx, sy);
        const f32x2 tx = add2_rm(sx, pack2(kx, kx)), ty = add2_rm(sy, pack2(ky, ky));
        const f32x2 flx = add2(tx, pack2(-kx, -kx)), fly = add2(ty, pack2(-ky, -ky));   // floor(s), exact
        const f32x2 m1 = pack2(-1.0f, -1.0f);
        const f32x2 fx = fma2(flx, m1, sx), fy = fma2(fly, m1, sy);                     // s - floor(s), exact
        unpack2(fx, fxs[2 * k], fxs[2 * k + 1]);
        unpack2(fy, fys[2 * k], fys[2 * k + 1]);
        float a0, a1;
        unpack2(tx, a0, a1); txi[2 * k] = __float_as_int(a0); txi[2 * k + 1] = __float_as_int(a1);
        unpack2(ty, a0, a1); tyi[2 * k] = __float_as_int(a0); tyi[2 * k + 1] = __float_as_int(a1);
    }
    // column slots: S[a][row][chunk]; slot a holds the column of relative index c_a (parity a)
    P8 S[2][2][NCHUNK];
    const uint32_t col0 = (uint32_t)txi[0] * (uint32_t)TB + bufk;   // byte offset of the first sample's left column
    int pc0 = 0, pc1 = 0;
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        const int kc = txi[d] - txi[0];                 // cell column relative to sample 0 (any sign)
        const int c0 = (kc + 1) & ~1, c1 = kc | 1;      // the even / odd one of {kc, kc + 1}
        const bool rowchg = d > 0 && tyi[d] != tyi[d - 1];
        const bool need0 = d == 0 || rowchg || c0 != pc0;
        const bool need1 = d == 0 || rowchg || c1 != pc1;
        pc0 = c0; pc1 = c1;
        const uint32_t rowa = (uint32_t)tyi[d] * (uint32_t)ROW + col0;
        const uint32_t b0 = rowa + (uint32_t)(c0 * TB), b1 = rowa + (uint32_t)(c1 * TB);
        const uint32_t a0 = b0 ^ ((b0 >> 3) & SWZ), a1 = b1 ^ ((b1 >> 3) & SWZ);
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            if (d == 0) {
                lds_tap<K, T, 0>(a0, c, S[0][0][c]); lds_tap<K, T, 1>(a0, c, S[0][1][c]);
                lds_tap<K, T, 0>(a1, c, S[1][0][c]); lds_tap<K, T, 1>(a1, c, S[1][1][c]);
            } else {
                lds_tap_pred<K, T, 0>(a0, c, S[0][0][c], need0); lds_tap_pred<K, T, 1>(a0, c, S[0][1][c], need0);
                lds_tap_pred<K, T, 0>(a1, c, S[1][0][c], need1); lds_tap_pred<K, T, 1>(a1, c, S[1][1][c], need1);
            }
        }
        // left taps sit in the slot of parity kc: swap the x weights instead of the data
        const bool odd = (kc & 1) != 0;
        const float gxd = 1.0f - fxs[d], gyd = 1.0f - fys[d];
        const float w0 = odd ? fxs[d] : gxd, w1 = odd ? gxd : fxs[d];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            float cg[GPC];
            blend_correlate<CPG>(S[0][0][c], S[1][0][c], S[0][1][c], S[1][1][c], w0 * gyd, w1 * gyd, w0 * fys[d], w1 * fys[d],
                                 rf + c * 4, cg);
#pragma unroll
            for (int g = 0; g < GPC; ++g) cor[c * GPC + g][d] = cg[g];
        }
    }
}

// ---- one view, exact path for footprints larger than the box (or degenerate positions): direct gather from global
//      memory with the reference's clamp / zero test / per-tap bounds weights -----------------------------------------
template <typename K, int CPG, typename T>
__device__ __forceinline__ void direct_view(const float* rt12, const void* src_b, int Hs, int Ws, float fxp, float fyp,
                                            const f32x2 (&hh)[K::DL / 2], const f32x2* rf,
                                            float (&cor)[K::NCHUNK * (8 / CPG)][K::DL]) {
    constexpr int DL = K::DL, NCHUNK = K::NCHUNK, TB = K::TB, GPC = 8 / CPG;
    const Homography h = load_homography(rt12);
    const float ax = fmaf(h.r00, fxp, fmaf(h.r01, fyp, h.r02));
    const float ay = fmaf(h.r10, fxp, fmaf(h.r11, fyp, h.r12));
    const float az = fmaf(h.r20, fxp, fmaf(h.r21, fyp, h.r22));
    const char* srcp = reinterpret_cast<const char*>(src_b);
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        float h0, h1;
        unpack2(hh[d / 2], h0, h1);
        const Taps t = make_taps(ax, ay, az, h, (d & 1) ? h1 : h0, Hs, Ws);
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            constexpr int KO = 8 * (int)sizeof(T);
            const P8 t00 = load_pairs<T>(srcp + (size_t)(unsigned)t.o00 * TB + c * KO);
            const P8 t01 = load_pairs<T>(srcp + (size_t)(unsigned)t.o01 * TB + c * KO);
            const P8 t10 = load_pairs<T>(srcp + (size_t)(unsigned)t.o10 * TB + c * KO);
            const P8 t11 = load_pairs<T>(srcp + (size_t)(unsigned)t.o11 * TB + c * KO);
            float cg[GPC];
            blend_correlate<CPG>(t00, t01, t10, t11, t.w00, t.w01, t.w10, t.w11, rf + c * 4, cg);
#pragma unroll
            for (int g = 0; g < GPC; ++g) cor[c * GPC + g][d] = cg[g];
        }
    }
}

// ---- epipolar attention of one view and its contribution to the aggregate ---------------------------------------------
// score[d] = sum over all G groups (reference cor_feat.sum(1), :1083); w = softmax over D of score / attn_temp, divided
// by sqrt(C); max and sum cross the LD hypothesis lanes of the pixel.  Returns the weights in w[] (for `weights`).
template <typename K, int GPL>
__device__ __forceinline__ void attend_accumulate(const float (&cor)[GPL][K::DL], float score_scale, float inv_sqrt_c,
                                                  float (&acc)[GPL][K::DL], float (&wsum)[K::DL], float (&w)[K::DL]) {
    constexpr int DL = K::DL, LD = K::LD;
    float score[DL];
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        float s = cor[0][d];
#pragma unroll
        for (int g = 1; g < GPL; ++g) s += cor[g][d];
        score[d] = s;
    }
    float mx = score[0];
#pragma unroll
    for (int d = 1; d < DL; ++d) mx = fmaxf(mx, score[d]);
#pragma unroll
    for (int m = 1; m < LD; m <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    float es = 0.0f;
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        w[d] = ex2_approx((score[d] - mx) * score_scale);
        es += w[d];
    }
#pragma unroll
    for (int m = 1; m < LD; m <<= 1) es += __shfl_xor_sync(0xffffffffu, es, m);
    const float norm = __fdividef(inv_sqrt_c, es);
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        w[d] *= norm;
        wsum[d] += w[d];
#pragma unroll
        for (int g = 0; g < GPL; ++g) acc[g][d] = fmaf(w[d], cor[g][d], acc[g][d]);
    }
}

// reference channels of one pixel as packed pairs, pre-scaled by 1/(C/G) so that the group mean is a plain sum
template <int C, int CPG, typename T>
__device__ __forceinline__ void load_ref(const void* ref, size_t pixel_index, f32x2 (&rf)[C / 2]) {
    const T* refp = reinterpret_cast<const T*>(ref) + pixel_index * C;
    const f32x2 sc = pack2(1.0f / CPG, 1.0f / CPG);
#pragma unroll
    for (int k = 0; k < C / 8; ++k) {
        const P8 r = load_pairs<T>(refp + k * 8);
#pragma unroll
        for (int q = 0; q < 4; ++q) rf[k * 4 + q] = mul2(r.q[q], sc);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int C, int CPG, int D, typename T>
__global__ void __launch_bounds__(BoxCfg<C, D, (int)sizeof(T)>::WARPS * 32, BoxCfg<C, D, (int)sizeof(T)>::MINB)
    epi_fwd_box_kernel(const __grid_constant__ EpiFwdParams p) {
    using K = BoxCfg<C, D, (int)sizeof(T)>;
    constexpr int DL = K::DL, LD = K::LD, TB = K::TB, NBUF = K::NBUF;
    constexpr int G = C / CPG, GPL = G;              // a lane owns all channels = all groups
    constexpr int NT = K::WARPS * 32;
    static_assert(8 % CPG == 0, "a chunk of 8 channels must hold whole groups");

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + K::BAR_OFF;
    int* bbox = reinterpret_cast<int*>(sm + K::BBOX_OFF);      // [view][4]: min x, min y, max x, max y (magic-offset bits)
    float* rt_s = reinterpret_cast<float*>(sm + K::RT_OFF);    // [view][12], repacked (see pixel_view)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pix = lane / LD, dl = lane % LD;
    const int b = blockIdx.z;
    const int Nsrc = p.Nsrc;
    {
        const float* rtg = p.rt + (size_t)b * Nsrc * 12;
        for (int i = tid; i < Nsrc * 12; i += NT) rt_s[i] = repack_rt(rtg + (i / 12) * 12, i % 12);
        if (tid < Nsrc) {
            bbox[tid * 4 + 0] = INT_MAX; bbox[tid * 4 + 1] = INT_MAX;
            bbox[tid * 4 + 2] = INT_MIN; bbox[tid * 4 + 3] = INT_MIN;
        }
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < NBUF; ++i) mbar_init(bars + 8u * i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    int x = blockIdx.x * K::TILE_W + (warp % K::WX) * K::PPW + pix;
    int y = blockIdx.y * K::TILE_H + (warp / K::WX);
    const bool live = (x < p.W) && (y < p.H);
    x = min(x, p.W - 1);  // dead lanes shadow a valid pixel: the bounding box is unaffected
    y = min(y, p.H - 1);
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix_off = (size_t)y * p.W + x;
    const float fxp = (float)x, fyp = (float)y;

    f32x2 hh[DL / 2];
    {
        float hyp[DL];
#pragma unroll
        for (int d = 0; d < DL; ++d) hyp[d] = ldg_stream(p.hypo + ((size_t)b * D + dl * DL + d) * plane + pix_off);
#pragma unroll
        for (int k = 0; k < DL / 2; ++k) hh[k] = pack2(hyp[2 * k], hyp[2 * k + 1]);
    }
    f32x2 rf[C / 2];
    load_ref<C, CPG, T>(p.ref, (size_t)b * plane + pix_off, rf);  // in flight across phase A
    __syncthreads();  // rt_s, bbox slots, mbarriers

    // ---- phase A: bounding box of every view's sample positions ------------------------------------------------
#pragma unroll 1
    for (int v = 0; v < Nsrc; ++v) {
        const PixelView pv = pixel_view(rt_s + v * 12, fxp, fyp);
        float lox, hix, loy, hiy;
#pragma unroll
        for (int k = 0; k < DL / 2; ++k) {
            f32x2 sx, sy;
            positions2(pv, hh[k], sx, sy);
            float x0, x1, y0, y1;
            unpack2(sx, x0, x1);
            unpack2(sy, y0, y1);
            if (k == 0) {
                lox = min_nan(x0, x1); hix = max_nan(x0, x1);
                loy = min_nan(y0, y1); hiy = max_nan(y0, y1);
            } else {
                lox = min_nan(lox, min_nan(x0, x1)); hix = max_nan(hix, max_nan(x0, x1));
                loy = min_nan(loy, min_nan(y0, y1)); hiy = max_nan(hiy, max_nan(y0, y1));
            }
        }
        // a NaN anywhere makes hix / hiy NaN -> 0x7fffffff after the add -> the box cannot fit.  ptxas turns a
        // warp-uniform-address shared atomic into REDUX + one ATOMS by an elected lane.
        int* slot = bbox + v * 4;
        atomicMin(slot + 0, __float_as_int(__fadd_rd(lox, kFloorMagic)));
        atomicMin(slot + 1, __float_as_int(__fadd_rd(loy, kFloorMagic)));
        atomicMax(slot + 2, __float_as_int(__fadd_rd(hix, kFloorMagic)));
        atomicMax(slot + 3, __float_as_int(__fadd_rd(hiy, kFloorMagic)));
    }
    __syncthreads();

    auto request = [&](int v) {  // one thread
        int bx, by;
        if (!(MVSTER_BOX_KO & 2) && box_fits<K>(*reinterpret_cast<const int4*>(bbox + v * 4), bx, by)) {
            const uint32_t bar = bars + 8u * (v % NBUF);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of a recycled buffer
            mbar_expect_tx(bar, (uint32_t)K::BUF_BYTES);
            tma_load_4d(smem_base + (uint32_t)K::BUF_BYTES * (v % NBUF), &p.tmap[v], bar, 0, bx, by, b);
        }
    };
    if (tid == 0) {
        for (int v = 0; v < Nsrc && v < NBUF; ++v) request(v);
    }

    float acc[GPL][DL], wsum[DL];
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        wsum[d] = 1e-8f;  // reference :1037
#pragma unroll
        for (int g = 0; g < GPL; ++g) acc[g][d] = 0.0f;
    }

    // ---- phase B: gather, correlate, attend, accumulate ------------------------------------------------------
    uint32_t phase_bits = 0;  // bit k: parity of the next completion of mbarrier k (uniform across the CTA)
#pragma unroll 1
    for (int v = 0; v < Nsrc; ++v) {
        int bx, by;
        const bool fit = box_fits<K>(*reinterpret_cast<const int4*>(bbox + v * 4), bx, by);
        float cor[GPL][DL];
        if (fit) {
            const PixelView pv = pixel_view(rt_s + v * 12, fxp, fyp);
            const int slot = v % NBUF;
            if (!(MVSTER_BOX_KO & 2)) mbar_wait(bars + 8u * slot, (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            if constexpr (MVSTER_BOX_SLOTS != 0)
                gather_view_slots<K, CPG, T>(pv, hh, rf, smem_base + (uint32_t)K::BUF_BYTES * slot, bx, by, cor);
            else
                gather_view<K, CPG, T>(pv, hh, rf, smem_base + (uint32_t)K::BUF_BYTES * slot, bx, by, cor);
        } else {
            direct_view<K, CPG, T>(p.rt + ((size_t)b * Nsrc + v) * 12,
                                   reinterpret_cast<const char*>(p.src[v]) + (size_t)b * p.Hs * p.Ws * TB, p.Hs, p.Ws, fxp,
                                   fyp, hh, rf, cor);
        }
        float w[DL];
        attend_accumulate<K, GPL>(cor, p.score_scale, p.inv_sqrt_c, acc, wsum, w);
        if (p.weights != nullptr && live) {
#pragma unroll
            for (int d = 0; d < DL; ++d)
                p.weights[(((size_t)b * Nsrc + v) * D + dl * DL + d) * plane + pix_off] = w[d];
        }
        if (v + NBUF < Nsrc) {  // more views than buffers: recycle this view's buffer once every warp has left it
            __syncthreads();
            if (tid == 0) request(v + NBUF);
        }
    }

    if (!live) return;
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        const float inv = __frcp_rn(wsum[d]);
        const int dd = dl * DL + d;
#pragma unroll
        for (int g = 0; g < GPL; ++g)
            if (!(MVSTER_BOX_KO & 8) || acc[g][d] * inv == 1234.5678f)
                stg_stream(p.out + (((size_t)b * G + g) * D + dd) * plane + pix_off, acc[g][d] * inv);
        if (p.wsum != nullptr) p.wsum[((size_t)b * D + dd) * plane + pix_off] = wsum[d];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: tensor maps (cached per calling thread: cuTensorMapEncodeTiled is ~1 us per view and eager callers pass
// the same feature buffers call after call) and launch
// ---------------------------------------------------------------------------------------------------------------------
struct MapKey {
    const void* ptr;
    int C, ES, B, Hs, Ws, bw, bh, swz;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && C == o.C && ES == o.ES && B == o.B && Hs == o.Hs && Ws == o.Ws && bw == o.bw &&
               bh == o.bh && swz == o.swz;
    }
};

// box {C, bw, bh, 1} over an NHWC feature map; swz: 0 none, 1 = 32B, 2 = 64B, 3 = 128B
static inline bool encode_box_map(CUtensorMap* out, const void* ptr, int C, int ES, int B, int Hs, int Ws, int bw, int bh,
                                  int swz) {
    constexpr int kSlots = 64;
    struct Entry { MapKey key; alignas(64) CUtensorMap map; bool used; };
    static thread_local Entry cache[kSlots] = {};
    static thread_local unsigned next = 0;
    const MapKey key{ptr, C, ES, B, Hs, Ws, bw, bh, swz};
    for (int i = 0; i < kSlots; ++i)
        if (cache[i].used && cache[i].key == key) { *out = cache[i].map; return true; }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    const cuuint64_t tby = (cuuint64_t)C * ES;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
    const cuuint64_t strides[3] = {tby, (cuuint64_t)Ws * tby, (cuuint64_t)Hs * Ws * tby};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const CUtensorMapSwizzle sw = swz == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                  : swz == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                  : swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    const CUtensorMapDataType dt = ES == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap m;
    if (enc(&m, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    Entry& e = cache[next++ % kSlots];
    e.key = key; e.map = m; e.used = true;
    *out = m;
    return true;
}

template <typename K, typename T>
static bool encode_view_maps(EpiFwdParams& p, int C) {
    constexpr int swz = K::TB == 16 ? 0 : (K::TB == 32 ? 1 : 2);
    for (int v = 0; v < p.Nsrc; ++v)
        if (!encode_box_map(&p.tmap[v], p.src[v], C, (int)sizeof(T), p.B, p.Hs, p.Ws, K::BW, K::BH, swz)) return false;
    return true;
}

template <int C, int CPG, int D, typename T>
static int launch_box(EpiFwdParams& p, cudaStream_t stream, bool* built) {
    using K = BoxCfg<C, D, (int)sizeof(T)>;
    *built = encode_view_maps<K, T>(p, C);
    if (!*built) return MVSTER_OK;
    static int smem_set[64] = {};
    const int st = ensure_dynamic_smem_bytes(epi_fwd_box_kernel<C, CPG, D, T>, K::SMEM, smem_set, "epi_fwd(box): cudaFuncSetAttribute");
    if (st != MVSTER_OK) return st;
    dim3 grid((p.W + K::TILE_W - 1) / K::TILE_W, (p.H + K::TILE_H - 1) / K::TILE_H, p.B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: grid too large");
    epi_fwd_box_kernel<C, CPG, D, T><<<grid, K::WARPS * 32, K::SMEM, stream>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_fwd(box) launch");
    return MVSTER_OK;
}

}  // namespace mvster
