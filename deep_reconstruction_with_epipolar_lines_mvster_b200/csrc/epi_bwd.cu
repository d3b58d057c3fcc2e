// K1 backward: hand-written gradient of the fused warp + correlation + epipolar attention + aggregation kernel
// with respect to the reference and source features (the sampling grid is not differentiated: it is built under
// no_grad in the reference, models/mvs4net_utils.py:31, and depth_hypo is detached at models/MVS4Net.py:116).
//
// Nothing but the forward's inputs, its output volume and the per-hypothesis weight sums is kept between the two
// passes; taps, correlations and attention weights are recomputed here (the reference's autograd keeps every
// [B,C,D,H,W] intermediate of every view alive instead).
//
//   out[g,d] = A[g,d] / S[d],  A = sum_v w_v[d] cor_v[g,d],  S = 1e-8 + sum_v w_v[d]
//   dA[g,d]  = gout[g,d] / S[d]                      dS[d] = -sum_g gout[g,d] out[g,d] / S[d]
//   dw_v[d]  = sum_g dA[g,d] cor_v[g,d] + dS[d]      w_v = softmax_d(score_v / T) / sqrt(C), score_v = sum_g cor_v
//   dscore_v[d] = p_v[d] (dp[d] - sum_d' p_v[d'] dp[d']) / T,   dp = dw_v / sqrt(C)
//   dcor_v[g,d] = w_v[d] dA[g,d] + dscore_v[d]
//   cor_v[g,d]  = (G/C) sum_{c in g} ref[c] warped_v[c,d]  ->  dref, dwarped; dwarped is scattered to the 4 taps.
//
// Scatter: one lane owns 8 channels of a pixel, so a tap is two RED.E.ADD.F32x4 vector reductions.  Horizontally
// adjacent pixels of a warp usually share a texel column (lane i's x0+1 is lane i+1's x0): the right-column
// contribution is handed to the neighbouring lane with shuffles and merged there, which removes up to half of
// the reductions before they reach L2 (warp-aggregated atomics).
#include "epi_bwd_common.cuh"

namespace mvster {

// ---------------------------------------------------------------------------------------------------------------------
// TMA-staged backward for the fine stages (fp32, C = 8 / 16, i.e. 32 / 64-byte texels).  ncu on the direct kernel above
// at the training shape: 56 % of the warp samples wait on the LDG.E.256 gathers of the recompute pass (sub-line texels
// through L1, the forward's old problem) while the reductions keep L2 at 47 %.  Here the recompute pass gathers from
// the same swizzled shared-memory box as the forward (one cp.async.bulk.tensor per CTA tile and view, next view
// requested before the current one is consumed), the blending / gradient math is packed fp32x2, and only the scatter
// (RED.E.ADD.F32x4 with the neighbour-lane merge) touches L1/L2.  A tile whose footprint exceeds the box gathers
// directly for that view.
// ---------------------------------------------------------------------------------------------------------------------
struct EpiBwdTmaParams {
    CUtensorMap tmap[MVSTER_MAX_SRC_VIEWS];
    EpiBwdParams q;
};

// CTAs per SM: 32-byte texels run best at 4 (128 registers, ~70 bytes of spills: stage 4 0.226 -> 0.216 ms), 64-byte
// texels at 3 (at 4: stage 3 0.121 -> 0.129 ms)
#ifndef MVSTER_BWD_TMA_MINB
#define MVSTER_BWD_TMA_MINB(C) ((C) == 8 ? 4 : 3)
#endif

// neighbour-lane merge of the right tap column in the TMA kernel (1 = on)
#ifndef MVSTER_BWD_TMA_NBR
#define MVSTER_BWD_TMA_NBR 1
#endif

template <int C, int CPG, int D>
__global__ void __launch_bounds__(128, MVSTER_BWD_TMA_MINB(C))
    epi_bwd_tma_kernel(const __grid_constant__ EpiBwdTmaParams pp) {
    const EpiBwdParams& p = pp.q;
    constexpr int L = C / 8, PPW = 32 / L, GPL = 8 / CPG, G = C / CPG, NT = 128;
    constexpr int TILE_H = 4 / L, TB = C * 4;  // 4 warps: L side by side (32 pixels), 4 / L rows
    constexpr int BW = TmaGeom<C>::BW, BH = 4 + TmaGeom<C>::BH_EXTRA;  // the forward's box (and tensor maps)
    constexpr int BUF_BYTES = BW * BH * TB, ROW_BYTES = BW * TB;
    constexpr uint32_t SWZ = (uint32_t)(TB / 16 - 1) << 4;
    static_assert(C == 8 || C == 16, "TMA backward: 32/64-byte texels");

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ctl = smem_base + 2u * BUF_BYTES;
    unsigned char* ctl_ptr = smem_raw + (ctl - smem_u32(smem_raw));
    int* bbox = reinterpret_cast<int*>(ctl_ptr + 16);
    float* rt_s = reinterpret_cast<float*>(ctl_ptr + TmaGeom<C>::CTL_BYTES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = lane % L, pix = lane / L;
    const int b = blockIdx.z;
    for (int i = tid; i < p.Nsrc * 12; i += NT) rt_s[i] = __ldg(p.rt + (size_t)b * p.Nsrc * 12 + i);
    int x = blockIdx.x * 32 + (warp % L) * PPW + pix;
    int y = blockIdx.y * TILE_H + warp / L;
    const bool live = (x < p.W) && (y < p.H);
    x = min(x, p.W - 1);  // dead lanes shadow a valid pixel: convergent shuffles, unchanged bounding box, zero gradients
    y = min(y, p.H - 1);
    if (tid == 0) {
        mbar_init(ctl, 1);
        mbar_init(ctl + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            bbox[i * 4 + 0] = INT_MAX; bbox[i * 4 + 1] = INT_MAX;
            bbox[i * 4 + 2] = INT_MIN; bbox[i * 4 + 3] = INT_MIN;
        }
    }
    __syncthreads();

    const size_t plane = (size_t)p.H * p.W;
    const size_t pix_off = (size_t)y * p.W + x;

    f32x2 rf[4], gref[4];
    {
        const P8 r = load_pairs<float>(reinterpret_cast<const float*>(p.ref) + (((size_t)b * plane + pix_off) * C + sub * 8));
        const f32x2 sc = pack2(1.0f / CPG, 1.0f / CPG);
#pragma unroll
        for (int q = 0; q < 4; ++q) { rf[q] = mul2(r.q[q], sc); gref[q] = pack2(0.0f, 0.0f); }
    }
    float hyp[D];
#pragma unroll
    for (int d = 0; d < D; ++d) hyp[d] = ldg_stream(p.hypo + ((size_t)b * D + d) * plane + pix_off);

    const float fxp = (float)x, fyp = (float)y;
    const float wlim = (float)p.Ws, hlim = (float)p.Hs;
    const size_t src_batch = (size_t)b * p.Hs * p.Ws * C + sub * 8;
    const bool has_prev = pix > 0, has_next = pix < PPW - 1;

    float nsx[D], nsy[D];
    int nbx = 0, nby = 0;
    bool nfit = false;
    uint32_t uses0 = 0, uses1 = 0;

    auto stage_view = [&](int v) {
        const Homography h = homography_from_smem(rt_s + v * 12);
        const float ax = fmaf(h.r00, fxp, fmaf(h.r01, fyp, h.r02));
        const float ay = fmaf(h.r10, fxp, fmaf(h.r11, fyp, h.r12));
        const float az = fmaf(h.r20, fxp, fmaf(h.r21, fyp, h.r22));
#pragma unroll
        for (int d = 0; d < D; ++d) sample_pos(ax, ay, az, h, hyp[d], wlim, hlim, nsx[d], nsy[d]);
        float lox = nsx[0], hix = nsx[0], loy = nsy[0], hiy = nsy[0];
#pragma unroll
        for (int d = 1; d < D; ++d) {
            lox = fminf(lox, nsx[d]); hix = fmaxf(hix, nsx[d]);
            loy = fminf(loy, nsy[d]); hiy = fmaxf(hiy, nsy[d]);
        }
        const int slot = v % 3;
        bbox_update(&bbox[slot * 4], lox, loy, hix, hiy, lane);
        if (tid == 0) {
            const int nx = (v + 1) % 3;
            bbox[nx * 4 + 0] = INT_MAX; bbox[nx * 4 + 1] = INT_MAX;
            bbox[nx * 4 + 2] = INT_MIN; bbox[nx * 4 + 3] = INT_MIN;
        }
        __syncthreads();  // bbox complete; every thread is done with buffer v&1 (view v-2)
        const int4 bb = *reinterpret_cast<const int4*>(&bbox[slot * 4]);
        nbx = bb.x; nby = bb.y;
        nfit = (bb.z - bb.x + 2 <= BW) && (bb.w - bb.y + 2 <= BH);
        if (nfit && tid == 0) {
            const uint32_t bar = ctl + 8u * (v & 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (uint32_t)BUF_BYTES);
            tma_load_4d(smem_base + (uint32_t)BUF_BYTES * (v & 1), &pp.tmap[v], bar, 0, nbx, nby, b);
        }
    };

    stage_view(0);  // needs the hypotheses only: the first box is in flight while the gradient planes are read

    float dS[D], dA[GPL][D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float inv_s = 1.0f / ldg_stream(p.wsum + ((size_t)b * D + d) * plane + pix_off);
        float part = 0.0f;
#pragma unroll
        for (int g = 0; g < GPL; ++g) {
            const size_t o = (((size_t)b * G + sub * GPL + g) * D + d) * plane + pix_off;
            const float go = live ? ldg_stream(p.gout + o) : 0.0f;
            const float ov = ldg_stream(p.out + o);
            dA[g][d] = go * inv_s;
            part = fmaf(go, ov, part);
        }
#pragma unroll
        for (int m = 1; m < L; m <<= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
        dS[d] = -part * inv_s;
    }

#pragma unroll 1
    for (int v = 0; v < p.Nsrc; ++v) {
        float sx[D], sy[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { sx[d] = nsx[d]; sy[d] = nsy[d]; }
        const int bx = nbx, by = nby;
        const bool fit = nfit;
        if (v + 1 < p.Nsrc) stage_view(v + 1);

        const float* srcp = reinterpret_cast<const float*>(p.src[v]) + src_batch;
        float* gsrc = p.grad_src[v] + src_batch;
        uint32_t buf = 0;
        if (fit) {
            const uint32_t parity = ((v & 1) ? uses1 : uses0) & 1u;
            mbar_wait(ctl + 8u * (v & 1), parity);
            if (v & 1) ++uses1; else ++uses0;
            buf = smem_base + (uint32_t)BUF_BYTES * (v & 1) + 32u * sub;
        }

        // ---- pass A: recompute warped features, correlations, attention ------------------------------------------
        P8 wv[D];
        float cor[GPL][D], score[D], dw[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float x0f, y0f;
            int x0, y0;
            floor_fi(sx[d], x0f, x0);
            floor_fi(sy[d], y0f, y0);
            const float fx = sx[d] - x0f, fy = sy[d] - y0f;
            P8 t00, t01, t10, t11;
            float w00, w01, w10, w11;
            if (fit) {
                const uint32_t xo = (uint32_t)(x0 - bx) * TB;
                const uint32_t base = buf + (uint32_t)(y0 - by) * ROW_BYTES + xo;
                const uint32_t mA = (xo >> 3) & SWZ, mB = ((xo + TB) >> 3) & SWZ;
                const uint32_t aL = base ^ mA, aR = (base + TB) ^ mB;
                lds_chunk8<float>(aL, t00);
                lds_chunk8<float>(aR, t01);
                lds_chunk8<float>(aL + ROW_BYTES, t10);
                lds_chunk8<float>(aR + ROW_BYTES, t11);
                const float gx = 1.0f - fx, gy = 1.0f - fy;
                w00 = gx * gy; w01 = fx * gy; w10 = gx * fy; w11 = fx * fy;
            } else {
                const bool vx0 = (unsigned)x0 < (unsigned)p.Ws, vx1 = (unsigned)(x0 + 1) < (unsigned)p.Ws;
                const bool vy0 = (unsigned)y0 < (unsigned)p.Hs, vy1 = (unsigned)(y0 + 1) < (unsigned)p.Hs;
                const int xc0 = min(max(x0, 0), p.Ws - 1), xc1 = min(x0 + 1, p.Ws - 1);
                const int yc0 = min(max(y0, 0), p.Hs - 1), yc1 = min(y0 + 1, p.Hs - 1);
                const float gx = vx0 ? 1.0f - fx : 0.0f, hx = vx1 ? fx : 0.0f;
                const float gy = vy0 ? 1.0f - fy : 0.0f, hy = vy1 ? fy : 0.0f;
                w00 = gx * gy; w01 = hx * gy; w10 = gx * hy; w11 = hx * hy;
                t00 = load_pairs<float>(srcp + (size_t)(yc0 * p.Ws + xc0) * C);
                t01 = load_pairs<float>(srcp + (size_t)(yc0 * p.Ws + xc1) * C);
                t10 = load_pairs<float>(srcp + (size_t)(yc1 * p.Ws + xc0) * C);
                t11 = load_pairs<float>(srcp + (size_t)(yc1 * p.Ws + xc1) * C);
            }
            const f32x2 p00 = pack2(w00, w00), p01 = pack2(w01, w01), p10 = pack2(w10, w10), p11 = pack2(w11, w11);
            f32x2 prod[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                f32x2 w = mul2(p00, t00.q[q]);
                w = fma2(p01, t01.q[q], w);
                w = fma2(p10, t10.q[q], w);
                w = fma2(p11, t11.q[q], w);
                wv[d].q[q] = w;
                prod[q] = mul2(rf[q], w);
            }
            float s = 0.0f, g_dot = 0.0f;
#pragma unroll
            for (int g = 0; g < GPL; ++g) {
                const float cg = group_sum<CPG>(prod, g);
                cor[g][d] = cg;
                s += cg;
                g_dot = fmaf(dA[g][d], cg, g_dot);
            }
            score[d] = s;
            dw[d] = g_dot;
        }
#pragma unroll
        for (int m = 1; m < L; m <<= 1) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                score[d] += __shfl_xor_sync(0xffffffffu, score[d], m);
                dw[d] += __shfl_xor_sync(0xffffffffu, dw[d], m);
            }
        }
        float mx = score[0];
#pragma unroll
        for (int d = 1; d < D; ++d) mx = fmaxf(mx, score[d]);
        float pr[D], es = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            pr[d] = exp2f((score[d] - mx) * p.score_scale);
            es += pr[d];
        }
        const float inv_es = 1.0f / es;
        float dot = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            pr[d] *= inv_es;
            dw[d] = (dw[d] + dS[d]) * p.inv_sqrt_c;  // dp[d]
            dot = fmaf(pr[d], dw[d], dot);
        }

        // ---- pass B: gradients, scatter ----------------------------------------------------------------------------
        // Two merges keep reductions away from L2 (its atomic units, not the gathers, bound this kernel: swapping the
        // LDG gathers for the staged box alone changed nothing): (1) the neighbour exchange hands a sample's right
        // column to the next pixel's lane when that is its left column; (2) consecutive hypotheses of a pixel that
        // stay in the same 2x2 texel cell (sub-texel hypothesis spacing at the fine stages) accumulate in registers
        // and are reduced once per cell instead of once per hypothesis.
        f32x2 aL0[4], aL1[4], aR0[4], aR1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) aL0[q] = aL1[q] = aR0[q] = aR1[q] = pack2(0.0f, 0.0f);
        bool fL0 = false, fL1 = false, fR0 = false, fR1 = false;
        int px0 = 0, py0 = 0;
        auto flush = [&]() {
            float* cell = gsrc + (size_t)(py0 * p.Ws + px0) * C;
            float r[8];
            if (fL0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) unpack2(aL0[q], r[2 * q], r[2 * q + 1]);
                red8(cell, r);
            }
            if (fR0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) unpack2(aR0[q], r[2 * q], r[2 * q + 1]);
                red8(cell + C, r);
            }
            if (fL1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) unpack2(aL1[q], r[2 * q], r[2 * q + 1]);
                red8(cell + (size_t)p.Ws * C, r);
            }
            if (fR1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) unpack2(aR1[q], r[2 * q], r[2 * q + 1]);
                red8(cell + (size_t)p.Ws * C + C, r);
            }
        };
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float dscore = p.inv_temp * pr[d] * (dw[d] - dot);
            const float w = pr[d] * p.inv_sqrt_c;
            float dc[GPL];
#pragma unroll
            for (int g = 0; g < GPL; ++g) dc[g] = fmaf(w, dA[g][d], dscore);  // dL/dcor_v[g,d]
            f32x2 dwp[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const f32x2 dcq = pack2(dc[(2 * q) / CPG], dc[(2 * q + 1) / CPG]);
                gref[q] = fma2(mul2(dcq, pack2(1.0f / CPG, 1.0f / CPG)), wv[d].q[q], gref[q]);
                dwp[q] = mul2(dcq, rf[q]);  // rf is pre-scaled by 1/CPG
            }
            float x0f, y0f;
            int x0, y0;
            floor_fi(sx[d], x0f, x0);
            floor_fi(sy[d], y0f, y0);
            const float fx = sx[d] - x0f, fy = sy[d] - y0f;
            const bool vx0 = (unsigned)x0 < (unsigned)p.Ws, vx1 = (unsigned)(x0 + 1) < (unsigned)p.Ws;
            const bool vy0 = (unsigned)y0 < (unsigned)p.Hs, vy1 = (unsigned)(y0 + 1) < (unsigned)p.Hs;
            const float gx = (vx0 && live) ? 1.0f - fx : 0.0f, hx = (vx1 && live) ? fx : 0.0f;
            const float gy = vy0 ? 1.0f - fy : 0.0f, hy = vy1 ? fy : 0.0f;
            const float wl0 = gx * gy, wr0 = hx * gy, wl1 = gx * hy, wr1 = hx * hy;
            // texel offsets (compared between lanes only where the weight is non-zero, i.e. in bounds)
            const int o00 = y0 * p.Ws + x0, o01 = o00 + 1, o10 = o00 + p.Ws, o11 = o10 + 1;
#if MVSTER_BWD_TMA_NBR
            const int nxt_ol0 = __shfl_down_sync(0xffffffffu, o00, L);
            const int nxt_ol1 = __shfl_down_sync(0xffffffffu, o10, L);
            const float nxt_wl0 = __shfl_down_sync(0xffffffffu, wl0, L);
            const float nxt_wl1 = __shfl_down_sync(0xffffffffu, wl1, L);
            const int prv_or0 = __shfl_up_sync(0xffffffffu, o01, L);
            const int prv_or1 = __shfl_up_sync(0xffffffffu, o11, L);
            const float prv_wr0 = __shfl_up_sync(0xffffffffu, wr0, L);
            const float prv_wr1 = __shfl_up_sync(0xffffffffu, wr1, L);
            const bool give0 = has_next && wr0 != 0.0f && nxt_wl0 != 0.0f && nxt_ol0 == o01;
            const bool give1 = has_next && wr1 != 0.0f && nxt_wl1 != 0.0f && nxt_ol1 == o11;
            const bool take0 = has_prev && prv_wr0 != 0.0f && wl0 != 0.0f && prv_or0 == o00;
            const bool take1 = has_prev && prv_wr1 != 0.0f && wl1 != 0.0f && prv_or1 == o10;
            const float tk0 = take0 ? prv_wr0 : 0.0f, tk1 = take1 ? prv_wr1 : 0.0f;
            const float kr0 = give0 ? 0.0f : wr0, kr1 = give1 ? 0.0f : wr1;  // right-column weights kept by this lane
#else
            const float kr0 = wr0, kr1 = wr1;
#endif
            if (d > 0 && (x0 != px0 || y0 != py0)) {
                flush();
                fL0 = fL1 = fR0 = fR1 = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) aL0[q] = aL1[q] = aR0[q] = aR1[q] = pack2(0.0f, 0.0f);
            }
            px0 = x0; py0 = y0;
            const f32x2 pl0 = pack2(wl0, wl0), pl1 = pack2(wl1, wl1);
            const f32x2 pr0 = pack2(kr0, kr0), pr1 = pack2(kr1, kr1);
#if MVSTER_BWD_TMA_NBR
            const f32x2 pt0 = pack2(tk0, tk0), pt1 = pack2(tk1, tk1);
#endif
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#if MVSTER_BWD_TMA_NBR
                float lo, hi;
                unpack2(dwp[q], lo, hi);
                const f32x2 prv = pack2(__shfl_up_sync(0xffffffffu, lo, L), __shfl_up_sync(0xffffffffu, hi, L));
                aL0[q] = fma2(pl0, dwp[q], fma2(pt0, prv, aL0[q]));
                aL1[q] = fma2(pl1, dwp[q], fma2(pt1, prv, aL1[q]));
#else
                aL0[q] = fma2(pl0, dwp[q], aL0[q]);
                aL1[q] = fma2(pl1, dwp[q], aL1[q]);
#endif
                aR0[q] = fma2(pr0, dwp[q], aR0[q]);
                aR1[q] = fma2(pr1, dwp[q], aR1[q]);
            }
            // a column is live once it has received a non-zero (hence in-bounds) weight; later zero-weight samples of
            // the same cell add exact zeros
            fL0 = fL0 || wl0 != 0.0f; fL1 = fL1 || wl1 != 0.0f;
            fR0 = fR0 || kr0 != 0.0f; fR1 = fR1 || kr1 != 0.0f;
        }
        flush();
    }

    if (live) {
        float* gp = p.grad_ref + (((size_t)b * plane + pix_off) * C + sub * 8);
        float g[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) unpack2(gref[q], g[2 * q], g[2 * q + 1]);
        float4* g4 = reinterpret_cast<float4*>(gp);
        g4[0] = make_float4(g[0], g[1], g[2], g[3]);
        g4[1] = make_float4(g[4], g[5], g[6], g[7]);
    }
}

template <int C, int CPG, int D>
static int launch_bwd_tma(const EpiBwdTmaParams& pp, cudaStream_t stream) {
    constexpr int NT = 128, TB = C * 4, TILE_H = 4 / (C / 8);
    constexpr int SMEM = 2 * TmaGeom<C>::BW * (4 + TmaGeom<C>::BH_EXTRA) * TB + 1024 + TmaGeom<C>::CTL_BYTES +
                         MVSTER_MAX_SRC_VIEWS * 48;
    static int attr_done[64] = {};  // largest size set per device
    if (SMEM > 48 * 1024) {
        const int st = ensure_dynamic_smem_bytes(epi_bwd_tma_kernel<C, CPG, D>, SMEM, attr_done, "epi_bwd(tma): cudaFuncSetAttribute");
        if (st != MVSTER_OK) return st;
    }
    const EpiBwdParams& p = pp.q;
    dim3 grid((p.W + 31) / 32, (p.H + TILE_H - 1) / TILE_H, p.B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd: grid too large");
    epi_bwd_tma_kernel<C, CPG, D><<<grid, NT, SMEM, stream>>>(pp);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_bwd(tma) launch");
    return MVSTER_OK;
}

// fine stages in fp32: TMA-staged kernel when the tensor maps can be built; returns -1 when not applicable
template <int C>
static int try_bwd_tma(const EpiBwdParams& p, int cpg, int D, cudaStream_t s) {
    if (D != 4 || getenv("MVSTER_NO_TMA") != nullptr || getenv("MVSTER_NO_TMA_BWD") != nullptr) return -1;
    static_assert(Split<C, 1, 4>::TILE_H == 4, "the backward tile reuses the forward's tensor-map box");
    static thread_local EpiBwdTmaParams pp;  // 64-byte aligned tensor maps; one per calling thread
    pp.q = p;
    if (!make_maps<C, 1, 4, float>(pp.tmap, p.src, p.Nsrc, p.B, p.Hs, p.Ws)) return -1;
    switch (cpg) {
        case 1: return launch_bwd_tma<C, 1, 4>(pp, s);
        case 2: return launch_bwd_tma<C, 2, 4>(pp, s);
        case 4: return launch_bwd_tma<C, 4, 4>(pp, s);
        case 8: return launch_bwd_tma<C, 8, 4>(pp, s);
        default: return -1;
    }
}

template <int C, int CPG, int D, typename T>
static int launch_bwd(const EpiBwdParams& p, cudaStream_t stream) {
    constexpr int DS = (D == 8) ? MVSTER_BWD_DSPLIT8 : 1;
    constexpr int PPW = 32 / ((C / 8) * DS);
    dim3 grid((p.W + PPW - 1) / PPW, (p.H + kBwdWarps - 1) / kBwdWarps, p.B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd: grid too large");
    epi_bwd_kernel<C, CPG, D, T, DS><<<grid, kBwdWarps * 32, 0, stream>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_bwd launch");
    return MVSTER_OK;
}

template <int C, int CPG, typename T>
static int bwd_d(const EpiBwdParams& p, int D, cudaStream_t s) {
    switch (D) {
        case 4: return launch_bwd<C, CPG, 4, T>(p, s);
        case 8: return launch_bwd<C, CPG, 8, T>(p, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd: D=%d not in {4,8}", D);
    }
}

template <int C, typename T>
static int bwd_cpg(const EpiBwdParams& p, int cpg, int D, cudaStream_t s) {
    switch (cpg) {
        case 1: return bwd_d<C, 1, T>(p, D, s);
        case 2: return bwd_d<C, 2, T>(p, D, s);
        case 4: return bwd_d<C, 4, T>(p, D, s);
        case 8: return bwd_d<C, 8, T>(p, D, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd: C/G=%d not in {1,2,4,8}", cpg);
    }
}

template <typename T>
static int bwd_c(const EpiBwdParams& p, int C, int cpg, int D, cudaStream_t s) {
    switch (C) {
        case 8: return bwd_cpg<8, T>(p, cpg, D, s);
        case 16: return bwd_cpg<16, T>(p, cpg, D, s);
        case 32: return bwd_cpg<32, T>(p, cpg, D, s);
        case 64: return bwd_cpg<64, T>(p, cpg, D, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd: C=%d not in {8,16,32,64}", C);
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_epi_bwd(const void* ref, const void* const* src, const float* rt, const float* hypo,
                              const float* out, const float* wsum, const float* gout, float* grad_ref,
                              float* const* grad_src, int B, int Nsrc, int C, int G, int D, int H, int W, int Hs,
                              int Ws, float attn_temp, int dtype, void* stream) {
    if (!ref || !src || !rt || !hypo || !out || !wsum || !gout || !grad_ref || !grad_src)
        return fail(MVSTER_ERR_BAD_ARG, "epi_bwd: null pointer");
    if (B <= 0 || Nsrc <= 0 || C <= 0 || G <= 0 || D <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0)
        return fail(MVSTER_ERR_BAD_ARG, "epi_bwd: non-positive dimension");
    if (Nsrc > MVSTER_MAX_SRC_VIEWS)
        return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd: %d source views > MVSTER_MAX_SRC_VIEWS=%d", Nsrc,
                    MVSTER_MAX_SRC_VIEWS);
    if (C % G != 0) return fail(MVSTER_ERR_BAD_ARG, "epi_bwd: C=%d not divisible by G=%d", C, G);
    if (!(attn_temp > 0.0f)) return fail(MVSTER_ERR_BAD_ARG, "epi_bwd: attn_temp must be > 0");
    if ((double)B * Hs * Ws * C >= 2147483648.0 || (double)H * W >= 2147483648.0)
        return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd: tensor too large for 32-bit texel offsets");
    const uintptr_t align = (dtype == MVSTER_BF16) ? 16 : 32;
    if (((uintptr_t)ref) % align) return fail(MVSTER_ERR_ALIGN, "epi_bwd: ref not %d-byte aligned", (int)align);
    if (((uintptr_t)grad_ref) % 16) return fail(MVSTER_ERR_ALIGN, "epi_bwd: grad_ref not 16-byte aligned");
    if (((uintptr_t)rt) % 16) return fail(MVSTER_ERR_ALIGN, "epi_bwd: rt not 16-byte aligned");
    EpiBwdParams p{};
    p.ref = ref;
    for (int v = 0; v < Nsrc; ++v) {
        if (!src[v] || !grad_src[v]) return fail(MVSTER_ERR_BAD_ARG, "epi_bwd: src/grad_src[%d] is null", v);
        if (((uintptr_t)src[v]) % align)
            return fail(MVSTER_ERR_ALIGN, "epi_bwd: src[%d] not %d-byte aligned", v, (int)align);
        if (((uintptr_t)grad_src[v]) % 16) return fail(MVSTER_ERR_ALIGN, "epi_bwd: grad_src[%d] not 16-byte aligned", v);
        p.src[v] = src[v];
        p.grad_src[v] = grad_src[v];
    }
    p.rt = rt; p.hypo = hypo; p.out = out; p.wsum = wsum; p.gout = gout; p.grad_ref = grad_ref;
    p.B = B; p.Nsrc = Nsrc; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws;
    p.score_scale = 1.4426950408889634f / attn_temp;
    p.inv_temp = 1.0f / attn_temp;
    p.inv_sqrt_c = (float)(1.0 / sqrt((double)C));
    DeviceGuard guard(grad_ref);
    if (guard.status != MVSTER_OK) return guard.status;
    const int cpg = C / G;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == MVSTER_F32 && (C == 8 || C == 16)) {
        const int st = (C == 8) ? try_bwd_tma<8>(p, cpg, D, s) : try_bwd_tma<16>(p, cpg, D, s);
        if (st >= 0) return st;
    }
    if (dtype == MVSTER_F32) return bwd_c<float>(p, C, cpg, D, s);
    if (dtype == MVSTER_BF16) return bwd_c<__nv_bfloat16>(p, C, cpg, D, s);
    return fail(MVSTER_ERR_BAD_ARG, "epi_bwd: unknown dtype %d", dtype);
}
