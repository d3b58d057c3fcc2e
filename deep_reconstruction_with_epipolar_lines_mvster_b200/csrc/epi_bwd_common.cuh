// K1 backward, shared between epi_bwd.cu (the shipped configuration) and epi_bwd_alt.cu (the reference's
// group_cor=False / attn_fuse_d=False options): parameter block and the direct-gather kernel.  See epi_bwd.cu for the
// derivation; the two option flags change it as follows.
//   VAR (group_cor=False, models/mvs4net_utils.py:1071): cor_v[c,d] = (ref[c] - warped_v[c,d])^2, G == C
//       dref[c] += 2 dcor (ref - warped),  dwarped[c,d] = -2 dcor (ref - warped)
//   !FUSE_D (attn_fuse_d=False, :1078-1081,1098): one weight per pixel and view, w_v = max_d softmax_d(score_v)
//       (no temperature, no sqrt(C)); S = 1e-8 + sum_v w_v is per pixel (wsum is [B,H,W]);
//       dA[g,d] = gout / S,  dS = -sum_{g,d} gout out / S,  dw_v = sum_{g,d} dA cor_v + dS,
//       dscore_v[d] = dw_v w_v ([d == argmax] - p_v[d]),  dcor_v[g,d] = w_v dA[g,d] + dscore_v[d]
#pragma once
#include <stdlib.h>

#include "epi_tma.cuh"

namespace mvster {

struct EpiBwdParams {
    const void* ref;
    const void* src[MVSTER_MAX_SRC_VIEWS];
    float* grad_src[MVSTER_MAX_SRC_VIEWS];
    const float* rt;
    const float* hypo;
    const float* out;
    const float* wsum;
    const float* gout;
    float* grad_ref;
    int B, Nsrc, H, W, Hs, Ws;
    float score_scale;  // log2(e) / attn_temp
    float inv_temp;     // 1 / attn_temp
    float inv_sqrt_c;
};

constexpr int kBwdWarps = 4;

#ifndef MVSTER_BWD_KO
#define MVSTER_BWD_KO 0   // development only, WRONG RESULTS: 1 = no reduction reaches memory (the arithmetic stays alive)
#endif
__device__ __forceinline__ void red8(float* p, const float* v) {
    if (MVSTER_BWD_KO && reinterpret_cast<uintptr_t>(p) != 1) return;
    red_add_v4(p, v[0], v[1], v[2], v[3]);
    red_add_v4(p + 4, v[4], v[5], v[6], v[7]);
}

#ifndef MVSTER_BWD_MINB
#define MVSTER_BWD_MINB 3
#endif
#ifndef MVSTER_BWD_MINB8
#define MVSTER_BWD_MINB8 3  // D = 8 (coarse stages): 168 registers + 16-48 bytes of spills at 12 warps per SM beat 248 registers at 8 (stage 2: 0.182 -> 0.152 ms)
#endif
// Hypothesis split of the direct kernel: DS lanes share a pixel's channel chunk and each owns D / DS hypotheses.  At
// the coarse stages (D = 8, 1/8 and 1/4 resolution) a launch is only a few waves of long serial per-lane chains
// (32 samples x ~400 instructions) at 12 warps per SM; splitting halves the chain and the hypothesis-indexed register
// arrays, so twice the warps run at a higher occupancy.  The softmax statistics and the reference gradient are
// combined across the DS lanes with shuffles.
#ifndef MVSTER_BWD_DSPLIT8
#define MVSTER_BWD_DSPLIT8 2
#endif
#ifndef MVSTER_BWD_MINB_SPLIT
#define MVSTER_BWD_MINB_SPLIT 4
#endif
template <int C, int CPG, int D, typename T, int DS, bool VAR = false, bool FUSE_D = true>
__global__ void __launch_bounds__(kBwdWarps * 32, DS > 1 ? MVSTER_BWD_MINB_SPLIT : ((D > 4) ? MVSTER_BWD_MINB8 : MVSTER_BWD_MINB))
    epi_bwd_kernel(const __grid_constant__ EpiBwdParams p) {
    static_assert(!VAR || CPG == 1, "the variance cost has one output channel per feature channel");
    static_assert((!VAR && FUSE_D) || DS == 1, "the option variants run without the hypothesis split");
    constexpr int CPL = 8;
    constexpr int L = C / CPL;        // lanes per channel sweep
    constexpr int LP = L * DS;        // lanes per pixel
    constexpr int DL = D / DS;        // hypotheses per lane
    constexpr int GPL = CPL / CPG;
    constexpr int PPW = 32 / LP;
    constexpr int G = C / CPG;
    static_assert(D % DS == 0 && LP <= 32, "bad hypothesis split");

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane % L;
    const int ds = (lane / L) % DS;
    const int pix = lane / LP;
    const int b = blockIdx.z;
    int x = blockIdx.x * PPW + pix;
    int y = blockIdx.y * kBwdWarps + warp;
    const bool live = (x < p.W) && (y < p.H);
    x = min(x, p.W - 1);
    y = min(y, p.H - 1);
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix_off = (size_t)y * p.W + x;

    const T* refp = reinterpret_cast<const T*>(p.ref) + (((size_t)b * plane + pix_off) * C + sub * CPL);
    F8 rf = load8<T>(refp);
#pragma unroll
    for (int c = 0; c < CPL; ++c) rf.v[c] *= (1.0f / CPG);
    float gref[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) gref[c] = 0.0f;

    float hyp[DL], dS[DL], dA[GPL][DL];
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        const int dg = ds * DL + d;
        hyp[d] = ldg_stream(p.hypo + ((size_t)b * D + dg) * plane + pix_off);
        // attn_fuse_d=False: one weight sum per pixel, wsum is [B,H,W]
        const float inv_s = 1.0f / ldg_stream(p.wsum + (FUSE_D ? ((size_t)b * D + dg) * plane : (size_t)b * plane) + pix_off);
        float part = 0.0f;
#pragma unroll
        for (int g = 0; g < GPL; ++g) {
            const size_t o = (((size_t)b * G + sub * GPL + g) * D + dg) * plane + pix_off;
            const float go = live ? ldg_stream(p.gout + o) : 0.0f;  // dead lanes contribute nothing
            const float ov = ldg_stream(p.out + o);
            dA[g][d] = go * inv_s;
            part = fmaf(go, ov, part);
        }
#pragma unroll
        for (int m = 1; m < L; m <<= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
        dS[d] = -part * inv_s;
    }
    if constexpr (!FUSE_D) {  // dS = -sum_{g,d} gout out / S, the same for every hypothesis
        float tot = 0.0f;
#pragma unroll
        for (int d = 0; d < DL; ++d) tot += dS[d];
#pragma unroll
        for (int d = 0; d < DL; ++d) dS[d] = tot;
    }

    const float fx = (float)x, fy = (float)y;
    const size_t src_batch = (size_t)b * p.Hs * p.Ws * C + sub * CPL;
    const bool has_prev = pix > 0;
    const bool has_next = pix < PPW - 1;

#pragma unroll 1
    for (int v = 0; v < p.Nsrc; ++v) {
        const Homography h = load_homography(p.rt + ((size_t)b * p.Nsrc + v) * 12);
        const T* srcp = reinterpret_cast<const T*>(p.src[v]) + src_batch;
        float* gsrc = p.grad_src[v] + src_batch;
        const float ax = fmaf(h.r00, fx, fmaf(h.r01, fy, h.r02));
        const float ay = fmaf(h.r10, fx, fmaf(h.r11, fy, h.r12));
        const float az = fmaf(h.r20, fx, fmaf(h.r21, fy, h.r22));

        // ---- pass A: recompute warped features, correlations, attention --------------------------------------
        float wv[DL][CPL];
        float cor[GPL][DL];
        float score[DL], dw[DL];
#pragma unroll
        for (int d = 0; d < DL; ++d) {
            const Taps t = make_taps(ax, ay, az, h, hyp[d], p.Hs, p.Ws);
#pragma unroll
            for (int c = 0; c < CPL; ++c) wv[d][c] = 0.0f;
            if (t.any) {
                const F8 a = load8<T>(srcp + (size_t)t.o00 * C);
                const F8 bq = load8<T>(srcp + (size_t)t.o01 * C);
                const F8 cq = load8<T>(srcp + (size_t)t.o10 * C);
                const F8 dq = load8<T>(srcp + (size_t)t.o11 * C);
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    wv[d][c] = fmaf(t.w00, a.v[c], fmaf(t.w01, bq.v[c], fmaf(t.w10, cq.v[c], t.w11 * dq.v[c])));
            }
            float s = 0.0f, g_dot = 0.0f;
#pragma unroll
            for (int g = 0; g < GPL; ++g) {
                float cg = 0.0f;
                if constexpr (VAR) {
                    const float diff = rf.v[g] - wv[d][g];
                    cg = diff * diff;
                } else {
#pragma unroll
                    for (int c = 0; c < CPG; ++c) cg = fmaf(rf.v[g * CPG + c], wv[d][g * CPG + c], cg);
                }
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
            for (int d = 0; d < DL; ++d) {
                score[d] += __shfl_xor_sync(0xffffffffu, score[d], m);
                dw[d] += __shfl_xor_sync(0xffffffffu, dw[d], m);
            }
        }
        float mx = score[0];
#pragma unroll
        for (int d = 1; d < DL; ++d) mx = fmaxf(mx, score[d]);
#pragma unroll
        for (int m = L; m < LP; m <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        float pr[DL], es = 0.0f;
#pragma unroll
        for (int d = 0; d < DL; ++d) {
            pr[d] = exp2f((score[d] - mx) * p.score_scale);
            es += pr[d];
        }
#pragma unroll
        for (int m = L; m < LP; m <<= 1) es += __shfl_xor_sync(0xffffffffu, es, m);
        const float inv_es = 1.0f / es;
        float dot = 0.0f;
        float wmax = 0.0f, dwv = 0.0f;  // !FUSE_D: the view's weight max_d p[d] and its gradient
        int dstar = 0;
        if constexpr (FUSE_D) {
#pragma unroll
            for (int d = 0; d < DL; ++d) {
                pr[d] *= inv_es;
                dw[d] = (dw[d] + dS[d]) * p.inv_sqrt_c;  // dp[d]
                dot = fmaf(pr[d], dw[d], dot);
            }
#pragma unroll
            for (int m = L; m < LP; m <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
        } else {
            dwv = dS[0];
#pragma unroll
            for (int d = 0; d < DL; ++d) {
                pr[d] *= inv_es;
                dwv += dw[d];
                if (score[d] > score[dstar]) dstar = d;  // first maximum, as torch.max(dim) returns it
            }
            wmax = inv_es;  // exp2(0) / es
        }

        // ---- pass B: gradients, scatter ----------------------------------------------------------------------
#pragma unroll
        for (int d = 0; d < DL; ++d) {
            const float dscore = FUSE_D ? p.inv_temp * pr[d] * (dw[d] - dot)
                                        : dwv * wmax * ((d == dstar ? 1.0f : 0.0f) - pr[d]);
            const float w = FUSE_D ? pr[d] * p.inv_sqrt_c : wmax;
            float dwarp[CPL];
#pragma unroll
            for (int g = 0; g < GPL; ++g) {
                const float dc = fmaf(w, dA[g][d], dscore);  // dL/dcor_v[g,d]
#pragma unroll
                for (int c = 0; c < CPG; ++c) {
                    const int cc = g * CPG + c;
                    if constexpr (VAR) {
                        const float t2 = 2.0f * dc * (rf.v[cc] - wv[d][cc]);
                        gref[cc] += t2;
                        dwarp[cc] = -t2;
                    } else {
                        gref[cc] = fmaf(dc * (1.0f / CPG), wv[d][cc], gref[cc]);
                        dwarp[cc] = dc * rf.v[cc];  // rf is pre-scaled by 1/CPG
                    }
                }
            }
            const Taps t = make_taps(ax, ay, az, h, hyp[d], p.Hs, p.Ws);
            // active taps: non-zero weight (implies in-bounds) on a live lane
            const float wl0 = live ? t.w00 : 0.0f, wr0 = live ? t.w01 : 0.0f;
            const float wl1 = live ? t.w10 : 0.0f, wr1 = live ? t.w11 : 0.0f;
            // neighbour exchange: lane+LP is the next pixel (same channel chunk, same hypotheses); hand my right column to it when
            // it is that lane's active left column, and take the previous pixel's right column likewise.
            const int nxt_ol0 = __shfl_down_sync(0xffffffffu, t.o00, LP);
            const int nxt_ol1 = __shfl_down_sync(0xffffffffu, t.o10, LP);
            const float nxt_wl0 = __shfl_down_sync(0xffffffffu, wl0, LP);
            const float nxt_wl1 = __shfl_down_sync(0xffffffffu, wl1, LP);
            const int prv_or0 = __shfl_up_sync(0xffffffffu, t.o01, LP);
            const int prv_or1 = __shfl_up_sync(0xffffffffu, t.o11, LP);
            const float prv_wr0 = __shfl_up_sync(0xffffffffu, wr0, LP);
            const float prv_wr1 = __shfl_up_sync(0xffffffffu, wr1, LP);
            const bool give0 = has_next && wr0 != 0.0f && nxt_wl0 != 0.0f && nxt_ol0 == t.o01;
            const bool give1 = has_next && wr1 != 0.0f && nxt_wl1 != 0.0f && nxt_ol1 == t.o11;
            const bool take0 = has_prev && prv_wr0 != 0.0f && wl0 != 0.0f && prv_or0 == t.o00;
            const bool take1 = has_prev && prv_wr1 != 0.0f && wl1 != 0.0f && prv_or1 == t.o10;
            const float tk0 = take0 ? prv_wr0 : 0.0f, tk1 = take1 ? prv_wr1 : 0.0f;
            float left0[CPL], left1[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const float prv = __shfl_up_sync(0xffffffffu, dwarp[c], LP);
                left0[c] = fmaf(wl0, dwarp[c], tk0 * prv);
                left1[c] = fmaf(wl1, dwarp[c], tk1 * prv);
            }
            if (wl0 != 0.0f) red8(gsrc + (size_t)t.o00 * C, left0);
            if (wl1 != 0.0f) red8(gsrc + (size_t)t.o10 * C, left1);
            if (wr0 != 0.0f && !give0) {
                float r[CPL];
#pragma unroll
                for (int c = 0; c < CPL; ++c) r[c] = wr0 * dwarp[c];
                red8(gsrc + (size_t)t.o01 * C, r);
            }
            if (wr1 != 0.0f && !give1) {
                float r[CPL];
#pragma unroll
                for (int c = 0; c < CPL; ++c) r[c] = wr1 * dwarp[c];
                red8(gsrc + (size_t)t.o11 * C, r);
            }
        }
    }

#pragma unroll
    for (int m = L; m < LP; m <<= 1)
#pragma unroll
        for (int c = 0; c < CPL; ++c) gref[c] += __shfl_xor_sync(0xffffffffu, gref[c], m);
    if (live && ds == 0) {
        float* gp = p.grad_ref + (((size_t)b * plane + pix_off) * C + sub * CPL);
        float4* g4 = reinterpret_cast<float4*>(gp);
        g4[0] = make_float4(gref[0], gref[1], gref[2], gref[3]);
        g4[1] = make_float4(gref[4], gref[5], gref[6], gref[7]);
    }
}

}  // namespace mvster
