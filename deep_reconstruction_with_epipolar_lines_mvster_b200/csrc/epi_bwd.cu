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
#include "common.cuh"

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

__device__ __forceinline__ void red8(float* p, const float* v) {
    red_add_v4(p, v[0], v[1], v[2], v[3]);
    red_add_v4(p + 4, v[4], v[5], v[6], v[7]);
}

template <int C, int CPG, int D, typename T>
__global__ void __launch_bounds__(kBwdWarps * 32, (D > 4) ? 2 : 3)
    epi_bwd_kernel(const __grid_constant__ EpiBwdParams p) {
    constexpr int CPL = 8;
    constexpr int L = C / CPL;
    constexpr int GPL = CPL / CPG;
    constexpr int PPW = 32 / L;
    constexpr int G = C / CPG;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane % L;
    const int pix = lane / L;
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

    float hyp[D], dS[D], dA[GPL][D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        hyp[d] = ldg_stream(p.hypo + ((size_t)b * D + d) * plane + pix_off);
        const float inv_s = 1.0f / ldg_stream(p.wsum + ((size_t)b * D + d) * plane + pix_off);
        float part = 0.0f;
#pragma unroll
        for (int g = 0; g < GPL; ++g) {
            const size_t o = (((size_t)b * G + sub * GPL + g) * D + d) * plane + pix_off;
            const float go = live ? ldg_stream(p.gout + o) : 0.0f;  // dead lanes contribute nothing
            const float ov = ldg_stream(p.out + o);
            dA[g][d] = go * inv_s;
            part = fmaf(go, ov, part);
        }
#pragma unroll
        for (int m = 1; m < L; m <<= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
        dS[d] = -part * inv_s;
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
        float wv[D][CPL];
        float cor[GPL][D];
        float score[D], dw[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
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
#pragma unroll
                for (int c = 0; c < CPG; ++c) cg = fmaf(rf.v[g * CPG + c], wv[d][g * CPG + c], cg);
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

        // ---- pass B: gradients, scatter ----------------------------------------------------------------------
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float dscore = p.inv_temp * pr[d] * (dw[d] - dot);
            const float w = pr[d] * p.inv_sqrt_c;
            float dwarp[CPL];
#pragma unroll
            for (int g = 0; g < GPL; ++g) {
                const float dc = fmaf(w, dA[g][d], dscore);  // dL/dcor_v[g,d]
#pragma unroll
                for (int c = 0; c < CPG; ++c) {
                    const int cc = g * CPG + c;
                    gref[cc] = fmaf(dc * (1.0f / CPG), wv[d][cc], gref[cc]);
                    dwarp[cc] = dc * rf.v[cc];  // rf is pre-scaled by 1/CPG
                }
            }
            const Taps t = make_taps(ax, ay, az, h, hyp[d], p.Hs, p.Ws);
            // active taps: non-zero weight (implies in-bounds) on a live lane
            const float wl0 = live ? t.w00 : 0.0f, wr0 = live ? t.w01 : 0.0f;
            const float wl1 = live ? t.w10 : 0.0f, wr1 = live ? t.w11 : 0.0f;
            // neighbour exchange: lane+L is the next pixel (same channel chunk); hand my right column to it when
            // it is that lane's active left column, and take the previous pixel's right column likewise.
            const int nxt_ol0 = __shfl_down_sync(0xffffffffu, t.o00, L);
            const int nxt_ol1 = __shfl_down_sync(0xffffffffu, t.o10, L);
            const float nxt_wl0 = __shfl_down_sync(0xffffffffu, wl0, L);
            const float nxt_wl1 = __shfl_down_sync(0xffffffffu, wl1, L);
            const int prv_or0 = __shfl_up_sync(0xffffffffu, t.o01, L);
            const int prv_or1 = __shfl_up_sync(0xffffffffu, t.o11, L);
            const float prv_wr0 = __shfl_up_sync(0xffffffffu, wr0, L);
            const float prv_wr1 = __shfl_up_sync(0xffffffffu, wr1, L);
            const bool give0 = has_next && wr0 != 0.0f && nxt_wl0 != 0.0f && nxt_ol0 == t.o01;
            const bool give1 = has_next && wr1 != 0.0f && nxt_wl1 != 0.0f && nxt_ol1 == t.o11;
            const bool take0 = has_prev && prv_wr0 != 0.0f && wl0 != 0.0f && prv_or0 == t.o00;
            const bool take1 = has_prev && prv_wr1 != 0.0f && wl1 != 0.0f && prv_or1 == t.o10;
            const float tk0 = take0 ? prv_wr0 : 0.0f, tk1 = take1 ? prv_wr1 : 0.0f;
            float left0[CPL], left1[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const float prv = __shfl_up_sync(0xffffffffu, dwarp[c], L);
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

    if (live) {
        float* gp = p.grad_ref + (((size_t)b * plane + pix_off) * C + sub * CPL);
        float4* g4 = reinterpret_cast<float4*>(gp);
        g4[0] = make_float4(gref[0], gref[1], gref[2], gref[3]);
        g4[1] = make_float4(gref[4], gref[5], gref[6], gref[7]);
    }
}

template <int C, int CPG, int D, typename T>
static int launch_bwd(const EpiBwdParams& p, cudaStream_t stream) {
    constexpr int PPW = 32 / (C / 8);
    dim3 grid((p.W + PPW - 1) / PPW, (p.H + kBwdWarps - 1) / kBwdWarps, p.B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd: grid too large");
    epi_bwd_kernel<C, CPG, D, T><<<grid, kBwdWarps * 32, 0, stream>>>(p);
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
    if (dtype == MVSTER_F32) return bwd_c<float>(p, C, cpg, D, s);
    if (dtype == MVSTER_BF16) return bwd_c<__nv_bfloat16>(p, C, cpg, D, s);
    return fail(MVSTER_ERR_BAD_ARG, "epi_bwd: unknown dtype %d", dtype);
}
