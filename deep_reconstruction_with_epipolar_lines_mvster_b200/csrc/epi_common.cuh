// Pieces shared by the K1 forward translation units (epi_fwd.cu: shipped configuration; epi_fwd_alt.cu: the
// group_cor=False / attn_fuse_d=False variants of the reference, forward only).
#pragma once

#include <cuda.h>
#include <limits.h>

#include "common.cuh"

namespace mvster {

#ifndef MVSTER_DIRECT_WARPS
#define MVSTER_DIRECT_WARPS 4   // warps per CTA of the direct-gather kernel
#endif
#ifndef MVSTER_DIRECT_MINB
#define MVSTER_DIRECT_MINB 4
#endif
#ifndef MVSTER_DIRECT_WXMAX
#define MVSTER_DIRECT_WXMAX 8   // at most this many warps side by side in x
#endif
constexpr int kWarps = MVSTER_DIRECT_WARPS;
constexpr int kThreads = kWarps * 32;

struct EpiFwdParams {
    CUtensorMap tmap[MVSTER_MAX_SRC_VIEWS];  // TMA variant only
    const void* ref;
    const void* src[MVSTER_MAX_SRC_VIEWS];
    const float* rt;
    const float* hypo;
    float* out;
    float* wsum;
    float* weights;
    int B, Nsrc, H, W, Hs, Ws;
    float score_scale;  // log2(e) / attn_temp
    float inv_sqrt_c;   // 1 / sqrt(C)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Horizontal sum of the channels of correlation group g (0 .. 8/CPG-1) of one 8-channel chunk of products.
template <int CPG>
__device__ __forceinline__ float group_sum(const f32x2 (&prod)[4], int g) {
    float lo, hi;
    if constexpr (CPG == 1) {
        unpack2(prod[g >> 1], lo, hi);
        return (g & 1) ? hi : lo;
    } else if constexpr (CPG == 2) {
        unpack2(prod[g], lo, hi);
        return lo + hi;
    } else if constexpr (CPG == 4) {
        unpack2(add2(prod[2 * g], prod[2 * g + 1]), lo, hi);
        return lo + hi;
    } else {
        unpack2(add2(add2(prod[0], prod[1]), add2(prod[2], prod[3])), lo, hi);
        return lo + hi;
    }
}

__device__ __forceinline__ Homography homography_from_smem(const float* rt_s) {
    const float4* hq = reinterpret_cast<const float4*>(rt_s);
    const float4 h0 = hq[0], h1 = hq[1], h2 = hq[2];
    Homography h;
    h.r00 = h0.x; h.r01 = h0.y; h.r02 = h0.z; h.t0 = h0.w;
    h.r10 = h1.x; h.r11 = h1.y; h.r12 = h1.z; h.t1 = h1.w;
    h.r20 = h2.x; h.r21 = h2.y; h.r22 = h2.z; h.t2 = h2.w;
    return h;
}

// blend four taps of one 8-channel chunk, combine with the reference chunk, add the per-group sums.
// VAR = false: group-wise correlation mean_c(ref * warped) (ref pre-scaled by 1/(C/G); reference :1066-1069)
// VAR = true : per-channel variance cost (ref - warped)^2 (group_cor=False, reference :1071; CPG must be 1)
template <int CPG, bool VAR = false>
__device__ __forceinline__ void blend_correlate(const P8& t00, const P8& t01, const P8& t10, const P8& t11, float w00,
                                                float w01, float w10, float w11, const f32x2* rf, float* cor_out) {
    const f32x2 p00 = pack2(w00, w00), p01 = pack2(w01, w01), p10 = pack2(w10, w10), p11 = pack2(w11, w11);
    f32x2 prod[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        f32x2 wv = mul2(p00, t00.q[q]);
        wv = fma2(p01, t01.q[q], wv);
        wv = fma2(p10, t10.q[q], wv);
        wv = fma2(p11, t11.q[q], wv);
        if constexpr (VAR) {
            const f32x2 diff = fma2(wv, pack2(-1.0f, -1.0f), rf[q]);  // ref - warped
            prod[q] = mul2(diff, diff);
        } else {
            prod[q] = mul2(rf[q], wv);  // ref * warped, per channel
        }
    }
#pragma unroll
    for (int g = 0; g < 8 / CPG; ++g) cor_out[g] = group_sum<CPG>(prod, g);
}

// ---------------------------------------------------------------------------------------------------------------------
// DIRECT kernel (any C, fp32 / bf16): L = C/8 lanes per pixel, each owning 8 channels, so that the lanes of a pixel
// read one texel as a single contiguous run - whole 128-byte lines per L1 wavefront for C >= 32.  The sample
// arithmetic is NOT repeated by every lane: lane j of a pixel computes the positions, tap indices and bilinear
// weights of hypotheses d = j, j+L, ... only ("owner"), and the D samples are then broadcast inside the pixel's
// lane group with width-L shuffles (8 values per sample).
// ---------------------------------------------------------------------------------------------------------------------
template <int C, int CPG, int D, typename T, bool VAR = false, bool FUSE_D = true>
__global__ void __launch_bounds__(kThreads, MVSTER_DIRECT_MINB) epi_fwd_direct_kernel(const __grid_constant__ EpiFwdParams p) {
    constexpr int L = C / 8, PPW = 32 / L, GPL = 8 / CPG, G = C / CPG;
    constexpr int NOWN = (D + L - 1) / L;               // samples whose coordinates this lane computes
    constexpr int WXM = MVSTER_DIRECT_WXMAX < kWarps ? MVSTER_DIRECT_WXMAX : kWarps;
    constexpr int WX = L < WXM ? L : WXM, TILE_W = PPW * WX, TILE_H = kWarps / WX;
    constexpr int TB = C * (int)sizeof(T);
    static_assert(L >= 1 && L <= 8 && 8 % CPG == 0, "C in {8,16,32,64}, C/G in {1,2,4,8}");
    static_assert(!VAR || CPG == 1, "the variance cost has one output channel per feature channel");

    extern __shared__ unsigned char smem_raw[];
    float* rt_s = reinterpret_cast<float*>(smem_raw);  // [Nsrc][12]
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int pix = lane / L, sub = lane % L;
    const int b = blockIdx.z;
    for (int i = tid; i < p.Nsrc * 12; i += kThreads) rt_s[i] = __ldg(p.rt + (size_t)b * p.Nsrc * 12 + i);
    int x = blockIdx.x * TILE_W + (warp % WX) * PPW + pix;
    int y = blockIdx.y * TILE_H + (warp / WX);
    const bool live = (x < p.W) && (y < p.H);
    x = min(x, p.W - 1);  // dead lanes shadow a valid pixel so that shuffles stay convergent
    y = min(y, p.H - 1);
    __syncthreads();

    const size_t plane = (size_t)p.H * p.W;
    const size_t pix_off = (size_t)y * p.W + x;

    f32x2 rf[4];
    {
        const P8 r = load_pairs<T>(reinterpret_cast<const T*>(p.ref) + (((size_t)b * plane + pix_off) * C + sub * 8));
        const f32x2 sc = pack2(1.0f / CPG, 1.0f / CPG);  // group mean as a plain sum (CPG = 1 for the variance cost)
#pragma unroll
        for (int q = 0; q < 4; ++q) rf[q] = mul2(r.q[q], sc);
    }
    float hyp[NOWN];
#pragma unroll
    for (int k = 0; k < NOWN; ++k) {
        const int d = min(sub + k * L, D - 1);
        hyp[k] = ldg_stream(p.hypo + ((size_t)b * D + d) * plane + pix_off);
    }
    float acc[GPL][D], wsum[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        wsum[d] = 1e-8f;  // reference :1037
#pragma unroll
        for (int g = 0; g < GPL; ++g) acc[g][d] = 0.0f;
    }
    const float fxp = (float)x, fyp = (float)y;
    const float wlim = (float)p.Ws, hlim = (float)p.Hs;
    const size_t lane_src_off = ((size_t)b * p.Hs * p.Ws * C + sub * 8) * sizeof(T);

#pragma unroll 1
    for (int v = 0; v < p.Nsrc; ++v) {
        // ---- owner phase: tap texel indices and weights of this lane's own hypotheses -----------------------------
        unsigned ti[NOWN][4];
        float tw[NOWN][4];
        {
            const Homography h = homography_from_smem(rt_s + v * 12);
            const float ax = fmaf(h.r00, fxp, fmaf(h.r01, fyp, h.r02));  // R * [x, y, 1]^T (reference :42)
            const float ay = fmaf(h.r10, fxp, fmaf(h.r11, fyp, h.r12));
            const float az = fmaf(h.r20, fxp, fmaf(h.r21, fyp, h.r22));
#pragma unroll
            for (int k = 0; k < NOWN; ++k) {
                float sx, sy;
                sample_pos(ax, ay, az, h, hyp[k], wlim, hlim, sx, sy);
                const float x0f = floorf(sx), y0f = floorf(sy);
                const float fx = sx - x0f, fy = sy - y0f;
                const int x0 = (int)x0f, y0 = (int)y0f;  // in [-1, Ws] x [-1, Hs] after the clamp
                const bool vx0 = (unsigned)x0 < (unsigned)p.Ws, vx1 = (unsigned)(x0 + 1) < (unsigned)p.Ws;
                const bool vy0 = (unsigned)y0 < (unsigned)p.Hs, vy1 = (unsigned)(y0 + 1) < (unsigned)p.Hs;
                const int xc0 = min(max(x0, 0), p.Ws - 1), xc1 = min(x0 + 1, p.Ws - 1);
                const int yc0 = min(max(y0, 0), p.Hs - 1), yc1 = min(y0 + 1, p.Hs - 1);
                const float gx = vx0 ? 1.0f - fx : 0.0f, hx = vx1 ? fx : 0.0f;
                const float gy = vy0 ? 1.0f - fy : 0.0f, hy = vy1 ? fy : 0.0f;
                tw[k][0] = gx * gy; tw[k][1] = hx * gy; tw[k][2] = gx * hy; tw[k][3] = hx * hy;
                const unsigned r0 = (unsigned)(yc0 * p.Ws), r1 = (unsigned)(yc1 * p.Ws);
                ti[k][0] = r0 + (unsigned)xc0; ti[k][1] = r0 + (unsigned)xc1;
                ti[k][2] = r1 + (unsigned)xc0; ti[k][3] = r1 + (unsigned)xc1;
            }
        }
        // ---- gather phase: every lane fetches its 8 channels of each sample's four taps ---------------------------
        const char* srcp = reinterpret_cast<const char*>(p.src[v]) + lane_src_off;
        float cor[GPL][D], score[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            constexpr int dummy = 0; (void)dummy;
            const int owner = d % L, k = d / L;  // compile-time after unrolling
            unsigned i0 = ti[k][0], i1 = ti[k][1], i2 = ti[k][2], i3 = ti[k][3];
            float w0 = tw[k][0], w1 = tw[k][1], w2 = tw[k][2], w3 = tw[k][3];
            if constexpr (L > 1) {
                i0 = __shfl_sync(0xffffffffu, i0, owner, L); i1 = __shfl_sync(0xffffffffu, i1, owner, L);
                i2 = __shfl_sync(0xffffffffu, i2, owner, L); i3 = __shfl_sync(0xffffffffu, i3, owner, L);
                w0 = __shfl_sync(0xffffffffu, w0, owner, L); w1 = __shfl_sync(0xffffffffu, w1, owner, L);
                w2 = __shfl_sync(0xffffffffu, w2, owner, L); w3 = __shfl_sync(0xffffffffu, w3, owner, L);
            }
            // texel index < 2^31 / C (checked on the host): one IMAD.WIDE.U32 per address
            const P8 t00 = load_pairs<T>(srcp + (size_t)i0 * TB), t01 = load_pairs<T>(srcp + (size_t)i1 * TB);
            const P8 t10 = load_pairs<T>(srcp + (size_t)i2 * TB), t11 = load_pairs<T>(srcp + (size_t)i3 * TB);
            float cg[GPL];
            blend_correlate<CPG, VAR>(t00, t01, t10, t11, w0, w1, w2, w3, rf, cg);
            float s = 0.0f;
#pragma unroll
            for (int g = 0; g < GPL; ++g) { cor[g][d] = cg[g]; s += cg[g]; }
            score[d] = s;
        }
        // sum over all G groups = over the L lanes of the pixel (reference cor_feat.sum(1), :1083)
#pragma unroll
        for (int m = 1; m < L; m <<= 1) {
#pragma unroll
            for (int d = 0; d < D; ++d) score[d] += __shfl_xor_sync(0xffffffffu, score[d], m);
        }
        float mx = score[0];
#pragma unroll
        for (int d = 1; d < D; ++d) mx = fmaxf(mx, score[d]);
        float e[D], es = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            e[d] = ex2_approx((score[d] - mx) * (FUSE_D ? p.score_scale : 1.4426950408889634f));
            es += e[d];
        }
        if constexpr (FUSE_D) {
            // w[d] = softmax_d(score / attn_temp) / sqrt(C)                                  (reference :1083)
            const float norm = __fdividef(p.inv_sqrt_c, es);
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float w = e[d] * norm;
                wsum[d] += w;
#pragma unroll
                for (int g = 0; g < GPL; ++g) acc[g][d] = fmaf(w, cor[g][d], acc[g][d]);
                if (p.weights != nullptr && sub == 0 && live)
                    p.weights[(((size_t)b * p.Nsrc + v) * D + d) * plane + pix_off] = w;
            }
        } else {
            // attn_fuse_d=False: one weight per pixel, w = max_d softmax_d(score) = 1 / sum_d exp(score_d - max)
            // (no temperature, no sqrt(C); reference :1079-1081,1098); wsum[0] carries the running sum
            const float w = __fdividef(1.0f, es);
            wsum[0] += w;
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int g = 0; g < GPL; ++g) acc[g][d] = fmaf(w, cor[g][d], acc[g][d]);
        }
    }

    if (!live) return;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float inv = __frcp_rn(wsum[FUSE_D ? d : 0]);
#pragma unroll
        for (int g = 0; g < GPL; ++g)
            stg_stream(p.out + (((size_t)b * G + sub * GPL + g) * D + d) * plane + pix_off, acc[g][d] * inv);
        if (FUSE_D && p.wsum != nullptr && sub == 0) p.wsum[((size_t)b * D + d) * plane + pix_off] = wsum[d];
    }
    if (!FUSE_D && p.wsum != nullptr && sub == 0) p.wsum[(size_t)b * plane + pix_off] = wsum[0];  // [B,H,W]
}

template <int C, int CPG, int D, typename T, bool VAR = false, bool FUSE_D = true>
static int launch_direct(const EpiFwdParams& p, cudaStream_t stream) {
    constexpr int WXM = MVSTER_DIRECT_WXMAX < kWarps ? MVSTER_DIRECT_WXMAX : kWarps;
    constexpr int L = C / 8, PPW = 32 / L, WX = L < WXM ? L : WXM, TILE_W = PPW * WX, TILE_H = kWarps / WX;
    dim3 grid((p.W + TILE_W - 1) / TILE_W, (p.H + TILE_H - 1) / TILE_H, p.B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: grid too large");
    epi_fwd_direct_kernel<C, CPG, D, T, VAR, FUSE_D><<<grid, kThreads, MVSTER_MAX_SRC_VIEWS * 48, stream>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_fwd launch");
    return MVSTER_OK;
}

}  // namespace mvster
