// K1 forward: fused homography warp + bilinear gather + group-wise correlation + epipolar attention (softmax over
// the D depth hypotheses) + view-weighted aggregation.  One kernel per cascade stage; the [B,C,D,H,W] warped
// volume, the repeated reference volume and the per-view correlation volumes of the reference
// (models/mvs4net_utils.py:1036,1051,1066-1069,1083-1085,1100) never exist in memory.
//
// Work decomposition (sm_100a):
//   * features are NHWC; a lane owns 8 consecutive channels of one reference pixel, so every bilinear tap is a
//     single 32-byte LDG.E.256 (fp32) / 16-byte LDG.E.128 (bf16) and the lanes of a warp read one contiguous
//     ~1 KB span of the source row (adjacent pixels hit adjacent texels);
//   * L = C/8 lanes cooperate on one pixel (L = 1, 2, 4, 8 for C = 8, 16, 32, 64); a lane's 8 channels cover
//     8/(C/G) whole correlation groups, so the group correlation is lane-local and only the per-hypothesis score
//     (sum over groups) crosses lanes, with log2(L) xor-shuffles;
//   * the softmax over D, the running sum of weights and the weighted volume accumulators live in registers;
//   * a CTA is 8 warps stacked in y (a (32/L) x 8 pixel tile) so that the y0+1 source row fetched for one
//     reference row is still in L1 when the next row needs it as y0.
#include "common.cuh"

namespace mvster {

struct EpiFwdParams {
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

constexpr int kWarpsPerCta = 8;

template <int C, int CPG, int D, typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 2) epi_fwd_kernel(const __grid_constant__ EpiFwdParams p) {
    constexpr int CPL = 8;          // channels per lane
    constexpr int L = C / CPL;      // lanes per pixel
    constexpr int GPL = CPL / CPG;  // correlation groups per lane
    constexpr int PPW = 32 / L;     // pixels per warp
    constexpr int G = C / CPG;
    static_assert(C % CPL == 0 && CPL % CPG == 0 && L >= 1 && L <= 32, "unsupported channel split");

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane % L;  // which 8-channel chunk of the pixel
    const int pix = lane / L;
    const int b = blockIdx.z;
    int x = blockIdx.x * PPW + pix;
    int y = blockIdx.y * kWarpsPerCta + warp;
    const bool live = (x < p.W) && (y < p.H);
    // dead lanes shadow a valid pixel so that warp shuffles stay convergent; their stores are masked
    x = min(x, p.W - 1);
    y = min(y, p.H - 1);

    const size_t plane = (size_t)p.H * p.W;
    const size_t pix_off = (size_t)y * p.W + x;

    // reference feature chunk: 8 channels, pre-scaled by 1/(C/G) so the group mean is a plain dot product
    const T* refp = reinterpret_cast<const T*>(p.ref) + (((size_t)b * plane + pix_off) * C + sub * CPL);
    F8 rf = load8<T>(refp);
#pragma unroll
    for (int c = 0; c < CPL; ++c) rf.v[c] *= (1.0f / CPG);

    float hyp[D];
#pragma unroll
    for (int d = 0; d < D; ++d) hyp[d] = ldg_stream(p.hypo + ((size_t)b * D + d) * plane + pix_off);

    float acc[GPL][D];
    float wsum[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        wsum[d] = 1e-8f;  // reference :1037
#pragma unroll
        for (int g = 0; g < GPL; ++g) acc[g][d] = 0.0f;
    }

    const float fx = (float)x, fy = (float)y;
    const size_t src_batch = (size_t)b * p.Hs * p.Ws * C + sub * CPL;

#pragma unroll 1
    for (int v = 0; v < p.Nsrc; ++v) {
        const Homography h = load_homography(p.rt + ((size_t)b * p.Nsrc + v) * 12);
        const T* srcp = reinterpret_cast<const T*>(p.src[v]) + src_batch;
        // R * [x, y, 1]^T, shared by all hypotheses of this pixel (reference :42)
        const float ax = fmaf(h.r00, fx, fmaf(h.r01, fy, h.r02));
        const float ay = fmaf(h.r10, fx, fmaf(h.r11, fy, h.r12));
        const float az = fmaf(h.r20, fx, fmaf(h.r21, fy, h.r22));

        float cor[GPL][D];
        float score[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const Taps t = make_taps(ax, ay, az, h, hyp[d], p.Hs, p.Ws);
            float wv[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) wv[c] = 0.0f;
            if (t.any) {
                const F8 a = load8<T>(srcp + (size_t)t.o00 * C);
                const F8 bq = load8<T>(srcp + (size_t)t.o01 * C);
                const F8 cq = load8<T>(srcp + (size_t)t.o10 * C);
                const F8 dq = load8<T>(srcp + (size_t)t.o11 * C);
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    wv[c] = fmaf(t.w00, a.v[c], fmaf(t.w01, bq.v[c], fmaf(t.w10, cq.v[c], t.w11 * dq.v[c])));
            }
            float s = 0.0f;
#pragma unroll
            for (int g = 0; g < GPL; ++g) {
                float cg = 0.0f;
#pragma unroll
                for (int c = 0; c < CPG; ++c) cg = fmaf(rf.v[g * CPG + c], wv[g * CPG + c], cg);
                cor[g][d] = cg;
                s += cg;
            }
            score[d] = s;
        }
        // sum over all G groups: reduce across the L lanes of this pixel (reference cor_feat.sum(1), :1083)
#pragma unroll
        for (int m = 1; m < L; m <<= 1) {
#pragma unroll
            for (int d = 0; d < D; ++d) score[d] += __shfl_xor_sync(0xffffffffu, score[d], m);
        }
        // softmax over D of score / attn_temp, then / sqrt(C)
        float mx = score[0];
#pragma unroll
        for (int d = 1; d < D; ++d) mx = fmaxf(mx, score[d]);
        float e[D];
        float es = 0.0f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            e[d] = exp2f((score[d] - mx) * p.score_scale);
            es += e[d];
        }
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
    }

    if (!live) return;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float inv = __frcp_rn(wsum[d]);
#pragma unroll
        for (int g = 0; g < GPL; ++g) {
            const int gg = sub * GPL + g;
            stg_stream(p.out + (((size_t)b * G + gg) * D + d) * plane + pix_off, acc[g][d] * inv);
        }
        if (p.wsum != nullptr && sub == 0) p.wsum[((size_t)b * D + d) * plane + pix_off] = wsum[d];
    }
}

template <int C, int CPG, int D, typename T>
static int launch_fwd(const EpiFwdParams& p, cudaStream_t stream) {
    constexpr int PPW = 32 / (C / 8);
    dim3 grid((p.W + PPW - 1) / PPW, (p.H + kWarpsPerCta - 1) / kWarpsPerCta, p.B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: grid too large");
    epi_fwd_kernel<C, CPG, D, T><<<grid, kWarpsPerCta * 32, 0, stream>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_fwd launch");
    return MVSTER_OK;
}

template <int C, int CPG, typename T>
static int dispatch_d(const EpiFwdParams& p, int D, cudaStream_t s) {
    switch (D) {
        case 4: return launch_fwd<C, CPG, 4, T>(p, s);
        case 8: return launch_fwd<C, CPG, 8, T>(p, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: D=%d not in {4,8}", D);
    }
}

template <int C, typename T>
static int dispatch_cpg(const EpiFwdParams& p, int cpg, int D, cudaStream_t s) {
    switch (cpg) {
        case 1: return dispatch_d<C, 1, T>(p, D, s);
        case 2: return dispatch_d<C, 2, T>(p, D, s);
        case 4: return dispatch_d<C, 4, T>(p, D, s);
        case 8: return dispatch_d<C, 8, T>(p, D, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: C/G=%d not in {1,2,4,8}", cpg);
    }
}

template <typename T>
static int dispatch_c(const EpiFwdParams& p, int C, int cpg, int D, cudaStream_t s) {
    switch (C) {
        case 8: return dispatch_cpg<8, T>(p, cpg, D, s);
        case 16: return dispatch_cpg<16, T>(p, cpg, D, s);
        case 32: return dispatch_cpg<32, T>(p, cpg, D, s);
        case 64: return dispatch_cpg<64, T>(p, cpg, D, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: C=%d not in {8,16,32,64}", C);
    }
}

int epi_fwd_dispatch(const EpiFwdParams& p, int C, int G, int D, int dtype, cudaStream_t s) {
    const int cpg = C / G;
    if (dtype == MVSTER_F32) return dispatch_c<float>(p, C, cpg, D, s);
    if (dtype == MVSTER_BF16) return dispatch_c<__nv_bfloat16>(p, C, cpg, D, s);
    return fail(MVSTER_ERR_BAD_ARG, "epi_fwd: unknown dtype %d", dtype);
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_epi_fwd(const void* ref, const void* const* src, const float* rt, const float* hypo, float* out,
                              float* wsum, float* weights, int B, int Nsrc, int C, int G, int D, int H, int W, int Hs,
                              int Ws, float attn_temp, int dtype, void* stream) {
    if (!ref || !src || !rt || !hypo || !out) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd: null pointer");
    if (B <= 0 || Nsrc <= 0 || C <= 0 || G <= 0 || D <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0)
        return fail(MVSTER_ERR_BAD_ARG, "epi_fwd: non-positive dimension");
    if (Nsrc > MVSTER_MAX_SRC_VIEWS)
        return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: %d source views > MVSTER_MAX_SRC_VIEWS=%d", Nsrc,
                    MVSTER_MAX_SRC_VIEWS);
    if (C % G != 0) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd: C=%d not divisible by G=%d", C, G);
    if (!(attn_temp > 0.0f)) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd: attn_temp must be > 0");
    if ((double)B * Hs * Ws * C >= 2147483648.0 || (double)H * W >= 2147483648.0)
        return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: tensor too large for 32-bit texel offsets");
    const uintptr_t align = (dtype == MVSTER_BF16) ? 16 : 32;
    if (((uintptr_t)ref) % align) return fail(MVSTER_ERR_ALIGN, "epi_fwd: ref not %d-byte aligned", (int)align);
    EpiFwdParams p{};
    p.ref = ref;
    for (int v = 0; v < Nsrc; ++v) {
        if (!src[v]) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd: src[%d] is null", v);
        if (((uintptr_t)src[v]) % align)
            return fail(MVSTER_ERR_ALIGN, "epi_fwd: src[%d] not %d-byte aligned", v, (int)align);
        p.src[v] = src[v];
    }
    if (((uintptr_t)rt) % 16) return fail(MVSTER_ERR_ALIGN, "epi_fwd: rt not 16-byte aligned");
    p.rt = rt; p.hypo = hypo; p.out = out; p.wsum = wsum; p.weights = weights;
    p.B = B; p.Nsrc = Nsrc; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws;
    p.score_scale = 1.4426950408889634f / attn_temp;
    p.inv_sqrt_c = (float)(1.0 / sqrt((double)C));
    DeviceGuard guard(out);
    if (guard.status != MVSTER_OK) return guard.status;
    return epi_fwd_dispatch(p, C, G, D, dtype, (cudaStream_t)stream);
}
