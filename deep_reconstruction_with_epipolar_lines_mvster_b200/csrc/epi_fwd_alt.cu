// K1 forward, the two variants of the reference that no shipped configuration uses (backward: epi_bwd_alt.cu):
//   group_cor=False : per-channel variance cost (ref - warped)^2, G == C   (models/mvs4net_utils.py:1071)
//   attn_fuse_d=False: one attention weight per pixel and view, max_d softmax_d(score), no temperature, no sqrt(C)
//                      (models/mvs4net_utils.py:1078-1081,1098)
// They reuse the direct-gather kernel of epi_common.cuh with two compile-time flags; fp32 features only.
#include "epi_common.cuh"

namespace mvster {

template <int C, int CPG, bool VAR, bool FUSE_D>
static int alt_d(const EpiFwdParams& p, int D, cudaStream_t s) {
    switch (D) {
        case 4: return launch_direct<C, CPG, 4, float, VAR, FUSE_D>(p, s);
        case 8: return launch_direct<C, CPG, 8, float, VAR, FUSE_D>(p, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd_mode: D=%d not in {4,8}", D);
    }
}

template <int C>
static int alt_c(const EpiFwdParams& p, int cpg, int D, bool var, bool fuse_d, cudaStream_t s) {
    if (var) return fuse_d ? alt_d<C, 1, true, true>(p, D, s) : alt_d<C, 1, true, false>(p, D, s);
    switch (cpg) {  // group correlation with the per-pixel weight
        case 1: return alt_d<C, 1, false, false>(p, D, s);
        case 2: return alt_d<C, 2, false, false>(p, D, s);
        case 4: return alt_d<C, 4, false, false>(p, D, s);
        case 8: return alt_d<C, 8, false, false>(p, D, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd_mode: C/G=%d not in {1,2,4,8}", cpg);
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_epi_fwd_mode(const void* ref, const void* const* src, const float* rt, const float* hypo,
                                   float* out, int B, int Nsrc, int C, int G, int D, int H, int W, int Hs, int Ws,
                                   float attn_temp, int dtype, int group_cor, int attn_fuse_d, void* stream) {
    return mvster_epi_fwd_mode_ex(ref, src, rt, hypo, out, nullptr, B, Nsrc, C, G, D, H, W, Hs, Ws, attn_temp, dtype,
                                  group_cor, attn_fuse_d, stream);
}

extern "C" int mvster_epi_fwd_mode_ex(const void* ref, const void* const* src, const float* rt, const float* hypo,
                                      float* out, float* wsum, int B, int Nsrc, int C, int G, int D, int H, int W,
                                      int Hs, int Ws, float attn_temp, int dtype, int group_cor, int attn_fuse_d,
                                      void* stream) {
    if (group_cor && attn_fuse_d)
        return mvster_epi_fwd(ref, src, rt, hypo, out, wsum, nullptr, B, Nsrc, C, G, D, H, W, Hs, Ws, attn_temp, dtype,
                              stream);
    if (!ref || !src || !rt || !hypo || !out) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd_mode: null pointer");
    if (B <= 0 || Nsrc <= 0 || C <= 0 || G <= 0 || D <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0)
        return fail(MVSTER_ERR_BAD_ARG, "epi_fwd_mode: non-positive dimension");
    if (Nsrc > MVSTER_MAX_SRC_VIEWS) return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd_mode: too many source views");
    if (dtype != MVSTER_F32) return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd_mode: the variants support fp32 features only");
    if (!group_cor && G != C) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd_mode: group_cor=0 needs G == C (got %d, %d)", G, C);
    if (C % G != 0) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd_mode: C=%d not divisible by G=%d", C, G);
    if (!(attn_temp > 0.0f)) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd_mode: attn_temp must be > 0");
    if ((double)B * Hs * Ws * C >= 2147483648.0 || (double)H * W >= 2147483648.0)
        return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd_mode: tensor too large for 32-bit texel offsets");
    if (((uintptr_t)ref) % 32 || ((uintptr_t)rt) % 16) return fail(MVSTER_ERR_ALIGN, "epi_fwd_mode: misaligned pointer");
    static thread_local EpiFwdParams p;
    p.ref = ref;
    for (int v = 0; v < Nsrc; ++v) {
        if (!src[v] || ((uintptr_t)src[v]) % 32) return fail(MVSTER_ERR_ALIGN, "epi_fwd_mode: src[%d] null or misaligned", v);
        p.src[v] = src[v];
    }
    p.rt = rt; p.hypo = hypo; p.out = out; p.wsum = wsum; p.weights = nullptr;
    p.B = B; p.Nsrc = Nsrc; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws;
    p.score_scale = 1.4426950408889634f / attn_temp;
    p.inv_sqrt_c = (float)(1.0 / sqrt((double)C));
    DeviceGuard guard(out);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    const bool var = !group_cor, fuse = attn_fuse_d != 0;
    const int cpg = C / G;
    switch (C) {
        case 8: return alt_c<8>(p, cpg, D, var, fuse, s);
        case 16: return alt_c<16>(p, cpg, D, var, fuse, s);
        case 32: return alt_c<32>(p, cpg, D, var, fuse, s);
        case 64: return alt_c<64>(p, cpg, D, var, fuse, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd_mode: C=%d not in {8,16,32,64}", C);
    }
}
