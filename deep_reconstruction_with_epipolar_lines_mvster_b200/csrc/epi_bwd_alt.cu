// K1 backward for the two reference options that no shipped configuration trains with:
//   group_cor=False  : variance cost (ref - warped)^2, G == C        (models/mvs4net_utils.py:1071)
//   attn_fuse_d=False: one weight per pixel and view, max_d softmax_d (models/mvs4net_utils.py:1078-1081,1098)
// The direct-gather backward kernel of epi_bwd_common.cuh with two compile-time flags (derivation in its header);
// fp32 features, no hypothesis split.  Forward that saves the matching weight sums: mvster_epi_fwd_mode_ex.
#include "epi_bwd_common.cuh"

namespace mvster {

template <int C, int CPG, int D, bool VAR, bool FUSE_D>
static int launch_bwd_alt(const EpiBwdParams& p, cudaStream_t stream) {
    constexpr int PPW = 32 / (C / 8);
    dim3 grid((p.W + PPW - 1) / PPW, (p.H + kBwdWarps - 1) / kBwdWarps, p.B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd_mode: grid too large");
    epi_bwd_kernel<C, CPG, D, float, 1, VAR, FUSE_D><<<grid, kBwdWarps * 32, 0, stream>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_bwd_mode launch");
    return MVSTER_OK;
}

template <int C, int CPG, bool VAR, bool FUSE_D>
static int alt_bwd_d(const EpiBwdParams& p, int D, cudaStream_t s) {
    switch (D) {
        case 4: return launch_bwd_alt<C, CPG, 4, VAR, FUSE_D>(p, s);
        case 8: return launch_bwd_alt<C, CPG, 8, VAR, FUSE_D>(p, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd_mode: D=%d not in {4,8}", D);
    }
}

template <int C>
static int alt_bwd_c(const EpiBwdParams& p, int cpg, int D, bool var, bool fuse_d, cudaStream_t s) {
    if (var) return fuse_d ? alt_bwd_d<C, 1, true, true>(p, D, s) : alt_bwd_d<C, 1, true, false>(p, D, s);
    switch (cpg) {  // group correlation with the per-pixel weight
        case 1: return alt_bwd_d<C, 1, false, false>(p, D, s);
        case 2: return alt_bwd_d<C, 2, false, false>(p, D, s);
        case 4: return alt_bwd_d<C, 4, false, false>(p, D, s);
        case 8: return alt_bwd_d<C, 8, false, false>(p, D, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd_mode: C/G=%d not in {1,2,4,8}", cpg);
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_epi_bwd_mode(const void* ref, const void* const* src, const float* rt, const float* hypo,
                                   const float* out, const float* wsum, const float* gout, float* grad_ref,
                                   float* const* grad_src, int B, int Nsrc, int C, int G, int D, int H, int W, int Hs,
                                   int Ws, float attn_temp, int dtype, int group_cor, int attn_fuse_d, void* stream) {
    if (group_cor && attn_fuse_d)
        return mvster_epi_bwd(ref, src, rt, hypo, out, wsum, gout, grad_ref, grad_src, B, Nsrc, C, G, D, H, W, Hs, Ws,
                              attn_temp, dtype, stream);
    if (!ref || !src || !rt || !hypo || !out || !wsum || !gout || !grad_ref || !grad_src)
        return fail(MVSTER_ERR_BAD_ARG, "epi_bwd_mode: null pointer");
    if (B <= 0 || Nsrc <= 0 || C <= 0 || G <= 0 || D <= 0 || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0)
        return fail(MVSTER_ERR_BAD_ARG, "epi_bwd_mode: non-positive dimension");
    if (Nsrc > MVSTER_MAX_SRC_VIEWS) return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd_mode: too many source views");
    if (dtype != MVSTER_F32) return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd_mode: the variants support fp32 features only");
    if (!group_cor && G != C) return fail(MVSTER_ERR_BAD_ARG, "epi_bwd_mode: group_cor=0 needs G == C (got %d, %d)", G, C);
    if (C % G != 0) return fail(MVSTER_ERR_BAD_ARG, "epi_bwd_mode: C=%d not divisible by G=%d", C, G);
    if (!(attn_temp > 0.0f)) return fail(MVSTER_ERR_BAD_ARG, "epi_bwd_mode: attn_temp must be > 0");
    if ((double)B * Hs * Ws * C >= 2147483648.0 || (double)H * W >= 2147483648.0)
        return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd_mode: tensor too large for 32-bit texel offsets");
    if (((uintptr_t)ref) % 32 || ((uintptr_t)rt) % 16 || ((uintptr_t)grad_ref) % 16)
        return fail(MVSTER_ERR_ALIGN, "epi_bwd_mode: misaligned pointer");
    EpiBwdParams p{};
    p.ref = ref;
    for (int v = 0; v < Nsrc; ++v) {
        if (!src[v] || !grad_src[v]) return fail(MVSTER_ERR_BAD_ARG, "epi_bwd_mode: src/grad_src[%d] is null", v);
        if (((uintptr_t)src[v]) % 32 || ((uintptr_t)grad_src[v]) % 16)
            return fail(MVSTER_ERR_ALIGN, "epi_bwd_mode: src/grad_src[%d] misaligned", v);
        p.src[v] = src[v];
        p.grad_src[v] = grad_src[v];
    }
    p.rt = rt; p.hypo = hypo; p.out = out; p.wsum = wsum; p.gout = gout; p.grad_ref = grad_ref;
    p.B = B; p.Nsrc = Nsrc; p.H = H; p.W = W; p.Hs = Hs; p.Ws = Ws;
    const bool fuse = attn_fuse_d != 0;
    // attn_fuse_d=False: plain softmax over D, no temperature, no 1/sqrt(C) (:1079)
    p.score_scale = fuse ? 1.4426950408889634f / attn_temp : 1.4426950408889634f;
    p.inv_temp = fuse ? 1.0f / attn_temp : 1.0f;
    p.inv_sqrt_c = fuse ? (float)(1.0 / sqrt((double)C)) : 1.0f;
    DeviceGuard guard(grad_ref);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    const int cpg = C / G;
    switch (C) {
        case 8: return alt_bwd_c<8>(p, cpg, D, !group_cor, fuse, s);
        case 16: return alt_bwd_c<16>(p, cpg, D, !group_cor, fuse, s);
        case 32: return alt_bwd_c<32>(p, cpg, D, !group_cor, fuse, s);
        case 64: return alt_bwd_c<64>(p, cpg, D, !group_cor, fuse, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_bwd_mode: C=%d not in {8,16,32,64}", C);
    }
}
