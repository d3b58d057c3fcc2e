// K3: the optimal-transport (Sinkhorn / Wasserstein) depth loss of MVSTER fused over D  (SURVEY.md section 8f rank 3).
//
// Replaces `sinkhorn` (models/mvs4net_utils.py:1164-1210) and the per-stage statistics around its call in
// `MVS4net_loss` / `Blend_loss` (models/MVS4Net.py:225-234, :272-281).  The reference builds [B,HW,D,D(+1)] cost and
// transport tensors and runs 2*iters logsumexp passes over them with ~15 eager kernels per iteration, keeping every
// intermediate for autograd.  Here one thread owns one pixel: the D x D(+1) problem lives in registers, the log-domain
// iterations are unrolled over D, and the backward sweep of the unrolled iterations runs in the SAME kernel, reading
// the per-iteration logsumexp values back from shared memory (thread-fastest layout, conflict-free).  Per pixel the
// kernel reads 4*(2D+1)+1 bytes and writes 4*D bytes of d(loss)/d(attn); nothing of size D*D reaches HBM unless the
// caller asks for the transport map.
//
// Math (i = prediction bin 0..D-1, j = ground-truth bin 0..NC-1, NC = D, or D+1 in `continuous` mode):
//   M[i][j]  = |i - j|;  continuous: M[i][D] = |gbd - i|, gbd = (1/gt - 1/hypo[0]) / (1/hypo[2] - 1/hypo[1]), 10 if unmasked
//   K        = M / eps                       (POSITIVE exponent, exactly as the reference: `D_map/eps`, :1200-1204)
//   log_mu_j = log(onehot(argmin_d |hypo_d - gt|)_j + 1e-12)   (continuous: onehot(D));  log_nu_i = log(attn_i + 1e-12)
//   u_0 = 0;  v_t[j] = log_mu_j - LSE_i(K_ij + u_{t-1}[i]);  u_t[i] = log_nu_i - LSE_j(K_ij + v_t[j])      t = 1..iters
//   T_ij = exp(K_ij + u_T[i] + v_T[j]);  loss_px = sum_ij T_ij M_ij;  loss = mean over masked pixels
// Backward (attn only: hypo_depth is detached upstream, models/MVS4Net.py:116; gt and mask carry no gradient):
//   gu_i = sum_j T_ij M_ij, gv_j = sum_i T_ij M_ij; for t = T..1:  g_ln += gu;  gv_j -= sum_i gu_i exp(K_ij + v_t[j] - Lu_t[i]);
//   gu_i = -sum_j gv_j exp(K_ij + u_{t-1}[i] - Lv_t[j]);  gv = 0     (Lu_t, Lv_t = the saved logsumexp values)
//   d loss_px / d attn_i = g_ln_i / (attn_i + 1e-12)
#include "common.cuh"

namespace mvster {

struct SinkhornParams {
    const float* gt;       // [B,H,W]
    const float* hypo;     // [B,D,H,W]
    const float* attn;     // [B,D,H,W]
    const uint8_t* mask;   // [B,H,W]
    float* grad;           // nullable [B,D,H,W]
    float* tmap;           // nullable [B,HW,D,NC]
    double* partials;      // [gridDim.x][3]: sum of masked per-pixel losses, masked count, masked out-of-range count
    size_t plane, total;
    int iters, inverse;
    float eps;
};

constexpr float kLogTiny = -27.6310211159285482f;  // log(1e-12)

// exp / log of the scaling passes.  MVSTER_SINKHORN_FAST=1: one MUFU each (ex2.approx / lg2.approx on pre-scaled
// arguments, relative error ~2^-22 plus |x| * 2^-24 from the scaling) instead of the ~8-instruction accurate expf /
// logf; every exponent here is <= 0 after the max subtraction, so the error is relative to a value <= 1.
#ifndef MVSTER_SINKHORN_FAST
#define MVSTER_SINKHORN_FAST 1
#endif
__device__ __forceinline__ float sk_exp(float x) {
#if MVSTER_SINKHORN_FAST
    return __expf(x);
#else
    return expf(x);
#endif
}
__device__ __forceinline__ float sk_log(float x) {
#if MVSTER_SINKHORN_FAST
    return __logf(x);
#else
    return logf(x);
#endif
}

template <int N>
__device__ __forceinline__ float lse(const float (&x)[N]) {
    float m = x[0];
#pragma unroll
    for (int k = 1; k < N; ++k) m = fmaxf(m, x[k]);
    if (fabsf(m) == INFINITY) m = 0.0f;  // torch.logsumexp: infinite maxima are replaced by 0 before the subtraction
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < N; ++k) s += sk_exp(x[k] - m);
    return sk_log(s) + m;
}

template <int D, bool CONT, int NT>
__global__ void __launch_bounds__(NT) sinkhorn_kernel(const __grid_constant__ SinkhornParams p) {
    constexpr int NC = CONT ? D + 1 : D;
    extern __shared__ float hist[];  // [iters][D + NC][NT]: Lu_t[i], then Lv_t[j]
    __shared__ double red[3][NT / 32];
    const int tid = threadIdx.x;
    const size_t idx = (size_t)blockIdx.x * NT + tid;
    const bool inside = idx < p.total;
    const size_t b = inside ? idx / p.plane : 0, pix = inside ? idx - b * p.plane : 0;
    const bool masked = inside && p.mask[idx] != 0;
    double my_loss = 0.0, my_cnt = 0.0, my_oor = 0.0;

    if (inside && (masked || p.tmap != nullptr)) {
        const float* hp = p.hypo + b * D * p.plane + pix;
        const float* ap = p.attn + b * D * p.plane + pix;
        const float gt = p.gt[idx];
        float hyp[D], ln[D], at[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            hyp[d] = ldg_stream(hp + (size_t)d * p.plane);
            at[d] = ldg_stream(ap + (size_t)d * p.plane) + 1e-12f;
            ln[d] = logf(at[d]);
        }
        // cost columns: K[i][j] = kt[|i-j|] for j < D (a Toeplitz matrix: D distinct values), plus the per-pixel column
        float kt[D], kc[D] = {}, mc[D] = {};
#pragma unroll
        for (int k = 0; k < D; ++k) kt[k] = (float)k / p.eps;
        int gi = 0;
        if constexpr (CONT) {
            const float itv = 1.0f / hyp[2] - 1.0f / hyp[1];                       // reference :1186
            float gbd = (1.0f / gt - 1.0f / hyp[0]) / itv;                         // :1187
            if (!masked) gbd = 10.0f;                                              // :1189
#pragma unroll
            for (int i = 0; i < D; ++i) { mc[i] = fabsf(gbd - (float)i); kc[i] = mc[i] / p.eps; }
        } else {
            float best = fabsf(hyp[0] - gt);                                       // :1176 (first minimum)
#pragma unroll
            for (int d = 1; d < D; ++d) {
                const float e = fabsf(hyp[d] - gt);
                if (e < best) { best = e; gi = d; }
            }
        }
        auto K = [&](int i, int j) -> float { return (CONT && j == D) ? kc[i] : kt[i > j ? i - j : j - i]; };
        auto M = [&](int i, int j) -> float { return (CONT && j == D) ? mc[i] : (float)(i > j ? i - j : j - i); };
        float lm[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) lm[j] = (j == (CONT ? D : gi)) ? 0.0f : kLogTiny;  // log(1 + 1e-12) == 0 in fp32

        float u[D], v[NC];
#pragma unroll
        for (int i = 0; i < D; ++i) u[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = 0.0f;
        const bool keep = p.grad != nullptr && masked;
#pragma unroll 1
        for (int t = 0; t < p.iters; ++t) {
            float* h = hist + (size_t)t * (D + NC) * NT + tid;
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                float x[D];
#pragma unroll
                for (int i = 0; i < D; ++i) x[i] = K(i, j) + u[i];
                const float l = lse<D>(x);
                v[j] = lm[j] - l;
                if (keep) h[(D + j) * NT] = l;
            }
#pragma unroll
            for (int i = 0; i < D; ++i) {
                float x[NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) x[j] = K(i, j) + v[j];
                const float l = lse<NC>(x);
                u[i] = ln[i] - l;
                if (keep) h[i * NT] = l;
            }
        }
        // transport map, per-pixel loss and the seeds of the backward sweep
        float gu[D], gv[NC], loss = 0.0f;
#pragma unroll
        for (int j = 0; j < NC; ++j) gv[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            gu[i] = 0.0f;
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float tm = sk_exp(K(i, j) + u[i] + v[j]);
                const float c = tm * M(i, j);
                gu[i] += c;
                gv[j] += c;
                if (p.tmap != nullptr) p.tmap[(idx * D + i) * NC + j] = tm;
            }
            loss += gu[i];
        }
        if (masked) {
            my_loss = (double)loss;
            my_cnt = 1.0;
            // range statistic of MVS4net_loss (models/MVS4Net.py:225-231): no hypothesis within one interval of gt
            bool any = false;
            if (p.inverse) {
                const float itv = fabsf(1.0f / hyp[2] - 1.0f / hyp[1]);
                const float ig = 1.0f / gt;
#pragma unroll
                for (int d = 0; d < D; ++d) any |= fabsf(1.0f / hyp[d] - ig) <= itv;
            } else {
                const float itv = fabsf(hyp[2] - hyp[1]);
#pragma unroll
                for (int d = 0; d < D; ++d) any |= fabsf(hyp[d] - gt) <= itv;
            }
            my_oor = any ? 0.0 : 1.0;
        }
        if (keep) {
            float gl[D];
#pragma unroll
            for (int i = 0; i < D; ++i) gl[i] = 0.0f;
#pragma unroll 1
            for (int t = p.iters - 1; t >= 0; --t) {
                const float* h = hist + (size_t)t * (D + NC) * NT + tid;
                float lu[D], lv[NC], up[D];
#pragma unroll
                for (int i = 0; i < D; ++i) lu[i] = h[i * NT];
#pragma unroll
                for (int j = 0; j < NC; ++j) { lv[j] = h[(D + j) * NT]; v[j] = lm[j] - lv[j]; }
                // u_t = log_nu - LSE_j(K + v_t)
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    gl[i] += gu[i];
#pragma unroll
                    for (int j = 0; j < NC; ++j) gv[j] -= gu[i] * sk_exp(K(i, j) + v[j] - lu[i]);
                }
                // v_t = log_mu - LSE_i(K + u_{t-1}),  u_{t-1} = log_nu - Lu_{t-1}  (u_0 = 0)
                if (t > 0) {
                    const float* hprev = h - (size_t)(D + NC) * NT;
#pragma unroll
                    for (int i = 0; i < D; ++i) up[i] = ln[i] - hprev[i * NT];
                } else {
#pragma unroll
                    for (int i = 0; i < D; ++i) up[i] = 0.0f;
                }
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    float g = 0.0f;
#pragma unroll
                    for (int j = 0; j < NC; ++j) g -= gv[j] * sk_exp(K(i, j) + up[i] - lv[j]);
                    gu[i] = g;
                }
#pragma unroll
                for (int j = 0; j < NC; ++j) gv[j] = 0.0f;
            }
            float* gp = p.grad + b * D * p.plane + pix;
#pragma unroll
            for (int d = 0; d < D; ++d) stg_stream(gp + (size_t)d * p.plane, gl[d] / at[d]);
        }
    }
    if (inside && !masked && p.grad != nullptr) {
        float* gp = p.grad + b * D * p.plane + pix;
#pragma unroll
        for (int d = 0; d < D; ++d) stg_stream(gp + (size_t)d * p.plane, 0.0f);
    }

    // deterministic block partials (fixed shuffle tree, fixed warp order)
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        my_loss += __shfl_xor_sync(0xffffffffu, my_loss, m);
        my_cnt += __shfl_xor_sync(0xffffffffu, my_cnt, m);
        my_oor += __shfl_xor_sync(0xffffffffu, my_oor, m);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = my_loss; red[1][tid >> 5] = my_cnt; red[2][tid >> 5] = my_oor; }
    __syncthreads();
    if (tid < 3) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += red[tid][w];
        p.partials[(size_t)blockIdx.x * 3 + tid] = s;
    }
}

// one block: out[0] = mean loss over masked pixels (NaN when there are none, like torch's mean of an empty tensor),
// out[1] = masked-pixel count, out[2] = out-of-range ratio
__global__ void __launch_bounds__(256) sinkhorn_finalize_kernel(const double* partials, unsigned nblocks, float* out) {
    __shared__ double red[3][8];
    double s[3] = {0.0, 0.0, 0.0};
    for (unsigned i = threadIdx.x; i < nblocks; i += 256) {
#pragma unroll
        for (int k = 0; k < 3; ++k) s[k] += partials[(size_t)i * 3 + k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], m);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int w = 0; w < 8; ++w) t[k] += red[k][w];
        out[0] = (float)(t[0] / t[1]);
        out[1] = (float)t[1];
        out[2] = (float)(t[2] / t[1]);
    }
}

// grad_attn = grad_px * (grad_loss / count): chain rule through the mean over masked pixels
__global__ void __launch_bounds__(256) sinkhorn_bwd_kernel(const float* grad_px, const float* stats, const float* grad_loss,
                                                           float* grad_attn, size_t total) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const float s = __ldg(grad_loss) / __ldg(stats + 1);
    grad_attn[i] = grad_px[i] * s;
}

template <int D, bool CONT, int NT>
static int launch_sinkhorn(const SinkhornParams& p, int smem, cudaStream_t s) {
    static int attr_done[64] = {};  // largest size set per device
    if (smem > 48 * 1024) {
        const int st = ensure_dynamic_smem_bytes(sinkhorn_kernel<D, CONT, NT>, smem, attr_done, "sinkhorn: cudaFuncSetAttribute");
        if (st != MVSTER_OK) return st;
    }
    sinkhorn_kernel<D, CONT, NT><<<(unsigned)((p.total + NT - 1) / NT), NT, smem, s>>>(p);
    return MVSTER_OK;
}

static int pick_threads(int D, int continuous, int iters, bool want_grad) {
    const int per_thread = want_grad ? iters * (2 * D + (continuous ? 1 : 0)) * 4 : 0;
    for (int nt = 128; nt >= 32; nt >>= 1)
        if ((size_t)per_thread * nt <= 200u * 1024u) return nt;
    return 0;
}

template <int D, bool CONT>
static int dispatch_threads(const SinkhornParams& p, int nt, cudaStream_t s) {
    const int smem = (p.grad != nullptr ? p.iters * (D + (CONT ? D + 1 : D)) * 4 : 0) * nt;
    switch (nt) {
        case 128: return launch_sinkhorn<D, CONT, 128>(p, smem, s);
        case 64: return launch_sinkhorn<D, CONT, 64>(p, smem, s);
        default: return launch_sinkhorn<D, CONT, 32>(p, smem, s);
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_sinkhorn_blocks(int B, int D, int H, int W, int iters, int continuous, int want_grad) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || iters < 0) return 0;
    const int nt = pick_threads(D, continuous, iters, want_grad != 0);
    if (nt == 0) return 0;
    const size_t total = (size_t)B * H * W;
    return (int)((total + nt - 1) / nt);
}

extern "C" int mvster_sinkhorn_fwd(const float* gt_depth, const float* hypo, const float* attn, const uint8_t* mask,
                                   int iters, float eps, int continuous, int inverse_depth, float* stats, float* grad_px,
                                   float* tmap, double* partials, int B, int D, int H, int W, void* stream) {
    if (!gt_depth || !hypo || !attn || !mask || !stats || !partials)
        return fail(MVSTER_ERR_BAD_ARG, "sinkhorn_fwd: null pointer");
    if (B <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "sinkhorn_fwd: non-positive dimension");
    if (iters < 0) return fail(MVSTER_ERR_BAD_ARG, "sinkhorn_fwd: iters must be >= 0");
    if (!(eps > 0.0f)) return fail(MVSTER_ERR_BAD_ARG, "sinkhorn_fwd: eps must be > 0");
    if (D != 4 && D != 8) return fail(MVSTER_ERR_UNSUPPORTED, "sinkhorn_fwd: D=%d not in {4,8}", D);
    const int nt = pick_threads(D, continuous, iters, grad_px != nullptr);
    if (nt == 0)
        return fail(MVSTER_ERR_UNSUPPORTED, "sinkhorn_fwd: iters=%d too many for the in-kernel backward (shared memory)",
                    iters);
    DeviceGuard guard(stats);
    if (guard.status != MVSTER_OK) return guard.status;
    SinkhornParams p{gt_depth, hypo, attn, mask, grad_px, tmap, partials, (size_t)H * W, (size_t)B * H * W,
                     iters, inverse_depth, eps};
    cudaStream_t s = (cudaStream_t)stream;
    int st;
    if (D == 4) st = continuous ? dispatch_threads<4, true>(p, nt, s) : dispatch_threads<4, false>(p, nt, s);
    else st = continuous ? dispatch_threads<8, true>(p, nt, s) : dispatch_threads<8, false>(p, nt, s);
    if (st != MVSTER_OK) return st;
    count_launch();
    MVSTER_CHECK_LAUNCH("sinkhorn launch");
    sinkhorn_finalize_kernel<<<1, 256, 0, s>>>(partials, (unsigned)((p.total + nt - 1) / nt), stats);
    count_launch();
    MVSTER_CHECK_LAUNCH("sinkhorn finalize launch");
    return MVSTER_OK;
}

extern "C" int mvster_sinkhorn_bwd(const float* grad_px, const float* stats, const float* grad_loss, float* grad_attn,
                                   int B, int D, int H, int W, void* stream) {
    if (!grad_px || !stats || !grad_loss || !grad_attn) return fail(MVSTER_ERR_BAD_ARG, "sinkhorn_bwd: null pointer");
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "sinkhorn_bwd: non-positive dimension");
    DeviceGuard guard(grad_attn);
    if (guard.status != MVSTER_OK) return guard.status;
    const size_t total = (size_t)B * D * H * W;
    sinkhorn_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad_px, stats, grad_loss,
                                                                                          grad_attn, total);
    count_launch();
    MVSTER_CHECK_LAUNCH("sinkhorn_bwd launch");
    return MVSTER_OK;
}
