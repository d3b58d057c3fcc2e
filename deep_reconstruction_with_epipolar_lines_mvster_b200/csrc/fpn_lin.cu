// FPN4 top-down step in its linearised form, the finest level's replacement for
// fpn_topdown_kernel.
//
// Reference (models/mvs4net_utils.py:488-495):   intra = up2(prev) + inner(lat);   out = out_conv(intra)
// with up2 = bilinear x2 (align_corners=True), inner = 1x1 convolution with bias (CL -> 64), out_conv = 3x3 convolution
// without bias (64 -> CO), zero padding.  Every step is linear, so the 64-channel full-resolution `intra` need not
// exist, not even in shared memory:
//     out[x] = sum_tap W[tap] intra[x + tap]                                            (taps inside the image only)
//            = sum_tap bilinear(P_tap)(x + tap)  +  sum_tap (W[tap] Wi) lat[x + tap]  +  sum_tap W[tap] bi
//     P_tap  = W[tap] prev          9 x CO channels at HALF resolution: one [pixels x 64] x [64 x 9 CO] GEMM (cuBLAS, by
//                                   the caller - a plain library GEMM) instead of a 64 -> CO 3x3 convolution at full
//                                   resolution: 64 x 9 CO FMAs per coarse pixel = a quarter of the work per fine pixel
//     W[tap] Wi                     a CL -> CO 3x3 convolution on the encoder map (composed in float64 on the host)
// Per fine pixel: 9 x 4 x CO (bilinear taps of P) + 9 x CL x CO (lateral) FMAs = 864 for CL = CO = 8 instead of
// 64 x 9 x CO + 64 x (CL + 4) = 5376.  The result differs from the reference's evaluation order by fp32 rounding
// only (parity bound of the FPN tests: 3e-5 of the output range against float64).
//
// A CTA owns a 32 x 8 output tile: the P rows its (tile + halo) pixels interpolate from (<= 7 x 20 coarse pixels x
// 9 CO channels, NHWC, copied with 16-byte loads into a padded shared-memory tile so that the lanes of a quarter-warp
// hit distinct bank groups) and the lateral tile (+1 halo, zero outside the image); a thread owns one pixel and all
// CO channels; the composed lateral weights and the bias terms travel as kernel parameters (uniform constant operands).
#include <string.h>

#include "common.cuh"

namespace mvster {

#ifndef MVSTER_LIN_PY8
#define MVSTER_LIN_PY8 1
#endif
constexpr int kLinTW = 32;      // output tile width; height 8 * PY (a thread owns PY vertically adjacent pixels)
constexpr int kLinRC = 20;      // coarse columns held per CTA
constexpr int kLinLW = 36;      // lateral tile row stride
constexpr int kLinThreads = 256;

template <int CL, int CO>
struct LinParams {
    float wc[9 * CL * CO];  // [tap][ci][co]  composed W[tap] Wi
    float bc[9 * CO];       // [tap][co]      W[tap] bi
    float bsum[CO];         // sum over the nine taps of bc (interior pixels)
    const float* P;         // [B, H/2, W/2, pc] NHWC: P[.., poff + tap*CO + co] = sum_c W[co,c,tap] prev[c]
    const float* lat;       // [B, CL, H, W] planar
    void* feat;             // [B, H, W, CO] NHWC, fp32 or bf16
    int B, H, W;
    int pc, poff;           // channels per coarse pixel of P (>= 9*CO, multiple of 4) and first channel used
    float sy, sx;           // align_corners=True source scale (Hl-1)/(H-1), (Wl-1)/(W-1)
};

template <int CL, int CO, int PY>
struct LinGeom {
    static constexpr int TH = 8 * PY;                  // tile height
    static constexpr int RR = (TH + 2) / 2 + 2;        // coarse rows under tile + halo (scale < 1/2): 7 / 11
    static constexpr int LH = TH + 2;                  // lateral tile rows
    static constexpr int PS = 9 * CO + 4;  // floats per coarse pixel; PS * 4 B = 16 B * odd: columns shift bank groups
    static constexpr int SMEM = (RR * kLinRC * PS + CL * LH * kLinLW) * 4;
};

__device__ __forceinline__ unsigned lin_pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned*>(&v);
}

template <int CL, int CO, int PY, typename OutT>
__global__ void __launch_bounds__(kLinThreads, (CO == 8 && PY == 1) ? 3 : 2) fpn_lin_kernel(const __grid_constant__ LinParams<CL, CO> p) {
    using Gm = LinGeom<CL, CO, PY>;
    constexpr int PS = Gm::PS, PC = 9 * CO, Q = PC / 4, RR = Gm::RR, LH = Gm::LH, TH = Gm::TH;
    extern __shared__ __align__(16) float lin_smem[];
    float* Ps = lin_smem;                    // [RR][kLinRC][PS]
    float* Ls = lin_smem + RR * kLinRC * PS; // [CL][LH][kLinLW]
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int tx0 = blockIdx.x * kLinTW, ty0 = blockIdx.y * TH;
    const int H = p.H, W = p.W, Hl = H / 2, Wl = W / 2;

    // coarse region that the tile + halo interpolates from (source = scale * dst is monotone in dst)
    const int lx0 = (int)(p.sx * (float)max(tx0 - 1, 0));
    const int lx1 = min((int)(p.sx * (float)min(tx0 + kLinTW, W - 1)) + 1, Wl - 1);
    const int ly0 = (int)(p.sy * (float)max(ty0 - 1, 0));
    const int ly1 = min((int)(p.sy * (float)min(ty0 + TH, H - 1)) + 1, Hl - 1);
    const int nc = min(lx1 - lx0 + 1, kLinRC), nr = min(ly1 - ly0 + 1, RR);
    {
        // both tiles are filled with cp.async: every copy of the CTA is in flight before anything is waited for (a
        // register-staged loop serialised the L2 latency of its ~20 iterations per thread)
        constexpr int RCQ = kLinRC * Q;  // compile-time divisors: no integer division in the copy loop
        const uint32_t ps_addr = (uint32_t)__cvta_generic_to_shared(Ps), ls_addr = (uint32_t)__cvta_generic_to_shared(Ls);
        for (int i = tid; i < nr * RCQ; i += kLinThreads) {
            const int r = i / RCQ, rem = i - r * RCQ;
            const int c = rem / Q, j = rem - c * Q;
            if (c >= nc) continue;
            const float* src = p.P + (((size_t)b * Hl + ly0 + r) * Wl + lx0 + c) * p.pc + p.poff + 4 * j;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ps_addr + (uint32_t)(((r * kLinRC + c) * PS + 4 * j) * 4)), "l"(src) : "memory");
        }
        constexpr int LHW = LH * (kLinTW + 2);
        for (int i = tid; i < CL * LHW; i += kLinThreads) {
            const int ci = i / LHW, rem = i - ci * LHW;
            const int ry = rem / (kLinTW + 2), rx = rem - ry * (kLinTW + 2);
            const int gy = ty0 - 1 + ry, gx = tx0 - 1 + rx;
            const bool inside = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
            // outside the image: source size 0 = zero fill (the clamped address is not read)
            const float* src = p.lat + (((size_t)b * CL + ci) * H + min(max(gy, 0), H - 1)) * W + min(max(gx, 0), W - 1);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ls_addr + (uint32_t)(((ci * LH + ry) * kLinLW + rx) * 4)), "l"(src), "r"(inside ? 4 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const int px = tid & 31, py = (tid >> 5) * PY;
    const int x = tx0 + px, y = ty0 + py;   // the thread's pixels: (x, y) .. (x, y + PY - 1)
    if (x >= W || y >= H) return;

    // the three source columns of the taps (kx = 0..2 <-> dx = -1, 0, 1): coarse column offsets and weights (ATen's
    // source = scale * dst arithmetic); a tap outside the image gets zero weights
    int ca[3], cb[3];
    float wxa[3], wxb[3];
    bool vx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int qx = x + k - 1;
        vx[k] = (unsigned)qx < (unsigned)W;
        const float fx = p.sx * (float)max(qx, 0);
        const int x0 = (int)fx;
        ca[k] = min(x0 - lx0, kLinRC - 1) * PS;
        cb[k] = min(x0 + (x0 < Wl - 1) - lx0, kLinRC - 1) * PS;
        const float l = fx - (float)x0;
        wxa[k] = vx[k] ? 1.0f - l : 0.0f;
        wxb[k] = vx[k] ? l : 0.0f;
    }
    // the PY + 2 source rows
    int ra[PY + 2], rb[PY + 2];
    float wya[PY + 2], wyb[PY + 2];
    bool vy[PY + 2];
#pragma unroll
    for (int k = 0; k < PY + 2; ++k) {
        const int qy = y + k - 1;
        vy[k] = (unsigned)qy < (unsigned)H;
        const float fy = p.sy * (float)min(max(qy, 0), H - 1);
        const int y0 = (int)fy;
        ra[k] = min(y0 - ly0, RR - 1) * (kLinRC * PS);
        rb[k] = min(y0 + (y0 < Hl - 1) - ly0, RR - 1) * (kLinRC * PS);
        const float l = fy - (float)y0;
        wya[k] = vy[k] ? 1.0f - l : 0.0f;
        wyb[k] = vy[k] ? l : 0.0f;
    }
    float acc[PY][CO];
#pragma unroll
    for (int q = 0; q < PY; ++q)
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[q][co] = p.bsum[co];

#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int tap = ky * 3 + kx;
            // bilinear(P_tap) at the tap's pixel, per owned pixel (row ky + q of the PY + 2 source rows)
#pragma unroll
            for (int q = 0; q < PY; ++q) {
                const int k = ky + q;
                const float w00 = wya[k] * wxa[kx], w01 = wya[k] * wxb[kx], w10 = wyb[k] * wxa[kx], w11 = wyb[k] * wxb[kx];
                const float* p00 = Ps + ra[k] + ca[kx] + tap * CO;
                const float* p01 = Ps + ra[k] + cb[kx] + tap * CO;
                const float* p10 = Ps + rb[k] + ca[kx] + tap * CO;
                const float* p11 = Ps + rb[k] + cb[kx] + tap * CO;
#pragma unroll
                for (int c4 = 0; c4 < CO; c4 += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(p00 + c4), bq = *reinterpret_cast<const float4*>(p01 + c4);
                    const float4 c = *reinterpret_cast<const float4*>(p10 + c4), d = *reinterpret_cast<const float4*>(p11 + c4);
                    acc[q][c4 + 0] = fmaf(w00, a.x, fmaf(w01, bq.x, fmaf(w10, c.x, fmaf(w11, d.x, acc[q][c4 + 0]))));
                    acc[q][c4 + 1] = fmaf(w00, a.y, fmaf(w01, bq.y, fmaf(w10, c.y, fmaf(w11, d.y, acc[q][c4 + 1]))));
                    acc[q][c4 + 2] = fmaf(w00, a.z, fmaf(w01, bq.z, fmaf(w10, c.z, fmaf(w11, d.z, acc[q][c4 + 2]))));
                    acc[q][c4 + 3] = fmaf(w00, a.w, fmaf(w01, bq.w, fmaf(w10, c.w, fmaf(w11, d.w, acc[q][c4 + 3]))));
                }
            }
            // composed lateral convolution: the tile is zero outside the image, every weight serves the PY pixels
            const float* lp = Ls + (py + ky) * kLinLW + px + kx;
#pragma unroll
            for (int ci = 0; ci < CL; ++ci) {
                float v[PY];
#pragma unroll
                for (int q = 0; q < PY; ++q) v[q] = lp[(ci * LH + q) * kLinLW];
#pragma unroll
                for (int co = 0; co < CO; ++co) {
                    const float wv = p.wc[(tap * CL + ci) * CO + co];
#pragma unroll
                    for (int q = 0; q < PY; ++q) acc[q][co] = fmaf(wv, v[q], acc[q][co]);
                }
            }
            // bias term: bsum holds all nine taps; take this one out again where it falls outside the image (borders)
#pragma unroll
            for (int q = 0; q < PY; ++q)
                if (!(vy[ky + q] && vx[kx])) {
#pragma unroll
                    for (int co = 0; co < CO; ++co) acc[q][co] -= p.bc[tap * CO + co];
                }
        }
    }

#pragma unroll
    for (int q = 0; q < PY; ++q) {
        if (y + q >= H) break;
        const size_t fo = (((size_t)b * H + y + q) * W + x) * CO;
        if constexpr (sizeof(OutT) == 4) {
            float* fp = static_cast<float*>(p.feat) + fo;
#pragma unroll
            for (int c4 = 0; c4 < CO; c4 += 4)
                *reinterpret_cast<float4*>(fp + c4) = make_float4(acc[q][c4], acc[q][c4 + 1], acc[q][c4 + 2], acc[q][c4 + 3]);
        } else {  // bf16, round to nearest even as torch's .to(bfloat16): 8 channels = one 16-byte store
            __nv_bfloat16* fp = static_cast<__nv_bfloat16*>(p.feat) + fo;
#pragma unroll
            for (int c4 = 0; c4 < CO; c4 += 8) {
                uint4 v;
                v.x = lin_pack_bf16x2(acc[q][c4], acc[q][c4 + 1]);
                v.y = lin_pack_bf16x2(acc[q][c4 + 2], acc[q][c4 + 3]);
                v.z = lin_pack_bf16x2(acc[q][c4 + 4], acc[q][c4 + 5]);
                v.w = lin_pack_bf16x2(acc[q][c4 + 6], acc[q][c4 + 7]);
                *reinterpret_cast<uint4*>(fp + c4) = v;
            }
        }
    }
}

template <int CL, int CO, typename OutT>
static int launch_lin(const float* P, int pc, int poff, const float* lat, void* feat, const float* wc, const float* bc,
                      int B, int H, int W, cudaStream_t s) {
    constexpr int PY = (CO == 8) ? MVSTER_LIN_PY8 : 1;  // pixels per thread; 2 (tile 32 x 16, weights shared) measured slower at (8,8): two CTAs per SM
    using Gm = LinGeom<CL, CO, PY>;
    static thread_local LinParams<CL, CO> p;
    static_assert(sizeof(LinParams<CL, CO>) <= 32000, "weights must fit the kernel-parameter space");
    memcpy(p.wc, wc, sizeof(p.wc));
    memcpy(p.bc, bc, sizeof(p.bc));
    for (int co = 0; co < CO; ++co) {
        double t = 0.0;
        for (int tap = 0; tap < 9; ++tap) t += (double)bc[tap * CO + co];
        p.bsum[co] = (float)t;
    }
    p.P = P; p.lat = lat; p.feat = feat; p.B = B; p.H = H; p.W = W; p.pc = pc; p.poff = poff;
    const int Hl = H / 2, Wl = W / 2;
    p.sy = H > 1 ? (float)(Hl - 1) / (float)(H - 1) : 0.f;  // ATen area_pixel_compute_scale, align_corners=True
    p.sx = W > 1 ? (float)(Wl - 1) / (float)(W - 1) : 0.f;
    constexpr int SMEM = Gm::SMEM;
    static int attr_done[64] = {};  // largest size set per device
    const int st = ensure_dynamic_smem_bytes(fpn_lin_kernel<CL, CO, PY, OutT>, SMEM, attr_done, "fpn_topdown_lin: cudaFuncSetAttribute");
    if (st != MVSTER_OK) return st;
    dim3 grid((W + kLinTW - 1) / kLinTW, (H + Gm::TH - 1) / Gm::TH, B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "fpn_topdown_lin: grid too large");
    fpn_lin_kernel<CL, CO, PY, OutT><<<grid, kLinThreads, SMEM, s>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("fpn_topdown_lin launch");
    return MVSTER_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Projection of an `intra` that is itself a top-down result, without forming it: with intra = up2(prev) + Wi lat + bi,
//     P = Wp intra = up2(Wp prev) + (Wp Wi) lat + Wp bi                      (NP = 9 * Cout of the NEXT level channels)
// Q = Wp prev comes from the caller's GEMM at the coarser resolution (NHWC, channels q_off .. q_off + NP - 1 of a pixel
// of NQ channels); a thread owns one pixel and all NP channels: CL x NP lateral FMAs with uniform constant weights, four
// bilinear taps of NP contiguous floats, NP contiguous floats out.
// ---------------------------------------------------------------------------------------------------------------------
template <int CL, int NP>
struct ProjUpParams {
    float wl[CL * NP];  // [ci][n]  Wp Wi
    float bl[NP];       //          Wp bi
    const float* Q;     // [B, H/2, W/2, NQ] NHWC
    const float* lat;   // [B, CL, H, W] planar
    float* P;           // [B, H, W, NP] NHWC
    int B, H, W, NQ, q_off;
    float sy, sx;
};

// NT channels per thread: the NP channels of a pixel are split over NP / NT CTAs (blockIdx.x % (NP / NT)), which
// triples the loads in flight - with all 72 channels in one thread the kernel waited on its gathers (ncu: 27 stall
// cycles on the long scoreboard per issued instruction, 25 % issue utilisation)
template <int CL, int NP, int NT>
__global__ void __launch_bounds__(128) fpn_proj_up_kernel(const __grid_constant__ ProjUpParams<CL, NP> p) {
    constexpr int NG = NP / NT;
    static_assert(NP % NT == 0 && NT % 4 == 0, "channel groups of whole float4s");
    const int H = p.H, W = p.W, Hl = H / 2, Wl = W / 2;
    const int g = blockIdx.x % NG;
    const int x = (blockIdx.x / NG) * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    // bilinear x2, align_corners=True (ATen: source = scale * dst)
    const float fy = p.sy * (float)y, fx = p.sx * (float)x;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < Hl - 1), x1 = x0 + (x0 < Wl - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float w00 = (1.0f - ly) * (1.0f - lx), w01 = (1.0f - ly) * lx, w10 = ly * (1.0f - lx), w11 = ly * lx;
    const float* qb = p.Q + (size_t)b * Hl * Wl * p.NQ + p.q_off + g * NT;
    const float4* q00 = reinterpret_cast<const float4*>(qb + ((size_t)y0 * Wl + x0) * p.NQ);
    const float4* q01 = reinterpret_cast<const float4*>(qb + ((size_t)y0 * Wl + x1) * p.NQ);
    const float4* q10 = reinterpret_cast<const float4*>(qb + ((size_t)y1 * Wl + x0) * p.NQ);
    const float4* q11 = reinterpret_cast<const float4*>(qb + ((size_t)y1 * Wl + x1) * p.NQ);
    float4 ta[NT / 4], tb[NT / 4], tc[NT / 4], td[NT / 4];
#pragma unroll
    for (int j = 0; j < NT / 4; ++j) { ta[j] = __ldg(q00 + j); tb[j] = __ldg(q01 + j); tc[j] = __ldg(q10 + j); td[j] = __ldg(q11 + j); }
    float lv[CL];
    const float* lp = p.lat + ((size_t)b * CL * H + y) * W + x;
#pragma unroll
    for (int ci = 0; ci < CL; ++ci) lv[ci] = __ldg(lp + (size_t)ci * H * W);
    float acc[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[n] = 0.0f;
    // the channel group is CTA-uniform but not a compile-time constant: one code copy per group keeps the weights
    // uniform constant operands
#pragma unroll
    for (int gg = 0; gg < NG; ++gg) {
        if (g != gg) continue;
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[n] = p.bl[gg * NT + n];
#pragma unroll
        for (int ci = 0; ci < CL; ++ci)
#pragma unroll
            for (int n = 0; n < NT; ++n) acc[n] = fmaf(p.wl[ci * NP + gg * NT + n], lv[ci], acc[n]);
    }
    float4* op = reinterpret_cast<float4*>(p.P + (((size_t)b * H + y) * W + x) * NP + g * NT);
#pragma unroll
    for (int j = 0; j < NT / 4; ++j) {
        float4 o;
        o.x = fmaf(w00, ta[j].x, fmaf(w01, tb[j].x, fmaf(w10, tc[j].x, fmaf(w11, td[j].x, acc[4 * j + 0]))));
        o.y = fmaf(w00, ta[j].y, fmaf(w01, tb[j].y, fmaf(w10, tc[j].y, fmaf(w11, td[j].y, acc[4 * j + 1]))));
        o.z = fmaf(w00, ta[j].z, fmaf(w01, tb[j].z, fmaf(w10, tc[j].z, fmaf(w11, td[j].z, acc[4 * j + 2]))));
        o.w = fmaf(w00, ta[j].w, fmaf(w01, tb[j].w, fmaf(w10, tc[j].w, fmaf(w11, td[j].w, acc[4 * j + 3]))));
        op[j] = o;
    }
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_fpn_project_up(const float* Q, int q_channels, int q_off, const float* lat, float* P,
                                     const float* wl_host, const float* bl_host, int B, int Clat, int NP, int H, int W,
                                     void* stream) {
    if (!Q || !lat || !P || !wl_host || !bl_host) return fail(MVSTER_ERR_BAD_ARG, "fpn_project_up: null pointer");
    if (B <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1))
        return fail(MVSTER_ERR_BAD_ARG, "fpn_project_up: H, W must be even and >= 2");
    if (q_off < 0 || q_off + NP > q_channels || (q_off & 3) || (q_channels & 3))
        return fail(MVSTER_ERR_BAD_ARG, "fpn_project_up: bad channel window of Q");
    if (((uintptr_t)Q) % 16 || ((uintptr_t)P) % 16) return fail(MVSTER_ERR_ALIGN, "fpn_project_up: Q and P must be 16-byte aligned");
    if (Clat != 16 || NP != 72)
        return fail(MVSTER_ERR_UNSUPPORTED, "fpn_project_up: no kernel for Clat=%d NP=%d (built: (16,72))", Clat, NP);
    DeviceGuard guard(P);
    if (guard.status != MVSTER_OK) return guard.status;
    static thread_local ProjUpParams<16, 72> p;
    memcpy(p.wl, wl_host, sizeof(p.wl));
    memcpy(p.bl, bl_host, sizeof(p.bl));
    p.Q = Q; p.lat = lat; p.P = P; p.B = B; p.H = H; p.W = W; p.NQ = q_channels; p.q_off = q_off;
    p.sy = (float)(H / 2 - 1) / (float)(H - 1);
    p.sx = (float)(W / 2 - 1) / (float)(W - 1);
    constexpr int NT = 24;  // channels per thread
    dim3 grid(((W + 31) / 32) * (72 / NT), (H + 3) / 4, B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "fpn_project_up: grid too large");
    fpn_proj_up_kernel<16, 72, NT><<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("fpn_project_up launch");
    return MVSTER_OK;
}

extern "C" int mvster_fpn_topdown_lin(const float* P, int p_channels, int p_off, const float* lat, void* feat,
                                      int feat_dtype, const float* wc_host, const float* bc_host, int B, int Clat,
                                      int Cout, int H, int W, void* stream) {
    if (!P || !lat || !feat || !wc_host || !bc_host) return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown_lin: null pointer");
    if (feat_dtype != MVSTER_F32 && feat_dtype != MVSTER_BF16)
        return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown_lin: feat_dtype must be MVSTER_F32 or MVSTER_BF16");
    if (B <= 0 || H < 2 || W < 2 || (H & 1) || (W & 1))
        return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown_lin: H, W must be even and >= 2");
    if (((uintptr_t)feat) % 16 || ((uintptr_t)P) % 16) return fail(MVSTER_ERR_ALIGN, "fpn_topdown_lin: P and feat must be 16-byte aligned");
    if (p_off < 0 || p_off + 9 * Cout > p_channels || (p_off & 3) || (p_channels & 3))
        return fail(MVSTER_ERR_BAD_ARG, "fpn_topdown_lin: bad channel window of P");
    DeviceGuard guard(feat);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    const bool bf = feat_dtype == MVSTER_BF16;
#define MVSTER_LIN_ARGS P, p_channels, p_off, lat, feat, wc_host, bc_host, B, H, W, s
    if (Clat == 8 && Cout == 8)
        return bf ? launch_lin<8, 8, __nv_bfloat16>(MVSTER_LIN_ARGS) : launch_lin<8, 8, float>(MVSTER_LIN_ARGS);
    if (Clat == 16 && Cout == 16)
        return bf ? launch_lin<16, 16, __nv_bfloat16>(MVSTER_LIN_ARGS) : launch_lin<16, 16, float>(MVSTER_LIN_ARGS);
#undef MVSTER_LIN_ARGS
    return fail(MVSTER_ERR_UNSUPPORTED, "fpn_topdown_lin: no kernel for Clat=%d Cout=%d (built: (8,8), (16,16))", Clat, Cout);
}
