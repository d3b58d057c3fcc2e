// K2b: photometric / geometric-consistency filter of test_mvs4.py.
//   reproject_with_depth          test_mvs4.py:612-649
//   check_geometric_consistency   test_mvs4.py:653-670
//   mask fusion in filter_depth   test_mvs4.py:716,738,744,746,749
// The reference builds dense float64 NumPy temporaries per (ref, src) pair and samples the source depth with
// cv2.remap; here one thread owns one reference pixel, walks all S source views of its reference view and keeps
// the vote count and depth sum in registers.  The projective math stays in float64 (the 1-px / 1 % thresholds are
// compared on float64 / float32 quantities exactly as NumPy does); the 3x3 / 3x1 matrices of every pair are
// pre-composed in float64 on the host (they are the same np.linalg.inv / matmul products the reference forms).
// cv2.remap(INTER_LINEAR) is emulated bit-for-bit in its coordinate handling: float32 map, fixed point with 5
// fractional bits (round-half-even), constant-0 border.
#include <cstddef>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace mvster {

// Per-(ref, src) pair camera chain, float64, pre-composed on the host from the same inverses / products the
// reference forms with np.linalg.inv / np.matmul (test_mvs4.py:619-647):
//   q      = F * ([x,y,1] * d_ref) + f          F = K_src R_rel K_ref^-1,  f = K_src t_rel     (rel  = E_src E_ref^-1)
//   (u,v)  = q.xy / q.z
//   X_ref' = Bk * ([u,v,1] * d_src) + tb        Bk = R_back K_src^-1,      tb = t_back         (back = E_ref E_src^-1)
//   q'     = G * ([u,v,1] * d_src) + g          G = K_ref Bk,              g = K_ref t_back
// Only the z row of X_ref' is needed (depth_reprojected).
struct alignas(16) PairCam {
    double F[9], f[3];
    double G[9], g[3];
    double Bz[3], tbz;
    float Gf[9], gf[3], Bzf[3], tbzf;  // float32 copies of the back-projection for the fast path (see check_pair)
    int src;  // source view index (into the depth stack), < 0 = skip
    int pad[3];
};
static_assert(sizeof(PairCam) % 16 == 0 && offsetof(PairCam, Gf) % 16 == 0,
              "PairCam is copied as doubles and read with 16-byte shared-memory loads");

__device__ __forceinline__ void mat3_vec(const double* m, double x, double y, double z, double& ox, double& oy,
                                         double& oz) {
    ox = m[0] * x + m[1] * y + m[2] * z;
    oy = m[3] * x + m[4] * y + m[5] * z;
    oz = m[6] * x + m[7] * y + m[8] * z;
}

// cv2.remap(INTER_LINEAR, BORDER_CONSTANT=0) on a float32 image: coordinates are converted to fixed point with
// INTER_BITS=5 (cvRound = round-half-even of x*32), tap = floor, weights = (frac/32) products in float32.
__device__ __forceinline__ float remap_linear(const float* __restrict__ img, int H, int W, float mx, float my) {
    if (!(isfinite(mx) && isfinite(my))) return 0.0f;
    // x * 32 is exact in float32 (a power of two), so rounding the float product equals rounding the double one
    const int ix = __float2int_rn(fminf(fmaxf(mx * 32.0f, -1e8f), 1e8f));
    const int iy = __float2int_rn(fminf(fmaxf(my * 32.0f, -1e8f), 1e8f));
    const int x0 = ix >> 5, y0 = iy >> 5;
    const float fx = (float)(ix & 31) * (1.0f / 32.0f);
    const float fy = (float)(iy & 31) * (1.0f / 32.0f);
    if (x0 < -1 || x0 >= W || y0 < -1 || y0 >= H) return 0.0f;
    const bool vx0 = x0 >= 0, vx1 = x0 + 1 < W, vy0 = y0 >= 0, vy1 = y0 + 1 < H;
    const float v00 = (vx0 && vy0) ? __ldg(img + (size_t)y0 * W + x0) : 0.0f;
    const float v01 = (vx1 && vy0) ? __ldg(img + (size_t)y0 * W + x0 + 1) : 0.0f;
    const float v10 = (vx0 && vy1) ? __ldg(img + (size_t)(y0 + 1) * W + x0) : 0.0f;
    const float v11 = (vx1 && vy1) ? __ldg(img + (size_t)(y0 + 1) * W + x0 + 1) : 0.0f;
    const float gx = 1.0f - fx, gy = 1.0f - fy;
    // no FMA contraction: OpenCV multiplies by a pre-rounded float weight table and adds left to right
    float acc = __fmul_rn(v00, __fmul_rn(gx, gy));
    acc = __fadd_rn(acc, __fmul_rn(v01, __fmul_rn(fx, gy)));
    acc = __fadd_rn(acc, __fmul_rn(v10, __fmul_rn(gx, fy)));
    acc = __fadd_rn(acc, __fmul_rn(v11, __fmul_rn(fx, fy)));
    return acc;
}

struct PairResult {
    bool mask;
    float depth_reprojected;  // before the [~mask] = 0 overwrite
    float x_src, y_src;
};

// 1 / x in float64 without the IEEE-division slow path: MUFU.RCP64H seed + two Newton steps (<= 2 ulp).  The
// quotients only feed values that are rounded to float32 (the remap coordinates) or compared with a threshold, so
// a last-bit difference from a correctly rounded division moves a result with probability ~1e-9.
__device__ __forceinline__ double fast_rcp64(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// remap_linear for the fused kernel in two phases, same results.  remap_taps: coordinates -> four tap values and the
// two fractions, with no conversion instructions (F2I / I2F run on the quarter-rate XU pipe) and NO branches, so that
// the loads of all the pixels a thread owns are in flight together (the kernel was latency-bound on them).
//   x * 32 + 1.5 * 2^23 rounds to the nearest-even integer in the mantissa (exact for |x * 32| <= 2^22; the clamp to
//   +-2^21 sends non-finite and far-away coordinates to a tap index outside any image of <= 32768 px, which is the
//   constant-0 border); the 5 fractional bits become a float through the 2^23 exponent trick; taps outside the image
//   are read from the clamped address and replaced by 0.
// remap_blend: OpenCV's weight products and left-to-right sum.
struct RemapTaps {
    float v00, v01, v10, v11, gx, fx, gy, fy;  // tap values, (1 - fx, fx, 1 - fy, fy) with out-of-image taps zeroed
};
__device__ __forceinline__ RemapTaps remap_taps(const float* __restrict__ stack, int img_off, int H, int W, float mx,
                                                float my) {
    // `stack + img_off` is the image; 32-bit element offsets (the host checks V * H * W < 2^31): one IMAD.WIDE per load
    constexpr float kMagic = 12582912.0f, kLim = 2097152.0f;  // 1.5 * 2^23, 2^21
    const float tx = fminf(fmaxf(mx * 32.0f, -kLim), kLim) + kMagic;
    const float ty = fminf(fmaxf(my * 32.0f, -kLim), kLim) + kMagic;
    const int ix = __float_as_int(tx) - 0x4B400000, iy = __float_as_int(ty) - 0x4B400000;
    const int x0 = ix >> 5, y0 = iy >> 5;
    RemapTaps t;
    const float fx = (__int_as_float(0x4B000000 | (ix & 31)) - 8388608.0f) * (1.0f / 32.0f);
    const float fy = (__int_as_float(0x4B000000 | (iy & 31)) - 8388608.0f) * (1.0f / 32.0f);
    const bool vx0 = (unsigned)x0 < (unsigned)W, vx1 = (unsigned)(x0 + 1) < (unsigned)W;
    const bool vy0 = (unsigned)y0 < (unsigned)H, vy1 = (unsigned)(y0 + 1) < (unsigned)H;
    const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
    const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
    const int o0 = img_off + yc0 * W, o1 = img_off + yc1 * W;
    t.v00 = __ldg(stack + (o0 + xc0));
    t.v01 = __ldg(stack + (o0 + xc1));
    t.v10 = __ldg(stack + (o1 + xc0));
    t.v11 = __ldg(stack + (o1 + xc1));
    // a tap outside the image counts as 0 (constant border): its weight is zeroed instead of its value (depth maps are
    // finite, so 0 * v == 0 * 0 up to the sign of zero, which no later operation distinguishes)
    t.gx = vx0 ? 1.0f - fx : 0.0f;
    t.fx = vx1 ? fx : 0.0f;
    t.gy = vy0 ? 1.0f - fy : 0.0f;
    t.fy = vy1 ? fy : 0.0f;
    return t;
}
__device__ __forceinline__ float remap_blend(const RemapTaps& t) {
    // no FMA contraction: OpenCV multiplies by a pre-rounded float weight table and adds left to right
    float acc = __fmul_rn(t.v00, __fmul_rn(t.gx, t.gy));
    acc = __fadd_rn(acc, __fmul_rn(t.v01, __fmul_rn(t.fx, t.gy)));
    acc = __fadd_rn(acc, __fmul_rn(t.v10, __fmul_rn(t.gx, t.fy)));
    acc = __fadd_rn(acc, __fmul_rn(t.v11, __fmul_rn(t.fx, t.fy)));
    return acc;
}

// The reference does everything in float64.  Here only what needs it does:
//   * reference pixel -> source pixel stays float64: the remap coordinate is rounded to 1/32 px, and a float32 chain
//     (error ~2e-4 px) would pick the neighbouring sub-pixel bin for ~2.6 % of the pixels;
//   * the back-projection runs in float32 (errors ~1e-4 px / ~1.5e-7 relative depth against thresholds of 1 px / 1 %),
//     and only a pixel whose squared distance is within 0.8 % of the threshold or whose relative depth difference is
//     within 0.03 % of its threshold (fp32 worst case: 7e-4 / 3e-5 of the thresholds at 640-px coordinates) is redone in float64 (EXACT = the reference arithmetic) - a fraction of a percent
//     of the pixels, so the masks are those of the float64 chain; depth_reprojected differs from the float64 value
//     by a few float32 ulps (tolerance of the parity tests: 2e-3).
// B200 issues DFMA at half rate and every float<->double conversion goes through the quarter-rate XU pipe: the
// all-float64 version was issue-bound at 82 % (profiles/r02_filter_ncu.md).
__device__ __forceinline__ PairResult check_pair(const PairCam& c, const float* __restrict__ depth_src, int H, int W,
                                                 int x, int y, float d_ref, double pix_thr2, float rel_thr) {
    PairResult r;
    const double dr = (double)d_ref;
    // step 1: reference pixel -> 3-D -> source pixel (:619-626)
    double qx, qy, qz;
    mat3_vec(c.F, (double)x * dr, (double)y * dr, dr, qx, qy, qz);
    qx += c.f[0]; qy += c.f[1]; qz += c.f[2];
    const double iq = fast_rcp64(qz);
    const double u = qx * iq, v = qy * iq;
    r.x_src = (float)u;  // :630-631
    r.y_src = (float)v;
    // step 2: sample the source depth and project back (:632-647)
    const float ds = remap_linear(depth_src, H, W, r.x_src, r.y_src);
    const float fxr = (float)x, fyr = (float)y;
    {   // float32 back-projection
        const float ux = r.x_src * ds, vx = r.y_src * ds;
        const float px = fmaf(c.Gf[0], ux, fmaf(c.Gf[1], vx, fmaf(c.Gf[2], ds, c.gf[0])));
        const float py = fmaf(c.Gf[3], ux, fmaf(c.Gf[4], vx, fmaf(c.Gf[5], ds, c.gf[1])));
        const float pz = fmaf(c.Gf[6], ux, fmaf(c.Gf[7], vx, fmaf(c.Gf[8], ds, c.gf[2])));
        r.depth_reprojected = fmaf(c.Bzf[0], ux, fmaf(c.Bzf[1], vx, fmaf(c.Bzf[2], ds, c.tbzf)));
        const float ip = fast_rcp(pz);
        const float dx = px * ip - fxr, dy = py * ip - fyr;
        const float dist2 = fmaf(dx, dx, dy * dy);
        const float rel = fabsf(r.depth_reprojected - d_ref) / d_ref;
        const float thr2 = (float)pix_thr2;
        r.mask = (dist2 < thr2) && (rel < rel_thr);
        const bool near = (fabsf(dist2 - thr2) < 0.008f * thr2) || (fabsf(rel - rel_thr) < 3e-4f * rel_thr);
        if (!near) return r;
    }
    // EXACT: the reference's float64 chain for the few pixels next to a threshold
    const double dsd = (double)ds;
    const double ux = u * dsd, vx = v * dsd;
    double px, py, pz;
    mat3_vec(c.G, ux, vx, dsd, px, py, pz);
    px += c.g[0]; py += c.g[1]; pz += c.g[2];
    r.depth_reprojected = (float)(c.Bz[0] * ux + c.Bz[1] * vx + c.Bz[2] * dsd + c.tbz);
    const double ip = 1.0 / pz;
    const float xr = (float)(px * ip), yr = (float)(py * ip);
    // :661-667: dist is float64 (float32 maps minus int64 grids) - compared squared, sqrt is monotone; the relative
    // depth difference is float32
    const double dx = (double)xr - (double)x, dy = (double)yr - (double)y;
    const double dist2 = dx * dx + dy * dy;
    const float rel = fabsf(r.depth_reprojected - d_ref) / d_ref;
    r.mask = (dist2 < pix_thr2) && (rel < rel_thr);
    return r;
}

__global__ void __launch_bounds__(256) geo_check_pair_kernel(const float* __restrict__ depth_ref,
                                                             const float* __restrict__ depth_src,
                                                             const __grid_constant__ PairCam cam,
                                                             double pix_thr2, float rel_thr,
                                                             uint8_t* __restrict__ mask,
                                                             float* __restrict__ depth_rep, float* __restrict__ x2d,
                                                             float* __restrict__ y2d, int H, int W) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    // the 560-byte camera block is a kernel parameter: it sits in the constant bank and is read with uniform loads
    const size_t i = (size_t)y * W + x;
    const PairResult r = check_pair(cam, depth_src, H, W, x, y, depth_ref[i], pix_thr2, rel_thr);
    mask[i] = r.mask ? 1 : 0;
    depth_rep[i] = r.mask ? r.depth_reprojected : 0.0f;  // :668
    x2d[i] = r.x_src;
    y2d[i] = r.y_src;
}

struct GeoFilterParams {
    const float* depths;
    const float* confs;
    const PairCam* cams;  // [R, S]
    const int* refs;      // [R]
    uint8_t* photo;
    uint8_t* geo;
    uint8_t* final_mask;
    float* depth_avg;
    int* geo_sum;
    int S, H, W;
    double pix_thr2;
    float rel_thr, photo_thr;
    int geo_thr;
};

// EXACT: the reference's float64 back-projection (:632-667) for the few pixels next to a threshold; out of line so
// that the hot loop keeps its registers
struct ExactResult {
    float depth_rep;
    int mask;
};
__device__ __noinline__ ExactResult exact_back_projection(const PairCam& c, double u, double v, float ds, int x, int y,
                                                          float d_ref, double pix_thr2, float rel_thr) {
    const double dsd = (double)ds;
    const double ux = u * dsd, vx = v * dsd;
    double px, py, pz;
    mat3_vec(c.G, ux, vx, dsd, px, py, pz);
    px += c.g[0]; py += c.g[1]; pz += c.g[2];
    ExactResult e;
    e.depth_rep = (float)(c.Bz[0] * ux + c.Bz[1] * vx + c.Bz[2] * dsd + c.tbz);
    const double ip = 1.0 / pz;
    const float xr = (float)(px * ip), yr = (float)(py * ip);
    const double dx = (double)xr - (double)x, dy = (double)yr - (double)y;
    const double dist2 = dx * dx + dy * dy;
    const float rel = fabsf(e.depth_rep - d_ref) / d_ref;
    e.mask = ((dist2 < pix_thr2) && (rel < rel_thr)) ? 1 : 0;
    return e;
}

#ifndef MVSTER_FILTER_MINB2
#define MVSTER_FILTER_MINB2 4
#endif
#ifndef MVSTER_FILTER_MINB
#define MVSTER_FILTER_MINB(PPT) ((PPT) == 4 ? 2 : ((PPT) == 2 ? MVSTER_FILTER_MINB2 : 5))
#endif
// Fused filter: a thread owns PPT vertically adjacent reference pixels and walks the S source views.  Per pair the
// camera block is read once per thread (16-byte shared-memory broadcasts) and serves PPT pixels; the float64 ray
// A = F [x, y, 1]^T is formed once per pair and stepped down the column with three additions (A is affine in y),
// so a pixel costs 3 + 4 + 2 float64 operations (q = A d + f, reciprocal, u / v) instead of 9 + 4 + 2, and the PPT
// independent chains hide each other's latency.  Same decisions as check_pair: float64 up to the remap coordinate,
// float32 back-projection with the float64 recheck next to the thresholds.
template <int PPT>
__global__ void __launch_bounds__(256, MVSTER_FILTER_MINB(PPT)) geo_filter_kernel(const GeoFilterParams p) {
    extern __shared__ __align__(16) double cam_s[];  // the S camera blocks of this reference view (uniform LDS broadcasts)
    const int r = blockIdx.z;
    {
        const double* src = reinterpret_cast<const double*>(p.cams + (size_t)r * p.S);
        const int n = p.S * (int)(sizeof(PairCam) / 8);
        for (int i = threadIdx.x; i < n; i += 256) cam_s[i] = src[i];
    }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yb = (blockIdx.y * 8 + (threadIdx.x >> 5)) * PPT;
    if (x >= p.W || yb >= p.H) return;
    const int H = p.H, W = p.W;
    const size_t plane = (size_t)H * W;
    const int ref = p.refs[r];
    const float* dref_p = p.depths + (size_t)ref * plane;
    float d_ref[PPT], inv_d[PPT], sum[PPT], yf[PPT];
    double dr[PPT];
    int votes[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int y = min(yb + k, H - 1);  // rows past the image repeat the last one and are not stored
        d_ref[k] = dref_p[(size_t)y * W + x];
        dr[k] = (double)d_ref[k];
        inv_d[k] = 1.0f / d_ref[k];
        yf[k] = (float)(yb + k);
        sum[k] = 0.0f;  // float32 running sum, like sum(list of float32 arrays) at :744
        votes[k] = 0;
    }
    const double xd = (double)x, yd = (double)yb;
    const float xf = (float)x;
    const float thr2 = (float)p.pix_thr2, rel_thr = p.rel_thr;
    const float thr2_band = 0.008f * thr2, rel_band = 3e-4f * rel_thr;
    const PairCam* cams = reinterpret_cast<const PairCam*>(cam_s);
#pragma unroll 1
    for (int s = 0; s < p.S; ++s) {
        const PairCam& c = cams[s];
        if (c.src < 0) continue;
        const double2* Fd = reinterpret_cast<const double2*>(c.F);      // F[0..8], f[0..2]: 12 contiguous doubles
        const double2 F01 = Fd[0], F23 = Fd[1], F45 = Fd[2], F67 = Fd[3], F8f0 = Fd[4], f12 = Fd[5];
        const float4* Gq = reinterpret_cast<const float4*>(c.Gf);       // Gf[0..8], gf[0..2], Bzf[0..2], tbzf
        const float4 G0 = Gq[0], G1 = Gq[1], G2 = Gq[2], G3 = Gq[3];
        const int src_off = c.src * (H * W);
        double ax = fma(F01.x, xd, fma(F01.y, yd, F23.x));
        double ay = fma(F23.y, xd, fma(F45.x, yd, F45.y));
        double az = fma(F67.x, xd, fma(F67.y, yd, F8f0.x));
        // phase 1, all pixels of the thread: reference pixel -> 3-D -> source pixel (:619-631) in float64, tap loads
        double u[PPT], v[PPT];
        RemapTaps taps[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (k > 0) { ax += F01.y; ay += F45.x; az += F67.y; }
            const double qx = fma(ax, dr[k], F8f0.y), qy = fma(ay, dr[k], f12.x), qz = fma(az, dr[k], f12.y);
            const double iq = fast_rcp64(qz);
            u[k] = qx * iq;
            v[k] = qy * iq;
            taps[k] = remap_taps(p.depths, src_off, H, W, (float)u[k], (float)v[k]);
        }
        // phase 2: source pixel -> 3-D -> reference pixel (:632-647) in float32, thresholds, float64 recheck
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const float xs = (float)u[k], ys = (float)v[k];
            const float ds = remap_blend(taps[k]);
            const float ux = xs * ds, vx = ys * ds;
            const float px = fmaf(G0.x, ux, fmaf(G0.y, vx, fmaf(G0.z, ds, G2.y)));
            const float py = fmaf(G0.w, ux, fmaf(G1.x, vx, fmaf(G1.y, ds, G2.z)));
            const float pz = fmaf(G1.z, ux, fmaf(G1.w, vx, fmaf(G2.x, ds, G2.w)));
            float dep = fmaf(G3.x, ux, fmaf(G3.y, vx, fmaf(G3.z, ds, G3.w)));
            const float ip = fast_rcp(pz);
            const float dx = fmaf(px, ip, -xf), dy = fmaf(py, ip, -yf[k]);
            const float dist2 = fmaf(dx, dx, dy * dy);
            const float rel = fabsf(dep - d_ref[k]) * inv_d[k];
            bool m = (dist2 < thr2) && (rel < rel_thr);
            if ((fabsf(dist2 - thr2) < thr2_band) || (fabsf(rel - rel_thr) < rel_band)) {
                const ExactResult e = exact_back_projection(c, u[k], v[k], ds, x, yb + k, d_ref[k], p.pix_thr2, rel_thr);
                m = e.mask != 0;
                dep = e.depth_rep;
            }
            votes[k] += m ? 1 : 0;
            sum[k] = __fadd_rn(sum[k], m ? dep : 0.0f);
        }
    }
    const float* conf_p = p.confs + (size_t)ref * plane;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int y = yb + k;
        if (y >= H) break;
        const size_t i = (size_t)y * W + x;
        const size_t o = (size_t)r * plane + i;
        const bool ph = conf_p[i] > p.photo_thr;   // :716
        const bool ge = votes[k] >= p.geo_thr;     // :746
        p.photo[o] = ph;
        p.geo[o] = ge;
        p.final_mask[o] = ph && ge;                // :749
        // :744 - float32 sum divided by an int32 count gives float64 in NumPy; rounded to float32 on output
        p.depth_avg[o] = (float)((double)__fadd_rn(sum[k], d_ref[k]) / (double)(votes[k] + 1));
        if (p.geo_sum != nullptr) p.geo_sum[o] = votes[k];
    }
}

// ---- host: float64 camera algebra (the same products np.linalg.inv / np.matmul form in the reference) -------
static void inv3(const double* m, double* o) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    const double id = 1.0 / det;
    o[0] = A * id; o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
    o[3] = B * id; o[4] = (a * i - c * g) * id;  o[5] = -(a * f - c * d) * id;
    o[6] = C * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}

static void inv4(const double* m, double* out) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { a[i][j] = m[i * 4 + j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r)
            if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        if (piv != c)
            for (int j = 0; j < 8; ++j) { double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        const double d = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = a[r][c];
            for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = a[i][4 + j];
}

static void mul4(const double* a, const double* b, double* o) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            o[i * 4 + j] = s;
        }
}

static void mul3(const double* a, const double* b, double* o) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) o[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

static void make_pair_cam(const double* Kr, const double* Er, const double* Ks, const double* Es, int src,
                          PairCam* c) {
    double Eri[16], Esi[16], rel[16], back[16], Kri[9], Ksi[9], Rrel[9], Rback[9], tmp[9], Bk[9];
    inv4(Er, Eri);
    inv4(Es, Esi);
    mul4(Es, Eri, rel);    // :622
    mul4(Er, Esi, back);   // :640
    inv3(Kr, Kri);         // :619
    inv3(Ks, Ksi);         // :637
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { Rrel[i * 3 + j] = rel[i * 4 + j]; Rback[i * 3 + j] = back[i * 4 + j]; }
    mul3(Rrel, Kri, tmp);
    mul3(Ks, tmp, c->F);   // K_src R_rel K_ref^-1
    mul3(Rback, Ksi, Bk);  // R_back K_src^-1
    mul3(Kr, Bk, c->G);    // K_ref R_back K_src^-1
    for (int i = 0; i < 3; ++i) {
        c->f[i] = Ks[i * 3] * rel[3] + Ks[i * 3 + 1] * rel[7] + Ks[i * 3 + 2] * rel[11];
        c->g[i] = Kr[i * 3] * back[3] + Kr[i * 3 + 1] * back[7] + Kr[i * 3 + 2] * back[11];
        c->Bz[i] = Bk[6 + i];
    }
    c->tbz = back[11];
    for (int i = 0; i < 9; ++i) c->Gf[i] = (float)c->G[i];
    for (int i = 0; i < 3; ++i) { c->gf[i] = (float)c->g[i]; c->Bzf[i] = (float)c->Bz[i]; }
    c->tbzf = (float)c->tbz;
    c->src = src;
    c->pad[0] = c->pad[1] = c->pad[2] = 0;
}

// The camera blocks are staged through the stream-ordered allocator.  Its default pool gives memory back to the
// driver at every synchronisation (release threshold 0), which turns each call into a fresh allocation; keep a few
// MB cached instead.  Only users of cudaMallocAsync on this device are affected.
static void keep_async_pool_warm(int dev) {
    static bool done[64] = {false};
    if (dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t threshold = 64ull << 20;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
    done[dev] = true;
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_geo_check_pair(const float* depth_ref, const double* K_ref, const double* E_ref,
                                     const float* depth_src, const double* K_src, const double* E_src,
                                     double condmask_pixel, double condmask_depth, uint8_t* mask,
                                     float* depth_reprojected, float* x2d_src, float* y2d_src, int H, int W,
                                     void* stream) {
    if (!depth_ref || !K_ref || !E_ref || !depth_src || !K_src || !E_src || !mask || !depth_reprojected ||
        !x2d_src || !y2d_src)
        return fail(MVSTER_ERR_BAD_ARG, "geo_check_pair: null pointer");
    if (H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "geo_check_pair: non-positive dimension");
    DeviceGuard guard(mask);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    PairCam cam;
    make_pair_cam(K_ref, E_ref, K_src, E_src, 0, &cam);
    dim3 grid((W + 31) / 32, (H + 7) / 8);
    // NumPy compares the float64 pixel distance with a Python float (float64) and the float32 relative depth
    // difference with the same scalar cast to float32
    geo_check_pair_kernel<<<grid, 256, 0, s>>>(depth_ref, depth_src, cam, condmask_pixel * condmask_pixel,
                                               (float)condmask_depth, mask,
                                               depth_reprojected, x2d_src, y2d_src, H, W);
    count_launch();
    MVSTER_CHECK_LAUNCH("geo_check_pair launch");
    return MVSTER_OK;
}

extern "C" int mvster_geo_filter(const float* depths, const float* confs, const double* K, const double* E,
                                 const int32_t* pairs, int V, int R, int S, double condmask_pixel,
                                 double condmask_depth, double photomask, int geomask, uint8_t* photo, uint8_t* geo,
                                 uint8_t* final_mask, float* depth_avg, int32_t* geo_sum, int H, int W,
                                 void* stream) {
    if (!depths || !confs || !K || !E || !pairs || !photo || !geo || !final_mask || !depth_avg)
        return fail(MVSTER_ERR_BAD_ARG, "geo_filter: null pointer");
    if (V <= 0 || R <= 0 || S < 0 || H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "geo_filter: bad dimension");
    if (R > 65535) return fail(MVSTER_ERR_UNSUPPORTED, "geo_filter: more than 65535 reference views per call");
    if (H > 32768 || W > 32768 || (double)V * H * W >= 2147483648.0)
        return fail(MVSTER_ERR_UNSUPPORTED, "geo_filter: depth maps larger than 32768 px per side or a stack of 2^31 px");
    DeviceGuard guard(depth_avg);
    if (guard.status != MVSTER_OK) return guard.status;
    cudaStream_t s = (cudaStream_t)stream;
    const int Sa = S > 0 ? S : 1;
    std::vector<PairCam> cams((size_t)R * Sa);
    std::vector<int> refs(R);
    for (int r = 0; r < R; ++r) {
        const int ref = pairs[(size_t)r * (1 + S)];
        if (ref < 0 || ref >= V) return fail(MVSTER_ERR_BAD_ARG, "geo_filter: pairs[%d][0]=%d out of range", r, ref);
        refs[r] = ref;
        for (int j = 0; j < Sa; ++j) {
            PairCam& c = cams[(size_t)r * Sa + j];
            const int src = j < S ? pairs[(size_t)r * (1 + S) + 1 + j] : -1;
            if (src >= V) return fail(MVSTER_ERR_BAD_ARG, "geo_filter: pairs[%d][%d]=%d out of range", r, j + 1, src);
            if (src < 0) { c = PairCam{}; c.src = -1; continue; }
            make_pair_cam(K + (size_t)ref * 9, E + (size_t)ref * 16, K + (size_t)src * 9, E + (size_t)src * 16, src, &c);
        }
    }
    const size_t cam_bytes = cams.size() * sizeof(PairCam), ref_bytes = refs.size() * sizeof(int);
    keep_async_pool_warm(guard.dev);
    char* dev = nullptr;
    cudaError_t e = cudaMallocAsync(&dev, cam_bytes + ref_bytes, s);
    if (e != cudaSuccess) return check_cuda(e, "geo_filter: cudaMallocAsync");
    e = cudaMemcpyAsync(dev, cams.data(), cam_bytes, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dev + cam_bytes, refs.data(), ref_bytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { cudaFreeAsync(dev, s); return check_cuda(e, "geo_filter: H2D"); }
    GeoFilterParams p{depths, confs, reinterpret_cast<const PairCam*>(dev), reinterpret_cast<const int*>(dev + cam_bytes),
                      photo, geo, final_mask, depth_avg, geo_sum, Sa, H, W, condmask_pixel * condmask_pixel, (float)condmask_depth,
                      (float)photomask, geomask};
    // reference pixels per thread (a column of PPT rows; CTA tile 32 x 8*PPT); MVSTER_FILTER_PPT=1|2 for A/B timing (default 4)
    static const int ppt = [] { const char* e = getenv("MVSTER_FILTER_PPT"); const int v = e ? atoi(e) : 4; return (v == 1 || v == 2) ? v : 4; }();
    dim3 grid((W + 31) / 32, (H + 8 * ppt - 1) / (8 * ppt), R);
    const size_t cam_smem = (size_t)Sa * sizeof(PairCam);
    if (cam_smem > 48 * 1024) { cudaFreeAsync(dev, s); return fail(MVSTER_ERR_UNSUPPORTED, "geo_filter: more than %d source views per reference view", (int)(48 * 1024 / sizeof(PairCam))); }
    if (ppt == 1) geo_filter_kernel<1><<<grid, 256, cam_smem, s>>>(p);
    else if (ppt == 2) geo_filter_kernel<2><<<grid, 256, cam_smem, s>>>(p);
    else geo_filter_kernel<4><<<grid, 256, cam_smem, s>>>(p);
    count_launch();
    cudaError_t le = cudaGetLastError();
    cudaFreeAsync(dev, s);
    if (le != cudaSuccess) return check_cuda(le, "geo_filter launch");
    return MVSTER_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// depth2pts (reference test_mvs4.py:206-218, pixel grid :220-229): back-projection of a fused depth map to world points
//     uv = K^-1 [x+0.5, y+0.5, 1]^T ;  X_cam = uv * depth ;  X_world = R^-1 (X_cam - t)          (float64 throughout)
// One thread per pixel, 4 bytes in, 24 bytes out, fully coalesced; stays on the GPU between the filter and the
// point-cloud writer instead of the reference's NumPy temporaries of shape [3, H*W] float64.
// ---------------------------------------------------------------------------------------------------------------------
namespace mvster {
struct Depth2PtsCam {
    double kinv[9], rinv[9], t[3];
};
__global__ void __launch_bounds__(256) depth2pts_kernel(const float* __restrict__ depth, double* __restrict__ xyz,
                                                        const __grid_constant__ Depth2PtsCam c, int H, int W) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)H * W) return;
    const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
    const double px = (double)x + 0.5, py = (double)y + 0.5;  // np.linspace(0.5, W - 0.5, W)
    const double d = (double)depth[i];
    const double cx = (c.kinv[0] * px + c.kinv[1] * py + c.kinv[2]) * d - c.t[0];
    const double cy = (c.kinv[3] * px + c.kinv[4] * py + c.kinv[5]) * d - c.t[1];
    const double cz = (c.kinv[6] * px + c.kinv[7] * py + c.kinv[8]) * d - c.t[2];
    xyz[3 * i + 0] = c.rinv[0] * cx + c.rinv[1] * cy + c.rinv[2] * cz;
    xyz[3 * i + 1] = c.rinv[3] * cx + c.rinv[4] * cy + c.rinv[5] * cz;
    xyz[3 * i + 2] = c.rinv[6] * cx + c.rinv[7] * cy + c.rinv[8] * cz;
}
}  // namespace mvster

extern "C" int mvster_depth2pts(const float* depth, const double* K, const double* E, double* xyz, int H, int W,
                                void* stream) {
    if (!depth || !K || !E || !xyz) return fail(MVSTER_ERR_BAD_ARG, "depth2pts: null pointer");
    if (H <= 0 || W <= 0) return fail(MVSTER_ERR_BAD_ARG, "depth2pts: non-positive dimension");
    DeviceGuard guard(xyz);
    if (guard.status != MVSTER_OK) return guard.status;
    Depth2PtsCam c;
    inv3(K, c.kinv);                                   // np.linalg.inv(cam_intrinsic), :209
    const double R[9] = {E[0], E[1], E[2], E[4], E[5], E[6], E[8], E[9], E[10]};
    inv3(R, c.rinv);                                   // np.linalg.inv(R), :214
    c.t[0] = E[3]; c.t[1] = E[7]; c.t[2] = E[11];
    const size_t n = (size_t)H * W;
    depth2pts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(depth, xyz, c, H, W);
    count_launch();
    MVSTER_CHECK_LAUNCH("depth2pts launch");
    return MVSTER_OK;
}
