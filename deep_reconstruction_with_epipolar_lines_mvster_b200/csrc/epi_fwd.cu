// K1 forward: fused homography warp + bilinear gather + group-wise correlation + epipolar attention (softmax over
// the D depth hypotheses) + view-weighted aggregation.  One kernel per cascade stage; the [B,C,D,H,W] warped
// volume, the repeated reference volume and the per-view correlation volumes of the reference
// (models/mvs4net_utils.py:1036,1051,1066-1069,1083-1085,1100) never exist in memory.
//
// Work decomposition (sm_100a).  Features are NHWC, so a bilinear tap is one contiguous channel vector.
// A lane owns CH channels (a multiple of 8) x DL depth hypotheses of one reference pixel; L = (C/CH)*(D/DL) lanes
// cooperate on a pixel.  Splitting the *hypotheses* across lanes (not only the channels) means that every lane
// does its own sample-position arithmetic - no redundant coordinate math for the wide coarse stages (C = 32/64).
//   - group correlation is lane-local (a lane's channels cover whole groups); the per-hypothesis score is summed
//     over the C/CH channel lanes, the softmax max / sum over the D/DL hypothesis lanes, both with xor-shuffles;
//   - bilinear blending and ref*warped products run as packed fp32x2 (FFMA2 / FMUL2), two channels per issue slot;
//   - the softmax over D, the running weight sum and the weighted volume accumulators live in registers.
//
// Two tap sources:
//   DIRECT (any C, fp32 or bf16): 32-byte LDG.E.256 (fp32) / 16-byte LDG.E.128 (bf16) per 8 channels straight from
//     global memory through L1, per-tap bounds weights.  Texels of C >= 32 fp32 channels are whole 128-byte
//     lines, which the L1 data pipe serves at full rate.
//   TMA (C = 8 or 16 fp32, i.e. 32/64-byte texels): direct gathers of sub-line texels waste the L1 data pipe
//     (a wavefront serves one 128-byte line; unaligned texel runs straddle lines - measured 22 wavefronts per
//     LDG.E.256).  A CTA owns a 32x8 / 32x4 pixel tile and, per source view, (1) reduces its sample positions to an
//     integer bounding box (REDUX.MIN/MAX + 4 shared atomics), (2) has one thread issue ONE cp.async.bulk.tensor.4d
//     of box {C, BW, BH, 1} into shared memory (mbarrier completion; TMA zero-fills outside the image, which IS
//     grid_sample's padding_mode='zeros', so the gather needs no bounds logic), (3) gathers with LDS.128 from the
//     hardware-swizzled box (SWIZZLE_32B/64B: 8 neighbouring texels hit 8 different bank groups).  The box of view
//     v+1 is requested before the gather of view v (two buffers).  A tile whose footprint exceeds the box falls
//     back to direct gathers for that view.
#include <stdlib.h>

#include "epi_tma.cuh"
#include "epi_fwd_box.cuh"

namespace mvster {

// ---------------------------------------------------------------------------------------------------------------------
// LINE kernel (fp32, C = 32: a texel is exactly one 128-byte line).  Through L1 a warp-wide gather of whole-line
// texels costs ~2 cycles per extra line touched by one instruction (measured: L1 data pipe 73 % busy at 2.2x the
// byte floor), shared memory costs 1 cycle per conflict-free 128 bytes.  So the coarse stage is staged by TMA as well:
//   - a lane owns ALL channels of DL = D/4 hypotheses of a pixel (4 lanes per pixel): no redundant sample arithmetic,
//     no tap broadcast shuffles;
//   - a lane reads its texel in 16-byte chunks in the order (j + lane) & 7.  The 8 lanes of an LDS.128 phase then
//     hit 8 different 16-byte bank groups WHATEVER texels they address (every texel starts on a 128-byte boundary,
//     the box is not swizzled): the gather is conflict-free for arbitrary sample positions;
//   - the reference channels and the accumulators live in the same rotated order; only the final store un-rotates.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef MVSTER_LINE_BW
#define MVSTER_LINE_BW 40
#endif
#ifndef MVSTER_LINE_BHX
#define MVSTER_LINE_BHX 4
#endif
#ifndef MVSTER_LINE_MINB
#define MVSTER_LINE_MINB 2
#endif
struct LineGeom {
    static constexpr int C = 32, TB = 128, LD = 4, PPW = 8, WARPS = 8, WX = 2, TILE_W = PPW * WX, TILE_H = WARPS / WX;
    static constexpr int BW = MVSTER_LINE_BW, BH = TILE_H + MVSTER_LINE_BHX;
    static constexpr int ROW_BYTES = BW * TB, BUF_BYTES = ROW_BYTES * BH;
    static constexpr int CTL_BYTES = 16 + 48;
    static constexpr int SMEM = 2 * BUF_BYTES + 1024 + CTL_BYTES + MVSTER_MAX_SRC_VIEWS * 48;
};

__device__ __forceinline__ void ldg128_pairs(const void* ptr, f32x2& a, f32x2& b) {
    asm volatile("ld.global.nc.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(ptr));
}

template <int CPG, int D>
__global__ void __launch_bounds__(LineGeom::WARPS * 32, MVSTER_LINE_MINB) epi_fwd_line_kernel(const __grid_constant__ EpiFwdParams p) {
    using Gm = LineGeom;
    constexpr int C = Gm::C, TB = Gm::TB, LD = Gm::LD, DL = D / LD, G = C / CPG;
    constexpr int GPC = 4 / CPG;  // correlation groups per 16-byte chunk
    constexpr int NT = Gm::WARPS * 32;
    static_assert(D % LD == 0 && (CPG == 1 || CPG == 2 || CPG == 4), "line kernel: D in {4,8}, C/G in {1,2,4}");

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ctl = smem_base + 2u * Gm::BUF_BYTES;
    unsigned char* ctl_ptr = smem_raw + (ctl - smem_u32(smem_raw));
    int* bbox = reinterpret_cast<int*>(ctl_ptr + 16);
    float* rt_s = reinterpret_cast<float*>(ctl_ptr + Gm::CTL_BYTES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pix = lane / LD, dl = lane % LD;
    const uint32_t rot = (uint32_t)lane & 7u;
    const int b = blockIdx.z;
    for (int i = tid; i < p.Nsrc * 12; i += NT) rt_s[i] = __ldg(p.rt + (size_t)b * p.Nsrc * 12 + i);
    int x = blockIdx.x * Gm::TILE_W + (warp % Gm::WX) * Gm::PPW + pix;
    int y = blockIdx.y * Gm::TILE_H + (warp / Gm::WX);
    const bool live = (x < p.W) && (y < p.H);
    x = min(x, p.W - 1);
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

    // reference channels in this lane's rotated chunk order, pre-scaled by 1/(C/G)
    f32x2 rf[8][2];
    {
        const char* refp = reinterpret_cast<const char*>(p.ref) + ((size_t)b * plane + pix_off) * TB;
        const f32x2 sc = pack2(1.0f / CPG, 1.0f / CPG);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            f32x2 a, c;
            ldg128_pairs(refp + ((((uint32_t)j + rot) & 7u) << 4), a, c);
            rf[j][0] = mul2(a, sc);
            rf[j][1] = mul2(c, sc);
        }
    }
    float hyp[DL];
#pragma unroll
    for (int d = 0; d < DL; ++d) hyp[d] = ldg_stream(p.hypo + ((size_t)b * D + dl * DL + d) * plane + pix_off);

    float acc[8 * GPC][DL], wsum[DL];
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        wsum[d] = 1e-8f;  // reference :1037
#pragma unroll
        for (int g = 0; g < 8 * GPC; ++g) acc[g][d] = 0.0f;
    }
    const float fxp = (float)x, fyp = (float)y;
    const float wlim = (float)p.Ws, hlim = (float)p.Hs;

    float nsx[DL], nsy[DL];
    int nbx = 0, nby = 0;
    bool nfit = false;
    uint32_t uses0 = 0, uses1 = 0;

    auto stage_view = [&](int v) {
        const Homography h = homography_from_smem(rt_s + v * 12);
        const float ax = fmaf(h.r00, fxp, fmaf(h.r01, fyp, h.r02));
        const float ay = fmaf(h.r10, fxp, fmaf(h.r11, fyp, h.r12));
        const float az = fmaf(h.r20, fxp, fmaf(h.r21, fyp, h.r22));
#pragma unroll
        for (int d = 0; d < DL; ++d) sample_pos(ax, ay, az, h, hyp[d], wlim, hlim, nsx[d], nsy[d]);
        float lox = nsx[0], hix = nsx[0], loy = nsy[0], hiy = nsy[0];
#pragma unroll
        for (int d = 1; d < DL; ++d) {
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
        __syncthreads();
        const int4 bb = *reinterpret_cast<const int4*>(&bbox[slot * 4]);
        nbx = bb.x; nby = bb.y;
        nfit = (bb.z - bb.x + 2 <= Gm::BW) && (bb.w - bb.y + 2 <= Gm::BH);
        if (nfit && tid == 0) {
            const uint32_t bar = ctl + 8u * (v & 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (uint32_t)Gm::BUF_BYTES);
            tma_load_4d(smem_base + (uint32_t)Gm::BUF_BYTES * (v & 1), &p.tmap[v], bar, 0, nbx, nby, b);
        }
    };

    // blend one 16-byte chunk of the four taps, multiply by the reference chunk, reduce to its groups
    auto chunk_cor = [&](const f32x2 (&t)[4][2], const float (&w)[4], int j, float* cg) {
        f32x2 wv0 = mul2(pack2(w[0], w[0]), t[0][0]), wv1 = mul2(pack2(w[0], w[0]), t[0][1]);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            const f32x2 pw = pack2(w[k], w[k]);
            wv0 = fma2(pw, t[k][0], wv0);
            wv1 = fma2(pw, t[k][1], wv1);
        }
        const f32x2 p0 = mul2(rf[j][0], wv0), p1 = mul2(rf[j][1], wv1);
        float a0, a1, b0, b1;
        if constexpr (CPG == 4) {
            unpack2(add2(p0, p1), a0, a1);
            cg[0] = a0 + a1;
        } else if constexpr (CPG == 2) {
            unpack2(p0, a0, a1); unpack2(p1, b0, b1);
            cg[0] = a0 + a1; cg[1] = b0 + b1;
        } else {
            unpack2(p0, cg[0], cg[1]); unpack2(p1, cg[2], cg[3]);
        }
    };

    stage_view(0);

#pragma unroll 1
    for (int v = 0; v < p.Nsrc; ++v) {
        float sx[DL], sy[DL];
#pragma unroll
        for (int d = 0; d < DL; ++d) { sx[d] = nsx[d]; sy[d] = nsy[d]; }
        const int bx = nbx, by = nby;
        const bool fit = nfit;
        if (v + 1 < p.Nsrc) stage_view(v + 1);

        float cor[8 * GPC][DL];
        if (fit) {
            const uint32_t parity = ((v & 1) ? uses1 : uses0) & 1u;
            mbar_wait(ctl + 8u * (v & 1), parity);
            if (v & 1) ++uses1; else ++uses0;
            const uint32_t buf = smem_base + (uint32_t)Gm::BUF_BYTES * (v & 1);
#pragma unroll
            for (int d = 0; d < DL; ++d) {
                float x0f, y0f;
                int x0i, y0i;
                floor_fi(sx[d], x0f, x0i);
                floor_fi(sy[d], y0f, y0i);
                const float fx = sx[d] - x0f, fy = sy[d] - y0f;
                const int rx = x0i - bx, ry = y0i - by;
                const uint32_t base = buf + (uint32_t)ry * Gm::ROW_BYTES + (uint32_t)rx * TB;
                const float gx = 1.0f - fx, gy = 1.0f - fy;
                const float w[4] = {gx * gy, fx * gy, gx * fy, fx * fy};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t a = base + ((((uint32_t)j + rot) & 7u) << 4);
                    f32x2 t[4][2];
                    lds_pairs(a, t[0][0], t[0][1]);
                    lds_pairs(a + TB, t[1][0], t[1][1]);
                    lds_pairs(a + Gm::ROW_BYTES, t[2][0], t[2][1]);
                    lds_pairs(a + Gm::ROW_BYTES + TB, t[3][0], t[3][1]);
                    float cg[GPC];
                    chunk_cor(t, w, j, cg);
#pragma unroll
                    for (int g = 0; g < GPC; ++g) cor[j * GPC + g][d] = cg[g];
                }
            }
        } else {
            // footprint larger than the box: direct gather (same rotated chunk order), per-tap bounds weights
            const char* srcp = reinterpret_cast<const char*>(p.src[v]) + (size_t)b * p.Hs * p.Ws * TB;
#pragma unroll
            for (int d = 0; d < DL; ++d) {
                const float x0f = floorf(sx[d]), y0f = floorf(sy[d]);
                const float fx = sx[d] - x0f, fy = sy[d] - y0f;
                const int x0 = (int)x0f, y0 = (int)y0f;
                const bool vx0 = (unsigned)x0 < (unsigned)p.Ws, vx1 = (unsigned)(x0 + 1) < (unsigned)p.Ws;
                const bool vy0 = (unsigned)y0 < (unsigned)p.Hs, vy1 = (unsigned)(y0 + 1) < (unsigned)p.Hs;
                const int xc0 = min(max(x0, 0), p.Ws - 1), xc1 = min(x0 + 1, p.Ws - 1);
                const int yc0 = min(max(y0, 0), p.Hs - 1), yc1 = min(y0 + 1, p.Hs - 1);
                const float gx = vx0 ? 1.0f - fx : 0.0f, hx = vx1 ? fx : 0.0f;
                const float gy = vy0 ? 1.0f - fy : 0.0f, hy = vy1 ? fy : 0.0f;
                const float w[4] = {gx * gy, hx * gy, gx * hy, hx * hy};
                const unsigned r0 = (unsigned)(yc0 * p.Ws), r1 = (unsigned)(yc1 * p.Ws);
                const char* a00 = srcp + (size_t)(r0 + (unsigned)xc0) * TB;
                const char* a01 = srcp + (size_t)(r0 + (unsigned)xc1) * TB;
                const char* a10 = srcp + (size_t)(r1 + (unsigned)xc0) * TB;
                const char* a11 = srcp + (size_t)(r1 + (unsigned)xc1) * TB;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t co = (((uint32_t)j + rot) & 7u) << 4;
                    f32x2 t[4][2];
                    ldg128_pairs(a00 + co, t[0][0], t[0][1]);
                    ldg128_pairs(a01 + co, t[1][0], t[1][1]);
                    ldg128_pairs(a10 + co, t[2][0], t[2][1]);
                    ldg128_pairs(a11 + co, t[3][0], t[3][1]);
                    float cg[GPC];
                    chunk_cor(t, w, j, cg);
#pragma unroll
                    for (int g = 0; g < GPC; ++g) cor[j * GPC + g][d] = cg[g];
                }
            }
        }

        // score[d] = sum over all G groups (all lane-local); softmax over D crosses the LD hypothesis lanes
        float score[DL];
#pragma unroll
        for (int d = 0; d < DL; ++d) {
            float s0 = cor[0][d], s1 = cor[1][d];
#pragma unroll
            for (int g = 2; g < 8 * GPC; g += 2) { s0 += cor[g][d]; s1 += cor[g + 1][d]; }
            score[d] = s0 + s1;
        }
        float mx = score[0];
#pragma unroll
        for (int d = 1; d < DL; ++d) mx = fmaxf(mx, score[d]);
#pragma unroll
        for (int m = 1; m < LD; m <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        float e[DL], es = 0.0f;
#pragma unroll
        for (int d = 0; d < DL; ++d) {
            e[d] = ex2_approx((score[d] - mx) * p.score_scale);
            es += e[d];
        }
#pragma unroll
        for (int m = 1; m < LD; m <<= 1) es += __shfl_xor_sync(0xffffffffu, es, m);
        const float norm = __fdividef(p.inv_sqrt_c, es);
#pragma unroll
        for (int d = 0; d < DL; ++d) {
            const float w = e[d] * norm;
            wsum[d] += w;
#pragma unroll
            for (int g = 0; g < 8 * GPC; ++g) acc[g][d] = fmaf(w, cor[g][d], acc[g][d]);
            if (p.weights != nullptr && live)
                p.weights[(((size_t)b * p.Nsrc + v) * D + dl * DL + d) * plane + pix_off] = w;
        }
    }

    if (!live) return;
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        const float inv = __frcp_rn(wsum[d]);
        const int dd = dl * DL + d;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g0 = (int)(((uint32_t)j + rot) & 7u) * GPC;  // un-rotate: chunk -> first group of the chunk
#pragma unroll
            for (int g = 0; g < GPC; ++g)
                stg_stream(p.out + (((size_t)b * G + g0 + g) * D + dd) * plane + pix_off, acc[j * GPC + g][d] * inv);
        }
        if (p.wsum != nullptr) p.wsum[((size_t)b * D + dd) * plane + pix_off] = wsum[d];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static bool make_line_maps(EpiFwdParams& p, int Nsrc, int B, int Hs, int Ws) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    constexpr int C = LineGeom::C;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)Ws * C * 4, (cuuint64_t)Hs * Ws * C * 4};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)LineGeom::BW, (cuuint32_t)LineGeom::BH, 1};
    for (int v = 0; v < Nsrc; ++v) {
        CUresult r = enc(&p.tmap[v], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(p.src[v]), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}

template <int CPG, int D>
static int launch_line(const EpiFwdParams& p, cudaStream_t stream) {
    using Gm = LineGeom;
    static int attr_done[64] = {};  // largest size set per device
    const int st = ensure_dynamic_smem_bytes(epi_fwd_line_kernel<CPG, D>, Gm::SMEM, attr_done, "epi_fwd(line): cudaFuncSetAttribute");
    if (st != MVSTER_OK) return st;
    dim3 grid((p.W + Gm::TILE_W - 1) / Gm::TILE_W, (p.H + Gm::TILE_H - 1) / Gm::TILE_H, p.B);
    if (grid.y > 65535u || grid.z > 65535u) return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: grid too large");
    epi_fwd_line_kernel<CPG, D><<<grid, Gm::WARPS * 32, Gm::SMEM, stream>>>(p);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_fwd(line) launch");
    return MVSTER_OK;
}

// environment switches, read once per process (the fallback tests)
struct FwdEnv {
    bool no_tma, no_line;
    FwdEnv() : no_tma(getenv("MVSTER_NO_TMA") != nullptr), no_line(getenv("MVSTER_NO_LINE") != nullptr) {}
};
static const FwdEnv& fwd_env() {
    static const FwdEnv e;
    return e;
}

template <int C, int CPG, int D>
static int dispatch_variant(EpiFwdParams& p, int dtype, bool allow_tma, cudaStream_t s) {
    const FwdEnv& env = fwd_env();
    if constexpr (C <= 16) {
        // fine stages (16/32/64-byte texels): TMA-staged box kernel when the tensor maps can be built
        if (allow_tma) {
            bool built = false;
            const int st = dtype == MVSTER_BF16 ? launch_box<C, CPG, D, __nv_bfloat16>(p, s, &built)
                                                : launch_box<C, CPG, D, float>(p, s, &built);
            if (built || st != MVSTER_OK) return st;
        }
    }
    if (dtype == MVSTER_BF16) return launch_direct<C, CPG, D, __nv_bfloat16>(p, s);
    if constexpr (C == 32 && CPG <= 4) {
        // whole-line texels: conflict-free rotated shared-memory gather (see epi_fwd_line_kernel)
        if (allow_tma && !env.no_line && make_line_maps(p, p.Nsrc, p.B, p.Hs, p.Ws))
            return launch_line<CPG, D>(p, s);
    }
    return launch_direct<C, CPG, D, float>(p, s);
}

template <int C, int CPG>
static int dispatch_d(EpiFwdParams& p, int D, int dtype, bool allow_tma, cudaStream_t s) {
    switch (D) {
        case 4: return dispatch_variant<C, CPG, 4>(p, dtype, allow_tma, s);
        case 8: return dispatch_variant<C, CPG, 8>(p, dtype, allow_tma, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: D=%d not in {4,8}", D);
    }
}

template <int C>
static int dispatch_cpg(EpiFwdParams& p, int cpg, int D, int dtype, bool allow_tma, cudaStream_t s) {
#ifdef MVSTER_FAST_BUILD  // development builds (scripts/build_variant.py): only the shipped (C, C/G) pairs
    constexpr int kCpg = C == 64 ? 8 : (C == 32 ? 4 : (C == 16 ? 4 : 2));
    if (cpg != kCpg) return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: MVSTER_FAST_BUILD has C/G=%d only for C=%d", kCpg, C);
    return dispatch_d<C, kCpg>(p, D, dtype, allow_tma, s);
#else
    switch (cpg) {
        case 1: return dispatch_d<C, 1>(p, D, dtype, allow_tma, s);
        case 2: return dispatch_d<C, 2>(p, D, dtype, allow_tma, s);
        case 4: return dispatch_d<C, 4>(p, D, dtype, allow_tma, s);
        case 8: return dispatch_d<C, 8>(p, D, dtype, allow_tma, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: C/G=%d not in {1,2,4,8}", cpg);
    }
#endif
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
    if (dtype != MVSTER_F32 && dtype != MVSTER_BF16) return fail(MVSTER_ERR_BAD_ARG, "epi_fwd: unknown dtype %d", dtype);
    if ((double)B * Hs * Ws * C >= 2147483648.0 || (double)H * W >= 2147483648.0)
        return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: tensor too large for 32-bit texel offsets");
    const uintptr_t align = (dtype == MVSTER_BF16) ? 16 : 32;
    if (((uintptr_t)ref) % align) return fail(MVSTER_ERR_ALIGN, "epi_fwd: ref not %d-byte aligned", (int)align);
    static thread_local EpiFwdParams p;  // holds the 64-byte aligned tensor maps; one per calling thread
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
    const bool allow_tma = !fwd_env().no_tma;
    const int cpg = C / G;
    cudaStream_t s = (cudaStream_t)stream;
    switch (C) {
        case 8: return dispatch_cpg<8>(p, cpg, D, dtype, allow_tma, s);
        case 16: return dispatch_cpg<16>(p, cpg, D, dtype, allow_tma, s);
        case 32: return dispatch_cpg<32>(p, cpg, D, dtype, allow_tma, s);
        case 64: return dispatch_cpg<64>(p, cpg, D, dtype, allow_tma, s);
        default: return fail(MVSTER_ERR_UNSUPPORTED, "epi_fwd: C=%d not in {8,16,32,64}", C);
    }
}
