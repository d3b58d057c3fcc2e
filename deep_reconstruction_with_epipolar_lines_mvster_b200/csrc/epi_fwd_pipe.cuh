// K1 forward, fine stages: persistent, software-pipelined variant of the box kernel (epi_fwd_box.cuh).
//
// Why: the box kernel's CTAs live for ~7 us, and a third of that is a latency chain nothing overlaps - hypothesis
// load -> bounding boxes -> barrier -> TMA round trip -> first gather (ncu: 47 % of the warp samples sit in the
// prologue / phase A / barrier regions that hold 25 % of the instructions; more resident CTAs do not help because the
// registers are gone at 16 warps per SM).  Here
//   * a tiny pre-pass (epi_tile_boxes_kernel) writes one conservative bounding box per (tile, view): the sample
//     position is projective in the pixel and monotone in the depth, so the 4 tile corners x {min, max depth of the
//     tile} bound every sample of the tile (+ a margin for fp32 rounding).  ~40 instructions per pixel-warp instead of
//     the ~240 of the exact per-sample reduction, and no barrier in the main kernel;
//   * CTAs are persistent (grid = resident CTAs): tile k+1's hypotheses / reference texel are prefetched into
//     registers while tile k is processed, and its source boxes are requested as soon as the ring slots free up -
//     a ring of NSLOT staging buffers with a "full" (TMA complete_tx) and an "empty" (one arrive per consumer warp)
//     mbarrier each.  Lane 0 of warp 0 is the producer: it tests the empty barrier without blocking after every view
//     and only blocks when its own warp is about to wait for a box that has not been requested yet.
// In the steady state there is no CTA-wide barrier and no exposed global-memory or TMA latency.
#pragma once

#include "epi_fwd_box.cuh"

#ifndef MVSTER_PIPE_BW
#define MVSTER_PIPE_BW 52      // box width for 16/32-byte texels (hull boxes are ~5 texels wider than exact ones)
#endif
#ifndef MVSTER_PIPE_BHX
#define MVSTER_PIPE_BHX 4
#endif
#ifndef MVSTER_PIPE_NSLOT
#define MVSTER_PIPE_NSLOT 4
#endif
#ifndef MVSTER_PIPE_BW64
#define MVSTER_PIPE_BW64 52    // 64-byte texels (C = 16 fp32)
#endif
#ifndef MVSTER_PIPE_BHX64
#define MVSTER_PIPE_BHX64 4
#endif
#ifndef MVSTER_PIPE_NSLOT64
#define MVSTER_PIPE_NSLOT64 3
#endif
#ifndef MVSTER_PIPE_CTAS
#define MVSTER_PIPE_CTAS 4     // resident CTAs per SM the grid is sized for (16/32-byte texels)
#endif
#ifndef MVSTER_PIPE_CTAS64
#define MVSTER_PIPE_CTAS64 2
#endif

namespace mvster {

template <int C, int D, int ES>
struct PipeCfg : BoxCfg<C, D, ES, (C * ES == 64 ? MVSTER_PIPE_BW64 : MVSTER_PIPE_BW),
                        (C * ES == 64 ? MVSTER_PIPE_BHX64 : MVSTER_PIPE_BHX),
                        (C * ES == 64 ? MVSTER_PIPE_NSLOT64 : MVSTER_PIPE_NSLOT)> {
    using Base = BoxCfg<C, D, ES, (C * ES == 64 ? MVSTER_PIPE_BW64 : MVSTER_PIPE_BW),
                        (C * ES == 64 ? MVSTER_PIPE_BHX64 : MVSTER_PIPE_BHX),
                        (C * ES == 64 ? MVSTER_PIPE_NSLOT64 : MVSTER_PIPE_NSLOT)>;
    static constexpr int NSLOT = Base::NBUF;
    static constexpr int CTAS = (C * ES == 64) ? MVSTER_PIPE_CTAS64 : MVSTER_PIPE_CTAS;
    // shared memory: ring | full[NSLOT] empty[NSLOT] (128 B) | slotinfo[NSLOT] int4 | boxq[2][MAXV] int4 | rt_s[B*Nsrc*12]
    static constexpr int FULL_OFF = NSLOT * Base::BUF_BYTES;
    static constexpr int EMPTY_OFF = FULL_OFF + 64;
    static constexpr int INFO_OFF = FULL_OFF + 128;
    static constexpr int BOXQ_OFF = INFO_OFF + NSLOT * 16;
    static constexpr int RT_OFF = BOXQ_OFF + 2 * MVSTER_MAX_SRC_VIEWS * 16;
    static constexpr int kMaxBatchViews = 128;  // rt of every (batch, view) lives in shared memory
    static int smem_bytes(int batch_views) { return RT_OFF + batch_views * 48 + 1024; }
    static_assert(NSLOT <= 8, "ring slots");
};

// ---------------------------------------------------------------------------------------------------------------------
// pre-pass: conservative bounding box of the sample positions of every (tile, view)
// ---------------------------------------------------------------------------------------------------------------------
// boxes[(tile * Nsrc + v)] = {min x, min y, max x, max y} as magic-offset float bits (see box_fits), or zeros when the
// geometry is degenerate for that tile (the main kernel then takes the exact direct path).
template <typename K>
__global__ void __launch_bounds__(K::WARPS * 32) epi_tile_boxes_kernel(const float* __restrict__ rt, const float* __restrict__ hypo,
                                                                       int4* __restrict__ boxes, int Nsrc, int H, int W,
                                                                       int tilesX, int tilesY) {
    constexpr int DL = K::DL, LD = K::LD, D = DL * LD;
    __shared__ unsigned s_min[K::WARPS], s_max[K::WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pix = lane / LD, dl = lane % LD;
    const int tx = blockIdx.x, ty = blockIdx.y, b = blockIdx.z;
    const int x = min(tx * K::TILE_W + (warp % K::WX) * K::PPW + pix, W - 1);
    const int y = min(ty * K::TILE_H + (warp / K::WX), H - 1);
    const size_t plane = (size_t)H * W;
    const float* hp = hypo + ((size_t)b * D + dl * DL) * plane + (size_t)y * W + x;
    // positive floats order like unsigned integers; negative / NaN values land above +inf and are rejected below
    unsigned lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int d = 0; d < DL; ++d) {
        const unsigned u = __float_as_uint(ldg_stream(hp + d * plane));
        lo = min(lo, u); hi = max(hi, u);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { s_min[warp] = lo; s_max[warp] = hi; }
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int w = 0; w < K::WARPS; ++w) { lo = min(lo, s_min[w]); hi = max(hi, s_max[w]); }
    const bool depth_ok = (hi < 0x7f800000u) && (lo > 0u);
    const float dsel = __uint_as_float((lane & 1) ? hi : lo);
    const int corner = (lane >> 1) & 3;
    const float cx = (float)min(tx * K::TILE_W + ((corner & 1) ? K::TILE_W - 1 : 0), W - 1);
    const float cy = (float)min(ty * K::TILE_H + ((corner & 2) ? K::TILE_H - 1 : 0), H - 1);
    const size_t tile = ((size_t)b * tilesY + ty) * tilesX + tx;
    for (int v0 = 0; v0 < Nsrc; v0 += 4) {
        const int v = v0 + (lane >> 3);
        const bool act = v < Nsrc;
        const Homography h = load_homography(rt + ((size_t)b * Nsrc + (act ? v : 0)) * 12);
        const float ax = fmaf(h.r00, cx, fmaf(h.r01, cy, h.r02));
        const float ay = fmaf(h.r10, cx, fmaf(h.r11, cy, h.r12));
        const float az = fmaf(h.r20, cx, fmaf(h.r21, cy, h.r22));
        const float px = fmaf(ax, dsel, h.t0), py = fmaf(ay, dsel, h.t1), pz = fmaf(az, dsel, h.t2);
        const float sx = px / pz, sy = py / pz;
        float lox = sx, hix = sx, loy = sy, hiy = sy;
        // the depth must not cross the camera plane anywhere in the tile: all eight pz of one sign, none tiny
        bool pos = pz > 1e-6f, neg = pz < -1e-6f;
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
            lox = min_nan(lox, __shfl_xor_sync(0xffffffffu, lox, m)); hix = max_nan(hix, __shfl_xor_sync(0xffffffffu, hix, m));
            loy = min_nan(loy, __shfl_xor_sync(0xffffffffu, loy, m)); hiy = max_nan(hiy, __shfl_xor_sync(0xffffffffu, hiy, m));
            pos = pos && (__shfl_xor_sync(0xffffffffu, (int)pos, m) != 0);
            neg = neg && (__shfl_xor_sync(0xffffffffu, (int)neg, m) != 0);
        }
        if (act && (lane & 7) == 0) {
            // margin: the kernel evaluates the same rational function per pixel in fp32 (a few ulp of 2^11 texels)
            constexpr float kMargin = 1.0f / 128.0f;
            int4 bb = make_int4(0, 0, 0, 0);
            if (depth_ok && (pos || neg)) {
                bb.x = __float_as_int(__fadd_rd(lox - kMargin, kFloorMagic));
                bb.y = __float_as_int(__fadd_rd(loy - kMargin, kFloorMagic));
                bb.z = __float_as_int(__fadd_rd(hix + kMargin, kFloorMagic));
                bb.w = __float_as_int(__fadd_rd(hiy + kMargin, kFloorMagic));
            }
            boxes[tile * Nsrc + v] = bb;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------------------
struct PipeArgs {
    const int4* boxes;
    int tilesX, tilesY, numTiles;
};

template <int C, int CPG, int D, typename T>
__global__ void __launch_bounds__(PipeCfg<C, D, (int)sizeof(T)>::WARPS * 32, PipeCfg<C, D, (int)sizeof(T)>::CTAS)
    epi_fwd_pipe_kernel(const __grid_constant__ EpiFwdParams p, const PipeArgs a) {
    using K = PipeCfg<C, D, (int)sizeof(T)>;
    constexpr int DL = K::DL, LD = K::LD, TB = K::TB, NSLOT = K::NSLOT;
    constexpr int G = C / CPG, GPL = G;
    constexpr int NT = K::WARPS * 32;
    static_assert(8 % CPG == 0, "a chunk of 8 channels must hold whole groups");

    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t full0 = smem_base + K::FULL_OFF, empty0 = smem_base + K::EMPTY_OFF;
    int4* slotinfo = reinterpret_cast<int4*>(sm + K::INFO_OFF);   // {box x, box y, staged?, -}
    int4* boxq = reinterpret_cast<int4*>(sm + K::BOXQ_OFF);       // [2][MAXV]: bounding boxes of the tile being requested
    float* rt_s = reinterpret_cast<float*>(sm + K::RT_OFF);       // [B][Nsrc][12], repacked

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pix = lane / LD, dl = lane % LD;
    const int Nsrc = p.Nsrc;
    const int stride = gridDim.x;
    const size_t plane = (size_t)p.H * p.W;
    const int tilesPerImage = a.tilesX * a.tilesY;

    for (int i = tid; i < p.B * Nsrc * 12; i += NT) rt_s[i] = repack_rt(p.rt + (i / 12) * 12, i % 12);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(full0 + 8u * i, 1);
            mbar_init(empty0 + 8u * i, K::WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // tile -> (batch, pixel) of this lane; dead lanes shadow a valid pixel
    auto locate = [&](int t, int& b, int& x, int& y, bool& live) {
        b = t / tilesPerImage;
        const int r = t - b * tilesPerImage;
        const int ty = r / a.tilesX, tx = r - ty * a.tilesX;
        x = tx * K::TILE_W + (warp % K::WX) * K::PPW + pix;
        y = ty * K::TILE_H + (warp / K::WX);
        live = (x < p.W) && (y < p.H);
        x = min(x, p.W - 1);
        y = min(y, p.H - 1);
    };

    // ---- producer state (meaningful in warp 0; uniform) -------------------------------------------------------------
    int r_k = 0, r_v = 0, r_slot = 0;   // next request: tile ordinal, view, ring slot
    uint32_t r_par = 1;                 // parity that means "slot is free" for the next request
    // issue requests in order up to and including (tile ordinal upto_k, view upto_v); block on a busy slot only if `block`
    auto pump = [&](int upto_k, int upto_v, bool block) {
        if (lane != 0) return;
        while (r_k < upto_k || (r_k == upto_k && r_v <= upto_v)) {
            const int t = blockIdx.x + r_k * stride;
            if (t >= a.numTiles) break;
            const uint32_t ebar = empty0 + 8u * r_slot;
            if (block) mbar_wait(ebar, r_par);
            else if (!mbar_test(ebar, r_par)) break;
            const int4 bb = boxq[(r_k & 1) * MVSTER_MAX_SRC_VIEWS + r_v];
            int bx, by;
            const bool fit = !(MVSTER_BOX_KO & 2) && box_fits<K>(bb, bx, by);
            const int b = t / tilesPerImage;
            slotinfo[r_slot] = make_int4(bx, by, fit ? 1 : 0, b);
            const uint32_t fbar = full0 + 8u * r_slot;
            if (fit) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the recycled slot
                mbar_expect_tx(fbar, (uint32_t)K::BUF_BYTES);
                tma_load_4d(smem_base + (uint32_t)K::BUF_BYTES * r_slot, &p.tmap[r_v], fbar, 0, bx, by, b);
            } else {
                mbar_arrive(fbar);
            }
            if (++r_slot == NSLOT) { r_slot = 0; r_par ^= 1u; }
            if (++r_v == Nsrc) { r_v = 0; ++r_k; }
        }
    };
    // boxes of tile ordinal k -> boxq[k & 1] (warp 0, asynchronous)
    auto fetch_boxes = [&](int k) {
        const int t = blockIdx.x + k * stride;
        if (warp == 0 && t < a.numTiles && lane < Nsrc) {
            const uint32_t dst = smem_u32(boxq + (k & 1) * MVSTER_MAX_SRC_VIEWS + lane);
            const int4* src = a.boxes + (size_t)t * Nsrc + lane;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
    };
    auto boxes_landed = [&]() {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    };

    // ---- prologue: first tile's inputs ------------------------------------------------------------------------------------
    int t = blockIdx.x;
    if (t >= a.numTiles) return;
    fetch_boxes(0);
    int b, x, y;
    bool live;
    locate(t, b, x, y, live);
    float nh[DL];
    P8 nr[C / 8];
    auto prefetch = [&](int bb, int xx, int yy) {
        const size_t po = (size_t)yy * p.W + xx;
#pragma unroll
        for (int d = 0; d < DL; ++d) nh[d] = ldg_stream(p.hypo + ((size_t)bb * D + dl * DL + d) * plane + po);
        const T* refp = reinterpret_cast<const T*>(p.ref) + ((size_t)bb * plane + po) * C;
#pragma unroll
        for (int k = 0; k < C / 8; ++k) nr[k] = load_pairs<T>(refp + k * 8);
    };
    prefetch(b, x, y);
    if (warp == 0) {
        boxes_landed();
        pump(0, Nsrc - 1, false);  // the first NSLOT requests go out; the rest follow as slots free up
    }

    int c_slot = 0;
    uint32_t c_par = 0;
#pragma unroll 1
    for (int k = 0; t < a.numTiles; ++k, t += stride) {
        // ---- this tile's inputs out of the prefetch registers; next tile's loads go out now ------------------------------
        f32x2 hh[DL / 2], rf[C / 2];
#pragma unroll
        for (int j = 0; j < DL / 2; ++j) hh[j] = pack2(nh[2 * j], nh[2 * j + 1]);
        {
            const f32x2 sc = pack2(1.0f / CPG, 1.0f / CPG);
#pragma unroll
            for (int j = 0; j < C / 8; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) rf[j * 4 + q] = mul2(nr[j].q[q], sc);
        }
        const int bcur = b, xcur = x, ycur = y;
        const bool livecur = live;
        const size_t pix_off = (size_t)ycur * p.W + xcur;
        const float fxp = (float)xcur, fyp = (float)ycur;
        const int tn = t + stride;
        if (tn < a.numTiles) {
            locate(tn, b, x, y, live);
            prefetch(b, x, y);
        }
        fetch_boxes(k + 1);
        bool boxes_pending = true;

        float acc[GPL][DL], wsum[DL];
#pragma unroll
        for (int d = 0; d < DL; ++d) {
            wsum[d] = 1e-8f;  // reference :1037
#pragma unroll
            for (int g = 0; g < GPL; ++g) acc[g][d] = 0.0f;
        }
        const float* rt_b = rt_s + (size_t)bcur * Nsrc * 12;

#pragma unroll 1
        for (int v = 0; v < Nsrc; ++v) {
            if (warp == 0) {
                // the request for (k, v) must be out before this warp waits for it
                pump(k, v, true);
            }
            mbar_wait(full0 + 8u * c_slot, c_par);
            const int4 info = slotinfo[c_slot];
            float cor[GPL][DL];
            if (info.z != 0) {
                const PixelView pv = pixel_view(rt_b + v * 12, fxp, fyp);
                gather_view<K, CPG, T>(pv, hh, rf, smem_base + (uint32_t)K::BUF_BYTES * c_slot, info.x, info.y, cor);
            } else {
                direct_view<K, CPG, T>(p.rt + ((size_t)bcur * Nsrc + v) * 12,
                                       reinterpret_cast<const char*>(p.src[v]) + (size_t)bcur * p.Hs * p.Ws * TB, p.Hs, p.Ws,
                                       fxp, fyp, hh, rf, cor);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8u * c_slot);  // this warp is done with the slot
            if (++c_slot == NSLOT) { c_slot = 0; c_par ^= 1u; }
            if (warp == 0) {
                if (boxes_pending) { boxes_landed(); boxes_pending = false; }
                pump(k + 1, Nsrc - 1, false);
            }
            float w[DL];
            attend_accumulate<K, GPL>(cor, p.score_scale, p.inv_sqrt_c, acc, wsum, w);
            if (p.weights != nullptr && livecur) {
#pragma unroll
                for (int d = 0; d < DL; ++d)
                    p.weights[(((size_t)bcur * Nsrc + v) * D + dl * DL + d) * plane + pix_off] = w[d];
            }
        }

        if (livecur) {
#pragma unroll
            for (int d = 0; d < DL; ++d) {
                const float inv = __frcp_rn(wsum[d]);
                const int dd = dl * DL + d;
#pragma unroll
                for (int g = 0; g < GPL; ++g)
                    if (!(MVSTER_BOX_KO & 8) || acc[g][d] * inv == 1234.5678f)
                        stg_stream(p.out + (((size_t)bcur * G + g) * D + dd) * plane + pix_off, acc[g][d] * inv);
                if (p.wsum != nullptr) p.wsum[((size_t)bcur * D + dd) * plane + pix_off] = wsum[d];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
// Scratch for the tile boxes: one grow-only device buffer per (device, stream), so that launches on different streams
// never share it and launches on one stream reuse it in stream order.  Allocation is not possible while the stream is
// being captured into a CUDA graph: callers capture after one eager call (CascadePlan.capture() does).
void* epi_workspace(cudaStream_t stream, size_t bytes, int* status);

struct SmCount {
    int n[64] = {};
    int get() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        int v = __atomic_load_n(&n[dev], __ATOMIC_ACQUIRE);
        if (v == 0) {
            if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
            __atomic_store_n(&n[dev], v, __ATOMIC_RELEASE);
        }
        return v;
    }
};

template <int C, int CPG, int D, typename T>
static int launch_pipe(EpiFwdParams& p, cudaStream_t stream, bool* built) {
    using K = PipeCfg<C, D, (int)sizeof(T)>;
    *built = false;
    if (p.B * p.Nsrc > K::kMaxBatchViews) return MVSTER_OK;
    const int tilesX = (p.W + K::TILE_W - 1) / K::TILE_W, tilesY = (p.H + K::TILE_H - 1) / K::TILE_H;
    if (tilesY > 65535 || p.B > 65535 || (double)tilesX * tilesY * p.B >= 2147483648.0) return MVSTER_OK;
    if (!encode_view_maps<K, T>(p, C)) return MVSTER_OK;
    const int numTiles = tilesX * tilesY * p.B;
    int wst = MVSTER_OK;
    int4* boxes = static_cast<int4*>(epi_workspace(stream, (size_t)numTiles * p.Nsrc * sizeof(int4), &wst));
    if (!boxes) return wst;
    *built = true;
    epi_tile_boxes_kernel<K><<<dim3(tilesX, tilesY, p.B), K::WARPS * 32, 0, stream>>>(p.rt, p.hypo, boxes, p.Nsrc, p.H, p.W,
                                                                                      tilesX, tilesY);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_fwd(tile boxes) launch");
    const int smem = K::smem_bytes(p.B * p.Nsrc);
    static int smem_set[64] = {};
    const int st = ensure_dynamic_smem_bytes(epi_fwd_pipe_kernel<C, CPG, D, T>, smem, smem_set, "epi_fwd(pipe): cudaFuncSetAttribute");
    if (st != MVSTER_OK) return st;
    static SmCount sms;
    const int grid = min(numTiles, sms.get() * K::CTAS);
    PipeArgs a{boxes, tilesX, tilesY, numTiles};
    epi_fwd_pipe_kernel<C, CPG, D, T><<<grid, K::WARPS * 32, smem, stream>>>(p, a);
    count_launch();
    MVSTER_CHECK_LAUNCH("epi_fwd(pipe) launch");
    return MVSTER_OK;
}

}  // namespace mvster
