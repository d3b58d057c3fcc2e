// TMA staging pieces shared by the K1 forward (epi_fwd.cu) and backward (epi_bwd.cu) kernels of the fine stages:
// tile / box geometry, mbarrier + cp.async.bulk.tensor wrappers, the swizzled shared-memory texel read, the XU-free
// floor, the CTA bounding-box reduction and the host-side tensor-map builder.
#pragma once

#include "epi_common.cuh"

#ifndef MVSTER_TMA_LD
#define MVSTER_TMA_LD 1
#endif
#ifndef MVSTER_TMA_MINB
#define MVSTER_TMA_MINB 4
#endif
#ifndef MVSTER_TMA_WARPS
#define MVSTER_TMA_WARPS 4
#endif
#ifndef MVSTER_CELL_REUSE
#define MVSTER_CELL_REUSE 1
#endif
#ifndef MVSTER_TMA_BW
#define MVSTER_TMA_BW 48
#endif
#ifndef MVSTER_TMA_BHX
#define MVSTER_TMA_BHX 6
#endif

namespace mvster {

// ---------------------------------------------------------------------------------------------------------------------
// lane decomposition of the TMA-staged kernel: one lane owns all C <= 16 channels and all D hypotheses of a pixel
// ---------------------------------------------------------------------------------------------------------------------
// 64-byte texels (C = 16 fp32): the per-sample gather is twice as long as for 32-byte texels, and two lanes per pixel
// (each owning half of the hypotheses) at 24 warps per SM beat one lane at 16 warps (stage 3: 0.374 -> 0.352 ms); for
// 32-byte texels the single-lane layout wins (A/B in gpurun_out/variants2.log).  ES = bytes per feature element.
template <int C, int ES>
struct WideTexel {
    static constexpr bool value = (C * ES == 64);
};

template <int C, int CPG, int D, int ES = 4>
struct Split {
    static constexpr int CH = C;                           // channels per lane
    static constexpr int GPL = CH / CPG;                   // correlation groups per lane
    static constexpr int LD = (WideTexel<C, ES>::value && D % 2 == 0) ? 2 : MVSTER_TMA_LD;  // lanes splitting the hypotheses
    static constexpr int DL = D / LD;                      // hypotheses per lane
    static constexpr int LC = 1, L = LD;                   // lanes per pixel
    static constexpr int PPW = 32 / L;                     // pixels per warp
    static constexpr int NCHUNK = CH / 8;                  // 8-channel chunks per lane
    static constexpr int WX = L;                           // warps side by side in x
    static constexpr int WARPS = (LD == 2 && MVSTER_TMA_LD == 1) ? 8 : MVSTER_TMA_WARPS;   // warps per CTA
    static constexpr int MINB = (LD == 2 && MVSTER_TMA_LD == 1) ? 3 : MVSTER_TMA_MINB;     // CTAs per SM
    static constexpr int TILE_W = 32, TILE_H = WARPS / WX;
    static_assert(CH % 8 == 0 && 8 % CPG == 0, "a lane's 8-channel chunks must hold whole groups");
};


template <int C, int ES = 4>
struct TmaGeom {
    static constexpr int TB = C * ES;  // texel bytes (16, 32 or 64)
    // staging box in texels; the width is a multiple of 8 so that the swizzle phase depends on x only.  Sized for
    // ~25 % scale change / a dozen texels of epipolar span across a tile; larger footprints take the direct path.
    static constexpr int BW = MVSTER_TMA_BW;
    static constexpr int CTL_BYTES = 16 /* 2 mbarriers */ + 48 /* 3 bbox slots */;
    static constexpr int BH_EXTRA = WideTexel<C, ES>::value ? 4 : MVSTER_TMA_BHX;  // box height = tile height + BH_EXTRA
};

// ---------------------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void lds_pairs(uint32_t addr, f32x2& a, f32x2& b) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
// floor(s) as a float and as an int for |s| < 2^22 without the XU pipe (FRND / F2I are quarter-rate): adding 1.5*2^23
// with round-toward-minus-infinity leaves floor(s) in the low mantissa bits.
#ifndef MVSTER_FAST_FLOOR
#define MVSTER_FAST_FLOOR 1
#endif
__device__ __forceinline__ void floor_fi(float s, float& f, int& i) {
#if MVSTER_FAST_FLOOR
    const float t = __fadd_rd(s, 12582912.0f);
    f = t - 12582912.0f;
    i = __float_as_int(t) - 0x4B400000;
#else
    f = floorf(s);
    i = (int)f;
#endif
}
__device__ __forceinline__ int floor_i(float s) {
#if MVSTER_FAST_FLOOR
    return __float_as_int(__fadd_rd(s, 12582912.0f)) - 0x4B400000;
#else
    return __float2int_rd(s);
#endif
}
#ifndef MVSTER_BBOX_ATOM
#define MVSTER_BBOX_ATOM 1
#endif
// CTA bounding box of the sample positions: ptxas turns a warp-uniform-address shared atomic into REDUX + one ATOMS
__device__ __forceinline__ void bbox_update(int* slot, float lox, float loy, float hix, float hiy, int lane) {
#if MVSTER_BBOX_ATOM
    atomicMin(slot + 0, floor_i(lox)); atomicMin(slot + 1, floor_i(loy));
    atomicMax(slot + 2, floor_i(hix)); atomicMax(slot + 3, floor_i(hiy));
#else
    const int wx0 = __reduce_min_sync(0xffffffffu, __float2int_rd(lox));
    const int wy0 = __reduce_min_sync(0xffffffffu, __float2int_rd(loy));
    const int wx1 = __reduce_max_sync(0xffffffffu, __float2int_rd(hix));
    const int wy1 = __reduce_max_sync(0xffffffffu, __float2int_rd(hiy));
    if (lane == 0) {
        atomicMin(slot + 0, wx0); atomicMin(slot + 1, wy0);
        atomicMax(slot + 2, wx1); atomicMax(slot + 3, wy1);
    }
#endif
}
// 8 channels of one texel from the staged box as 4 packed fp32 pairs.  `a` is the (swizzled) address of the chunk's
// first 16 bytes.  fp32: 32 bytes = two LDS.128, the second half sits at a ^ 16 in both swizzle modes;
// bf16: 16 bytes = one LDS.128, widened to fp32 (exact).
template <typename T>
__device__ __forceinline__ void lds_chunk8(uint32_t a, P8& t) {
    if constexpr (sizeof(T) == 4) {
        lds_pairs(a, t.q[0], t.q[1]);
        lds_pairs(a ^ 16u, t.q[2], t.q[3]);
    } else {
        uint32_t x, y, z, w;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a));
        t.q[0] = pack2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
        t.q[1] = pack2(__uint_as_float(y << 16), __uint_as_float(y & 0xffff0000u));
        t.q[2] = pack2(__uint_as_float(z << 16), __uint_as_float(z & 0xffff0000u));
        t.q[3] = pack2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    }
}

// Same with a compile-time byte offset folded into the load instructions ([reg + imm]): the XOR that finds the second
// half is computed once per texel column and shared by both rows.
template <int OFF>
__device__ __forceinline__ void lds_pairs_at(uint32_t addr, f32x2& a, f32x2& b) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2+%3];" : "=l"(a), "=l"(b) : "r"(addr), "n"(OFF));
}
template <typename T, int OFF>
__device__ __forceinline__ void lds_chunk8_at(uint32_t a, P8& t) {
    if constexpr (sizeof(T) == 4) {
        const uint32_t a2 = a ^ 16u;
        lds_pairs_at<OFF>(a, t.q[0], t.q[1]);
        lds_pairs_at<OFF>(a2, t.q[2], t.q[3]);
    } else {
        lds_chunk8<T>(a + OFF, t);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: tensor maps of the source feature maps (one per view), box {C, BW, TILE_H + BH_EXTRA, 1}
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else
            cudaGetLastError();
    }
    return fn;
}

template <int C, int CPG, int D, typename T>
static inline bool make_maps(CUtensorMap* tmap, const void* const* src, int Nsrc, int B, int Hs, int Ws) {
    constexpr int ES = (int)sizeof(T), TBY = C * ES;
    using S = Split<C, CPG, D, ES>;
    constexpr int TILE_H = S::TILE_H;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)TBY, (cuuint64_t)Ws * TBY, (cuuint64_t)Hs * Ws * TBY};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)TmaGeom<C, ES>::BW, (cuuint32_t)(TILE_H + TmaGeom<C, ES>::BH_EXTRA), 1};
    // the swizzle span equals the texel size: 8 neighbouring texels land in 8 different 16-byte bank groups
    const CUtensorMapSwizzle swz = TBY == 16 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                   : (TBY == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B);
    const CUtensorMapDataType dt = ES == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    for (int v = 0; v < Nsrc; ++v) {
        CUresult r = enc(&tmap[v], dt, 4, const_cast<void*>(src[v]), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}


}  // namespace mvster
