// The per-pixel arithmetic of the stagenet tail (models/mvs4net_utils.py:1109-1156), shared by the streaming tail
// kernel (tail.cu) and the fused regulariser-tail kernel (regtail.cu) so that both produce bit-identical results
// from identical logits.
#pragma once

#include <string.h>

#include "common.cuh"

namespace mvster {

struct TailOut {
    float depth, conf, inv_min, inv_max;
};

// l: raw logits, h: depth hypotheses, at: softmax_D(l) (written).  D is a compile-time constant >= 2.
template <int D>
__device__ __forceinline__ TailOut tail_pixel(const float (&l)[D], const float (&h)[D], int mode, float split_itv,
                                              float (&at)[D]) {
    float lmax = l[0], lsum = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) { lmax = fmaxf(lmax, l[d]); lsum += l[d]; }
    float e[D], es = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) { e[d] = expf(l[d] - lmax); es += e[d]; }
    // arg-max over the softmax values, first maximum wins (torch.max semantics, reference :1129)
    float best = -1.f, reg = 0.f;
    int bi = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float a = e[d] / es;
        at[d] = a;
        if (a > best) { best = a; bi = d; }
        reg = fmaf(a, h[d], reg);
    }
    float depth = h[0];
#pragma unroll
    for (int d = 1; d < D; ++d) depth = (bi == d) ? h[d] : depth;
    if (mode == MVSTER_DEPTH_REGRESS) depth = reg;
    TailOut r;
    r.depth = depth;
    r.conf = lmax / lsum;  // photometric confidence on the raw logits: max / sum (reference :1109-1113,1138)
    // last_depth_itv = 1/hypo[:,2] - 1/hypo[:,1]  (reference :1152)
    const float itv = 1.0f / h[D > 2 ? 2 : 1] - 1.0f / h[1];
    const float inv = 1.0f / depth;
    r.inv_min = inv + split_itv * itv;
    r.inv_max = inv - split_itv * itv;
    return r;
}

}  // namespace mvster
