// K0: relative homographies M_v = P_v @ inv(P_0) for every (batch, source view), in float64 on the device.
// Replaces the per-view prologue of stagenet.forward (models/mvs4net_utils.py:1047-1050) and the
// torch.inverse + matmul at the top of homo_warping (:32-34), which the reference redoes for every source view
// with a cuSOLVER launch each.  One tiny kernel per stage, no host synchronisation.
#include "common.cuh"

namespace mvster {

// P = E with rows 0..2 replaced by K[:3,:3] @ E[:3,:4]; the bottom row of E is kept (reference :1047-1050)
__device__ void compose_projection(const float* view /* [2,4,4] */, double P[4][4]) {
    const float* E = view;
    const float* K = view + 16;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += (double)K[i * 4 + k] * (double)E[k * 4 + j];
            P[i][j] = s;
        }
    for (int j = 0; j < 4; ++j) P[3][j] = (double)E[12 + j];
}

// Gauss-Jordan with partial pivoting; a singular matrix yields inf/nan exactly like an unchecked LU would.
__device__ void invert4(const double A[4][4], double inv[4][4]) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            a[i][j] = A[i][j];
            a[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(a[c][c]);
        for (int r = c + 1; r < 4; ++r)
            if (fabs(a[r][c]) > best) { best = fabs(a[r][c]); piv = r; }
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
        for (int j = 0; j < 4; ++j) inv[i][j] = a[i][4 + j];
}

__device__ void write_rt(const double Ps[4][4], const double Pri[4][4], float* rt) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += Ps[i][k] * Pri[k][j];
            rt[i * 4 + j] = (float)s;
        }
}

__global__ void compose_homographies_kernel(const float* __restrict__ proj, float* __restrict__ rt, int B, int N) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nsrc = N - 1;
    if (idx >= B * nsrc) return;
    const int b = idx / nsrc, v = idx % nsrc + 1;
    double Pr[4][4], Pri[4][4], Ps[4][4];
    compose_projection(proj + ((size_t)b * N) * 32, Pr);
    invert4(Pr, Pri);
    compose_projection(proj + ((size_t)b * N + v) * 32, Ps);
    write_rt(Ps, Pri, rt + (size_t)idx * 12);
}

__global__ void compose_pair_kernel(const float* __restrict__ src_proj, const float* __restrict__ ref_proj,
                                    float* __restrict__ rt, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double Pr[4][4], Pri[4][4], Ps[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            Pr[i][j] = (double)ref_proj[b * 16 + i * 4 + j];
            Ps[i][j] = (double)src_proj[b * 16 + i * 4 + j];
        }
    invert4(Pr, Pri);
    write_rt(Ps, Pri, rt + (size_t)b * 12);
}

}  // namespace mvster

using namespace mvster;

extern "C" int mvster_compose_homographies(const float* proj, float* rt, int B, int N, void* stream) {
    if (!proj || !rt) return fail(MVSTER_ERR_BAD_ARG, "compose_homographies: null pointer");
    if (B <= 0 || N < 2) return fail(MVSTER_ERR_BAD_ARG, "compose_homographies: need B>0 and N>=2 (got %d, %d)", B, N);
    DeviceGuard guard(rt);
    if (guard.status != MVSTER_OK) return guard.status;
    const int n = B * (N - 1);
    compose_homographies_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(proj, rt, B, N);
    count_launch();
    MVSTER_CHECK_LAUNCH("compose_homographies launch");
    return MVSTER_OK;
}

extern "C" int mvster_compose_homography_pair(const float* src_proj, const float* ref_proj, float* rt, int B,
                                              void* stream) {
    if (!src_proj || !ref_proj || !rt) return fail(MVSTER_ERR_BAD_ARG, "compose_homography_pair: null pointer");
    if (B <= 0) return fail(MVSTER_ERR_BAD_ARG, "compose_homography_pair: B must be > 0");
    DeviceGuard guard(rt);
    if (guard.status != MVSTER_OK) return guard.status;
    compose_pair_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(src_proj, ref_proj, rt, B);
    count_launch();
    MVSTER_CHECK_LAUNCH("compose_homography_pair launch");
    return MVSTER_OK;
}
