/*
 * mvster_b200.h - C ABI of the B200-native MVSTER cost-volume hot path (libmvster_b200.so).
 *
 * The reference (olivier-2018/Deep_reconstruction_with_epipolar_lines_MVSTER) is pure Python/PyTorch and has no
 * FFI of its own; its seam for this path is a set of Python call signatures.  Every entry point below names the
 * reference interface (file:line, relative to the reference tree) it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all sizes are ints, all buffers are raw pointers.
 *   - "dev" pointers are CUDA device pointers owned by the caller; kernels never allocate or free them.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Launches are asynchronous;
 *     a launch-time failure is returned immediately, an execution fault surfaces at the caller's next sync.
 *   - every function returns MVSTER_OK (0) or an mvster_status; mvster_last_error() returns a thread-local
 *     human-readable message for the last non-zero status on the calling thread.
 *   - the device is taken from the output pointer (cudaPointerGetAttributes) and restored on return, so the
 *     library is safe under nn.DataParallel (one host thread per GPU, reference test_mvs4.py:393).
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns MVSTER_ERR_NO_DEVICE.
 *
 * Feature layout: NHWC ("channels_last"), i.e. [B, H, W, C] with C contiguous; fp32 or bf16.
 */
#ifndef MVSTER_B200_H
#define MVSTER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVSTER_ABI_VERSION 1
#define MVSTER_MAX_SRC_VIEWS 15 /* source views per launch (reference NviewGen <= 16) */

typedef enum mvster_status {
    MVSTER_OK = 0,
    MVSTER_ERR_BAD_ARG = 1,     /* null pointer, non-positive size, ...                         */
    MVSTER_ERR_UNSUPPORTED = 2, /* (C, G, D) combination or mode without a compiled kernel       */
    MVSTER_ERR_ALIGN = 3,       /* a feature pointer is not 32-byte (fp32) / 16-byte (bf16) aligned */
    MVSTER_ERR_CUDA = 4,        /* a CUDA runtime call failed; message holds cudaGetErrorString  */
    MVSTER_ERR_NO_DEVICE = 5    /* no usable CUDA device / pointer is not device memory          */
} mvster_status;

typedef enum mvster_dtype { MVSTER_F32 = 0, MVSTER_BF16 = 1 } mvster_dtype;

typedef enum mvster_depth_mode {
    MVSTER_DEPTH_ARGMAX = 0, /* reference default: models/mvs4net_utils.py:1129-1130               */
    MVSTER_DEPTH_REGRESS = 1 /* depth_regression, models/module.py:935-941 (= commented :1133)     */
} mvster_depth_mode;

int mvster_version(void);
const char* mvster_last_error(void);
/* number of kernel launches issued by this library on the calling process since load (for bench accounting) */
uint64_t mvster_launch_count(void);

/* ---- K0: relative homographies ---------------------------------------------------------------------------
 * Replaces the per-view prologue of stagenet.forward (models/mvs4net_utils.py:1047-1050: P = E with rows 0..2 :=
 * K[:3,:3] @ E[:3,:4]) and homo_warping's `proj = src_proj @ inverse(ref_proj)` (:32-34), done once per stage for
 * all views in float64 on the device (no host sync).
 *   proj  dev [B, N, 2, 4, 4] fp32   ([:, v, 0] = extrinsic 4x4, [:, v, 1, :3, :3] = intrinsic)
 *   rt    dev [B, N-1, 12]    fp32   row-major 3x4 [R | t] of M_v = P_v @ inv(P_0), v = 1..N-1
 */
int mvster_compose_homographies(const float* proj, float* rt, int B, int N, void* stream);
/* same from already-composed 4x4 projections (the arguments of homo_warping, :21): src_proj, ref_proj dev [B,4,4] */
int mvster_compose_homography_pair(const float* src_proj, const float* ref_proj, float* rt, int B, void* stream);

/* ---- K1 forward: fused homography warp + group correlation + epipolar attention + view aggregation ---------
 * Replaces stagenet.forward steps 1-2 (models/mvs4net_utils.py:1030-1102) including homo_warping (:21-67); the
 * [B,C,D,H,W] warped volume is never materialised.
 *   ref      dev [B, H, W, C]      reference-view features, `dtype`
 *   src      HOST array of Nsrc dev pointers, each [B, Hs, Ws, C], `dtype`
 *   rt       dev [B, Nsrc, 12]     from mvster_compose_homographies
 *   hypo     dev [B, D, H, W]      depth hypotheses, fp32
 *   out      dev [B, G, D, H, W]   aggregated correlation volume (regnet input), fp32, written in full
 *   wsum     dev [B, D, H, W]      optional (NULL ok): 1e-8 + sum_v w_v, needed by the backward
 *   weights  dev [B, Nsrc, D, H, W] optional (NULL ok): per-view attention weights (reference `cor_weight`, :1083)
 * Supported: C in {8,16,32,64}, C/G in {1,2,4,8}, D in {4,8}; group_cor=True, attn_fuse_d=True (every shipped config;
 * the other two combinations: mvster_epi_fwd_mode below).
 */
int mvster_epi_fwd(const void* ref, const void* const* src, const float* rt, const float* hypo, float* out,
                   float* wsum, float* weights, int B, int Nsrc, int C, int G, int D, int H, int W, int Hs, int Ws,
                   float attn_temp, int dtype, void* stream);

/* Same forward with the two reference options that no shipped configuration uses:
 *   group_cor = 0   : per-channel variance cost (ref - warped)^2, G must equal C  (models/mvs4net_utils.py:1071)
 *   attn_fuse_d = 0 : one weight per pixel and view, max_D softmax_D(score), no temperature, no sqrt(C)
 *                     (models/mvs4net_utils.py:1078-1081,1098)
 * fp32 features only; group_cor = attn_fuse_d = 1 forwards to mvster_epi_fwd. */
int mvster_epi_fwd_mode(const void* ref, const void* const* src, const float* rt, const float* hypo, float* out, int B,
                        int Nsrc, int C, int G, int D, int H, int W, int Hs, int Ws, float attn_temp, int dtype,
                        int group_cor, int attn_fuse_d, void* stream);
/* mvster_epi_fwd_mode that also stores what mvster_epi_bwd_mode needs: wsum (NULL ok) is [B, D, H, W] when
 * attn_fuse_d = 1 and [B, H, W] when attn_fuse_d = 0 (cor_weight_sum is per pixel there, :1080). */
int mvster_epi_fwd_mode_ex(const void* ref, const void* const* src, const float* rt, const float* hypo, float* out,
                           float* wsum, int B, int Nsrc, int C, int G, int D, int H, int W, int Hs, int Ws,
                           float attn_temp, int dtype, int group_cor, int attn_fuse_d, void* stream);

/* ---- K1 backward -----------------------------------------------------------------------------------------
 * Replaces autograd through the same lines (grid_sampler_2d_backward, softmax_backward, ...).  Gradients flow to
 * the reference and source features only (the sampling grid is built under no_grad, :31; depth_hypo is detached at
 * models/MVS4Net.py:116).
 *   out, wsum  dev  the forward's outputs (saved; nothing else is stored between forward and backward)
 *   gout       dev [B, G, D, H, W] fp32
 *   grad_ref   dev [B, H, W, C]   fp32, written in full
 *   grad_src   HOST array of Nsrc dev pointers, each [B, Hs, Ws, C] fp32, ACCUMULATED with atomics: the caller
 *              zero-initialises them
 */
int mvster_epi_bwd(const void* ref, const void* const* src, const float* rt, const float* hypo, const float* out,
                   const float* wsum, const float* gout, float* grad_ref, float* const* grad_src, int B, int Nsrc,
                   int C, int G, int D, int H, int W, int Hs, int Ws, float attn_temp, int dtype, void* stream);

/* Backward of mvster_epi_fwd_mode_ex: autograd of models/mvs4net_utils.py:1066-1100 with group_cor=False (gradient
 * of (ref - warped)^2, :1071) and / or attn_fuse_d=False (gradient through max_D softmax_D, :1079, which reaches the
 * first maximum only, as torch.max(dim) does).  Arguments as mvster_epi_bwd; wsum as written by mvster_epi_fwd_mode_ex;
 * fp32 features only; group_cor = attn_fuse_d = 1 forwards to mvster_epi_bwd. */
int mvster_epi_bwd_mode(const void* ref, const void* const* src, const float* rt, const float* hypo, const float* out,
                        const float* wsum, const float* gout, float* grad_ref, float* const* grad_src, int B, int Nsrc,
                        int C, int G, int D, int H, int W, int Hs, int Ws, float attn_temp, int dtype, int group_cor,
                        int attn_fuse_d, void* stream);

/* ---- training-mode BatchNorm (+ ReLU) of the regulariser blocks ----------------------------------------------------
 * nn.BatchNorm3d / nn.BatchNorm2d in train() followed by ReLU, as inside ConvBnReLU3D / Deconv3d / Conv2d
 * (models/mvs4net_utils.py:123-130, 231-258, 884-926; autograd of the same in the backward), over PLANAR fp32
 * activations x [N, C, S] (S = D*H*W or H*W contiguous, S % 4 == 0):
 *   fwd: batch mean / biased variance per channel, y = relu?(gamma (x - mean) invstd + beta); mean and invstd [C] are
 *        returned for the backward; running_mean / running_var (nullable) get torch's momentum update (unbiased var)
 *   bwd: g = dy [y > 0];  dbeta = sum g;  dgamma = sum g xhat;  dx = gamma invstd (g - mean(g) - xhat mean(g xhat))
 * gamma / beta may be NULL (affine=False).  workspace: dev, mvster_bn_train_workspace_bytes(N, C, S) bytes, 8-byte aligned.
 * Partial sums are kept per (plane, chunk) in double and summed in a fixed order: bit-reproducible. */
long long mvster_bn_train_workspace_bytes(int N, int C, long long S);
int mvster_bn_train_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* invstd,
                        float* running_mean, float* running_var, float momentum, float eps, int relu, int N, int C,
                        long long S, void* workspace, void* stream);
int mvster_bn_train_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* mean,
                        const float* invstd, float* dx, float* dgamma, float* dbeta, int relu, int N, int C, long long S,
                        void* workspace, void* stream);

/* ---- weight gradient of the regulariser's 3-D convolutions (training mode) -----------------------------------------
 * Autograd of nn.Conv3d / nn.ConvTranspose3d inside ConvBnReLU3D / Deconv3d (models/mvs4net_utils.py:123-130, 884-926):
 *     dW[a][b][kd][ky][kx] = sum_{n,d,y,x} A[n,a,d,y,x] * B[n,b, d+kd-KD/2, s*y+ky-1, s*x+kx-1]     (B zero outside)
 *   Conv3d, kernel (KD,3,3), padding (KD/2,1,1), stride (1,s,s): A = grad_output [N,Cout,D,HA,WA], B = input [N,Cin,D,HB,WB]
 *   ConvTranspose3d, kernel (1,3,3), padding (0,1,1), output_padding (0,1,1), stride (1,2,2):
 *                                                     A = input [N,Cin,D,HA,WA], B = grad_output [N,Cout,D,2HA,2WA], s = 2
 * dW is written in PyTorch's weight layout ([CA,CB,KD,3,3]).  KD in {1,3}, s in {1,2}, CA % 8 == 0, planar fp32.
 * workspace: dev, mvster_conv3d_wgrad_workspace_bytes(...) bytes.  Bit-reproducible (fixed-order partial sums). */
long long mvster_conv3d_wgrad_workspace_bytes(int N, int CA, int CB, int KD, int D, int HA, int WA);
int mvster_conv3d_wgrad(const float* A, const float* B, float* dw, int N, int CA, int CB, int KD, int D, int HA, int WA,
                        int HB, int WB, int stride, void* workspace, void* stream);

/* Weight and bias gradient of the regulariser's last layer, prob = nn.Conv3d(8, 1, 1) (models/mvs4net_utils.py:914):
 *   dw[c] = sum_{n,s} x[n,c,s] g[n,0,s],  db[0] = sum g  (db nullable);  x [N,8,S], g [N,1,S] planar fp32, S % 4 == 0. */
long long mvster_conv1x1_wgrad_workspace_bytes(int N, int C, long long S);
int mvster_conv1x1_wgrad(const float* x, const float* g, float* dw, float* db, int N, int C, long long S,
                         void* workspace, void* stream);

/* ---- homo_warping compatibility (models/mvs4net_utils.py:21-67): materialises [B, C, D, H, W] fp32 -------- */
int mvster_homo_warp(const void* src, const float* rt /* dev [B,12] */, const float* hypo, float* warped, int B,
                     int C, int D, int H, int W, int Hs, int Ws, int dtype, void* stream);

/* ---- hypothesis schedule (models/mvs4net_utils.py:79-85 and :87-94) --------------------------------------
 *   depth_values dev [B, nvals] fp32 (only [:,0] and [:,nvals-1] are used, as in the reference)
 *   inv_min/max  dev [B, H/2, W/2] fp32  (previous stage's inverse_min_depth / inverse_max_depth)
 *   hypo         dev [B, D, H, W] fp32
 */
int mvster_init_inverse_range(const float* depth_values, int nvals, float* hypo, int B, int D, int H, int W,
                              void* stream);
int mvster_schedule_inverse_range(const float* inv_min, const float* inv_max, float* hypo, int B, int D, int H,
                                  int W, void* stream);

/* ---- K2a: depth / confidence tail (models/mvs4net_utils.py:1109-1156) --------------------------------------
 *   logits dev [B, D, H, W] regnet output;  hypo dev [B, D, H, W]
 *   attn   dev [B, D, H, W] softmax_D(logits)                           (ret_dict["attn_weight"])
 *   depth  dev [B, H, W]    hypo[argmax attn] or sum attn*hypo            (ret_dict["depth"])
 *   conf   dev [B, H, W]    max_D logits / sum_D logits; NULL in training (ret_dict["photometric_confidence"])
 *   inv_min / inv_max dev [B, H, W]: 1/depth +- split_itv * (1/hypo[:,2] - 1/hypo[:,1]); NULL when !inverse_depth
 */
int mvster_tail(const float* logits, const float* hypo, float split_itv, int depth_mode, float* attn, float* depth,
                float* conf, float* inv_min, float* inv_max, int B, int D, int H, int W, void* stream);
/* ---- K2a': last layers of the cost regulariser fused with the tail ---------------
 * Replaces, in eval mode, reg2d.forward's `x = conv0 + self.conv11(x); x = self.prob(x)` (models/mvs4net_utils.py:
 * 923-926; conv11 = ConvTranspose3d(16,8,(1,3,3),stride (1,2,2)) + BatchNorm3d + ReLU, prob = Conv3d(8,1,1)) AND the
 * stagenet tail (:1109-1156) with one kernel: the full-resolution [B,8,D,H,W] activation and the logits never
 * reach HBM.
 *   low     dev  [B, 16, D, H/2, W/2] input of conv11 (NCDHW, fp32)
 *   skip    dev  [B,  8, D, H,   W  ] conv0 output (NCDHW, fp32)
 *   w       HOST [3, 3, 16, 8]  conv11 weight as [ky][kx][ci][co] with the BatchNorm scale folded in
 *   params  HOST [17]           BatchNorm shift[8] (beta - mean*scale), prob weight[8], prob bias
 *   the remaining arguments are those of mvster_tail.  D in {4, 8}; H, W even.
 * The weights are host pointers because they travel as kernel parameters (constant bank), read by FFMA directly.
 */
int mvster_regtail(const float* low, const float* skip, const float* w, const float* params, const float* hypo,
                   float split_itv, int depth_mode, float* attn, float* depth, float* conf, float* inv_min,
                   float* inv_max, int B, int D, int H, int W, void* stream);
/* ---- direct fp32 convolutions for reg2d's high-resolution, few-channel layers ---------
 * Replaces, in eval mode, ConvBnReLU3D / (ConvTranspose3d + BatchNorm3d + ReLU) blocks of reg2d
 * (models/mvs4net_utils.py:889-912; building blocks :123-130) for the layers whose whole filter bank fits in the
 * kernel-parameter constant bank.  BatchNorm is folded by the caller: y = relu(conv(x, w) + bias) [+ skip].
 *   x     dev  [B, Cin, D, H, W] fp32 NCDHW;   y dev [B, Cout, D, H', W']
 *   w     HOST [kd, 3, 3, Cin, Cout]  ([kd][ky][kx][ci][co]);   bias HOST [Cout]
 *   mode  0: Conv3d k=(kd,3,3) stride 1 padding (kd/2,1,1), H'=H W'=W            (H, W even)
 *         1: Conv3d k=(1,3,3) stride (1,2,2) padding (0,1,1), H'=H/2 W'=W/2       (H even, W % 4 == 0)
 *         2: ConvTranspose3d k=(1,3,3) stride (1,2,2) padding (0,1,1) output_padding (0,1,1), H'=2H W'=2W;
 *            skip (dev, [B,Cout,D,2H,2W], nullable) is added AFTER the ReLU (reg2d: x = conv2 + conv9(x))
 * Compiled (Cin, Cout, kd, mode): (4,8,1,0) (8,8,1,0) (8,16,1,1) (16,16,3,0) (16,32,1,1) (32,16,1,2) (16,8,1,2);
 * anything else returns MVSTER_ERR_UNSUPPORTED and the caller keeps the layer on cuDNN.
 */
int mvster_conv3d_small(const float* x, const float* w, const float* bias, const float* skip, float* y, int B, int Cin,
                        int Cout, int D, int H, int W, int kd, int mode, int relu, void* stream);
/* 5x5 stride-2 padding-2 Conv2d + BatchNorm + ReLU of FPN4 (conv1.0 / conv2.0 / conv3.0, models/mvs4net_utils.py:438,444,
 * 446) with the folded weights in DEVICE memory: x dev [B,Cin,H,W], w_dev dev [5,5,Cin,Cout], bias_dev dev [Cout],
 * y dev [B,Cout,H/2,W/2]; Cin in {8,16,32}, Cout % 16 == 0, H % 4 == 0, W % 4 == 0. */
int mvster_conv2d_mid5(const float* x, const float* w_dev, const float* bias_dev, float* y, int B, int Cin, int Cout, int H,
                       int W, int relu, void* stream);

/* One slice of `Cout` output channels of a strided (mode 1) / transposed (mode 2, optional skip) reg2d layer whose filter
 * bank exceeds the kernel-parameter space: reg2d.conv5 (32->64, slices of 16) and conv7 (64->32, slices of 8),
 * models/mvs4net_utils.py:899,903.  w_host HOST [1,3,3,Cin,Cout], bias_host HOST [Cout] hold the slice; y (and skip) are
 * the full Cout_total-channel tensors, the slice writes channels co_off .. co_off+Cout-1. */
int mvster_conv3d_small_slice(const float* x, const float* w_host, const float* bias_host, const float* skip, float* y,
                              int B, int Cin, int Cout, int Cout_total, int co_off, int D, int H, int W, int mode,
                              int relu, void* stream);

/* ---- 32/64-channel stride-1 layers of reg2d / FPN4 with the folded weights in DEVICE memory ----
 * Replaces, in eval mode, reg2d.conv4 / conv6 (ConvBnReLU3D (3,3,3), models/mvs4net_utils.py:897,901) and FPN4
 * conv3.1 / conv3.2 (Conv2d 3x3 + BatchNorm + ReLU, :447-448; pass D = 1, kd = 1) - filter banks of 110 .. 442 KB that do
 * not fit the kernel-parameter space.  BatchNorm is folded by the caller: y = relu(conv(x, w) + bias).
 *   x dev [B, Cin, D, H, W] fp32;  w_dev dev [kd, 3, 3, Cin, Cout];  bias_dev dev [Cout];  y dev [B, Cout, D, H, W]
 *   Cin in {32, 64}, Cout % 16 == 0, kd in {1, 3}, H and W even; stride 1, padding (kd/2, 1, 1).
 */
int mvster_conv3d_mid(const float* x, const float* w_dev, const float* bias_dev, float* y, int B, int Cin, int Cout,
                      int D, int H, int W, int kd, int relu, void* stream);

/* ---- FPN4: few-channel 2-D convolutions and the fused top-down step ----------------------
 * mvster_conv2d_small: Conv2d + BatchNorm2d (folded) + ReLU blocks of FPN4's encoder (models/mvs4net_utils.py:431-449,
 * building block :231-258), NCHW planar fp32.  3x3 stride 1 padding 1 (H, W even) or 5x5 stride 2 padding 2 (H even,
 * W % 4 == 0).  One launch computes output channels co_off .. co_off+Cout-1 of a Cout_total-channel y:
 *   x dev [B,Cin,H,W];  y dev [B,Cout_total,H',W'];  w HOST [k,k,Cin,Cout] (slice);  bias HOST [Cout]
 * Compiled (Cin, Cout, k): (3,8,3) (8,8,3) (16,16,3) (32,16,3) (8,16,5) (16,16,5).
 */
int mvster_conv2d_small(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                        int Cout_total, int co_off, int H, int W, int ksize, int stride, int relu, void* stream);
/* mvster_conv3d_mid_tc: the same 32/64-channel stride-1 layers as mvster_conv3d_mid (reg2d conv4 / conv6,
 * models/mvs4net_utils.py:896-903; FPN4 conv3.1 / conv3.2 / out2, :456-468) as an implicit GEMM on the tensor cores
 * with fp32-grade accuracy: operands split into two TF32 numbers each, products accumulated as lo*hi + hi*lo + hi*hi
 * in fp32 (3xTF32).  w_hi / w_lo dev [kd,3,3,Cin,Cout] are made once per layer by mvster_tf32_split from the folded
 * fp32 weights.  Compiled (kd, Cin, Cout): (1,32,32) (1,64,64) (1,64,32) (3,32,32) (3,64,64); any H, W. */
int mvster_conv3d_mid_tc(const float* x, const float* w_hi, const float* w_lo, const float* bias, float* y, int B,
                         int Cin, int Cout, int D, int H, int W, int kd, int relu, void* stream);
/* hi[i] = rna_tf32(w[i]), lo[i] = rna_tf32(w[i] - hi[i]) for n device floats */
int mvster_tf32_split(const float* w, float* hi, float* lo, long long n, void* stream);
/* mvster_conv3d_mid_umma: the same layers on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
 * tensor memory, 3xTF32 operand split, four round-robin accumulators against the accumulator's truncation).
 * w_packed is made once per layer by mvster_umma_pack_weights from the folded fp32 weights [kd,3,3,Cin,Cout]
 * (same number of floats x 2).  Compiled (kd, Cin, Cout): (1,32,32) (1,64,64) (1,64,32) (3,32,32) (3,64,64); any H, W. */
int mvster_conv3d_mid_umma(const float* x, const float* w_packed, const float* bias, float* y, int B, int Cin, int Cout,
                           int D, int H, int W, int kd, int relu, void* stream);
int mvster_umma_pack_weights(const float* w, float* packed, int kd, int Cin, int Cout, void* stream);
/* mvster_fpn_topdown: one pyramid level of FPN4.forward (models/mvs4net_utils.py:488-495),
 *     intra = interpolate(prev, x2, bilinear, align_corners=True) + inner(lat);  feat = out_conv(intra)
 * with the 64-channel intra tile kept in shared memory; feat is written NHWC (what mvster_epi_fwd reads).
 *   prev      dev  [B,64,H/2,W/2] planar (nullable when intra_in is given)
 *   lat       dev  [B,Clat,H,W]   planar encoder map (nullable when intra_in is given)
 *   intra_in  dev  [B,64,H,W]     nullable: reload the tile instead of computing it (second output-channel slice)
 *   intra_out dev  [B,64,H,W]     nullable: store intra (needed by the next finer level)
 *   feat      dev  [B,H,W,Cout_total] NHWC; this launch writes channels co_off .. co_off+Cout-1
 *   w_out HOST [3,3,64,Cout] (slice of out_conv.weight as [ky][kx][c64][co]); w_in HOST [Clat,64]; b_in HOST [64]
 * Compiled (Clat, Cout): (8,8), (16,8).  H, W even.
 */
int mvster_fpn_topdown(const float* prev, const float* lat, const float* intra_in, float* intra_out, float* feat,
                       const float* w_out, const float* w_in, const float* b_in, int B, int Clat, int Cout,
                       int Cout_total, int co_off, int H, int W, void* stream);
/* mvster_fpn_topdown_ex: the same level with a choice of the emitted feature type - feat_dtype MVSTER_F32 or
 * MVSTER_BF16 (round to nearest even of the fp32 result; co_off and Cout_total then multiples of 8).  With MVSTER_BF16
 * FPN4 (models/mvs4net_utils.py:426-509) hands K1 its bf16 NHWC feature maps directly, without a cast pass. */
int mvster_fpn_topdown_ex(const float* prev, const float* lat, const float* intra_in, float* intra_out, void* feat,
                          int feat_dtype, const float* w_out, const float* w_in, const float* b_in, int B, int Clat,
                          int Cout, int Cout_total, int co_off, int H, int W, void* stream);

/* mvster_fpn_topdown_lin: the same pyramid level (models/mvs4net_utils.py:488-495) evaluated through its linearity,
 *     feat[x] = sum_tap bilinear(P_tap)(x + tap) + sum_tap (W[tap] Wi) lat[x + tap] + sum_tap W[tap] bi   (taps inside the image)
 * without the 64-channel `intra`:
 *   P     dev  [B,H/2,W/2,p_channels] NHWC fp32; P[.., p_off + tap*Cout + co] = sum_c out_conv.weight[co,c,tap] prev[c]
 *                                 (a GEMM over the coarser level's intra, done by the caller, or mvster_fpn_project_up)
 *   lat   dev  [B,Clat,H,W] planar encoder map
 *   feat  dev  [B,H,W,Cout] NHWC, feat_dtype MVSTER_F32 or MVSTER_BF16
 *   wc    HOST [9,Clat,Cout] composed lateral weights W[tap] Wi;  bc HOST [9,Cout] bias terms W[tap] bi
 * Compiled (Clat, Cout): (8,8), (16,16).  H, W even.  Differs from mvster_fpn_topdown by fp32 rounding only. */
int mvster_fpn_topdown_lin(const float* P, int p_channels, int p_off, const float* lat, void* feat, int feat_dtype,
                           const float* wc, const float* bc, int B, int Clat, int Cout, int H, int W, void* stream);

/* mvster_fpn_project_up: the projection P = Wp intra of a level whose intra = up2(prev) + inner(lat) is not formed:
 *     P[x] = bilinear(Q)(x) + (Wp Wi) lat[x] + Wp bi,     Q = Wp prev at the coarser resolution
 *   Q   dev  [B,H/2,W/2,q_channels] NHWC (channels q_off .. q_off+NP-1 are used);  lat dev [B,Clat,H,W] planar
 *   P   dev  [B,H,W,NP] NHWC;  wl HOST [Clat,NP];  bl HOST [NP].   Compiled (Clat, NP): (16,72). */
int mvster_fpn_project_up(const float* Q, int q_channels, int q_off, const float* lat, float* P, const float* wl,
                          const float* bl, int B, int Clat, int NP, int H, int W, void* stream);
/* backward of the tail w.r.t. the logits: softmax backward of g_attn (+ the regression term of g_depth when
 * depth_mode == MVSTER_DEPTH_REGRESS); g_attn / g_depth may be NULL (treated as zero) */
int mvster_tail_bwd(const float* attn, const float* hypo, const float* depth, const float* g_attn,
                    const float* g_depth, int depth_mode, float* g_logits, int B, int D, int H, int W, void* stream);

/* ---- K2b: geometric consistency filter (test_mvs4.py:612-670) and mask fusion (:716-749) ------------------
 * mvster_geo_check_pair == check_geometric_consistency for one (ref, src) pair:
 *   depth_ref, depth_src dev [H, W] fp32; K_* HOST 3x3 double; E_* HOST 4x4 double (as read by
 *   read_camera_parameters, test_mvs4.py:143-151)
 *   mask dev [H, W] uint8; depth_reprojected, x2d_src, y2d_src dev [H, W] fp32
 * mvster_geo_filter fuses all S source views of R reference views in one launch:
 *   depths, confs dev [V, H, W] fp32; K HOST [V, 9] double; E HOST [V, 16] double
 *   pairs HOST [R, 1+S] int32: (ref view, S source views); a negative source id is skipped
 *   photo, geo, final dev [R, H, W] uint8; depth_avg dev [R, H, W] fp32; geo_sum dev [R, H, W] int32 (NULL ok)
 */
int mvster_geo_check_pair(const float* depth_ref, const double* K_ref, const double* E_ref, const float* depth_src,
                          const double* K_src, const double* E_src, double condmask_pixel, double condmask_depth,
                          uint8_t* mask, float* depth_reprojected, float* x2d_src, float* y2d_src, int H, int W,
                          void* stream);
int mvster_geo_filter(const float* depths, const float* confs, const double* K, const double* E, const int32_t* pairs,
                      int V, int R, int S, double condmask_pixel, double condmask_depth, double photomask, int geomask,
                      uint8_t* photo, uint8_t* geo, uint8_t* final_mask, float* depth_avg, int32_t* geo_sum, int H,
                      int W, void* stream);

/* ---- depth2pts: fused depth map -> world points, reference test_mvs4.py:206-229 ---------
 *   depth dev [H, W] fp32;  K HOST [9], E HOST [16] float64 row-major;  xyz dev [H*W, 3] float64 (row-major pixel order)
 *   X_world = R^-1 (K^-1 [x+0.5, y+0.5, 1]^T * depth - t), all in float64 as the reference's NumPy arithmetic.
 */
int mvster_depth2pts(const float* depth, const double* K, const double* E, double* xyz, int H, int W, void* stream);

/* ---- K3: fused Sinkhorn / Wasserstein depth loss ----------------------------------------
 * Replaces `sinkhorn(gt_depth, hypo_depth, attn_weight, mask, iters, eps, continuous)` (models/mvs4net_utils.py:
 * 1164-1210) as called per stage by MVS4net_loss / Blend_loss (models/MVS4Net.py:234, :281), together with the
 * `range_err_ratio` statistic next to that call (:225-232).  One thread per pixel runs the 2*iters log-domain scaling
 * passes on its D x D(+1) problem in registers AND the backward sweep of the unrolled iterations, so the call returns
 * the loss and d(loss)/d(attn_weight); the reference's [B,HW,D,D] tensors never exist.
 *   gt_depth dev [B,H,W] fp32; hypo, attn dev [B,D,H,W] fp32; mask dev [B,H,W] uint8 (the reference's `mask > 0.5`)
 *   iters, eps, continuous: as in the reference (train_mvs4.py:73-75: --ot_iter, --ot_eps, --ot_continous)
 *   inverse_depth: selects the 1/depth form of range_err_ratio (models/MVS4Net.py:226-228)
 *   stats    dev [3] fp32, written: mean loss over masked pixels (NaN if none, like torch), masked count, range_err_ratio
 *   grad_px  dev [B,D,H,W] fp32, optional (NULL ok): d(per-pixel loss)/d(attn), zero at unmasked pixels;
 *            mvster_sinkhorn_bwd turns it into the gradient of the mean
 *   tmap     dev [B,H*W,D,D(+1)] fp32, optional (NULL ok): the transport map T_map (first return value of the reference)
 *   partials dev workspace of 3 * mvster_sinkhorn_blocks(...) doubles (deterministic two-pass reduction, no atomics)
 * D in {4, 8}.  mvster_sinkhorn_blocks returns 0 when (D, iters) does not fit the in-kernel backward.
 */
int mvster_sinkhorn_blocks(int B, int D, int H, int W, int iters, int continuous, int want_grad);
int mvster_sinkhorn_fwd(const float* gt_depth, const float* hypo, const float* attn, const uint8_t* mask, int iters,
                        float eps, int continuous, int inverse_depth, float* stats, float* grad_px, float* tmap,
                        double* partials, int B, int D, int H, int W, void* stream);
/* grad_attn[i] = grad_px[i] * grad_loss[0] / stats[1]   (grad_loss dev [1]: upstream gradient of the scalar loss) */
int mvster_sinkhorn_bwd(const float* grad_px, const float* stats, const float* grad_loss, float* grad_attn, int B,
                        int D, int H, int W, void* stream);

/* ---- layout helper: NCHW fp32 -> NHWC fp32/bf16 (the FPN emits NCHW, models/mvs4net_utils.py:504-507) ----- */
int mvster_nchw_to_nhwc(const float* in, void* out, int B, int C, int H, int W, int out_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVSTER_B200_H */
