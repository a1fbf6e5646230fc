/*
 * dfmir_b200.h — C ABI of libdfmir_b200.so (sm_100a).
 *
 * Drop-in boundary for the translation + registration hot path of heyblackC/DFMIR.  The reference
 * has no FFI of its own (pure PyTorch); each entry point below names the
 * reference call site (file:line, relative to the reference repo) whose library kernels it replaces.
 *
 * Conventions (all entry points):
 *   - plain C types only; every pointer is a DEVICE pointer borrowed from the caller (PyTorch's
 *     allocator owns all memory), except `shape` arrays which are host ints;
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*), no hidden allocation, no
 *     implicit synchronisation, re-entrant;
 *   - returns 0 on success, a negative DFMIR_ERR_* code otherwise; dfmir_last_error() returns a
 *     thread-local message for the last failing call;
 *   - scratch memory is caller-provided: query `*_workspace_bytes`, pass pointer + size;
 *   - planar tensors are the reference's NCHW / NCDHW fp32; "nhwc" tensors are channels-last
 *     (N, spatial..., C) fp32 — the layout the convolution stacks keep internally.
 */
#ifndef DFMIR_B200_H
#define DFMIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFMIR_ABI_VERSION 7

#define DFMIR_INTERP_LINEAR 0
#define DFMIR_INTERP_NEAREST 1
/* coordinate arithmetic: 0 reproduces the CPU ATen path (true division, the oracle),
 * 1 reproduces the CUDA ATen path (multiply by the fp32 reciprocal of S-1). */
#define DFMIR_COORD_IEEE_DIV 0
#define DFMIR_COORD_RCP_MUL 1

const char* dfmir_last_error(void);
int dfmir_abi_version(void);
/* number of kernel launches issued by this library since the last reset (bench.py gpu_launches) */
long long dfmir_launch_count(void);
void dfmir_launch_count_reset(void);

/* ---- K3: SpatialTransformer.forward — models/voxelmorph/torchvoxelmorph/layers.py:30-48
 * src (B,C,*S), flow (B,nd,*S) in voxel units (ij order) -> out (B,C,*S); zeros outside.
 * idx_out (nullable): int32 (B,nd,*S) — floor() (linear) or nearbyint() (nearest) of the
 * un-normalised sampling coordinate, the "deformation-field indices" of the parity target. */
int dfmir_warp_fwd(const float* src, const float* flow, float* out, int32_t* idx_out, int B, int C, int nd,
                   const int* shape, int interp, int coord_mode, void* stream);
/* d_src (nullable) must be zero-filled by the caller (atomically accumulated); d_flow (nullable). */
int dfmir_warp_bwd(const float* grad_out, const float* src, const float* flow, float* d_src, float* d_flow,
                   int B, int C, int nd, const int* shape, int coord_mode, void* stream);

/* ---- K4: VecInt.forward — layers.py:64-68 (and the negated flow of vxm networks.py:1125,1129-1130)
 * vel (B,nd,*S).  steps: (keep_all ? nsteps : 2) slabs of (Bv,nd,*S), Bv = bidir ? 2B : B; the
 * integrated field is slab keep_all ? nsteps-1 : (nsteps-1)&1; rows B..2B-1 hold integrate(-vel). */
int dfmir_vecint_fwd(const float* vel, float* steps, int B, int nd, const int* shape, int nsteps, int bidir,
                     int keep_all, int coord_mode, void* stream);
/* work: 2 slabs (Bv,nd,*S) scratch; d_vel (B,nd,*S) overwritten. */
int dfmir_vecint_bwd(const float* grad_out, const float* vel, const float* steps, float* work, float* d_vel,
                     int B, int nd, const int* shape, int nsteps, int bidir, int coord_mode, void* stream);

/* ---- K4+K5 fused: ONE cooperative launch for integrate -> resize x2 -> warp -> NCC + Grad
 * (vxm networks.py:1129-1139 tail of VxmDense.forward + util/losses.py:92-130, 183-261).
 * vel (B,nd,*half) -> steps (nsteps,B,nd,*half) every squaring step (kept for the backward), flow_full
 * (B,nd,*2half), warped (B,C,*2half), out[4] = {ncc loss, sum cc, voxel count, grad loss}.
 * ncc_reduction / grad_penalty / grad_mult as in dfmir_ncc_fwd / dfmir_grad_loss_fwd; win in {5,7,9}.
 * flow_full and warped are bit-identical to dfmir_vecint_fwd + dfmir_resize_linear_fwd + dfmir_warp_fwd. */
size_t dfmir_fused_reg_workspace_bytes(void);
int dfmir_fused_reg_fwd(const float* vel, const float* moving, const float* fixed, float* steps, float* flow_full,
                        float* warped, float* out, void* ws, size_t ws_bytes, int B, int C, int nd,
                        const int* half_shape, int nsteps, int win, float eps, int ncc_reduction, int grad_penalty,
                        float grad_mult, int coord_mode, void* stream);

/* ---- ResizeTransform.forward — layers.py:85-97 (F.interpolate align_corners=True + rescale)
 * y = post_mul * interp(pre_mul * x); x (BC,*in_shape) -> y (BC,*out_shape). */
int dfmir_resize_linear_fwd(const float* x, float* y, int BC, int nd, const int* in_shape, const int* out_shape,
                            float pre_mul, float post_mul, void* stream);
int dfmir_resize_linear_bwd(const float* gy, float* gx, int BC, int nd, const int* in_shape,
                            const int* out_shape, float pre_mul, float post_mul, void* stream);

/* ---- K5: NCC_Loss.forward — util/losses.py:183-261 (reduction 0: -sqrt(mean cc)) and
 * vxm losses.py:15-67 (reduction 1: -mean cc).  I,J (B,1,*S) planar, nd in {2,3}; `win` = window
 * edge (uniform, odd: 3..11); mask (nullable) float (B,1,*S).  out: 3 floats {loss, sum, norm}. */
size_t dfmir_ncc_workspace_bytes(int B, int nd, const int* shape, int win);
int dfmir_ncc_fwd(const float* I, const float* J, const float* mask, float* out, void* ws, size_t ws_bytes,
                  int B, int nd, const int* shape, int win, float eps, int reduction, void* stream);
/* gradient wrt I (swap I and J for the gradient wrt J); grad_loss: device scalar */
int dfmir_ncc_bwd(const float* I, const float* J, const float* mask, const float* fwd_out,
                  const float* grad_loss, float* dI, void* ws, size_t ws_bytes, int B, int nd, const int* shape,
                  int win, float eps, int reduction, void* stream);

/* ---- Grad_Loss — util/losses.py:92-130; smooothing_loss — models/registration_model.py:25-32
 * flow as (planes = B*C, *shape); penalty 1 = l1, 2 = l2; loss: 1 float. */
size_t dfmir_grad_loss_workspace_bytes(void);
int dfmir_grad_loss_fwd(const float* flow, float* loss, void* ws, size_t ws_bytes, int planes, int nd,
                        const int* shape, int penalty, float loss_mult, void* stream);
int dfmir_grad_loss_bwd(const float* flow, const float* grad_loss, float* d_flow, int planes, int nd,
                        const int* shape, int penalty, float loss_mult, void* stream);

/* ---- calculate_L1_loss — models/registration_model.py:255-263, masks of :160-161
 * mask: nullable uint8; or fused mask (mu > thr) | (mv > thr) when mu/mv given; neither = mean.
 * out: 2 floats {loss, sum(mask)}; an empty mask yields loss 0 without a host sync. */
size_t dfmir_l1_masked_workspace_bytes(void);
int dfmir_l1_masked_fwd(const float* a, const float* b, const uint8_t* mask, const float* mu, const float* mv,
                        float thr, float* out, void* ws, size_t ws_bytes, long long n, void* stream);
int dfmir_l1_masked_bwd(const float* a, const float* b, const uint8_t* mask, const float* mu, const float* mv,
                        float thr, const float* fwd_out, const float* grad_loss, float* da, float* db,
                        long long n, void* stream);

/* ---- K1/K2: nn.Conv2d / nn.Conv3d — models/networks.py:983,995,1016,1023,1201,1214 (ResnetGenerator),
 * models/voxelmorph/torchvoxelmorph/networks.py:1515 (ConvBlock) and :1077 (flow head).
 * Activations are channels-last with explicit element strides {n, spatial..., c} (so padded buffers,
 * interior views and the planar flow output need no copy); zero padding only (reflection padding is
 * materialised by the producer, see dfmir_instnorm_fwd / dfmir_pad_reflect_fwd).
 * Weights: forward / wgrad layout [tap][Cin][Cout], tap = (kd*KH + kh)*KW + kw; dgrad layout
 * [tap][Cout][Cin].  The descriptor always describes the FORWARD convolution. */
#define DFMIR_ACT_NONE 0
#define DFMIR_ACT_LEAKY 1 /* LeakyReLU(0.2), vxm networks.py:1516 */
#define DFMIR_ACT_TANH 2  /* models/networks.py:1024 */
#define DFMIR_ACT_RELU 3
typedef struct dfmir_conv_desc {
  int nd;                 /* 2 or 3 */
  int N, Cin, Cout;
  int in_shape[3], out_shape[3], kernel[3], pad[3];
  int stride;             /* same on every axis (1 or 2 in the reference) */
  int act;                /* epilogue activation of the forward (DFMIR_ACT_*) */
  long long x_strides[5]; /* input  element strides {n, spatial[nd], c} */
  long long y_strides[5]; /* output element strides {n, spatial[nd], c} */
} dfmir_conv_desc;
/* engine: 0 = fp32 CUDA cores (exact fp32 accumulate, any shape); 1 = tcgen05 tensor cores (TF32
 * operands, fp32 accumulate) — only for shapes dfmir_conv_umma_supported() accepts. */
int dfmir_conv_fwd(const float* x, const float* w, const float* bias, float* y, const dfmir_conv_desc* d,
                   void* stream);
int dfmir_conv_dgrad(const float* dy, const float* wt, float* dx, const dfmir_conv_desc* d, void* stream);
/* tcgen05 path (2-D / 3-D, stride 1, reduction-side channels a multiple of 4 and >= 16, any output channel count,
 * channels-last operands with 16-byte aligned strides; TF32 operands, fp32 accumulate in TMEM).
 * fwd weights: [tap][Cout][Cin]; dgrad weights: [tap][Cin][Cout] (each K-major for its product). */
int dfmir_conv_umma_supported(const dfmir_conv_desc* d, int dgrad);
int dfmir_conv_umma_fwd(const float* x, const float* w, const float* bias, float* y, const dfmir_conv_desc* d,
                        void* stream);
/* Forward with the InstanceNorm statistics of its result as an epilogue by-product (models/networks.py:984,996,1020,
 * 1201-1215: InstanceNorm2d follows every generator convolution): stat_rows receives [N][rows][Cout] float2 {sum, sum of
 * squares} over row groups of <= 32 voxels; dfmir_instnorm_fwd_rows consumes them.  dfmir_conv_umma_stat_rows: rows per
 * image for this shape, 0 when the kernel chosen for it has no such epilogue. */
int dfmir_conv_umma_stat_rows(const dfmir_conv_desc* d);
int dfmir_conv_umma_fwd_stats(const float* x, const float* w, const float* bias, float* y, const dfmir_conv_desc* d,
                              float* stat_rows, void* stream);
int dfmir_conv_umma_dgrad(const float* dy, const float* w, float* dx, const dfmir_conv_desc* d, void* stream);
/* dx += data gradient (the ResnetBlock input already holds the residual branch's gradient, models/networks.py:1218-1221):
 * covered by the CTA-pair kernel (2-D 3x3, Cin a multiple of 256); _supported tells, else DFMIR_ERR_UNSUPPORTED. */
int dfmir_conv_umma_dgrad_acc_supported(const dfmir_conv_desc* d);
int dfmir_conv_umma_dgrad_acc(const float* dy, const float* w, float* dx, const dfmir_conv_desc* d, void* stream);
/* tcgen05 weight gradient (2-D / 3-D, stride 1, channels-last x and dy, channel counts multiples of 4; >= 16 unless
 * the kernel is 3x3 / 3x3x3): split-K over output voxels, TF32 operands, fp32 accumulate in TMEM, red.add into
 * dw [tap][Cin][Cout] / db [Cout] (nullable) — both ACCUMULATED into: zero-fill them first. */
int dfmir_conv_umma_wgrad_supported(const dfmir_conv_desc* d);
int dfmir_conv_umma_wgrad(const float* x, const float* dy, float* dw, float* db, const dfmir_conv_desc* d,
                          void* stream);
/* dw [tap][Cin][Cout] and db [Cout] (nullable) are accumulated into: zero-fill them first. */
int dfmir_conv_wgrad(const float* x, const float* dy, float* dw, float* db, const dfmir_conv_desc* d,
                     void* stream);
/* dx = dy * act'(y) for the epilogue activations, contiguous arrays of n elements */
int dfmir_act_bwd(const float* y, const float* dy, float* dx, long long n, int act, void* stream);

/* ---- InstanceNorm2d(affine=False) [+ReLU] [+skip add] [+ReflectionPad2d] — models/networks.py:984,
 * 996,1020 and ResnetBlock :1193-1221.  x (N,H,W,C) channels-last; y (N,H+2p,W+2p,C) with a mirrored
 * halo of width p = out_pad; res (nullable) (N,H+2rp,W+2rp,C) is read at its interior.
 * stats (N,C,2) = {mean, rstd}; ws: dfmir_instnorm_workspace_bytes. */
size_t dfmir_instnorm_workspace_bytes(int N, int C);
int dfmir_instnorm_fwd(const float* x, const float* res, float* y, float* stats, void* ws, size_t ws_bytes, int N,
                       int H, int W, int C, float eps, int relu, int out_pad, int res_pad, void* stream);
/* Same layer, statistics reduced from the row sums written by dfmir_conv_umma_fwd_stats (no pass over x for them) */
int dfmir_instnorm_fwd_rows(const float* x, const float* res, float* y, float* stats, const float* stat_rows,
                            int rows_per_image, int N, int H, int W, int C, float eps, int relu, int out_pad, int res_pad,
                            void* stream);
/* dy (N,H+2p,W+2p,C) -> dx (N,H,W,C); dres (nullable, (N,H+2rp,W+2rp,C), halo zeroed here) */
int dfmir_instnorm_bwd(const float* dy, const float* x, const float* stats, float* dx, float* dres, void* ws,
                       size_t ws_bytes, int N, int H, int W, int C, int relu, int out_pad, int res_pad,
                       void* stream);
/* Same, plus dbias (C) = sum of dx over samples and pixels: dx is the output gradient of the convolution that
 * produced x (nn.Conv2d(..., bias=True) followed by InstanceNorm2d, models/networks.py:983-984, 1201-1202), so this is
 * that convolution's bias gradient without another pass over dx.  dbias NULL: identical to dfmir_instnorm_bwd. */
int dfmir_instnorm_bwd_bias(const float* dy, const float* x, const float* stats, float* dx, float* dres, float* dbias,
                            void* ws, size_t ws_bytes, int N, int H, int W, int C, int relu, int out_pad, int res_pad,
                            void* stream);
/* ---- nn.ReflectionPad2d — models/networks.py:982,1022 */
int dfmir_pad_reflect_fwd(const float* x, float* y, int N, int H, int W, int C, int pad, void* stream);
int dfmir_pad_reflect_bwd(const float* dy, float* dx, int N, int H, int W, int C, int pad, void* stream);
/* ---- Downsample / Upsample (anti-aliased) — models/networks.py:37-60, 73-93; channels-last */
int dfmir_blur_down_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);
int dfmir_blur_down_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream);
int dfmir_blur_up_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);
int dfmir_blur_up_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream);
/* ---- space-to-depth view of a channels-last activation: the stride-2 3^nd encoder convolutions of VoxelMorph's U-Net
 * (vxm networks.py:1514-1515, ConvBlock stride=2; Unet.downarm :52-56) become stride-1 2^nd convolutions over it and run
 * on the tcgen05 kernels.  src (N,*2*out_shape,C) -> dst (N,*out_shape,2^nd*C), channel = parity(d,h,w)*C + c;
 * inverse != 0: the inverse permutation (= the adjoint, used for the data gradient). */
int dfmir_space_to_depth(const float* src, float* dst, int N, int nd, const int* out_shape, int C, int inverse, void* stream);
/* ---- nn.Upsample(x2, nearest) + torch.cat([x, skip], 1) — vxm networks.py:99-102; channels-last.
 * a (N,*shape/2,C1), b (N,*shape,C2) -> y (N,*shape,C1+C2) */
int dfmir_upsample_concat_fwd(const float* a, const float* b, float* y, int N, int nd, const int* shape, int C1,
                              int C2, void* stream);
int dfmir_upsample_concat_bwd(const float* dy, float* da, float* db, int N, int nd, const int* shape, int C1,
                              int C2, void* stream);
/* same with a channel stride Cs >= C1 + C2 of y / dy (extra channels written as zero): keeps the pixel stride a
 * multiple of 16 bytes so the next convolution can load the tensor with TMA (34 -> 36 channels in VoxelMorph) */
int dfmir_upsample_concat_padded_fwd(const float* a, const float* b, float* y, int N, int nd, const int* shape, int C1,
                                     int C2, int Cs, void* stream);
int dfmir_upsample_concat_padded_bwd(const float* dy, float* da, float* db, int N, int nd, const int* shape, int C1,
                                     int C2, int Cs, void* stream);

/* ---- K6: PatchNCELoss.forward — models/patchnce.py:14-55.  q,k (B*P, D); S (B,P,P) scratch that the
 * forward leaves holding dLoss/dS; loss (B*P).  k is treated as detached (patchnce.py:17). */
int dfmir_patchnce_fwd(const float* q, const float* k, float* S, float* loss, int B, int P, int D, float T,
                       void* stream);
int dfmir_patchnce_bwd(const float* S, const float* k, const float* g, float* work, float* dq, int B, int P, int D,
                       void* stream);
/* tensor-core variant (K6): the two products run on the tcgen05 kernel through dfmir_bmm_nt_umma with 3xTF32-split
 * operands (hi*hi + hi*lo + lo*hi: fp32-class accuracy, the reference's torch.bmm is fp32).
 * dfmir_tf32_split3: x (rows, D) -> out (rows, 3D), a-style [hi|hi|lo] or b-style [hi|lo|hi].
 * dfmir_bmm_nt_umma: C[b] (M,N) = A[b] (M,K) * B[b]^T (N,K), dense row-major, M % 256 == 0 (N > 64), K % 4 == 0. */
int dfmir_tf32_split3(const float* x, float* out, long long rows, int D, int b_style, void* stream);
/* the same split stacked along the rows (out (3, rows, D)), for products that reduce over the rows (dW = dy^T x) */
int dfmir_tf32_split3_rows(const float* x, float* out, long long rows, int D, int b_style, void* stream);
int dfmir_bmm_nt_umma(const float* A, const float* B, float* C, int batch, int M, int N, int K, void* stream);
int dfmir_patchnce_tc_fwd(const float* q3, const float* k3, float* S, float* loss, int B, int P, int D3, float T,
                          void* stream);
int dfmir_patchnce_scale(const float* S, const float* g, float* work, int B, int P, void* stream);
/* ---- PatchSampleF — models/networks.py:597-624: patch gather, MLP products, L2 normalisation */
int dfmir_gather_patches_fwd(const float* feat, const long long* ids, float* out, int B, int P, int C, int Wd,
                             const long long* strides, void* stream);
int dfmir_gather_patches_bwd(const float* dout, const long long* ids, float* dfeat, int B, int P, int C, int Wd,
                             const long long* strides, void* stream);
int dfmir_l2norm_fwd(const float* x, float* y, float* norms, int rows, int D, void* stream);
int dfmir_l2norm_bwd(const float* x, const float* norms, const float* dy, float* dx, int rows, int D, void* stream);
/* strided batched product C[b](m,n) (+)= alpha * sum_k A[b](m,k) B[b](k,n) (+ bias[n]) (ReLU);
 * strides {batch,row,col} in elements.  nn.Linear of PatchSampleF.create_mlp (networks.py:587-595). */
int dfmir_gemm(const float* A, const float* B, const float* bias, float* C, int batch, int M, int N, int K,
               const long long* sA, const long long* sB, const long long* sC, float alpha, int accumulate, int relu,
               void* stream);

/* ---- fused multi-tensor Adam — torch.optim.Adam of models/registration_model.py:114-115,135 (optimizer_G / _F / _R:
 * lr opt.lr, betas (opt.beta1, opt.beta2), eps 1e-8) stepped at :168-171.  One launch per optimizer.
 * tensors: device array of 40-byte records {float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
 * int64 numel}; work: device array of n_work int pairs {tensor, chunk} (chunks of dfmir_adam_chunk_elems() elements);
 * step: device float, incremented before the update; lr_dev: device float, or NULL to use lr_host.  Hyper-parameters
 * are doubles: the step's scalars (1 - beta, bias corrections) are formed in double like torch.optim.Adam's. */
int dfmir_adam_chunk_elems(void);
int dfmir_adam_multi(const void* tensors, const void* work, int n_work, float* step, const float* lr_dev, double lr_host,
                     double beta1, double beta2, double eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFMIR_B200_H */
