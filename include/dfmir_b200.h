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

#define DFMIR_ABI_VERSION 1

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

#ifdef __cplusplus
}
#endif
#endif /* DFMIR_B200_H */
