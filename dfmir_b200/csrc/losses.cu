// K5: registration losses — windowed local NCC, deformation-field Grad / smoothing, masked L1.
//
// Replaces util/losses.py:183-261 (NCC_Loss: five box-filter F.conv{2,3}d + elementwise),
// util/losses.py:92-130 + models/registration_model.py:25-32 (Grad_Loss / smooothing_loss) and
// models/registration_model.py:255-263 (calculate_L1_loss) of the reference.
//
// NCC: the reference runs the 9^nd box filter as a dense convolution (729 MAC/voxel in 3-D).
// Here a CTA owns an 8x32 (y,x) column of outputs and marches along z: each slice is staged in
// shared memory with its halo, the window sums are built separably (x, then y, then a 9-deep ring
// of slice sums for z) and cc is reduced on the fly, so I and J are read from HBM once (plus halo)
// and nothing but per-CTA partial sums is written.  All window sums are fresh (no running-sum
// drift).  Reductions are deterministic: per-CTA partials in fp64, one finalize CTA.
#include "common.cuh"
#include "ncc.cuh"
#include "dfmir_b200.h"

namespace {
using namespace nccdev;

template <int WIN>
__global__ void __launch_bounds__(NT)
ncc_fwd_kernel(const float* __restrict__ I, const float* __restrict__ J, const float* __restrict__ mask,
               double* __restrict__ partials, BoxGeom g, float eps) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double sred[2][NT / 32];
  const long long hw = (long long)g.H * g.W, vol = hw * g.D;
  IJLoader ld{I, J, vol, hw, g.W};
  float wsz = (float)g.win * (float)g.win * (float)g.wz;
  CcReduce cs{mask, vol, hw, g.W, wsz, eps, 0.0, 0.0};
  box_march<5, WIN>(g, ld, cs, smem, blockIdx.x, blockIdx.y, blockIdx.z);
  double a = warp_sum_d(cs.acc_cc), m = warp_sum_d(cs.acc_m);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sred[0][w] = a; sred[1][w] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tm = 0;
    for (int i = 0; i < NT / 32; ++i) { ta += sred[0][i]; tm += sred[1][i]; }
    const long long cta = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partials[2 * cta] = ta; partials[2 * cta + 1] = tm;
  }
}

// out[0] = loss, out[1] = sum(cc [*mask]), out[2] = normaliser (voxel count or sum(mask))
__global__ void ncc_finalize_kernel(const double* __restrict__ partials, long long n, float* __restrict__ out,
                                    double count, int has_mask, int reduction) {
  __shared__ double s0[32], s1[32];
  double a = 0, m = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) { a += partials[2 * i]; m += partials[2 * i + 1]; }
  a = warp_sum_d(a); m = warp_sum_d(m);
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a; s1[threadIdx.x >> 5] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0; m = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s0[i]; m += s1[i]; }
    const double norm = has_mask ? m : count;
    float loss;
    if (has_mask && m == 0.0) loss = 0.f;                       // util/losses.py:255-256
    else if (reduction == 0) loss = -sqrtf((float)(a / norm));  // util/losses.py:254,259
    else loss = -(float)(a / norm);                             // vxm/losses.py:67
    out[0] = loss; out[1] = (float)a; out[2] = (float)norm;
  }
}

// ---- backward pass 1: per-voxel coefficient fields  A, A*uJ, Bi, Bi*uI
struct CoefWrite {
  const float* mask; float* coef; long long vol, hw, total; int W; float wsz, eps, gscale;
  __device__ __forceinline__ void consume(int b, int z, int y, int x, const float* s) {
    const CcTerms t = cc_terms(s, wsz, eps);
    const long long o = b * vol + z * hw + (long long)y * W + x;
    float gc = gscale;
    if (mask) gc *= mask[o];
    const float A = gc * 2.f * t.cross / t.denom;
    const float Bi = -gc * t.cross * t.cross * t.jvar / (t.denom * t.denom);
    coef[o] = A; coef[total + o] = A * t.uJ; coef[2 * total + o] = Bi; coef[3 * total + o] = Bi * t.uI;
  }
};

template <int WIN>
__global__ void __launch_bounds__(NT)
ncc_bwd_coef_kernel(const float* __restrict__ I, const float* __restrict__ J, const float* __restrict__ mask,
                    const float* __restrict__ fwd_out, const float* __restrict__ grad_loss,
                    float* __restrict__ coef, BoxGeom g, float eps, int reduction) {
  extern __shared__ __align__(16) float smem[];
  const long long hw = (long long)g.H * g.W, vol = hw * g.D;
  IJLoader ld{I, J, vol, hw, g.W};
  const float wsz = (float)g.win * (float)g.win * (float)g.wz;
  // d loss / d cc_p (uniform): reduction 0: -1/(2*sqrt(mean)*norm) = 1/(2*loss*norm); 1: -1/norm
  const float loss = fwd_out[0], norm = fwd_out[2];
  float gs;
  if (norm == 0.f) gs = 0.f;
  else if (reduction == 0) gs = 1.0f / (2.0f * loss * norm);
  else gs = -1.0f / norm;
  gs *= grad_loss[0];
  CoefWrite cs{mask, coef, vol, hw, (long long)g.B * vol, g.W, wsz, eps, gs};
  box_march<5, WIN>(g, ld, cs, smem, blockIdx.x, blockIdx.y, blockIdx.z);
}

// ---- backward pass 2: box-filter the coefficient fields and combine
struct CoefLoader {
  const float* coef; long long vol, hw, total; int W;
  __device__ __forceinline__ void load(int b, int z, int y, int x, bool inb, float* q) const {
    if (inb) {
      const long long o = b * vol + z * hw + (long long)y * W + x;
      q[0] = coef[o]; q[1] = coef[total + o]; q[2] = coef[2 * total + o]; q[3] = coef[3 * total + o];
    } else {
      q[0] = q[1] = q[2] = q[3] = 0.f;
    }
  }
};
struct GradCombine {
  const float* I; const float* J; float* dI; long long vol, hw; int W;
  __device__ __forceinline__ void consume(int b, int z, int y, int x, const float* s) {
    const long long o = b * vol + z * hw + (long long)y * W + x;
    dI[o] = J[o] * s[0] - s[1] + 2.f * (I[o] * s[2] - s[3]);
  }
};

template <int WIN>
__global__ void __launch_bounds__(NT)
ncc_bwd_apply_kernel(const float* __restrict__ I, const float* __restrict__ J, const float* __restrict__ coef,
                     float* __restrict__ dI, BoxGeom g) {
  extern __shared__ __align__(16) float smem[];
  const long long hw = (long long)g.H * g.W, vol = hw * g.D;
  CoefLoader ld{coef, vol, hw, (long long)g.B * vol, g.W};
  GradCombine cs{I, J, dI, vol, hw, g.W};
  box_march<4, WIN>(g, ld, cs, smem, blockIdx.x, blockIdx.y, blockIdx.z);
}

#define DISPATCH_WIN(win, ...)                                 \
  switch (win) {                                               \
    case 3: { constexpr int WIN = 3; __VA_ARGS__; } break;     \
    case 5: { constexpr int WIN = 5; __VA_ARGS__; } break;     \
    case 7: { constexpr int WIN = 7; __VA_ARGS__; } break;     \
    case 9: { constexpr int WIN = 9; __VA_ARGS__; } break;     \
    case 11: { constexpr int WIN = 11; __VA_ARGS__; } break;   \
    default: dfmir_set_error("ncc: window %d not supported (3,5,7,9,11)", win); return DFMIR_ERR_UNSUPPORTED; \
  }

}  // namespace

extern "C" size_t dfmir_ncc_workspace_bytes(int B, int nd, const int* shape, int win) {
  BoxGeom g;
  if (make_box(g, B, nd, shape, win) != 0) return 0;
  const dim3 gr = box_grid(g);
  const size_t partials = sizeof(double) * 2 * (size_t)gr.x * gr.y * gr.z;
  const size_t coef = sizeof(float) * 4 * (size_t)B * g.D * g.H * g.W;  // backward scratch
  return (partials > coef ? partials : coef) + 256;
}

// out: 3 floats on device {loss, sum(cc*mask), normaliser}
extern "C" int dfmir_ncc_fwd(const float* I, const float* J, const float* mask, float* out, void* ws,
                             size_t ws_bytes, int B, int nd, const int* shape, int win, float eps,
                             int reduction, void* stream) {
  BoxGeom g;
  DFMIR_CHECK_ARG(make_box(g, B, nd, shape, win) == 0, "dfmir_ncc_fwd: bad geometry (nd=%d; 2 or 3 supported)", nd);
  DFMIR_CHECK_ARG(I && J && out && ws, "dfmir_ncc_fwd: null pointer");
  DFMIR_CHECK_ARG(ws_bytes >= dfmir_ncc_workspace_bytes(B, nd, shape, win), "dfmir_ncc_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 gr = box_grid(g);
  double* partials = (double*)ws;
  DISPATCH_WIN(win, {
    const size_t sm = box_smem_bytes<5, WIN>(g.wz);
    DFMIR_CUDA(cudaFuncSetAttribute(ncc_fwd_kernel<WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    ncc_fwd_kernel<WIN><<<gr, NT, sm, st>>>(I, J, mask, partials, g, eps);
  });
  DFMIR_CHECK_LAUNCH("dfmir_ncc_fwd");
  const double count = (double)B * g.D * g.H * g.W;
  ncc_finalize_kernel<<<1, 256, 0, st>>>(partials, (long long)gr.x * gr.y * gr.z, out, count, mask != nullptr, reduction);
  DFMIR_CHECK_LAUNCH("dfmir_ncc_fwd(finalize)");
  return DFMIR_OK;
}

// Gradient wrt I (call with I and J swapped for the gradient wrt J). fwd_out = the 3 floats of the
// forward; grad_loss = device scalar (upstream gradient).
extern "C" int dfmir_ncc_bwd(const float* I, const float* J, const float* mask, const float* fwd_out,
                             const float* grad_loss, float* dI, void* ws, size_t ws_bytes, int B, int nd,
                             const int* shape, int win, float eps, int reduction, void* stream) {
  BoxGeom g;
  DFMIR_CHECK_ARG(make_box(g, B, nd, shape, win) == 0, "dfmir_ncc_bwd: bad geometry (nd=%d)", nd);
  DFMIR_CHECK_ARG(I && J && fwd_out && grad_loss && dI && ws, "dfmir_ncc_bwd: null pointer");
  DFMIR_CHECK_ARG(ws_bytes >= dfmir_ncc_workspace_bytes(B, nd, shape, win), "dfmir_ncc_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 gr = box_grid(g);
  float* coef = (float*)ws;
  DISPATCH_WIN(win, {
    const size_t sm5 = box_smem_bytes<5, WIN>(g.wz), sm4 = box_smem_bytes<4, WIN>(g.wz);
    DFMIR_CUDA(cudaFuncSetAttribute(ncc_bwd_coef_kernel<WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm5));
    DFMIR_CUDA(cudaFuncSetAttribute(ncc_bwd_apply_kernel<WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4));
    ncc_bwd_coef_kernel<WIN><<<gr, NT, sm5, st>>>(I, J, mask, fwd_out, grad_loss, coef, g, eps, reduction);
    ncc_bwd_apply_kernel<WIN><<<gr, NT, sm4, st>>>(I, J, coef, dI, g);
  });
  DFMIR_CHECK_LAUNCH("dfmir_ncc_bwd");
  return DFMIR_OK;
}

// =====================================================================================
// Grad_Loss / smooothing_loss
// =====================================================================================
namespace {

struct GGeom { int P; int S[3]; int nd; long long nvox; };  // P = B*C planes

// I: int when P * nvox < 2^31 (64-bit div/mod per element dominated this memory-bound kernel), else long long
template <typename I>
__global__ void __launch_bounds__(256)
grad_fwd_kernel(const float* __restrict__ x, double* __restrict__ partials, GGeom g, int penalty) {
  __shared__ double sred[3][8];
  float acc[3] = {0.f, 0.f, 0.f};
  const I nvox = (I)g.nvox;
  const I total = (I)g.P * nvox;
  const I st1 = g.S[2], st0 = (I)g.S[1] * g.S[2];
  for (I it = (I)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (I)gridDim.x * blockDim.x) {
    const I v = it % nvox;
    const int px = (int)(v % g.S[2]);
    const int py = (int)((v / g.S[2]) % g.S[1]);
    const int pz = (int)(v / st0);
    const float c = x[it];
    if (px + 1 < g.S[2]) { float d = fabsf(x[it + 1] - c); acc[2] += penalty == 2 ? d * d : d; }
    if (py + 1 < g.S[1]) { float d = fabsf(x[it + st1] - c); acc[1] += penalty == 2 ? d * d : d; }
    if (pz + 1 < g.S[0]) { float d = fabsf(x[it + st0] - c); acc[0] += penalty == 2 ? d * d : d; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double s = warp_sum_d((double)acc[k]);
    if (lane == 0) sred[k][w] = s;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0;
    for (int i = 0; i < 8; ++i) s += sred[threadIdx.x][i];
    partials[3 * (long long)blockIdx.x + threadIdx.x] = s;
  }
}

// volumes are stored with their spatial dims right-aligned in S[3] (2-D: S = {1,H,W}); the axis
// means are added in the reference's order and divided by nd (util/losses.py:105-116).
__global__ void grad_finalize_kernel(const double* __restrict__ partials, int n, float* __restrict__ loss,
                                     GGeom g, float loss_mult) {
  // launched with 96 threads: warp k sums the partials of axis k
  __shared__ double s[3];
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double a = 0;
  for (int i = lane; i < n; i += 32) a += partials[3 * (long long)i + k];
  a = warp_sum_d(a);
  if (lane == 0) s[k] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float d = 0.f;
    for (int k2 = 2; k2 >= 3 - g.nd; --k2) {  // x, y, (z): mean(dx) + mean(dy) + mean(dz)
      const double cnt = (double)g.P * (double)(g.nvox / g.S[k2]) * (double)(g.S[k2] - 1);
      d += (float)(s[k2] / cnt);
    }
    loss[0] = d / (float)g.nd * loss_mult;
  }
}

template <typename I>
__global__ void __launch_bounds__(256)
grad_bwd_kernel(const float* __restrict__ x, const float* __restrict__ grad_loss, float* __restrict__ dx,
                GGeom g, int penalty, float loss_mult) {
  const I nvox = (I)g.nvox;
  const I total = (I)g.P * nvox;
  const I strd[3] = {(I)g.S[1] * g.S[2], (I)g.S[2], 1};
  float coef[3];
  for (int k = 0; k < 3; ++k) {
    const double cnt = (double)g.P * (double)(g.nvox / g.S[k]) * (double)(g.S[k] - 1);
    coef[k] = (k >= 3 - g.nd && cnt > 0) ? (float)((double)grad_loss[0] * loss_mult / ((double)g.nd * cnt)) : 0.f;
  }
  for (I it = (I)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (I)gridDim.x * blockDim.x) {
    const I v = it % nvox;
    int p[3];
    p[2] = (int)(v % g.S[2]);
    p[1] = (int)((v / g.S[2]) % g.S[1]);
    p[0] = (int)(v / strd[0]);
    const float c = x[it];
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (coef[k] == 0.f) continue;
      if (p[k] > 0) {
        const float d = c - x[it - strd[k]];
        r += coef[k] * (penalty == 2 ? 2.f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)));
      }
      if (p[k] + 1 < g.S[k]) {
        const float d = x[it + strd[k]] - c;
        r -= coef[k] * (penalty == 2 ? 2.f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)));
      }
    }
    dx[it] = r;
  }
}

int make_ggeom(GGeom& g, int planes, int nd, const int* shape) {
  if (nd < 2 || nd > 3 || planes < 1) return -1;
  g.P = planes; g.nd = nd; g.nvox = 1;
  for (int d = 0; d < 3; ++d) {
    const int src = d - (3 - nd);
    g.S[d] = src >= 0 ? shape[src] : 1;
    if (g.S[d] <= 0) return -1;
    g.nvox *= g.S[d];
  }
  return 0;
}

inline int red_grid(long long items) {
  long long blocks = (items + 256 * 8 - 1) / (256 * 8);
  const long long cap = (long long)dfmir_num_sms() * 8;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace

extern "C" size_t dfmir_grad_loss_workspace_bytes(void) { return sizeof(double) * 3 * 148 * 8 * 2 + 256; }

extern "C" int dfmir_grad_loss_fwd(const float* flow, float* loss, void* ws, size_t ws_bytes, int planes,
                                   int nd, const int* shape, int penalty, float loss_mult, void* stream) {
  GGeom g;
  DFMIR_CHECK_ARG(make_ggeom(g, planes, nd, shape) == 0, "dfmir_grad_loss_fwd: bad geometry (nd=%d)", nd);
  DFMIR_CHECK_ARG(penalty == 1 || penalty == 2, "dfmir_grad_loss_fwd: penalty must be 1 (l1) or 2 (l2)");
  DFMIR_CHECK_ARG(flow && loss && ws, "dfmir_grad_loss_fwd: null pointer");
  const int grid = red_grid((long long)planes * g.nvox);
  DFMIR_CHECK_ARG(ws_bytes >= sizeof(double) * 3 * (size_t)grid, "dfmir_grad_loss_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if ((long long)planes * g.nvox < (1LL << 31)) grad_fwd_kernel<int><<<grid, 256, 0, st>>>(flow, (double*)ws, g, penalty);
  else grad_fwd_kernel<long long><<<grid, 256, 0, st>>>(flow, (double*)ws, g, penalty);
  DFMIR_CHECK_LAUNCH("dfmir_grad_loss_fwd");
  grad_finalize_kernel<<<1, 96, 0, st>>>((const double*)ws, grid, loss, g, loss_mult);
  DFMIR_CHECK_LAUNCH("dfmir_grad_loss_fwd(finalize)");
  return DFMIR_OK;
}

extern "C" int dfmir_grad_loss_bwd(const float* flow, const float* grad_loss, float* d_flow, int planes,
                                   int nd, const int* shape, int penalty, float loss_mult, void* stream) {
  GGeom g;
  DFMIR_CHECK_ARG(make_ggeom(g, planes, nd, shape) == 0, "dfmir_grad_loss_bwd: bad geometry (nd=%d)", nd);
  DFMIR_CHECK_ARG(penalty == 1 || penalty == 2, "dfmir_grad_loss_bwd: penalty must be 1 (l1) or 2 (l2)");
  DFMIR_CHECK_ARG(flow && grad_loss && d_flow, "dfmir_grad_loss_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const long long items = (long long)planes * g.nvox;
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  if (items < (1LL << 31)) grad_bwd_kernel<int><<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(flow, grad_loss, d_flow, g, penalty, loss_mult);
  else grad_bwd_kernel<long long><<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(flow, grad_loss, d_flow, g, penalty, loss_mult);
  DFMIR_CHECK_LAUNCH("dfmir_grad_loss_bwd");
  return DFMIR_OK;
}

// =====================================================================================
// masked L1  (registration_model.py:255-263; mask = (u > thr) | (v > thr), :160-161)
// =====================================================================================
namespace {

__device__ __forceinline__ float mask_at(const uint8_t* mask, const float* mu, const float* mv, float thr,
                                         long long i) {
  if (mask) return mask[i] ? 1.f : 0.f;
  if (mu) return (mu[i] > thr || mv[i] > thr) ? 1.f : 0.f;
  return 1.f;
}

__global__ void __launch_bounds__(256)
l1m_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ mask,
               const float* __restrict__ mu, const float* __restrict__ mv, float thr,
               double* __restrict__ partials, long long n) {
  __shared__ double sred[2][8];
  float s = 0.f, m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float mk = mask_at(mask, mu, mv, thr, i);
    s += fabsf(a[i] - b[i]) * mk;
    m += mk;
  }
  double ds = warp_sum_d((double)s), dm = warp_sum_d((double)m);
  if ((threadIdx.x & 31) == 0) { sred[0][threadIdx.x >> 5] = ds; sred[1][threadIdx.x >> 5] = dm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ds = 0; dm = 0;
    for (int i = 0; i < 8; ++i) { ds += sred[0][i]; dm += sred[1][i]; }
    partials[2 * (long long)blockIdx.x] = ds; partials[2 * (long long)blockIdx.x + 1] = dm;
  }
}

// out = {loss, sum(mask)}
__global__ void l1m_finalize_kernel(const double* __restrict__ partials, int n, float* __restrict__ out,
                                    int has_mask, double count) {
  if (threadIdx.x == 0) {
    double s = 0, m = 0;
    for (int i = 0; i < n; ++i) { s += partials[2 * (long long)i]; m += partials[2 * (long long)i + 1]; }
    if (!has_mask) m = count;                        // plain mean (registration_model.py:257-258)
    out[0] = m == 0.0 ? 0.f : (float)(s / m);        // empty mask -> 0 (registration_model.py:259-260)
    out[1] = (float)m;
  }
}

__global__ void __launch_bounds__(256)
l1m_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ mask,
               const float* __restrict__ mu, const float* __restrict__ mv, float thr,
               const float* __restrict__ fwd_out, const float* __restrict__ grad_loss,
               float* __restrict__ da, float* __restrict__ db, long long n) {
  const float msum = fwd_out[1];
  const float gs = msum == 0.f ? 0.f : grad_loss[0] / msum;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    const float r = gs * sg * mask_at(mask, mu, mv, thr, i);
    if (da) da[i] = r;
    if (db) db[i] = -r;
  }
}

}  // namespace

extern "C" size_t dfmir_l1_masked_workspace_bytes(void) { return sizeof(double) * 2 * 148 * 8 * 2 + 256; }

extern "C" int dfmir_l1_masked_fwd(const float* a, const float* b, const uint8_t* mask, const float* mu,
                                   const float* mv, float thr, float* out, void* ws, size_t ws_bytes,
                                   long long n, void* stream) {
  DFMIR_CHECK_ARG(a && b && out && ws && n >= 0, "dfmir_l1_masked_fwd: null pointer / bad n");
  DFMIR_CHECK_ARG((mu == nullptr) == (mv == nullptr), "dfmir_l1_masked_fwd: mu and mv must be given together");
  const int grid = red_grid(n);
  DFMIR_CHECK_ARG(ws_bytes >= sizeof(double) * 2 * (size_t)grid, "dfmir_l1_masked_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  l1m_fwd_kernel<<<grid, 256, 0, st>>>(a, b, mask, mu, mv, thr, (double*)ws, n);
  DFMIR_CHECK_LAUNCH("dfmir_l1_masked_fwd");
  l1m_finalize_kernel<<<1, 32, 0, st>>>((const double*)ws, grid, out, mask || mu, (double)n);
  DFMIR_CHECK_LAUNCH("dfmir_l1_masked_fwd(finalize)");
  return DFMIR_OK;
}

extern "C" int dfmir_l1_masked_bwd(const float* a, const float* b, const uint8_t* mask, const float* mu,
                                   const float* mv, float thr, const float* fwd_out, const float* grad_loss,
                                   float* da, float* db, long long n, void* stream) {
  DFMIR_CHECK_ARG(a && b && fwd_out && grad_loss && n >= 0, "dfmir_l1_masked_bwd: null pointer / bad n");
  if (!da && !db) return DFMIR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  if (blocks < 1) blocks = 1;
  l1m_bwd_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(a, b, mask, mu, mv, thr, fwd_out, grad_loss, da, db, n);
  DFMIR_CHECK_LAUNCH("dfmir_l1_masked_bwd");
  return DFMIR_OK;
}
