// K6 (fp32 CUDA-core path): PatchNCE contrastive loss and the small dense products around it.
//
// Replaces models/patchnce.py:14-55 (two torch.bmm, masked_fill, cat, CrossEntropyLoss) and the
// nn.Linear layers of PatchSampleF (models/networks.py:587-595, 617) of the reference.
//   logits row i of image b:  [ q_i.k_i , q_i.k_0 , ... , (-10 at j == i) , ... ] / T ; loss = CE vs class 0
// Since l_pos is the diagonal of the per-image product S = Q K^T, one batched GEMM feeds a row-wise
// log-softmax kernel which also leaves dLoss/dS in place for the backward product dQ = dS K.
#include "common.cuh"
#include "dfmir_b200.h"

namespace {

struct GemmP {
  int M, N, K, batch;
  long long sAb, sAm, sAk;
  long long sBb, sBk, sBn;
  long long sCb, sCm, sCn;
  float alpha;
  int accumulate, relu;
  int ksplit;   // > 1: blockIdx.z = b * ksplit + slice; partial sums are combined with atomics (C pre-zeroed)
};

// C[b](m,n) (+)= alpha * sum_k A[b](m,k) B[b](k,n) (+ bias[n]) ; optional ReLU.  64x64x16 tiles, 4x4 per thread.
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ bias,
                 float* __restrict__ C, GemmP p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int b = blockIdx.z / p.ksplit, ks = blockIdx.z - b * p.ksplit;
  const int klen = ((p.K + p.ksplit - 1) / p.ksplit + BK - 1) / BK * BK;
  const int kbeg = ks * klen, kend = min(p.K, kbeg + klen);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const float* Ab = A + b * p.sAb;
  const float* Bb = B + b * p.sBb;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tx = t % 16, ty = t / 16;
  // pick the load mapping that walks the contiguous axis of each operand
  const bool a_kfast = p.sAk == 1, b_nfast = p.sBn == 1;
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = t + j * 256;
      int kk, mm;
      if (a_kfast) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < p.M && k < kend) ? __ldg(Ab + m * p.sAm + k * p.sAk) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = t + j * 256;
      int kk, nn;
      if (b_nfast) { nn = e % BN; kk = e / BN; } else { kk = e % BK; nn = e / BK; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < p.N && k < kend) ? __ldg(Bb + k * p.sBk + n * p.sBn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Cb = C + b * p.sCb;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = p.alpha * acc[i][j];
      if (bias) v += __ldg(bias + n);
      float* dst = Cb + m * p.sCm + n * p.sCn;
      if (p.ksplit > 1) { atomicAdd(dst, v); continue; }
      if (p.accumulate) v += *dst;
      if (p.relu) v = v > 0.f ? v : 0.f;
      *dst = v;
    }
  }
}

// One warp per row of S (B, P, P).  Writes loss[row] and overwrites S with dLoss/dS for a unit
// upstream gradient: (softmax - onehot(class 0)) / T, the class-0 (positive) term living on the
// diagonal, the masked negative at j == i contributing nothing (masked_fill_ has no gradient).
__global__ void __launch_bounds__(256)
patchnce_rows_kernel(float* __restrict__ S, float* __restrict__ loss, int rows, int P, float invT) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int i = row % P;
  float* s = S + (long long)row * P;
  const float pos = s[i] * invT;
  const float diag = -10.0f * invT;
  float mx = fmaxf(pos, diag);
  for (int j = lane; j < P; j += 32)
    if (j != i) mx = fmaxf(mx, s[j] * invT);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < P; j += 32)
    if (j != i) sum += expf(s[j] * invT - mx);
  sum = warp_sum(sum);
  sum += expf(pos - mx) + expf(diag - mx);
  const float lse = mx + logf(sum);
  if (lane == 0) loss[row] = lse - pos;
  for (int j = lane; j < P; j += 32) {
    const float v = j == i ? (expf(pos - lse) - 1.0f) : expf(s[j] * invT - lse);
    s[j] = v * invT;
  }
}

// G[row][:] *= g[row]
__global__ void __launch_bounds__(256)
row_scale_kernel(const float* __restrict__ G, const float* __restrict__ g, float* __restrict__ out, long long rows,
                 int P) {
  const long long total = rows * P;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
    out[e] = G[e] * g[e / P];
}

// y[r][:] = x[r][:] / (||x[r]||_2 + 1e-7)  (models/networks.py:493-502), one warp per row
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norms, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (long long)row * D;
  float ss = 0.f;
  for (int j = lane; j < D; j += 32) ss += xr[j] * xr[j];
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float inv = 1.0f / (nrm + 1e-7f);
  for (int j = lane; j < D; j += 32) y[(long long)row * D + j] = xr[j] * inv;
  if (lane == 0) norms[row] = nrm;
}

// dx = dy / (n + eps) - x * (dy . x) / (n * (n + eps)^2)
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ norms, const float* __restrict__ dy,
                  float* __restrict__ dx, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (long long)row * D;
  const float* gr = dy + (long long)row * D;
  float dot = 0.f;
  for (int j = lane; j < D; j += 32) dot += xr[j] * gr[j];
  dot = warp_sum(dot);
  const float nrm = norms[row], d = nrm + 1e-7f;
  const float c = nrm > 0.f ? dot / (nrm * d * d) : 0.f;
  for (int j = lane; j < D; j += 32) dx[(long long)row * D + j] = gr[j] / d - xr[j] * c;
}

// out[(b*P + i), :] = feat[b, ids[i], :] for a channels-last feature map given by element strides
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ feat, const long long* __restrict__ ids, float* __restrict__ out, int B,
                   int P, int C, int Wd, long long sN, long long sH, long long sW, long long sC) {
  const long long total = (long long)B * P * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long r = e / C;
    const int i = (int)(r % P), b = (int)(r / P);
    const long long id = ids[i];
    const long long h = id / Wd, w = id - h * Wd;
    out[e] = feat[b * sN + h * sH + w * sW + c * sC];
  }
}

// dfeat[b, ids[i], :] += dout[(b*P + i), :]   (ids are distinct within one draw; atomics keep it safe anyway)
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ dout, const long long* __restrict__ ids, float* __restrict__ dfeat,
                    int B, int P, int C, int Wd, long long sN, long long sH, long long sW, long long sC) {
  const long long total = (long long)B * P * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long r = e / C;
    const int i = (int)(r % P), b = (int)(r / P);
    const long long id = ids[i];
    const long long h = id / Wd, w = id - h * Wd;
    atomicAdd(dfeat + b * sN + h * sH + w * sW + c * sC, dout[e]);
  }
}

int launch_gemm(const float* A, const float* B, const float* bias, float* C, GemmP p, cudaStream_t st,
                const char* who) {
  if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return DFMIR_OK;
  const int tiles = ((p.M + 63) / 64) * ((p.N + 63) / 64) * p.batch;
  p.ksplit = 1;
  // small outputs with a long reduction (the weight gradients of the MLP: K = B*P rows): split K over
  // the SMs, partial sums combined with fp32 atomics into a zero-filled dense C
  const bool dense_c = p.sCn == 1 && p.sCm == p.N && (p.batch == 1 || p.sCb == (long long)p.M * p.N);
  if (!bias && !p.relu && dense_c && tiles < dfmir_num_sms() && p.K >= 512) {
    int ks = (2 * dfmir_num_sms() + tiles - 1) / tiles;
    if (ks > p.K / 128) ks = p.K / 128;
    if (ks > 1) {
      p.ksplit = ks;
      if (!p.accumulate) DFMIR_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)p.batch * p.M * p.N, st));
    }
  }
  dim3 grid((p.M + 63) / 64, (p.N + 63) / 64, p.batch * p.ksplit);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(A, B, bias, C, p);
  DFMIR_CHECK_LAUNCH(who);
  return DFMIR_OK;
}

}  // namespace

// Generic strided batched product: C[b](m,n) (+)= alpha * sum_k A[b](m,k) B[b](k,n) (+ bias[n]) (ReLU).
// strides: {batch, row, col} in elements for A (m,k), B (k,n), C (m,n).
extern "C" int dfmir_gemm(const float* A, const float* B, const float* bias, float* C, int batch, int M, int N, int K,
                          const long long* sA, const long long* sB, const long long* sC, float alpha, int accumulate,
                          int relu, void* stream) {
  DFMIR_CHECK_ARG(A && B && C && sA && sB && sC, "dfmir_gemm: null pointer");
  DFMIR_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "dfmir_gemm: negative size");
  GemmP p{M, N, K, batch, sA[0], sA[1], sA[2], sB[0], sB[1], sB[2], sC[0], sC[1], sC[2], alpha, accumulate, relu, 1};
  return launch_gemm(A, B, bias, C, p, (cudaStream_t)stream, "dfmir_gemm");
}

// q, k: (B*P, D) row-major, rows grouped per image.  S: scratch / saved (B, P, P).  loss: (B*P).
extern "C" int dfmir_patchnce_fwd(const float* q, const float* k, float* S, float* loss, int B, int P, int D, float T,
                                  void* stream) {
  DFMIR_CHECK_ARG(q && k && S && loss, "dfmir_patchnce_fwd: null pointer");
  DFMIR_CHECK_ARG(B > 0 && P > 0 && D > 0 && T > 0.f, "dfmir_patchnce_fwd: bad sizes (B=%d P=%d D=%d T=%g)", B, P, D, T);
  cudaStream_t st = (cudaStream_t)stream;
  GemmP p{P, P, D, B, (long long)P * D, D, 1, (long long)P * D, 1, D, (long long)P * P, P, 1, 1.0f, 0, 0, 1};
  int rc = launch_gemm(q, k, nullptr, S, p, st, "dfmir_patchnce_fwd(gemm)");
  if (rc) return rc;
  const int rows = B * P;
  patchnce_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(S, loss, rows, P, 1.0f / T);
  DFMIR_CHECK_LAUNCH("dfmir_patchnce_fwd(rows)");
  return DFMIR_OK;
}

// dq = diag(g) dS k ; S as left by the forward; work: (B, P, P) scratch; g: (B*P) upstream gradient.
extern "C" int dfmir_patchnce_bwd(const float* S, const float* k, const float* g, float* work, float* dq, int B, int P,
                                  int D, void* stream) {
  DFMIR_CHECK_ARG(S && k && g && work && dq, "dfmir_patchnce_bwd: null pointer");
  DFMIR_CHECK_ARG(B > 0 && P > 0 && D > 0, "dfmir_patchnce_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)B * P;
  long long blocks = (rows * P + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  row_scale_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(S, g, work, rows, P);
  DFMIR_CHECK_LAUNCH("dfmir_patchnce_bwd(scale)");
  GemmP p{P, D, P, B, (long long)P * P, P, 1, (long long)P * D, D, 1, (long long)P * D, D, 1, 1.0f, 0, 0, 1};
  return launch_gemm(work, k, nullptr, dq, p, st, "dfmir_patchnce_bwd(gemm)");
}

extern "C" int dfmir_l2norm_fwd(const float* x, float* y, float* norms, int rows, int D, void* stream) {
  DFMIR_CHECK_ARG(x && y && norms && rows >= 0 && D > 0, "dfmir_l2norm_fwd: bad argument");
  if (rows == 0) return DFMIR_OK;
  l2norm_fwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, y, norms, rows, D);
  DFMIR_CHECK_LAUNCH("dfmir_l2norm_fwd");
  return DFMIR_OK;
}

extern "C" int dfmir_l2norm_bwd(const float* x, const float* norms, const float* dy, float* dx, int rows, int D,
                                void* stream) {
  DFMIR_CHECK_ARG(x && norms && dy && dx && rows >= 0 && D > 0, "dfmir_l2norm_bwd: bad argument");
  if (rows == 0) return DFMIR_OK;
  l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, norms, dy, dx, rows, D);
  DFMIR_CHECK_LAUNCH("dfmir_l2norm_bwd");
  return DFMIR_OK;
}

// PatchSampleF gather (models/networks.py:603-611): feat given by element strides {n, h, w, c}; ids int64 (P).
extern "C" int dfmir_gather_patches_fwd(const float* feat, const long long* ids, float* out, int B, int P, int C, int Wd,
                                        const long long* strides, void* stream) {
  DFMIR_CHECK_ARG(feat && ids && out && strides, "dfmir_gather_patches_fwd: null pointer");
  DFMIR_CHECK_ARG(B > 0 && P > 0 && C > 0 && Wd > 0, "dfmir_gather_patches_fwd: bad sizes");
  const long long total = (long long)B * P * C;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  gather_rows_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(
      feat, ids, out, B, P, C, Wd, strides[0], strides[1], strides[2], strides[3]);
  DFMIR_CHECK_LAUNCH("dfmir_gather_patches_fwd");
  return DFMIR_OK;
}

// dfeat must be zero-filled by the caller.
extern "C" int dfmir_gather_patches_bwd(const float* dout, const long long* ids, float* dfeat, int B, int P, int C,
                                        int Wd, const long long* strides, void* stream) {
  DFMIR_CHECK_ARG(dout && ids && dfeat && strides, "dfmir_gather_patches_bwd: null pointer");
  DFMIR_CHECK_ARG(B > 0 && P > 0 && C > 0 && Wd > 0, "dfmir_gather_patches_bwd: bad sizes");
  const long long total = (long long)B * P * C;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  scatter_rows_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(
      dout, ids, dfeat, B, P, C, Wd, strides[0], strides[1], strides[2], strides[3]);
  DFMIR_CHECK_LAUNCH("dfmir_gather_patches_bwd");
  return DFMIR_OK;
}

// ---- tensor-core variant of the two PatchNCE products (conv_umma.cu: dfmir_bmm_nt_umma on 3xTF32-split operands)
extern "C" int dfmir_bmm_nt_umma(const float* A, const float* B, float* C, int batch, int M, int N, int K, void* stream);

// q3 (B*P, 3D) a-style split of q, k3 (B*P, 3D) b-style split of k (dfmir_tf32_split3) -> S, loss as dfmir_patchnce_fwd
extern "C" int dfmir_patchnce_tc_fwd(const float* q3, const float* k3, float* S, float* loss, int B, int P, int D3, float T,
                                     void* stream) {
  DFMIR_CHECK_ARG(q3 && k3 && S && loss, "dfmir_patchnce_tc_fwd: null pointer");
  DFMIR_CHECK_ARG(B > 0 && P > 0 && D3 > 0 && T > 0.f, "dfmir_patchnce_tc_fwd: bad sizes");
  int rc = dfmir_bmm_nt_umma(q3, k3, S, B, P, P, D3, stream);
  if (rc) return rc;
  const int rows = B * P;
  patchnce_rows_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(S, loss, rows, P, 1.0f / T);
  DFMIR_CHECK_LAUNCH("dfmir_patchnce_tc_fwd(rows)");
  return DFMIR_OK;
}

// work[row][:] = S[row][:] * g[row]  (first half of dfmir_patchnce_bwd; the product dq = work * k then runs on the
// tensor-core kernel as dfmir_bmm_nt_umma(work3, kT3))
extern "C" int dfmir_patchnce_scale(const float* S, const float* g, float* work, int B, int P, void* stream) {
  DFMIR_CHECK_ARG(S && g && work && B > 0 && P > 0, "dfmir_patchnce_scale: bad argument");
  const long long rows = (long long)B * P;
  long long blocks = (rows * P + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  row_scale_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(S, g, work, rows, P);
  DFMIR_CHECK_LAUNCH("dfmir_patchnce_scale");
  return DFMIR_OK;
}
