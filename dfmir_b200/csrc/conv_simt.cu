// K1/K2 (fp32 CUDA-core path): N-D convolution as an implicit GEMM, forward / data-gradient /
// weight-gradient, for every layer shape of the reference's two networks.
//
// Replaces the nn.Conv2d / nn.Conv3d calls of models/networks.py:983,995,1016,1023,1201,1214
// (ResnetGenerator) and models/voxelmorph/torchvoxelmorph/networks.py:1515,1077 (ConvBlock, flow).
// This is the exact-fp32 path: it serves the layers whose GEMM shape cannot feed a tensor core
// (Cin = 1 stem, Cout = 1 head, the 2/16/34-channel VoxelMorph layers, the nd-channel flow conv)
// and is the parity mode for all the others (conv_umma.cu is the tcgen05 path for those).
//
// Layout: activations channels-last with explicit element strides (n, d, h, w, c) so that padded
// buffers, interior views and the planar flow output need no copy; weights [tap][Cin][Cout]
// (tap = (kd*KH + kh)*KW + kw) for the forward, [tap][Cout][Cin] for the data gradient.
// GEMM view: M = N*OD*OH*OW output positions, N = Cout, K = taps*Cin.
#include "common.cuh"
#include "dfmir_b200.h"

// direct kernels for the Cin = 1 / Cout = 1 7x7 layers (conv_thin.cu); return 1 when they handled the call
int dfmir_thin_fwd(const float* x, const float* w, const float* bias, float* y, const dfmir_conv_desc* d, cudaStream_t st, int* rc);
int dfmir_thin_dgrad(const float* dy, const float* wt, float* dx, const dfmir_conv_desc* d, cudaStream_t st, int* rc);
int dfmir_thin_wgrad(const float* x, const float* dy, float* dw, float* db, const dfmir_conv_desc* d, cudaStream_t st, int* rc);

namespace {

struct ConvP {
  int N, Cin, Cout;
  int I[3], O[3], Kk[3], pad[3];
  int stride, transposed, act;
  long long xs[5];  // input element strides: n, d, h, w, c
  long long ys[5];  // output element strides
  long long M;      // N * O0 * O1 * O2
  int K;            // taps * Cin
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DFMIR_ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  if (act == DFMIR_ACT_TANH) return tanhf(v);
  if (act == DFMIR_ACT_RELU) return v > 0.f ? v : 0.f;
  return v;
}

// Source coordinate along one axis for output coordinate origin `o` and tap `t`.
//   forward:    i = o*stride - pad + t            (o pre-multiplied: origin = o*stride - pad)
//   transposed: q = o + pad - t, valid iff q % stride == 0, i = q / stride (origin = o + pad)
__device__ __forceinline__ bool src_coord(int origin, int t, int stride, int transposed, int I, int& i) {
  if (!transposed) {
    i = origin + t;
  } else {
    const int q = origin - t;
    if (q < 0) return false;
    if (stride > 1) {
      if (q % stride) return false;
      i = q / stride;
    } else {
      i = q;
    }
  }
  return i >= 0 && i < I;
}

// ------------------------------------------------------------------ forward / dgrad
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_simt_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ y, ConvP p) {
  constexpr int T = (BM / TM) * (BN / TN);
  static_assert(T % BK == 0, "thread count must be a multiple of BK");
  constexpr int A_ROWS_PER_PASS = T / BK;
  constexpr int A_PER_T = BM / A_ROWS_PER_PASS;
  constexpr int B_PER_T = (BK * BN + T - 1) / T;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ long long rowx[BM];   // input base offset (sample), -1 for rows beyond M
  __shared__ long long rowy[BM];   // output base offset
  __shared__ int roworg[BM][3];    // per-axis origin coordinate

  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  for (int r = t; r < BM; r += T) {
    const long long m = m0 + r;
    if (m < p.M) {
      long long q = m;
      const int ox = (int)(q % p.O[2]); q /= p.O[2];
      const int oy = (int)(q % p.O[1]); q /= p.O[1];
      const int oz = (int)(q % p.O[0]); q /= p.O[0];
      const int n = (int)q;
      rowx[r] = (long long)n * p.xs[0];
      rowy[r] = (long long)n * p.ys[0] + oz * p.ys[1] + oy * p.ys[2] + ox * p.ys[3];
      if (!p.transposed) {
        roworg[r][0] = oz * p.stride - p.pad[0];
        roworg[r][1] = oy * p.stride - p.pad[1];
        roworg[r][2] = ox * p.stride - p.pad[2];
      } else {
        roworg[r][0] = oz + p.pad[0];
        roworg[r][1] = oy + p.pad[1];
        roworg[r][2] = ox + p.pad[2];
      }
    } else {
      rowx[r] = -1; rowy[r] = -1;
      roworg[r][0] = roworg[r][1] = roworg[r][2] = 0;
    }
  }
  __syncthreads();

  const int ak = t % BK;      // this thread's k column of the A tile (fixed for the whole kernel)
  const int ar0 = t / BK;
  float areg[A_PER_T], breg[B_PER_T];
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int tx = t % (BN / TN), ty = t / (BN / TN);
  const int nkt = (p.K + BK - 1) / BK;

  auto load_tiles = [&](int kt) {
    const int k = kt * BK + ak;
    int tz = 0, tyy = 0, txx = 0, ci = 0;
    const bool kin = k < p.K;
    if (kin) {
      int tap = k / p.Cin;
      ci = k - tap * p.Cin;
      txx = tap % p.Kk[2]; tap /= p.Kk[2];
      tyy = tap % p.Kk[1]; tz = tap / p.Kk[1];
    }
#pragma unroll
    for (int j = 0; j < A_PER_T; ++j) {
      const int r = ar0 + j * A_ROWS_PER_PASS;
      float v = 0.f;
      const long long base = rowx[r];
      if (kin && base >= 0) {
        int iz, iy, ix;
        if (src_coord(roworg[r][0], tz, p.stride, p.transposed, p.I[0], iz) &&
            src_coord(roworg[r][1], tyy, p.stride, p.transposed, p.I[1], iy) &&
            src_coord(roworg[r][2], txx, p.stride, p.transposed, p.I[2], ix))
          v = __ldg(x + base + iz * p.xs[1] + iy * p.xs[2] + ix * p.xs[3] + ci * p.xs[4]);
      }
      areg[j] = v;
    }
#pragma unroll
    for (int j = 0; j < B_PER_T; ++j) {
      const int e = t + j * T;
      float v = 0.f;
      if (e < BK * BN) {
        const int bk = e / BN, bn = e - bk * BN;
        const int kk = kt * BK + bk, co = n0 + bn;
        if (kk < p.K && co < p.Cout) v = __ldg(w + (long long)kk * p.Cout + co);
      }
      breg[j] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int j = 0; j < A_PER_T; ++j) As[ak][ar0 + j * A_ROWS_PER_PASS] = areg[j];
#pragma unroll
    for (int j = 0; j < B_PER_T; ++j) {
      const int e = t + j * T;
      if (e < BK * BN) Bs[e / BN][e % BN] = breg[j];
    }
  };

  load_tiles(0);
  for (int kt = 0; kt < nkt; ++kt) {
    store_tiles();
    __syncthreads();
    if (kt + 1 < nkt) load_tiles(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long yb = rowy[ty * TM + i];
    if (yb < 0) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co < p.Cout) {
        float v = acc[i][j];
        if (bias) v += __ldg(bias + co);
        y[yb + co * p.ys[4]] = apply_act(v, p.act);
      }
    }
  }
}

// ------------------------------------------------------------------ weight gradient
// dW[k][co] += sum over positions of x(pos, k) * dy(pos, co); k = tap*Cin + ci.  Positions are
// split over gridDim.z; partial results are combined with fp32 atomics (dW, db pre-zeroed).
template <int BKD, int BN, int BR, int TM, int TN>
__global__ void __launch_bounds__((BKD / TM) * (BN / TN))
conv_wgrad_simt_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                       float* __restrict__ db, ConvP p, long long chunk) {
  constexpr int T = (BKD / TM) * (BN / TN);
  static_assert(T % BKD == 0 || BKD % T == 0, "tile/thread mismatch");
  __shared__ __align__(16) float Xs[BR][BKD + 4];
  __shared__ __align__(16) float Gs[BR][BN];
  __shared__ long long rowx[BR];
  __shared__ long long rowy[BR];
  __shared__ int roworg[BR][3];

  const int t = threadIdx.x;
  const int k0 = blockIdx.x * BKD;
  const int n0 = blockIdx.y * BN;
  const long long mbeg = (long long)blockIdx.z * chunk;
  const long long mend = min(p.M, mbeg + chunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bacc = 0.f;
  const int tx = t % (BN / TN), ty = t / (BN / TN);

  for (long long mb = mbeg; mb < mend; mb += BR) {
    for (int r = t; r < BR; r += T) {
      const long long m = mb + r;
      if (m < mend) {
        long long q = m;
        const int ox = (int)(q % p.O[2]); q /= p.O[2];
        const int oy = (int)(q % p.O[1]); q /= p.O[1];
        const int oz = (int)(q % p.O[0]); q /= p.O[0];
        const int n = (int)q;
        rowx[r] = (long long)n * p.xs[0];
        rowy[r] = (long long)n * p.ys[0] + oz * p.ys[1] + oy * p.ys[2] + ox * p.ys[3];
        roworg[r][0] = oz * p.stride - p.pad[0];
        roworg[r][1] = oy * p.stride - p.pad[1];
        roworg[r][2] = ox * p.stride - p.pad[2];
      } else {
        rowx[r] = -1; rowy[r] = -1;
        roworg[r][0] = roworg[r][1] = roworg[r][2] = 0;
      }
    }
    __syncthreads();
    for (int e = t; e < BR * BKD; e += T) {
      const int r = e / BKD, kk = e - r * BKD;
      const int k = k0 + kk;
      float v = 0.f;
      const long long base = rowx[r];
      if (k < p.K && base >= 0) {
        int tap = k / p.Cin;
        const int ci = k - tap * p.Cin;
        const int txx = tap % p.Kk[2]; tap /= p.Kk[2];
        const int tyy = tap % p.Kk[1]; const int tz = tap / p.Kk[1];
        const int iz = roworg[r][0] + tz, iy = roworg[r][1] + tyy, ix = roworg[r][2] + txx;
        if (iz >= 0 && iz < p.I[0] && iy >= 0 && iy < p.I[1] && ix >= 0 && ix < p.I[2])
          v = __ldg(x + base + iz * p.xs[1] + iy * p.xs[2] + ix * p.xs[3] + ci * p.xs[4]);
      }
      Xs[r][kk] = v;
    }
    for (int e = t; e < BR * BN; e += T) {
      const int r = e / BN, c = e - r * BN;
      const int co = n0 + c;
      float v = 0.f;
      const long long base = rowy[r];
      if (co < p.Cout && base >= 0) v = __ldg(dy + base + co * p.ys[4]);
      Gs[r][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = Xs[r][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Gs[r][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (db && blockIdx.x == 0 && t < BN) {
#pragma unroll
      for (int r = 0; r < BR; ++r) bacc += Gs[r][t];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int k = k0 + ty * TM + i;
    if (k >= p.K) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co < p.Cout) atomicAdd(dw + (long long)k * p.Cout + co, acc[i][j]);
    }
  }
  if (db && blockIdx.x == 0 && t < BN && n0 + t < p.Cout) atomicAdd(db + n0 + t, bacc);
}

// ------------------------------------------------------------------ weight gradient, few output channels
// Cout <= CO (the VoxelMorph full-resolution layers: 34 -> 16 and the flow head 16 -> nd).  The GEMM view is
// K x Cout with Cout tiny and M = millions of positions, so each thread owns ONE row k = (tap, ci) of dW with all
// its CO accumulators in registers and walks the positions of its chunk: one coalesced x load (consecutive
// threads = consecutive channels of a tap) feeds CO FMAs; the dy rows of a batch of positions are staged in shared
// memory and read as broadcast float4.  Position chunks over gridDim.x, combined with fp32 atomics.
template <int CO, int PB>
__global__ void __launch_bounds__(256)
conv_wgrad_fewco_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                        float* __restrict__ db, ConvP p, long long chunk) {
  __shared__ __align__(16) float gs[PB][CO];
  __shared__ long long xbase[PB];
  __shared__ int org[PB][3];
  const int t = threadIdx.x;
  const int k = blockIdx.y * 256 + t;
  const bool kin = k < p.K;
  int tz = 0, ty = 0, tx = 0, ci = 0;
  if (kin) {
    int tap = k / p.Cin;
    ci = k - tap * p.Cin;
    tx = tap % p.Kk[2]; tap /= p.Kk[2];
    ty = tap % p.Kk[1]; tz = tap / p.Kk[1];
  }
  const long long mbeg = (long long)blockIdx.x * chunk;
  const long long mend = min(p.M, mbeg + chunk);
  float acc[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) acc[j] = 0.f;
  float bacc = 0.f;
  for (long long mb = mbeg; mb < mend; mb += PB) {
    __syncthreads();
    for (int e = t; e < PB * CO; e += 256) {
      const int r = e / CO, c = e - r * CO;
      const long long m = mb + r;
      float v = 0.f;
      if (m < mend && c < p.Cout) {
        long long q = m;
        const int ox = (int)(q % p.O[2]); q /= p.O[2];
        const int oy = (int)(q % p.O[1]); q /= p.O[1];
        const int oz = (int)(q % p.O[0]); q /= p.O[0];
        v = __ldg(dy + q * p.ys[0] + oz * p.ys[1] + oy * p.ys[2] + ox * p.ys[3] + c * p.ys[4]);
      }
      gs[r][c] = v;
    }
    if (t < PB) {
      const long long m = mb + t;
      if (m < mend) {
        long long q = m;
        const int ox = (int)(q % p.O[2]); q /= p.O[2];
        const int oy = (int)(q % p.O[1]); q /= p.O[1];
        const int oz = (int)(q % p.O[0]); q /= p.O[0];
        xbase[t] = q * p.xs[0];
        org[t][0] = oz * p.stride - p.pad[0]; org[t][1] = oy * p.stride - p.pad[1]; org[t][2] = ox * p.stride - p.pad[2];
      } else {
        xbase[t] = -1; org[t][0] = org[t][1] = org[t][2] = 0;
      }
    }
    __syncthreads();
    if (kin) {
#pragma unroll 4
      for (int r = 0; r < PB; ++r) {
        const long long base = xbase[r];
        const int iz = org[r][0] + tz, iy = org[r][1] + ty, ix = org[r][2] + tx;
        float xv = 0.f;
        if (base >= 0 && iz >= 0 && iz < p.I[0] && iy >= 0 && iy < p.I[1] && ix >= 0 && ix < p.I[2])
          xv = __ldg(x + base + iz * p.xs[1] + iy * p.xs[2] + ix * p.xs[3] + ci * p.xs[4]);
        const float4* g4 = reinterpret_cast<const float4*>(gs[r]);
#pragma unroll
        for (int j = 0; j < CO / 4; ++j) {
          const float4 g = g4[j];
          acc[4 * j] = fmaf(xv, g.x, acc[4 * j]); acc[4 * j + 1] = fmaf(xv, g.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(xv, g.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(xv, g.w, acc[4 * j + 3]);
        }
      }
    }
    if (db && blockIdx.y == 0 && t < CO) {
#pragma unroll 4
      for (int r = 0; r < PB; ++r) bacc += gs[r][t];
    }
  }
  if (kin) {
#pragma unroll
    for (int j = 0; j < CO; ++j)
      if (j < p.Cout) atomicAdd(dw + (long long)k * p.Cout + j, acc[j]);
  }
  if (db && blockIdx.y == 0 && t < p.Cout && t < CO) atomicAdd(db + t, bacc);
}

// dx = dy * act'(y) on contiguous arrays (LeakyReLU(0.2) / tanh / ReLU of a conv epilogue)
__device__ __forceinline__ float act_grad(float yv, float g, int act) {
  if (act == DFMIR_ACT_LEAKY) return yv > 0.f ? g : 0.2f * g;
  if (act == DFMIR_ACT_TANH) return g * (1.f - yv * yv);
  if (act == DFMIR_ACT_RELU) return yv > 0.f ? g : 0.f;
  return g;
}
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long n,
               int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = act_grad(y[i], dy[i], act);
}
// 128-bit variant (16-byte aligned arrays): n4 float4 elements
__global__ void __launch_bounds__(256)
act_bwd_v4_kernel(const float4* __restrict__ y, const float4* __restrict__ dy, float4* __restrict__ dx, long long n4, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = y[i], g = dy[i];
    dx[i] = make_float4(act_grad(a.x, g.x, act), act_grad(a.y, g.y, act), act_grad(a.z, g.z, act), act_grad(a.w, g.w, act));
  }
}

int fill_geom(ConvP& p, const dfmir_conv_desc* d, const char* who) {
  if (!d) { dfmir_set_error("%s: null descriptor", who); return DFMIR_ERR_ARG; }
  if (d->nd != 2 && d->nd != 3) { dfmir_set_error("%s: nd must be 2 or 3, got %d", who, d->nd); return DFMIR_ERR_ARG; }
  if (d->N < 1 || d->Cin < 1 || d->Cout < 1 || d->stride < 1) {
    dfmir_set_error("%s: bad sizes N=%d Cin=%d Cout=%d stride=%d", who, d->N, d->Cin, d->Cout, d->stride);
    return DFMIR_ERR_ARG;
  }
  p.N = d->N; p.Cin = d->Cin; p.Cout = d->Cout; p.stride = d->stride; p.act = d->act; p.transposed = 0;
  const int sh = 3 - d->nd;
  p.I[0] = p.O[0] = p.Kk[0] = 1; p.pad[0] = 0;
  for (int a = 0; a < d->nd; ++a) {
    p.I[a + sh] = d->in_shape[a]; p.O[a + sh] = d->out_shape[a];
    p.Kk[a + sh] = d->kernel[a]; p.pad[a + sh] = d->pad[a];
    if (d->in_shape[a] < 1 || d->out_shape[a] < 1 || d->kernel[a] < 1 || d->pad[a] < 0) {
      dfmir_set_error("%s: bad spatial geometry on axis %d", who, a); return DFMIR_ERR_ARG;
    }
    // forward output size of a (strided, zero-padded) convolution; a smaller out_shape is the same convolution with less
    // trailing padding (pad = leading padding: the space-to-depth form of a stride-2 layer pads one side only)
    const int want = (d->in_shape[a] + 2 * d->pad[a] - d->kernel[a]) / d->stride + 1;
    if (d->out_shape[a] > want) {
      dfmir_set_error("%s: out_shape[%d]=%d inconsistent with in=%d k=%d pad=%d stride=%d (expected at most %d)", who, a,
                      d->out_shape[a], d->in_shape[a], d->kernel[a], d->pad[a], d->stride, want);
      return DFMIR_ERR_ARG;
    }
  }
  // strides arrive as (n, spatial..., c): spread to (n, d, h, w, c)
  p.xs[0] = d->x_strides[0]; p.ys[0] = d->y_strides[0];
  p.xs[1] = p.ys[1] = 0;
  for (int a = 0; a < d->nd; ++a) { p.xs[1 + a + sh] = d->x_strides[1 + a]; p.ys[1 + a + sh] = d->y_strides[1 + a]; }
  p.xs[4] = d->x_strides[1 + d->nd]; p.ys[4] = d->y_strides[1 + d->nd];
  p.M = (long long)p.N * p.O[0] * p.O[1] * p.O[2];
  p.K = p.Kk[0] * p.Kk[1] * p.Kk[2] * p.Cin;
  return DFMIR_OK;
}

template <int BM, int BN, int BK, int TM, int TN>
void launch_conv(const float* x, const float* w, const float* b, float* y, const ConvP& p, int ncol, cudaStream_t st) {
  dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((ncol + BN - 1) / BN));
  conv_simt_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(x, w, b, y, p);
}

int run_conv(const float* x, const float* w, const float* b, float* y, const ConvP& p, cudaStream_t st,
             const char* who) {
  if (p.M == 0) return DFMIR_OK;
  const int nc = p.Cout;
  if (nc > 16) launch_conv<128, 64, 16, 8, 4>(x, w, b, y, p, nc, st);
  else if (nc > 4) launch_conv<128, 16, 16, 4, 4>(x, w, b, y, p, nc, st);
  else launch_conv<256, 4, 16, 2, 4>(x, w, b, y, p, nc, st);
  DFMIR_CHECK_LAUNCH(who);
  return DFMIR_OK;
}

}  // namespace

// Forward: y = act(conv(x, w) + bias).  w: [tap][Cin][Cout].
extern "C" int dfmir_conv_fwd(const float* x, const float* w, const float* bias, float* y,
                              const dfmir_conv_desc* d, void* stream) {
  ConvP p;
  int rc = fill_geom(p, d, "dfmir_conv_fwd");
  if (rc) return rc;
  DFMIR_CHECK_ARG(x && w && y, "dfmir_conv_fwd: null pointer");
  if (p.M == 0) return DFMIR_OK;
  if (dfmir_thin_fwd(x, w, bias, y, d, (cudaStream_t)stream, &rc)) return rc;
  return run_conv(x, w, bias, y, p, (cudaStream_t)stream, "dfmir_conv_fwd");
}

// Data gradient: dx = conv_transpose(dy, w).  wt: [tap][Cout][Cin] (per-tap transpose of the forward
// weights).  The descriptor is the FORWARD descriptor; x_strides describe dx, y_strides describe dy.
extern "C" int dfmir_conv_dgrad(const float* dy, const float* wt, float* dx, const dfmir_conv_desc* d,
                                void* stream) {
  ConvP f;
  int rc = fill_geom(f, d, "dfmir_conv_dgrad");
  if (rc) return rc;
  DFMIR_CHECK_ARG(dy && wt && dx, "dfmir_conv_dgrad: null pointer");
  if (f.M == 0) return DFMIR_OK;
  if (dfmir_thin_dgrad(dy, wt, dx, d, (cudaStream_t)stream, &rc)) return rc;
  ConvP p = f;
  p.transposed = 1; p.act = DFMIR_ACT_NONE;
  p.Cin = f.Cout; p.Cout = f.Cin;
  for (int a = 0; a < 3; ++a) { p.I[a] = f.O[a]; p.O[a] = f.I[a]; }
  for (int a = 0; a < 5; ++a) { p.xs[a] = f.ys[a]; p.ys[a] = f.xs[a]; }
  p.M = (long long)p.N * p.O[0] * p.O[1] * p.O[2];
  p.K = p.Kk[0] * p.Kk[1] * p.Kk[2] * p.Cin;
  return run_conv(dy, wt, nullptr, dx, p, (cudaStream_t)stream, "dfmir_conv_dgrad");
}

// Weight / bias gradient.  dw: [tap][Cin][Cout] and db: [Cout] (nullable) are ACCUMULATED into
// (the caller zero-fills them).
extern "C" int dfmir_conv_wgrad(const float* x, const float* dy, float* dw, float* db,
                                const dfmir_conv_desc* d, void* stream) {
  ConvP p;
  int rc = fill_geom(p, d, "dfmir_conv_wgrad");
  if (rc) return rc;
  DFMIR_CHECK_ARG(x && dy && dw, "dfmir_conv_wgrad: null pointer");
  if (p.M == 0) return DFMIR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dfmir_thin_wgrad(x, dy, dw, db, d, st, &rc)) return rc;
  if (p.Cout <= 16 && p.M >= 4096) {
    // few output channels, many positions: one dW row per thread (see conv_wgrad_fewco_kernel)
    const int gy = (p.K + 255) / 256;
    long long blocks = (8LL * dfmir_num_sms() + gy - 1) / gy;
    long long chunk = (p.M + blocks - 1) / blocks;
    chunk = (chunk + 31) / 32 * 32;
    if (chunk < 256) chunk = 256;
    const unsigned gxf = (unsigned)((p.M + chunk - 1) / chunk);
    if (p.Cout <= 4) conv_wgrad_fewco_kernel<4, 32><<<dim3(gxf, gy), 256, 0, st>>>(x, dy, dw, db, p, chunk);
    else conv_wgrad_fewco_kernel<16, 32><<<dim3(gxf, gy), 256, 0, st>>>(x, dy, dw, db, p, chunk);
    DFMIR_CHECK_LAUNCH("dfmir_conv_wgrad(few output channels)");
    return DFMIR_OK;
  }
  constexpr int BKD = 64, BR = 16;
  const int gx = (p.K + BKD - 1) / BKD;
  auto splits_for = [&](int gy) {
    long long want = (4LL * dfmir_num_sms() + (long long)gx * gy - 1) / ((long long)gx * gy);
    long long maxs = (p.M + 4 * BR - 1) / (4 * BR);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    return (int)want;
  };
  if (p.Cout > 16) {
    const int gy = (p.Cout + 63) / 64, gz = splits_for(gy);
    long long chunk = (p.M + gz - 1) / gz; chunk = (chunk + BR - 1) / BR * BR;
    conv_wgrad_simt_kernel<BKD, 64, BR, 4, 4><<<dim3(gx, gy, (unsigned)((p.M + chunk - 1) / chunk)), 256, 0, st>>>(x, dy, dw, db, p, chunk);
  } else if (p.Cout > 4) {
    const int gy = (p.Cout + 15) / 16, gz = splits_for(gy);
    long long chunk = (p.M + gz - 1) / gz; chunk = (chunk + BR - 1) / BR * BR;
    conv_wgrad_simt_kernel<BKD, 16, BR, 2, 2><<<dim3(gx, gy, (unsigned)((p.M + chunk - 1) / chunk)), 256, 0, st>>>(x, dy, dw, db, p, chunk);
  } else {
    const int gy = (p.Cout + 3) / 4, gz = splits_for(gy);
    long long chunk = (p.M + gz - 1) / gz; chunk = (chunk + BR - 1) / BR * BR;
    conv_wgrad_simt_kernel<BKD, 4, BR, 1, 1><<<dim3(gx, gy, (unsigned)((p.M + chunk - 1) / chunk)), 256, 0, st>>>(x, dy, dw, db, p, chunk);
  }
  DFMIR_CHECK_LAUNCH("dfmir_conv_wgrad");
  return DFMIR_OK;
}

extern "C" int dfmir_act_bwd(const float* y, const float* dy, float* dx, long long n, int act, void* stream) {
  DFMIR_CHECK_ARG(y && dy && dx && n >= 0, "dfmir_act_bwd: null pointer / bad n");
  DFMIR_CHECK_ARG(act >= DFMIR_ACT_NONE && act <= DFMIR_ACT_RELU, "dfmir_act_bwd: unknown activation %d", act);
  if (n == 0) return DFMIR_OK;
  const long long cap = (long long)dfmir_num_sms() * 16;
  if ((n & 3) == 0 && ((((uintptr_t)y) | ((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0) {
    const long long blocks = (n / 4 + 255) / 256;
    act_bwd_v4_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>((const float4*)y, (const float4*)dy, (float4*)dx, n / 4, act);
  } else {
    const long long blocks = (n + 255) / 256;
    act_bwd_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, n, act);
  }
  DFMIR_CHECK_LAUNCH("dfmir_act_bwd");
  return DFMIR_OK;
}
