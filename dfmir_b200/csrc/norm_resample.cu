// Memory-bound layers around the convolutions of the translation net and the registration U-Net,
// channels-last fp32:
//   InstanceNorm2d(affine=False) [+ ReLU] [+ residual add] [+ ReflectionPad2d of the result]
//       models/networks.py:984,996,1020 (norm + ReLU), :1193-1214 (ResnetBlock: pad, norm, skip add :1220)
//   ReflectionPad2d                     models/networks.py:982,1022
//   Downsample (anti-aliased blur-pool) models/networks.py:37-60   (reflect pad 1, [1,2,1]^2/16, stride 2)
//   Upsample   (anti-aliased blur-up)   models/networks.py:73-93   (replicate pad, conv_transpose [1,3,3,1]^2/16, crop)
//   nn.Upsample(nearest x2) + torch.cat models/voxelmorph/torchvoxelmorph/networks.py:99-102
// Each is one pass over HBM (stats: one read; apply: one read + one write); the normalised tensor
// is written directly in the padded layout the next convolution's loader wants.
#include "common.cuh"
#include "dfmir_b200.h"

namespace {

// i -> (i0, i1, i2, i3) for extents (d0, d1, d2, *), i0 fastest.  32-bit divisions when the element count allows: a 64-bit
// divide by a run-time divisor costs ~100 instructions, and these kernels move only 4 - 16 bytes per thread.
__device__ __forceinline__ void unravel4(long long i, bool small, int d0, int d1, int d2, int& i0, int& i1, int& i2, int& i3) {
  if (small) {
    unsigned q = (unsigned)i, t;
    t = q / (unsigned)d0; i0 = (int)(q - t * (unsigned)d0); q = t;
    t = q / (unsigned)d1; i1 = (int)(q - t * (unsigned)d1); q = t;
    t = q / (unsigned)d2; i2 = (int)(q - t * (unsigned)d2); i3 = (int)t;
  } else {
    long long q = i;
    i0 = (int)(q % d0); q /= d0;
    i1 = (int)(q % d1); q /= d1;
    i2 = (int)(q % d2); i3 = (int)(q / d2);
  }
}


__device__ __forceinline__ int reflect_idx(int i, int n) {  // ReflectionPad: no edge repeat
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

inline int ew_grid(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------ instance norm: statistics
// sums[(n*C + c)*2 + {0,1}] += sum x, sum x^2 over a chunk of pixels (fp64 accumulation).
__global__ void __launch_bounds__(256)
in_stats_kernel(const float* __restrict__ x, double* __restrict__ sums, int HW, int C, int chunk) {
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
  const float* xb = x + (long long)n * HW * C;
  // thread -> channel (fastest), pixel lane
  const int cl = C < 256 ? C : 256;
  const int c_lane = threadIdx.x % cl, p_lane = threadIdx.x / cl, pl = 256 / cl;
  if (p_lane >= pl) return;
  for (int c = c_lane; c < C; c += cl) {
    double s = 0, ss = 0;
    for (int p = p0 + p_lane; p < p1; p += pl) {
      const float v = xb[(long long)p * C + c];
      s += (double)v; ss += (double)v * (double)v;
    }
    atomicAdd(sums + ((long long)n * C + c) * 2, s);
    atomicAdd(sums + ((long long)n * C + c) * 2 + 1, ss);
  }
}

// sums[i] = sum over chunks of partials[chunk][i]  (i < 2*N*C), chunks summed in a fixed order (deterministic)
__global__ void in_sum_partials_kernel(const double* __restrict__ partials, double* __restrict__ sums, int n2, int nch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  double a = 0;
  for (int c = 0; c < nch; ++c) a += partials[(long long)c * n2 + i];
  sums[i] = a;
}

// stats[(n*C + c)*2] = mean, [..+1] = 1/sqrt(var + eps)  (biased variance, as F.instance_norm)
__global__ void in_finalize_kernel(const double* __restrict__ sums, float* __restrict__ stats, int NC, int HW,
                                   float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const double m = sums[2 * i] / HW;
  double var = sums[2 * i + 1] / HW - m * m;
  if (var < 0) var = 0;
  stats[2 * i] = (float)m;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// Statistics from the per-row-group sums a convolution epilogue wrote (dfmir_conv_umma_fwd_stats):
// rows [n][rows_per_image][C] of {sum, sum of squares} over <= 32 voxels each -> stats (mean, rstd), summed in fp64 in a
// fixed order.  grid (C / 32, N); thread = (channel of the 32-channel group, one of 8 row lanes).
__global__ void __launch_bounds__(1024)
in_finalize_rows_kernel(const float2* __restrict__ rows, float* __restrict__ stats, int rows_per_image, int C, int HW, float eps) {
  // 32 channels x 32 row lanes: with 128 rows per image a lane adds four independent 8-byte loads (the 8-lane version
  // walked 16 dependent-latency trips: 19 us per launch for 8 MB)
  __shared__ double red[32][32][2];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl, n = blockIdx.y;
  double s = 0, ss = 0;
  if (c < C) {
    const float2* rb = rows + (long long)n * rows_per_image * C + c;
    int r = rl;
    for (; r + 96 < rows_per_image; r += 128) {
      const float2 v0 = rb[(long long)r * C], v1 = rb[(long long)(r + 32) * C], v2 = rb[(long long)(r + 64) * C], v3 = rb[(long long)(r + 96) * C];
      s += ((double)v0.x + (double)v1.x) + ((double)v2.x + (double)v3.x);
      ss += ((double)v0.y + (double)v1.y) + ((double)v2.y + (double)v3.y);
    }
    for (; r < rows_per_image; r += 32) {
      const float2 v = rb[(long long)r * C];
      s += (double)v.x; ss += (double)v.y;
    }
  }
  red[rl][cl][0] = s; red[rl][cl][1] = ss;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int l = 1; l < 32; ++l) { s += red[l][cl][0]; ss += red[l][cl][1]; }
    const double m = s / HW;
    double var = ss / HW - m * m;
    if (var < 0) var = 0;
    stats[((long long)n * C + c) * 2] = (float)m;
    stats[((long long)n * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// y[n, hp, wp, c] = act((x[n,h,w,c] - mean) * rstd) (+ res[n, h+rp, w+rp, c]); (h,w) = reflect(hp-p, wp-p)
__global__ void __launch_bounds__(256)
in_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ res,
                float* __restrict__ y, int N, int H, int W, int C, int relu, int p, int rp) {
  const int HP = H + 2 * p, WP = W + 2 * p;
  const long long total = (long long)N * HP * WP * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, wp, hp, n;
    unravel4(i, total < (1LL << 32), C, WP, HP, c, wp, hp, n);
    const int h = reflect_idx(hp - p, H), w = reflect_idx(wp - p, W);
    const float mean = __ldg(stats + ((long long)n * C + c) * 2), rstd = __ldg(stats + ((long long)n * C + c) * 2 + 1);
    float v = (x[(((long long)n * H + h) * W + w) * C + c] - mean) * rstd;
    if (relu) v = v > 0.f ? v : 0.f;
    if (res) v += res[(((long long)n * (H + 2 * rp) + h + rp) * (W + 2 * rp) + w + rp) * C + c];
    y[i] = v;
  }
}

// Sum of the padded-gradient entries that alias interior pixel (h,w): the adjoint of reflect_idx.
__device__ __forceinline__ int fold_list(int h, int H, int p, int* out) {
  int n = 0;
  out[n++] = h + p;
  if (p > 0) {
    if (h >= 1 && h <= p) out[n++] = p - h;
    if (h <= H - 2 && h >= H - 1 - p) out[n++] = 2 * (H - 1) - h + p;
  }
  return n;
}

// pass 1: g = fold(dy) (* relu mask); store g in dx (and in the interior of dres); accumulate
// sum g, sum g*xhat per (n,c).
__global__ void __launch_bounds__(256)
in_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                     float* __restrict__ dx, float* __restrict__ dres, double* __restrict__ sums, int H, int W, int C,
                     int relu, int p, int rp, int chunk) {
  const int n = blockIdx.y;
  const int HW = H * W;
  const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
  const int HP = H + 2 * p, WP = W + 2 * p;
  const int cl = C < 256 ? C : 256;
  const int c_lane = threadIdx.x % cl, p_lane = threadIdx.x / cl, pl = 256 / cl;
  if (p_lane >= pl) return;
  for (int c = c_lane; c < C; c += cl) {
    const float mean = stats[((long long)n * C + c) * 2], rstd = stats[((long long)n * C + c) * 2 + 1];
    double s1 = 0, s2 = 0;
    for (int px = p0 + p_lane; px < p1; px += pl) {
      const int h = px / W, w = px - h * W;
      int hl[3], wl[3];
      const int nh = fold_list(h, H, p, hl), nw = fold_list(w, W, p, wl);
      float g = 0.f;
      for (int a = 0; a < nh; ++a)
        for (int b = 0; b < nw; ++b) g += dy[(((long long)n * HP + hl[a]) * WP + wl[b]) * C + c];
      const long long o = ((long long)n * HW + px) * C + c;
      if (dres) dres[(((long long)n * (H + 2 * rp) + h + rp) * (W + 2 * rp) + w + rp) * C + c] = g;
      const float xh = (x[o] - mean) * rstd;
      if (relu && !(xh > 0.f)) g = 0.f;
      dx[o] = g;
      s1 += (double)g; s2 += (double)g * (double)xh;
    }
    atomicAdd(sums + ((long long)n * C + c) * 2, s1);
    atomicAdd(sums + ((long long)n * C + c) * 2 + 1, s2);
  }
}

// pass 2: dx = rstd * (g - mean(g) - xhat * mean(g * xhat))
__global__ void __launch_bounds__(256)
in_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats, const double* __restrict__ sums,
                    float* __restrict__ dx, int N, int HW, int C) {
  const long long total = (long long)N * HW * C;
  const double inv = 1.0 / HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((long long)HW * C));
    const long long sc = ((long long)n * C + c) * 2;
    const float mean = __ldg(stats + sc), rstd = __ldg(stats + sc + 1);
    const float m1 = (float)(sums[sc] * inv), m2 = (float)(sums[sc + 1] * inv);
    const float xh = (x[i] - mean) * rstd;
    dx[i] = rstd * (dx[i] - m1 - xh * m2);
  }
}

// ------------------------------------------------------------------ instance norm, 128-bit path (C % 4 == 0)
// Same arithmetic as the scalar kernels above, four channels per thread: every global access is a
// coalesced float4, pixel coordinates come from the grid (no per-element div/mod chains).
__device__ __forceinline__ void acc4(float* a, const float4 v) { a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w; }

__global__ void __launch_bounds__(256)
in_stats_v4_kernel(const float4* __restrict__ x, double* __restrict__ sums, int HW, int C4, int chunk) {
  __shared__ double red[256][8];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4, npl = 256 / C4;
  double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  if (pl < npl) {
    const float4* xb = x + (long long)n * HW * C4 + c4;
    int px = p0 + pl;
    while (px < p1) {
      float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < 4 && px < p1; ++k, px += 4 * npl) {      // 4 x 4 pixels in fp32, then into fp64
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (px + u * npl < p1) ? xb[(long long)(px + u * npl) * C4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          fs[0] += v[u].x; fs[1] += v[u].y; fs[2] += v[u].z; fs[3] += v[u].w;
          fq[0] = fmaf(v[u].x, v[u].x, fq[0]); fq[1] = fmaf(v[u].y, v[u].y, fq[1]);
          fq[2] = fmaf(v[u].z, v[u].z, fq[2]); fq[3] = fmaf(v[u].w, v[u].w, fq[3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[j] += (double)fs[j]; ss[j] += (double)fq[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[threadIdx.x][2 * j] = s[j]; red[threadIdx.x][2 * j + 1] = ss[j]; }
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < npl; ++l)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[threadIdx.x][j] += red[l * C4 + c4][j];
    // per-CTA partial sums (no atomics, no zero-fill): in_sum_partials_kernel adds the gridDim.x chunks
    double* o = sums + (((long long)blockIdx.x * gridDim.y + n) * C4 * 4 + c4 * 4) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = red[threadIdx.x][j];
  }
}

// grid (chunks, N); thread = (channel quad c4, pixel lane): the (mean, rstd) of its four channels are
// loaded once, then it walks the padded pixels of its chunk four at a time (four loads in flight).
__global__ void __launch_bounds__(256)
in_apply_v4_kernel(const float4* __restrict__ x, const float4* __restrict__ stats, const float4* __restrict__ res,
                   float4* __restrict__ y, int H, int W, int C4, int relu, int p, int rp, int chunk) {
  const int HP = H + 2 * p, WP = W + 2 * p, RW = W + 2 * rp;
  const int total = HP * WP;
  // images in descending order: the producing convolution wrote image N - 1 last (those lines are still in L2), and the
  // consuming convolution starts with image 0, which this kernel then writes last
  const int n = gridDim.y - 1 - blockIdx.y;
  const int p0 = blockIdx.x * chunk, p1 = min(total, p0 + chunk);
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4, npl = 256 / C4;
  if (pl >= npl) return;
  const float4 s0 = __ldg(stats + ((long long)n * C4 + c4) * 2), s1 = __ldg(stats + ((long long)n * C4 + c4) * 2 + 1);
  const float4* xb = x + (long long)n * H * W * C4 + c4;
  const float4* rb = res ? res + (long long)n * (H + 2 * rp) * RW * C4 + c4 : nullptr;
  float4* yb = y + (long long)n * total * C4 + c4;
  for (int pp = p0 + pl; pp < p1; pp += 4 * npl) {
    float4 v[4], r[4];
    int ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int q = pp + k * npl;
      ok[k] = q < p1;
      if (ok[k]) {
        const int hp = q / WP, wp = q - hp * WP;
        const int h = reflect_idx(hp - p, H), w = reflect_idx(wp - p, W);
        v[k] = xb[((long long)h * W + w) * C4];
        if (rb) r[k] = rb[((long long)(h + rp) * RW + w + rp) * C4];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      float4 t = v[k];
      t.x = (t.x - s0.x) * s0.y; t.y = (t.y - s0.z) * s0.w; t.z = (t.z - s1.x) * s1.y; t.w = (t.w - s1.z) * s1.w;
      if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
      if (rb) { t.x += r[k].x; t.y += r[k].y; t.z += r[k].z; t.w += r[k].w; }
      yb[(long long)(pp + k * npl) * C4] = t;
    }
  }
}

__global__ void __launch_bounds__(256)
in_bwd_reduce_v4_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ stats,
                        float4* __restrict__ dx, float4* __restrict__ dres, double* __restrict__ sums, int H, int W, int C4,
                        int relu, int p, int rp, int chunk) {
  __shared__ double red[256][8];
  // descending image order (the tail of the data-gradient convolution's output is still in L2); the apply pass then
  // runs ascending and starts on the images this kernel wrote last
  const int n = gridDim.y - 1 - blockIdx.y;
  const int HW = H * W;
  const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
  const int HP = H + 2 * p, WP = W + 2 * p;
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4, npl = 256 / C4;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (pl < npl) {
    const float4 s0 = __ldg(stats + ((long long)n * C4 + c4) * 2), s1 = __ldg(stats + ((long long)n * C4 + c4) * 2 + 1);
    const float4* dyb = dy + (long long)n * HP * WP * C4 + c4;
    const float4* xb = x + (long long)n * HW * C4 + c4;
    float4* dxb = dx + (long long)n * HW * C4 + c4;
    float f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
    int flush = 0;
    for (int px = p0 + pl; px < p1; px += 4 * npl) {
      // four pixels per trip: the centre gradient and x loads of all four are issued before any is used
      float4 gv[4], xv[4];
      int hh[4], ww[4], ok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = px + k * npl;
        ok[k] = q < p1;
        hh[k] = ok[k] ? q / W : 0; ww[k] = ok[k] ? q - hh[k] * W : 0;
        if (ok[k]) {
          gv[k] = dyb[((long long)(hh[k] + p) * WP + ww[k] + p) * C4];
          xv[k] = xb[(long long)q * C4];
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        const int h = hh[k], w = ww[k];
        float g[4] = {gv[k].x, gv[k].y, gv[k].z, gv[k].w};
        if (p > 0 && (h <= p || w <= p || h >= H - 1 - p || w >= W - 1 - p)) {
          // halo pixels of the reflected padding alias this interior pixel: add their gradients
          int hl[3], wl[3];
          const int nh = fold_list(h, H, p, hl), nw = fold_list(w, W, p, wl);
          for (int a = 0; a < nh; ++a)
            for (int b = 0; b < nw; ++b)
              if (a | b) acc4(g, dyb[((long long)hl[a] * WP + wl[b]) * C4]);
        }
        if (dres) dres[(((long long)n * (H + 2 * rp) + h + rp) * (W + 2 * rp) + w + rp) * C4 + c4] = make_float4(g[0], g[1], g[2], g[3]);
        const float xh[4] = {(xv[k].x - s0.x) * s0.y, (xv[k].y - s0.z) * s0.w, (xv[k].z - s1.x) * s1.y, (xv[k].w - s1.z) * s1.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (relu && !(xh[j] > 0.f)) g[j] = 0.f;
          f1[j] += g[j]; f2[j] = fmaf(g[j], xh[j], f2[j]);
        }
        dxb[(long long)(px + k * npl) * C4] = make_float4(g[0], g[1], g[2], g[3]);
      }
      if (++flush == 4) {      // fp32 partial sums over at most 16 pixels, then into fp64
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[2 * j] += (double)f1[j]; s[2 * j + 1] += (double)f2[j]; f1[j] = 0.f; f2[j] = 0.f; }
        flush = 0;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s[2 * j] += (double)f1[j]; s[2 * j + 1] += (double)f2[j]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = s[j];
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < npl; ++l)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[threadIdx.x][j] += red[l * C4 + c4][j];
    // per-CTA partial sums (no atomics, no zero-fill): in_sum_partials_kernel adds the gridDim.x chunks
    double* o = sums + (((long long)blockIdx.x * gridDim.y + n) * C4 * 4 + c4 * 4) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = red[threadIdx.x][j];
  }
}

// grid (chunks, N); thread = (channel quad, pixel lane), statistics hoisted out of the pixel loop
// bsum (nullable): per-CTA sums of the result over the CTA's pixels, [chunk][n][C] floats - the bias gradient of the
// convolution that produced x (its dy IS this dx), finished by in_bias_finalize_kernel.
__global__ void __launch_bounds__(256)
in_bwd_apply_v4_kernel(const float4* __restrict__ x, const float4* __restrict__ stats, const double* __restrict__ sums,
                       float4* __restrict__ dx, float* __restrict__ bsum, int HW, int C4, int chunk) {
  __shared__ float bred[256][4];
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
  const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4, npl = 256 / C4;
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  if (pl < npl) {
  const float4 s0 = __ldg(stats + ((long long)n * C4 + c4) * 2), s1 = __ldg(stats + ((long long)n * C4 + c4) * 2 + 1);
  const double* sm = sums + ((long long)n * C4 * 4 + c4 * 4) * 2;
  const double inv = 1.0 / HW;
  const float mean[4] = {s0.x, s0.z, s1.x, s1.z}, rstd[4] = {s0.y, s0.w, s1.y, s1.w};
  float m1[4], m2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { m1[j] = (float)(sm[2 * j] * inv); m2[j] = (float)(sm[2 * j + 1] * inv); }
  const float4* xb = x + (long long)n * HW * C4 + c4;
  float4* gb = dx + (long long)n * HW * C4 + c4;
  for (int pp = p0 + pl; pp < p1; pp += 4 * npl) {
    float4 xv[4], gv[4];
    int ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int q = pp + k * npl;
      ok[k] = q < p1;
      if (ok[k]) { xv[k] = xb[(long long)q * C4]; gv[k] = gb[(long long)q * C4]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      const float xs[4] = {xv[k].x, xv[k].y, xv[k].z, xv[k].w}, gs[4] = {gv[k].x, gv[k].y, gv[k].z, gv[k].w};
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (xs[j] - mean[j]) * rstd[j];
        r[j] = rstd[j] * (gs[j] - m1[j] - xh * m2[j]);
      }
      gb[(long long)(pp + k * npl) * C4] = make_float4(r[0], r[1], r[2], r[3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) bs[j] += r[j];
    }
  }
  }
  if (bsum == nullptr) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) bred[threadIdx.x][j] = bs[j];
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < npl; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) bs[j] += bred[l * C4 + c4][j];
    float* o = bsum + ((long long)blockIdx.x * gridDim.y + n) * C4 * 4 + c4 * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = bs[j];
  }
}

// db[c] += sum over the (chunk, sample) rows of the per-CTA sums.  grid (C / 32, row splits); thread = (channel of a
// 32-channel group, one of 8 row lanes): coalesced 128-byte reads, shared-memory sum of the lanes, one atomic per
// channel and block (db zero-filled by the caller).
__global__ void __launch_bounds__(256)
in_bias_finalize_kernel(const float* __restrict__ bsum, float* __restrict__ db, int rows, int C) {
  __shared__ float red[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float acc = 0.f;
  if (c < C)
    for (int r = blockIdx.y * 8 + rl; r < rows; r += 8 * gridDim.y) acc += bsum[(long long)r * C + c];
  red[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int l = 1; l < 8; ++l) acc += red[l][cl];
    atomicAdd(db + c, acc);
  }
}

// chunk of pixels per CTA for the elementwise instance-norm passes: ~16 CTAs per SM
static int apply_chunk(int pixels, int N, int* nchunks) {
  int want = (16 * dfmir_num_sms() + N - 1) / N;
  int chunk = (pixels + want - 1) / want;
  if (chunk < 64) chunk = 64;
  *nchunks = (pixels + chunk - 1) / chunk;
  return chunk;
}

static inline bool in_v4_ok(int C, const void* a, const void* b, const void* c, const void* d) {
  const uintptr_t m = (uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d;
  return C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0 && (m & 15) == 0;
}

// ------------------------------------------------------------------ reflection pad (stand-alone)
__global__ void __launch_bounds__(256)
pad_reflect_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int p) {
  const int HP = H + 2 * p, WP = W + 2 * p;
  const long long total = (long long)N * HP * WP * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, wp, hp, n;
    unravel4(i, total < (1LL << 32), C, WP, HP, c, wp, hp, n);
    y[i] = x[(((long long)n * H + reflect_idx(hp - p, H)) * W + reflect_idx(wp - p, W)) * C + c];
  }
}

__global__ void __launch_bounds__(256)
pad_reflect_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C, int p) {
  const int HP = H + 2 * p, WP = W + 2 * p;
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, w, h, n;
    unravel4(i, total < (1LL << 32), C, W, H, c, w, h, n);
    int hl[3], wl[3];
    const int nh = fold_list(h, H, p, hl), nw = fold_list(w, W, p, wl);
    float g = 0.f;
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) g += dy[(((long long)n * HP + hl[a]) * WP + wl[b]) * C + c];
    dx[i] = g;
  }
}

// ------------------------------------------------------------------ blur-pool down (x0.5) / blur-up (x2)
// Templated on the element type: float4 (four channels per thread, 128-bit accesses) when C % 4 == 0,
// float otherwise.  Adjoint tap lists live in three fixed register slots per axis (weight 0 = unused).
struct V1 { typedef float T; };
__device__ __forceinline__ float vzero(float) { return 0.f; }
__device__ __forceinline__ float4 vzero(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vfma(float& a, const float x, const float w) { a = fmaf(x, w, a); }
__device__ __forceinline__ void vfma(float4& a, const float4 x, const float w) {
  a.x = fmaf(x.x, w, a.x); a.y = fmaf(x.y, w, a.y); a.z = fmaf(x.z, w, a.z); a.w = fmaf(x.w, w, a.w);
}
struct Adj3 { int o[3]; float w[3]; };

template <typename T>
__global__ void __launch_bounds__(256)
blur_down_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int OH, int OW) {
  const long long total = (long long)N * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, ow, oh, n;
    unravel4(i, total < (1LL << 32), C, OW, OH, c, ow, oh, n);
    const T* xb = x + (long long)n * H * W * C + c;
    T acc = vzero(T());
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int h = reflect_idx(2 * oh - 1 + a, H);
      const float fa = a == 1 ? 0.5f : 0.25f;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int w = reflect_idx(2 * ow - 1 + b, W);
        const float fb = b == 1 ? 0.5f : 0.25f;
        vfma(acc, xb[((long long)h * W + w) * C], fa * fb);
      }
    }
    y[i] = acc;
  }
}

// adjoint taps of the blur-pool along one axis: outputs o (weight w) that read input index h.
//   direct: 2o-1+a = h;  reflected: index -1 aliases 1 (o = 0, a = 0), index H aliases H-2 (H odd: o = OH-1, a = 2)
__device__ __forceinline__ Adj3 blur_down_adj(int h, int H, int OH) {
  Adj3 r;
  r.o[0] = r.o[1] = r.o[2] = 0; r.w[0] = r.w[1] = r.w[2] = 0.f;
  if ((h & 1) == 0) {
    if (h / 2 < OH) { r.o[0] = h / 2; r.w[0] = 0.5f; }
  } else {
    r.o[0] = (h - 1) / 2; r.w[0] = 0.25f;
    if ((h + 1) / 2 < OH) { r.o[1] = (h + 1) / 2; r.w[1] = 0.25f; }
  }
  if (h == 1) { r.o[2] = 0; r.w[2] = 0.25f; }
  else if (h == H - 2 && (H & 1)) { r.o[2] = OH - 1; r.w[2] = 0.25f; }
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256)
blur_down_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int OH, int OW) {
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, w, h, n;
    unravel4(i, total < (1LL << 32), C, W, H, c, w, h, n);
    const Adj3 ah = blur_down_adj(h, H, OH), aw = blur_down_adj(w, W, OW);
    const T* gb = dy + (long long)n * OH * OW * C + c;
    T acc = vzero(T());
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (ah.w[a] == 0.f) continue;
#pragma unroll
      for (int b = 0; b < 3; ++b)
        if (aw.w[b] != 0.f) vfma(acc, gb[((long long)ah.o[a] * OW + aw.o[b]) * C], ah.w[a] * aw.w[b]);
    }
    dx[i] = acc;
  }
}

// The same adjoint for even H and W, one 2 x 2 block of dx per thread from the four dy values that touch it: per axis
//   dx[2i] = 0.5 g[i],  dx[2i+1] = 0.25 (g[i] + g[i+1])  (g[OH] = 0),  dx[1] += 0.25 g[0]  (the reflected index -1)
// - 4 loads and 4 stores per thread instead of up to 9 gathers per stored element (2.0 ms / step at 1.3 TB/s).
__device__ __forceinline__ float4 vadd(const float4 a, const float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float vadd(const float a, const float b) { return a + b; }
__device__ __forceinline__ float4 vscale(const float4 a, const float w) { return make_float4(a.x * w, a.y * w, a.z * w, a.w * w); }
__device__ __forceinline__ float vscale(const float a, const float w) { return a * w; }

template <typename T>
__global__ void __launch_bounds__(256)
blur_down_bwd_even_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int OH, int OW) {
  const long long total = (long long)N * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, j, ii, n;
    unravel4(i, total < (1LL << 32), C, OW, OH, c, j, ii, n);
    const T* gb = dy + (long long)n * OH * OW * C + c;
    const T z = vzero(T());
    const bool hj = j + 1 < OW, hi = ii + 1 < OH;
    const T g00 = gb[((long long)ii * OW + j) * C];
    const T g01 = hj ? gb[((long long)ii * OW + j + 1) * C] : z;
    const T g10 = hi ? gb[((long long)(ii + 1) * OW + j) * C] : z;
    const T g11 = (hi && hj) ? gb[((long long)(ii + 1) * OW + j + 1) * C] : z;
    // w axis: u(row, 2j), u(row, 2j + 1)
    const float ew = j == 0 ? 0.5f : 0.25f;                 // 0.25 g[j] + (j == 0: the reflected tap adds another 0.25 g[0])
    const T u0e = vscale(g00, 0.5f), u0o = vadd(vscale(g00, ew), vscale(g01, 0.25f));
    const T u1e = vscale(g10, 0.5f), u1o = vadd(vscale(g10, ew), vscale(g11, 0.25f));
    // h axis
    const float eh = ii == 0 ? 0.5f : 0.25f;
    T* ob = dx + (long long)n * H * W * C + c;
    const long long r0 = (long long)(2 * ii) * W + 2 * j, r1 = r0 + W;
    ob[r0 * C] = vscale(u0e, 0.5f);
    ob[(r0 + 1) * C] = vscale(u0o, 0.5f);
    ob[r1 * C] = vadd(vscale(u0e, eh), vscale(u1e, 0.25f));
    ob[(r1 + 1) * C] = vadd(vscale(u0o, eh), vscale(u1o, 0.25f));
  }
}

// per axis: y[2m] = (x[clamp(m-1)] + 3 x[m]) / 4 ; y[2m+1] = (3 x[m] + x[clamp(m+1)]) / 4
template <typename T>
__global__ void __launch_bounds__(256)
blur_up_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
  const int OH = 2 * H, OW = 2 * W;
  const long long total = (long long)N * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, ow, oh, n;
    unravel4(i, total < (1LL << 32), C, OW, OH, c, ow, oh, n);
    const int mh = oh >> 1, mw = ow >> 1;
    const int h0 = (oh & 1) ? mh : max(mh - 1, 0), h1 = (oh & 1) ? min(mh + 1, H - 1) : mh;
    const int w0 = (ow & 1) ? mw : max(mw - 1, 0), w1 = (ow & 1) ? min(mw + 1, W - 1) : mw;
    const float fh0 = (oh & 1) ? 0.75f : 0.25f, fh1 = 1.f - fh0;
    const float fw0 = (ow & 1) ? 0.75f : 0.25f, fw1 = 1.f - fw0;
    const T* xb = x + (long long)n * H * W * C + c;
    // same association as the scalar formula: fh0 * (fw0 a + fw1 b) + fh1 * (fw0 c + fw1 d)
    T r0 = vzero(T()), r1 = vzero(T()), acc = vzero(T());
    vfma(r0, xb[((long long)h0 * W + w0) * C], fw0); vfma(r0, xb[((long long)h0 * W + w1) * C], fw1);
    vfma(r1, xb[((long long)h1 * W + w0) * C], fw0); vfma(r1, xb[((long long)h1 * W + w1) * C], fw1);
    vfma(acc, r0, fh0); vfma(acc, r1, fh1);
    y[i] = acc;
  }
}

// adjoint of the blur-up along one axis: outputs that read input m.  2m and 2m+1 (weight 3/4 each);
// 2m+2 reads x[m] as clamp(m+1-1)... i.e. y[2(m+1)] uses x[m] with 1/4 (if m+1 < H, else the clamp makes
// y[2m+1] read x[m] twice); y[2(m-1)+1] = y[2m-1] uses x[m] with 1/4 (if m > 0, else y[2m] reads it twice).
struct Adj4 { int o[4]; float w[4]; };
__device__ __forceinline__ Adj4 blur_up_adj(int m, int H) {
  Adj4 r;
  r.o[0] = 2 * m; r.w[0] = 0.75f;
  r.o[1] = 2 * m + 1; r.w[1] = 0.75f;
  r.o[2] = (m + 1 <= H - 1) ? 2 * m + 2 : 2 * m + 1; r.w[2] = 0.25f;
  r.o[3] = (m - 1 >= 0) ? 2 * m - 1 : 2 * m; r.w[3] = 0.25f;
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256)
blur_up_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C) {
  const int OW = 2 * W, OH = 2 * H;
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, w, h, n;
    unravel4(i, total < (1LL << 32), C, W, H, c, w, h, n);
    const Adj4 ah = blur_up_adj(h, H), aw = blur_up_adj(w, W);
    const T* gb = dy + (long long)n * OH * OW * C + c;
    T acc = vzero(T());
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) vfma(acc, gb[((long long)ah.o[a] * OW + aw.o[b]) * C], ah.w[a] * aw.w[b]);
    dx[i] = acc;
  }
}

// ------------------------------------------------------------------ nearest x2 upsample + concat (U-Net skip)
struct UcGeom { int N, nd, C1, C2; int S[3]; int Cs; };  // S = full-resolution spatial dims (right-aligned); Cs = channel stride of y (>= C1+C2, the rest zero)

__global__ void __launch_bounds__(256)
upcat_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, UcGeom g) {
  const int C = g.Cs;
  const long long vox = (long long)g.S[0] * g.S[1] * g.S[2];
  const long long total = (long long)g.N * vox * C;
  const int L0 = g.S[0] > 1 ? g.S[0] / 2 : 1, L1 = g.S[1] / 2, L2 = g.S[2] / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long q = i;
    const int c = (int)(q % C); q /= C;
    const int xw = (int)(q % g.S[2]); q /= g.S[2];
    const int yh = (int)(q % g.S[1]); q /= g.S[1];
    const int zd = (int)(q % g.S[0]); q /= g.S[0];
    const int n = (int)q;
    float v;
    if (c < g.C1) {
      const int lz = g.S[0] > 1 ? zd >> 1 : 0;
      v = a[((((long long)n * L0 + lz) * L1 + (yh >> 1)) * L2 + (xw >> 1)) * g.C1 + c];
    } else if (c < g.C1 + g.C2) {
      v = b[((((long long)n * g.S[0] + zd) * g.S[1] + yh) * g.S[2] + xw) * g.C2 + (c - g.C1)];
    } else {
      v = 0.f;      // channel padding (keeps the pixel stride a multiple of 16 bytes for the TMA loads of the next conv)
    }
    y[i] = v;
  }
}

// 128-bit variant (C1 % 4 == 0, Cs % 4 == 0): one float4 of one output voxel per thread.  The first C1 / 4 quads are
// copies of the parent voxel's quads; the remaining ones are assembled from the skip tensor (whole quads when C2 % 4 == 0)
// and the zero padding.
template <typename IDX>      // unsigned when the element count fits 32 bits (64-bit divides dominated this copy kernel)
__global__ void __launch_bounds__(256)
upcat_fwd_v4_kernel(const float4* __restrict__ a, const float* __restrict__ b, float4* __restrict__ y, UcGeom g) {
  const int Q = g.Cs >> 2, Q1 = g.C1 >> 2;
  const long long total = (long long)g.N * g.S[0] * g.S[1] * g.S[2] * Q;
  const int L0 = g.S[0] > 1 ? g.S[0] / 2 : 1, L1 = g.S[1] / 2, L2 = g.S[2] / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const IDX pix = (IDX)i / (IDX)Q;
    const int j = (int)((IDX)i - pix * (IDX)Q);
    float4 v;
    if (j < Q1) {
      IDX q = pix;
      const int xw = (int)(q % (IDX)g.S[2]); q /= (IDX)g.S[2];
      const int yh = (int)(q % (IDX)g.S[1]); q /= (IDX)g.S[1];
      const int zd = (int)(q % (IDX)g.S[0]); q /= (IDX)g.S[0];
      const int lz = g.S[0] > 1 ? zd >> 1 : 0;
      v = a[(((((long long)q * L0 + lz) * L1 + (yh >> 1)) * L2 + (xw >> 1)) * Q1) + j];
    } else {
      const int c0 = (j << 2) - g.C1;              // first skip channel of this quad
      const float* bp = b + (long long)pix * g.C2 + c0;
      if ((g.C2 & 3) == 0 && c0 + 4 <= g.C2) {
        v = *reinterpret_cast<const float4*>(bp);
      } else {
        v.x = c0 < g.C2 ? bp[0] : 0.f; v.y = c0 + 1 < g.C2 ? bp[1] : 0.f;
        v.z = c0 + 2 < g.C2 ? bp[2] : 0.f; v.w = c0 + 3 < g.C2 ? bp[3] : 0.f;
      }
    }
    y[i] = v;
  }
}

// da (128-bit, C1 % 4 == 0, Cs % 4 == 0): one quad of one parent voxel per thread, the 2^nd children summed in a fixed order
template <typename IDX>
__global__ void __launch_bounds__(256)
upcat_bwd_a_v4_kernel(const float4* __restrict__ dy, float4* __restrict__ da, UcGeom g) {
  const int Q = g.Cs >> 2, Q1 = g.C1 >> 2;
  const int L0 = g.S[0] > 1 ? g.S[0] / 2 : 1, L1 = g.S[1] / 2, L2 = g.S[2] / 2;
  const long long total = (long long)g.N * L0 * L1 * L2 * Q1;
  const int nz = g.S[0] > 1 ? 2 : 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    IDX q = (IDX)i;
    const int j = (int)(q % (IDX)Q1); q /= (IDX)Q1;
    const int lx = (int)(q % (IDX)L2); q /= (IDX)L2;
    const int ly = (int)(q % (IDX)L1); q /= (IDX)L1;
    const int lz = (int)(q % (IDX)L0); q /= (IDX)L0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dz = 0; dz < nz; ++dz)
#pragma unroll
      for (int dyy = 0; dyy < 2; ++dyy)
#pragma unroll
        for (int dxx = 0; dxx < 2; ++dxx) {
          const int zd = g.S[0] > 1 ? 2 * lz + dz : 0;
          const float4 t = dy[(((((long long)q * g.S[0] + zd) * g.S[1] + 2 * ly + dyy) * g.S[2] + 2 * lx + dxx) * Q) + j];
          acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
    da[i] = acc;
  }
}

// db = dy[..., C1 : C1 + C2]
__global__ void __launch_bounds__(256)
upcat_bwd_b_kernel(const float* __restrict__ dy, float* __restrict__ db, long long pixels, int C1, int C2, int Cs) {
  const long long total = pixels * C2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / C2;
    db[i] = dy[pix * Cs + C1 + (int)(i - pix * C2)];
  }
}

// da = sum over the 2^nd children of dy[..., :C1];  db = dy[..., C1:]
__global__ void __launch_bounds__(256)
upcat_bwd_kernel(const float* __restrict__ dy, float* __restrict__ da, float* __restrict__ db, UcGeom g) {
  const int C = g.Cs;
  const int L0 = g.S[0] > 1 ? g.S[0] / 2 : 1, L1 = g.S[1] / 2, L2 = g.S[2] / 2;
  const long long na = (long long)g.N * L0 * L1 * L2 * g.C1;
  const long long nb = (long long)g.N * g.S[0] * g.S[1] * g.S[2] * g.C2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += (long long)gridDim.x * blockDim.x) {
    if (i < na) {
      long long q = i;
      const int c = (int)(q % g.C1); q /= g.C1;
      const int lx = (int)(q % L2); q /= L2;
      const int ly = (int)(q % L1); q /= L1;
      const int lz = (int)(q % L0); q /= L0;
      const int n = (int)q;
      float acc = 0.f;
      const int nz = g.S[0] > 1 ? 2 : 1;
      for (int dz = 0; dz < nz; ++dz)
        for (int dyy = 0; dyy < 2; ++dyy)
          for (int dxx = 0; dxx < 2; ++dxx) {
            const int zd = g.S[0] > 1 ? 2 * lz + dz : 0;
            acc += dy[((((long long)n * g.S[0] + zd) * g.S[1] + 2 * ly + dyy) * g.S[2] + 2 * lx + dxx) * C + c];
          }
      if (da) da[i] = acc;
    } else if (db) {
      const long long j = i - na;
      const int c = (int)(j % g.C2);
      const long long pos = j / g.C2;
      db[j] = dy[pos * C + g.C1 + c];
    }
  }
}

// ------------------------------------------------------------------ space-to-depth (stride-2 convolutions on the tensor cores)
// y[n, jd, jh, jw, ((pd*2 + ph)*2 + pw)*C + c] = x[n, 2jd+pd, 2jh+ph, 2jw+pw, c]   (2-D: no d axis, 4C channels)
// A stride-2 3^nd convolution with pad 1 over x (vxm networks.py:1514-1515, the U-Net encoder) is a stride-1 2^nd
// convolution over y with one leading zero row per axis, which the tcgen05 implicit-GEMM kernels run as they are.
// V = elements per access (the channel count is a multiple of V).  inverse = the adjoint / inverse permutation.
template <typename T>
__global__ void __launch_bounds__(256)
s2d_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int OD, int OH, int OW, int CV, int nd3, int inverse) {
  const int P = nd3 ? 8 : 4;
  const long long total = (long long)N * OD * OH * OW * P * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long q = i;
    const int c = (int)(q % CV); q /= CV;
    const int par = (int)(q % P); q /= P;
    const int jw = (int)(q % OW); q /= OW;
    const int jh = (int)(q % OH); q /= OH;
    const int jd = (int)(q % OD); q /= OD;
    const int n = (int)q;
    const int pw = par & 1, ph = (par >> 1) & 1, pd = nd3 ? (par >> 2) : 0;
    const int ID = nd3 ? 2 * OD : 1;
    const long long full = ((((long long)n * ID + (nd3 ? 2 * jd + pd : 0)) * (2 * OH) + 2 * jh + ph) * (2 * OW) + 2 * jw + pw) * CV + c;
    if (inverse) dst[full] = src[i]; else dst[i] = src[full];
  }
}

}  // namespace

extern "C" int dfmir_space_to_depth(const float* src, float* dst, int N, int nd, const int* out_shape, int C, int inverse, void* stream) {
  DFMIR_CHECK_ARG(src && dst && N > 0 && C > 0 && (nd == 2 || nd == 3) && out_shape, "dfmir_space_to_depth: bad argument");
  const int OD = nd == 3 ? out_shape[0] : 1, OH = out_shape[nd - 2], OW = out_shape[nd - 1];
  DFMIR_CHECK_ARG(OD > 0 && OH > 0 && OW > 0, "dfmir_space_to_depth: bad shape");
  const long long total = (long long)N * OD * OH * OW * (nd == 3 ? 8 : 4) * C;
  cudaStream_t st = (cudaStream_t)stream;
  const uintptr_t al = (uintptr_t)src | (uintptr_t)dst;
  if (C % 4 == 0 && (al & 15) == 0)
    s2d_kernel<float4><<<ew_grid(total / 4), 256, 0, st>>>((const float4*)src, (float4*)dst, N, OD, OH, OW, C / 4, nd == 3, inverse);
  else if (C % 2 == 0 && (al & 7) == 0)
    s2d_kernel<float2><<<ew_grid(total / 2), 256, 0, st>>>((const float2*)src, (float2*)dst, N, OD, OH, OW, C / 2, nd == 3, inverse);
  else
    s2d_kernel<float><<<ew_grid(total), 256, 0, st>>>(src, dst, N, OD, OH, OW, C, nd == 3, inverse);
  DFMIR_CHECK_LAUNCH("dfmir_space_to_depth");
  return DFMIR_OK;
}

// [sums: 2*N*C doubles][per-CTA partials: chunks * 2*N*C doubles, chunks * N <= 16 * SMs + 2 * N]
extern "C" size_t dfmir_instnorm_workspace_bytes(int N, int C) {
  return sizeof(double) * 2 * (size_t)C * ((size_t)N + 16 * (size_t)dfmir_num_sms() + 2 * (size_t)N) + 256;
}

static int in_chunk(int HW, int N, int* nchunks) {
  // enough CTAs for ~8 per SM, at least 128 pixels per CTA (finer chunks measured slower: 7.5 vs 7.2 ms/step for
  // the backward reduction, and the partial-sum pass grows with the chunk count)
  int want = (8 * dfmir_num_sms() + N - 1) / N;
  int chunk = (HW + want - 1) / want;
  if (chunk < 128) chunk = 128;
  *nchunks = (HW + chunk - 1) / chunk;
  return chunk;
}

extern "C" int dfmir_instnorm_fwd(const float* x, const float* res, float* y, float* stats, void* ws, size_t ws_bytes,
                                  int N, int H, int W, int C, float eps, int relu, int out_pad, int res_pad,
                                  void* stream) {
  DFMIR_CHECK_ARG(x && y && stats && ws, "dfmir_instnorm_fwd: null pointer");
  DFMIR_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0, "dfmir_instnorm_fwd: bad sizes");
  DFMIR_CHECK_ARG(out_pad >= 0 && out_pad < H && out_pad < W && res_pad >= 0, "dfmir_instnorm_fwd: bad padding");
  DFMIR_CHECK_ARG(ws_bytes >= dfmir_instnorm_workspace_bytes(N, C), "dfmir_instnorm_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = (double*)ws;
  double* partials = sums + 2 * (size_t)N * C;
  int nch; const int chunk = in_chunk(H * W, N, &nch);
  const bool v4 = in_v4_ok(C, x, y, res, stats) && N <= 65535 && H + 2 * out_pad <= 65535;
  if (v4) {
    in_stats_v4_kernel<<<dim3(nch, N), 256, 0, st>>>((const float4*)x, partials, H * W, C / 4, chunk);
    DFMIR_CHECK_LAUNCH("dfmir_instnorm_fwd(stats)");
    in_sum_partials_kernel<<<(2 * N * C + 255) / 256, 256, 0, st>>>(partials, sums, 2 * N * C, nch);
  } else {
    DFMIR_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)N * C, st));
    in_stats_kernel<<<dim3(nch, N), 256, 0, st>>>(x, sums, H * W, C, chunk);
  }
  DFMIR_CHECK_LAUNCH("dfmir_instnorm_fwd(stats)");
  in_finalize_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(sums, stats, N * C, H * W, eps);
  DFMIR_CHECK_LAUNCH("dfmir_instnorm_fwd(finalize)");
  if (v4) {
    int ach; const int achunk = apply_chunk((H + 2 * out_pad) * (W + 2 * out_pad), N, &ach);
    in_apply_v4_kernel<<<dim3(ach, N), 256, 0, st>>>((const float4*)x, (const float4*)stats, (const float4*)res, (float4*)y, H, W,
                                                     C / 4, relu, out_pad, res_pad, achunk);
  } else {
    const long long total = (long long)N * (H + 2 * out_pad) * (W + 2 * out_pad) * C;
    in_apply_kernel<<<ew_grid(total), 256, 0, st>>>(x, stats, res, y, N, H, W, C, relu, out_pad, res_pad);
  }
  DFMIR_CHECK_LAUNCH("dfmir_instnorm_fwd(apply)");
  return DFMIR_OK;
}

// The same layer with the statistics taken from the producing convolution's epilogue (dfmir_conv_umma_fwd_stats)
// instead of a pass over x: stat_rows [N][rows_per_image][C] float2.
extern "C" int dfmir_instnorm_fwd_rows(const float* x, const float* res, float* y, float* stats, const float* stat_rows,
                                       int rows_per_image, int N, int H, int W, int C, float eps, int relu, int out_pad,
                                       int res_pad, void* stream) {
  DFMIR_CHECK_ARG(x && y && stats && stat_rows && rows_per_image > 0, "dfmir_instnorm_fwd_rows: null pointer");
  DFMIR_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0, "dfmir_instnorm_fwd_rows: bad sizes");
  DFMIR_CHECK_ARG(out_pad >= 0 && out_pad < H && out_pad < W && res_pad >= 0, "dfmir_instnorm_fwd_rows: bad padding");
  DFMIR_CHECK_ARG(N <= 65535, "dfmir_instnorm_fwd_rows: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  in_finalize_rows_kernel<<<dim3((C + 31) / 32, N), 1024, 0, st>>>((const float2*)stat_rows, stats, rows_per_image, C, H * W, eps);
  DFMIR_CHECK_LAUNCH("dfmir_instnorm_fwd_rows(finalize)");
  if (in_v4_ok(C, x, y, res, stats) && H + 2 * out_pad <= 65535) {
    int ach; const int achunk = apply_chunk((H + 2 * out_pad) * (W + 2 * out_pad), N, &ach);
    in_apply_v4_kernel<<<dim3(ach, N), 256, 0, st>>>((const float4*)x, (const float4*)stats, (const float4*)res, (float4*)y, H, W,
                                                     C / 4, relu, out_pad, res_pad, achunk);
  } else {
    const long long total = (long long)N * (H + 2 * out_pad) * (W + 2 * out_pad) * C;
    in_apply_kernel<<<ew_grid(total), 256, 0, st>>>(x, stats, res, y, N, H, W, C, relu, out_pad, res_pad);
  }
  DFMIR_CHECK_LAUNCH("dfmir_instnorm_fwd_rows(apply)");
  return DFMIR_OK;
}

extern "C" int dfmir_instnorm_bwd(const float* dy, const float* x, const float* stats, float* dx, float* dres,
                                  void* ws, size_t ws_bytes, int N, int H, int W, int C, int relu, int out_pad,
                                  int res_pad, void* stream) {
  return dfmir_instnorm_bwd_bias(dy, x, stats, dx, dres, nullptr, ws, ws_bytes, N, H, W, C, relu, out_pad, res_pad, stream);
}

extern "C" int dfmir_instnorm_bwd_bias(const float* dy, const float* x, const float* stats, float* dx, float* dres,
                                       float* dbias, void* ws, size_t ws_bytes, int N, int H, int W, int C, int relu,
                                       int out_pad, int res_pad, void* stream) {
  DFMIR_CHECK_ARG(dy && x && stats && dx && ws, "dfmir_instnorm_bwd: null pointer");
  DFMIR_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0, "dfmir_instnorm_bwd: bad sizes");
  DFMIR_CHECK_ARG(ws_bytes >= dfmir_instnorm_workspace_bytes(N, C), "dfmir_instnorm_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = (double*)ws;
  double* partials = sums + 2 * (size_t)N * C;
  if (dres && res_pad > 0)
    DFMIR_CUDA(cudaMemsetAsync(dres, 0, sizeof(float) * (size_t)N * (H + 2 * res_pad) * (W + 2 * res_pad) * C, st));
  int nch; const int chunk = in_chunk(H * W, N, &nch);
  const bool v4 = in_v4_ok(C, x, dy, dx, dres) && (((uintptr_t)stats) & 15) == 0 && N <= 65535;
  if (v4) {
    in_bwd_reduce_v4_kernel<<<dim3(nch, N), 256, 0, st>>>((const float4*)dy, (const float4*)x, (const float4*)stats, (float4*)dx,
                                                          (float4*)dres, partials, H, W, C / 4, relu, out_pad, res_pad, chunk);
    DFMIR_CHECK_LAUNCH("dfmir_instnorm_bwd(reduce)");
    in_sum_partials_kernel<<<(2 * N * C + 255) / 256, 256, 0, st>>>(partials, sums, 2 * N * C, nch);
    DFMIR_CHECK_LAUNCH("dfmir_instnorm_bwd(sum)");
    int ach; const int achunk = apply_chunk(H * W, N, &ach);
    // the reduction's partial sums are consumed: their space holds the per-CTA bias sums (ach * N * C floats)
    float* bsum = dbias ? (float*)partials : nullptr;
    in_bwd_apply_v4_kernel<<<dim3(ach, N), 256, 0, st>>>((const float4*)x, (const float4*)stats, sums, (float4*)dx, bsum, H * W, C / 4, achunk);
    DFMIR_CHECK_LAUNCH("dfmir_instnorm_bwd(apply)");
    if (dbias) {
      DFMIR_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * C, st));
      const int rows = ach * N;
      in_bias_finalize_kernel<<<dim3((C + 31) / 32, rows >= 512 ? 16 : 1), 256, 0, st>>>(bsum, dbias, rows, C);
      DFMIR_CHECK_LAUNCH("dfmir_instnorm_bwd(bias)");
    }
    return DFMIR_OK;
  }
  DFMIR_CHECK_ARG(dbias == nullptr, "dfmir_instnorm_bwd_bias: the bias by-product needs the 128-bit path (C %% 4 == 0, 16-byte aligned tensors)");
  DFMIR_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)N * C, st));
  in_bwd_reduce_kernel<<<dim3(nch, N), 256, 0, st>>>(dy, x, stats, dx, dres, sums, H, W, C, relu, out_pad, res_pad, chunk);
  DFMIR_CHECK_LAUNCH("dfmir_instnorm_bwd(reduce)");
  in_bwd_apply_kernel<<<ew_grid((long long)N * H * W * C), 256, 0, st>>>(x, stats, sums, dx, N, H * W, C);
  DFMIR_CHECK_LAUNCH("dfmir_instnorm_bwd(apply)");
  return DFMIR_OK;
}

extern "C" int dfmir_pad_reflect_fwd(const float* x, float* y, int N, int H, int W, int C, int pad, void* stream) {
  DFMIR_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0, "dfmir_pad_reflect_fwd: bad argument");
  DFMIR_CHECK_ARG(pad >= 0 && pad < H && pad < W, "dfmir_pad_reflect_fwd: pad %d must be smaller than the image", pad);
  const long long total = (long long)N * (H + 2 * pad) * (W + 2 * pad) * C;
  pad_reflect_fwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, C, pad);
  DFMIR_CHECK_LAUNCH("dfmir_pad_reflect_fwd");
  return DFMIR_OK;
}

extern "C" int dfmir_pad_reflect_bwd(const float* dy, float* dx, int N, int H, int W, int C, int pad, void* stream) {
  DFMIR_CHECK_ARG(dy && dx && N > 0 && H > 0 && W > 0 && C > 0, "dfmir_pad_reflect_bwd: bad argument");
  DFMIR_CHECK_ARG(pad >= 0 && pad < H && pad < W, "dfmir_pad_reflect_bwd: pad %d must be smaller than the image", pad);
  pad_reflect_bwd_kernel<<<ew_grid((long long)N * H * W * C), 256, 0, (cudaStream_t)stream>>>(dy, dx, N, H, W, C, pad);
  DFMIR_CHECK_LAUNCH("dfmir_pad_reflect_bwd");
  return DFMIR_OK;
}

extern "C" int dfmir_blur_down_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  DFMIR_CHECK_ARG(x && y && N > 0 && H > 1 && W > 1 && C > 0, "dfmir_blur_down_fwd: bad argument");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  if (C % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0)
    blur_down_fwd_kernel<float4><<<ew_grid((long long)N * OH * OW * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)x, (float4*)y, N, H, W, C / 4, OH, OW);
  else
    blur_down_fwd_kernel<float><<<ew_grid((long long)N * OH * OW * C), 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, C, OH, OW);
  DFMIR_CHECK_LAUNCH("dfmir_blur_down_fwd");
  return DFMIR_OK;
}

extern "C" int dfmir_blur_down_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream) {
  DFMIR_CHECK_ARG(dy && dx && N > 0 && H > 1 && W > 1 && C > 0, "dfmir_blur_down_bwd: bad argument");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const bool even = (H % 2 == 0) && (W % 2 == 0);          // the layers of the generator (256 -> 128 -> 64)
  if (even && C % 4 == 0 && ((((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0)
    blur_down_bwd_even_kernel<float4><<<ew_grid((long long)N * OH * OW * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dy, (float4*)dx, N, H, W, C / 4, OH, OW);
  else if (even)
    blur_down_bwd_even_kernel<float><<<ew_grid((long long)N * OH * OW * C), 256, 0, (cudaStream_t)stream>>>(dy, dx, N, H, W, C, OH, OW);
  else if (C % 4 == 0 && ((((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0)
    blur_down_bwd_kernel<float4><<<ew_grid((long long)N * H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dy, (float4*)dx, N, H, W, C / 4, OH, OW);
  else
    blur_down_bwd_kernel<float><<<ew_grid((long long)N * H * W * C), 256, 0, (cudaStream_t)stream>>>(dy, dx, N, H, W, C, OH, OW);
  DFMIR_CHECK_LAUNCH("dfmir_blur_down_bwd");
  return DFMIR_OK;
}

extern "C" int dfmir_blur_up_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  DFMIR_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0, "dfmir_blur_up_fwd: bad argument");
  if (C % 4 == 0 && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0)
    blur_up_fwd_kernel<float4><<<ew_grid((long long)N * 4 * H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)x, (float4*)y, N, H, W, C / 4);
  else
    blur_up_fwd_kernel<float><<<ew_grid((long long)N * 4 * H * W * C), 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, C);
  DFMIR_CHECK_LAUNCH("dfmir_blur_up_fwd");
  return DFMIR_OK;
}

extern "C" int dfmir_blur_up_bwd(const float* dy, float* dx, int N, int H, int W, int C, void* stream) {
  DFMIR_CHECK_ARG(dy && dx && N > 0 && H > 0 && W > 0 && C > 0, "dfmir_blur_up_bwd: bad argument");
  if (C % 4 == 0 && ((((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0)
    blur_up_bwd_kernel<float4><<<ew_grid((long long)N * H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)dy, (float4*)dx, N, H, W, C / 4);
  else
    blur_up_bwd_kernel<float><<<ew_grid((long long)N * H * W * C), 256, 0, (cudaStream_t)stream>>>(dy, dx, N, H, W, C);
  DFMIR_CHECK_LAUNCH("dfmir_blur_up_bwd");
  return DFMIR_OK;
}

static int make_uc(UcGeom& g, int N, int nd, const int* shape, int C1, int C2) {
  if ((nd != 2 && nd != 3) || N < 1 || C1 < 1 || C2 < 0) return -1;
  g.N = N; g.nd = nd; g.C1 = C1; g.C2 = C2; g.Cs = C1 + C2;
  g.S[0] = 1;
  for (int a = 0; a < nd; ++a) {
    g.S[a + 3 - nd] = shape[a];
    if (shape[a] < 2 || (shape[a] & 1)) return -1;  // x2 nearest upsample: full-res dims are even
  }
  return 0;
}

extern "C" int dfmir_upsample_concat_padded_fwd(const float* a, const float* b, float* y, int N, int nd, const int* shape,
                                                int C1, int C2, int Cs, void* stream) {
  UcGeom g;
  DFMIR_CHECK_ARG(make_uc(g, N, nd, shape, C1, C2) == 0, "dfmir_upsample_concat_fwd: bad geometry (even full-res dims, nd 2|3)");
  DFMIR_CHECK_ARG(a && y && (b || C2 == 0), "dfmir_upsample_concat_fwd: null pointer");
  DFMIR_CHECK_ARG(Cs >= C1 + C2, "dfmir_upsample_concat_fwd: channel stride %d smaller than %d + %d channels", Cs, C1, C2);
  g.Cs = Cs;
  const long long total = (long long)N * g.S[0] * g.S[1] * g.S[2] * Cs;
  if (C1 % 4 == 0 && Cs % 4 == 0 && ((((uintptr_t)a) | ((uintptr_t)y)) & 15) == 0 && (C2 % 4 != 0 || (((uintptr_t)b) & 15) == 0))
  {
    if (total / 4 < (1LL << 32)) upcat_fwd_v4_kernel<unsigned><<<ew_grid(total / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)a, b, (float4*)y, g);
    else upcat_fwd_v4_kernel<long long><<<ew_grid(total / 4), 256, 0, (cudaStream_t)stream>>>((const float4*)a, b, (float4*)y, g);
  }
  else
    upcat_fwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(a, b, y, g);
  DFMIR_CHECK_LAUNCH("dfmir_upsample_concat_fwd");
  return DFMIR_OK;
}

extern "C" int dfmir_upsample_concat_padded_bwd(const float* dy, float* da, float* db, int N, int nd, const int* shape,
                                                int C1, int C2, int Cs, void* stream) {
  UcGeom g;
  DFMIR_CHECK_ARG(make_uc(g, N, nd, shape, C1, C2) == 0, "dfmir_upsample_concat_bwd: bad geometry");
  DFMIR_CHECK_ARG(dy && (da || db), "dfmir_upsample_concat_bwd: null pointer");
  DFMIR_CHECK_ARG(Cs >= C1 + C2, "dfmir_upsample_concat_bwd: channel stride %d smaller than %d + %d channels", Cs, C1, C2);
  g.Cs = Cs;
  const long long vox = (long long)g.S[0] * g.S[1] * g.S[2];
  const long long total = (long long)N * (vox >> nd) * C1 + (long long)N * vox * C2;
  if (C1 % 4 == 0 && Cs % 4 == 0 && ((((uintptr_t)dy) | ((uintptr_t)da)) & 15) == 0) {
    const long long ta = (long long)N * (vox >> nd) * (C1 / 4);
    if (da && ta < (1LL << 32)) upcat_bwd_a_v4_kernel<unsigned><<<ew_grid(ta), 256, 0, (cudaStream_t)stream>>>((const float4*)dy, (float4*)da, g);
    else if (da) upcat_bwd_a_v4_kernel<long long><<<ew_grid(ta), 256, 0, (cudaStream_t)stream>>>((const float4*)dy, (float4*)da, g);
    if (db && C2 > 0) {
      if (da) DFMIR_CHECK_LAUNCH("dfmir_upsample_concat_bwd");
      upcat_bwd_b_kernel<<<ew_grid((long long)N * vox * C2), 256, 0, (cudaStream_t)stream>>>(dy, db, (long long)N * vox, C1, C2, Cs);
    }
  } else {
    upcat_bwd_kernel<<<ew_grid(total), 256, 0, (cudaStream_t)stream>>>(dy, da, db, g);
  }
  DFMIR_CHECK_LAUNCH("dfmir_upsample_concat_bwd");
  return DFMIR_OK;
}

extern "C" int dfmir_upsample_concat_fwd(const float* a, const float* b, float* y, int N, int nd, const int* shape,
                                         int C1, int C2, void* stream) {
  return dfmir_upsample_concat_padded_fwd(a, b, y, N, nd, shape, C1, C2, C1 + C2, stream);
}

extern "C" int dfmir_upsample_concat_bwd(const float* dy, float* da, float* db, int N, int nd, const int* shape,
                                         int C1, int C2, void* stream) {
  return dfmir_upsample_concat_padded_bwd(dy, da, db, N, nd, shape, C1, C2, C1 + C2, stream);
}
