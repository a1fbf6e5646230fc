/*
 * dfmir_oracle.c — CPU restatement of the reference's memory-bound registration ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dfmir_b200/ may import, link or call this file; it is
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
 * the checker.  Plain scalar C, IEEE fp32, compiled with -ffp-contract=off so that every
 * operation rounds exactly like the reference's PyTorch CPU path.
 *
 * Pinned against the reference itself: oracle/gen_golden.py imports /root/reference, runs
 * SpatialTransformer / VecInt / ResizeTransform / NCC_Loss / Grad_Loss / smooothing_loss /
 * calculate_L1_loss on seeded inputs and commits the results to tests/golden/; tests/test_oracle.py
 * checks this file against those vectors (bit-exact for indices and normalised coordinates).
 *
 * Reference (paths relative to the reference repo):
 *   orc_warp      models/voxelmorph/torchvoxelmorph/layers.py:30-48 + ATen grid_sampler (align_corners=True, zeros)
 *   orc_vecint    models/voxelmorph/torchvoxelmorph/layers.py:64-68
 *   orc_resize    models/voxelmorph/torchvoxelmorph/layers.py:85-97 + ATen upsample_linear (align_corners=True)
 *   orc_ncc       util/losses.py:183-261 ; reduction 1: models/voxelmorph/torchvoxelmorph/losses.py:15-67
 *   orc_grad      util/losses.py:92-116 ; models/registration_model.py:25-32
 *   orc_l1_masked models/registration_model.py:255-263
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* un-normalised sampling coordinate: layers.py:32-37 then grid_sampler_unnormalize */
static float unnorm_coord(int i, float f, int S, int rcp_mul) {
  const float sm1 = (float)(S - 1);
  const float loc = (float)i + f;
  float q;
  if (rcp_mul) {
    const float r = 1.0f / sm1; /* CUDA ATen divides by a scalar as a reciprocal multiply */
    q = loc * r;
  } else {
    q = loc / sm1;
  }
  const float n = 2.0f * (q - 0.5f);
  return ((n + 1.0f) / 2.0f) * sm1;
}

static int safe_int(float v) {
  if (!(v > -1.0e9f)) return -1000000000;
  if (v > 1.0e9f) return 1000000000;
  return (int)v;
}

/* normalised coordinates (the tensor the reference hands to F.grid_sample), (B,nd,*S), ij order */
void orc_normalized_grid(const float* flow, float* ngrid, int B, int nd, const int* S) {
  long long nvox = 1;
  for (int d = 0; d < nd; ++d) nvox *= S[d];
  for (int b = 0; b < B; ++b)
    for (long long v = 0; v < nvox; ++v) {
      long long r = v;
      int pos[3] = {0, 0, 0};
      for (int d = nd - 1; d >= 0; --d) { pos[d] = (int)(r % S[d]); r /= S[d]; }
      for (int d = 0; d < nd; ++d) {
        const long long o = ((long long)b * nd + d) * nvox + v;
        const float loc = (float)pos[d] + flow[o];
        ngrid[o] = 2.0f * (loc / (float)(S[d] - 1) - 0.5f);
      }
    }
}

/* interp: 0 linear, 1 nearest.  idx_out (nullable) int32 (B,nd,*S). */
void orc_warp(const float* src, const float* flow, float* out, int32_t* idx_out, int B, int C, int nd,
              const int* S, int interp, int rcp_mul) {
  long long nvox = 1;
  for (int d = 0; d < nd; ++d) nvox *= S[d];
  const int ncorner = 1 << nd;
  for (int b = 0; b < B; ++b)
    for (long long v = 0; v < nvox; ++v) {
      long long r = v;
      int pos[3] = {0, 0, 0}, i0[3] = {0, 0, 0};
      float w0[3], w1[3];
      for (int d = nd - 1; d >= 0; --d) { pos[d] = (int)(r % S[d]); r /= S[d]; }
      for (int d = 0; d < nd; ++d) {
        const float f = flow[((long long)b * nd + d) * nvox + v];
        const float ix = unnorm_coord(pos[d], f, S[d], rcp_mul);
        if (interp == 1) {
          i0[d] = safe_int(nearbyintf(ix));
        } else {
          const float fl = floorf(ix);
          i0[d] = safe_int(fl);
          w1[d] = ix - fl;
          /* ATen's vectorised 2-D CPU kernel forms the far weight as 1 - w (GridSamplerKernel.cpp,
           * compute_interp_params); the generic 3-D / CUDA kernels use (i0 + 1) - ix. */
          w0[d] = nd == 2 ? 1.0f - w1[d] : (fl + 1.0f) - ix;
        }
        if (idx_out) idx_out[((long long)b * nd + d) * nvox + v] = i0[d];
      }
      for (int c = 0; c < C; ++c) {
        const float* plane = src + ((long long)b * C + c) * nvox;
        float acc = 0.f;
        if (interp == 1) {
          long long off = 0; int ok = 1;
          for (int d = 0; d < nd; ++d) { ok = ok && i0[d] >= 0 && i0[d] < S[d]; off = off * S[d] + i0[d]; }
          acc = ok ? plane[off] : 0.f;
        } else {
          for (int k = 0; k < ncorner; ++k) {
            long long off = 0; int ok = 1; float w = 1.f;
            for (int d = 0; d < nd; ++d) {
              const int id = i0[d] + ((k >> (nd - 1 - d)) & 1);
              ok = ok && id >= 0 && id < S[d];
              off = off * S[d] + id;
            }
            for (int d = nd - 1; d >= 0; --d) {
              const float wd = ((k >> (nd - 1 - d)) & 1) ? w1[d] : w0[d];
              w = (d == nd - 1) ? wd : w * wd;
            }
            if (ok) acc += plane[off] * w;
          }
        }
        out[((long long)b * C + c) * nvox + v] = acc;
      }
    }
}

/* vec (B,nd,*S) -> out (B,nd,*S);  layers.py:64-68 */
void orc_vecint(const float* vec, float* out, int B, int nd, const int* S, int nsteps, int rcp_mul) {
  long long nvox = 1;
  for (int d = 0; d < nd; ++d) nvox *= S[d];
  const long long n = (long long)B * nd * nvox;
  float* cur = (float*)malloc(sizeof(float) * n);
  float* tmp = (float*)malloc(sizeof(float) * n);
  const float scale = 1.0f / (float)(1 << nsteps);
  for (long long i = 0; i < n; ++i) cur[i] = vec[i] * scale;
  for (int k = 0; k < nsteps; ++k) {
    orc_warp(cur, cur, tmp, NULL, B, nd, nd, S, 0, rcp_mul);
    for (long long i = 0; i < n; ++i) cur[i] = cur[i] + tmp[i];
  }
  memcpy(out, cur, sizeof(float) * n);
  free(cur); free(tmp);
}

/* y = post_mul * interp(pre_mul * x), linear, align_corners=True */
void orc_resize(const float* x, float* y, int BC, int nd, const int* I, const int* O, float pre_mul,
                float post_mul) {
  long long nin = 1, nout = 1;
  float sc[3] = {0, 0, 0};
  for (int d = 0; d < nd; ++d) {
    nin *= I[d]; nout *= O[d];
    sc[d] = O[d] > 1 ? (float)(I[d] - 1) / (float)(O[d] - 1) : 0.f;
  }
  for (int p = 0; p < BC; ++p)
    for (long long v = 0; v < nout; ++v) {
      long long r = v;
      int i0[3], i1[3]; float l0[3], l1[3];
      for (int d = nd - 1; d >= 0; --d) {
        const int o = (int)(r % O[d]); r /= O[d];
        const float s = sc[d] * (float)o;
        int a = (int)s;
        if (a > I[d] - 1) a = I[d] - 1;
        i0[d] = a; i1[d] = a + (a < I[d] - 1 ? 1 : 0);
        l1[d] = s - (float)a; l0[d] = 1.0f - l1[d];
      }
      const float* xp = x + (long long)p * nin;
      /* nested lerp, innermost axis first (ATen upsample_trilinear3d order) */
      float acc[8];
      const int nc = 1 << nd;
      for (int k = 0; k < nc; ++k) {
        long long off = 0;
        for (int d = 0; d < nd; ++d) off = off * I[d] + (((k >> (nd - 1 - d)) & 1) ? i1[d] : i0[d]);
        acc[k] = pre_mul * xp[off];
      }
      int m = nc;
      for (int d = nd - 1; d >= 0; --d) {
        m >>= 1;
        for (int k = 0; k < m; ++k) acc[k] = l0[d] * acc[2 * k] + l1[d] * acc[2 * k + 1];
      }
      y[(long long)p * nout + v] = post_mul * acc[0];
    }
}

/* Local NCC. I,J (B,1,*S) nd in {2,3}; direct window sums (row-major accumulation), zero padding.
 * out[0]=loss, out[1]=sum(cc*mask), out[2]=normaliser; cc_out (nullable) receives cc per voxel. */
void orc_ncc(const float* I, const float* J, const float* mask, float* out, float* cc_out, int B, int nd,
             const int* S, int win, float eps, int reduction) {
  const int D = nd == 3 ? S[0] : 1, H = S[nd - 2], W = S[nd - 1];
  const int r = win / 2, rz = nd == 3 ? r : 0;
  const float wsz = nd == 3 ? (float)(win * win * win) : (float)(win * win);
  double acc = 0, macc = 0;
  for (int b = 0; b < B; ++b)
    for (int z = 0; z < D; ++z)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          float sI = 0, sJ = 0, sII = 0, sJJ = 0, sIJ = 0;
          for (int dz = -rz; dz <= rz; ++dz) {
            const int zz = z + dz;
            if (zz < 0 || zz >= D) continue;
            for (int dy = -r; dy <= r; ++dy) {
              const int yy = y + dy;
              if (yy < 0 || yy >= H) continue;
              for (int dx = -r; dx <= r; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                const long long o = (((long long)b * D + zz) * H + yy) * W + xx;
                const float i = I[o], j = J[o];
                sI += i; sJ += j; sII += i * i; sJJ += j * j; sIJ += i * j;
              }
            }
          }
          const float uI = sI / wsz, uJ = sJ / wsz;
          const float cross = sIJ - uJ * sI - uI * sJ + uI * uJ * wsz;
          const float ivar = sII - 2 * uI * sI + uI * uI * wsz;
          const float jvar = sJJ - 2 * uJ * sJ + uJ * uJ * wsz;
          const float cc = cross * cross / (ivar * jvar + eps);
          const long long o = (((long long)b * D + z) * H + y) * W + x;
          if (cc_out) cc_out[o] = cc;
          if (mask) { acc += (double)(cc * mask[o]); macc += (double)mask[o]; }
          else acc += (double)cc;
        }
  const double norm = mask ? macc : (double)B * D * H * W;
  float loss;
  if (mask && macc == 0.0) loss = 0.f;
  else if (reduction == 0) loss = -sqrtf((float)(acc / norm));
  else loss = -(float)(acc / norm);
  out[0] = loss; out[1] = (float)acc; out[2] = (float)norm;
}

/* Grad_Loss / smooothing_loss: x as (P planes, *S); penalty 1 l1, 2 l2 */
float orc_grad(const float* x, int P, int nd, const int* S, int penalty, float loss_mult) {
  long long nvox = 1, stride[3];
  for (int d = 0; d < nd; ++d) nvox *= S[d];
  long long s = 1;
  for (int d = nd - 1; d >= 0; --d) { stride[d] = s; s *= S[d]; }
  float total = 0.f;
  for (int d = nd - 1; d >= 0; --d) { /* x axis first: mean(dx) + mean(dy) + mean(dz) */
    double acc = 0; long long cnt = 0;
    for (int p = 0; p < P; ++p)
      for (long long v = 0; v < nvox; ++v) {
        const int pos = (int)((v / stride[d]) % S[d]);
        if (pos + 1 >= S[d]) continue;
        float df = fabsf(x[(long long)p * nvox + v + stride[d]] - x[(long long)p * nvox + v]);
        if (penalty == 2) df = df * df;
        acc += (double)df; ++cnt;
      }
    total += (float)(acc / (double)cnt);
  }
  return total / (float)nd * loss_mult;
}

/* masked L1: mask nullable uint8; or (mu > thr) | (mv > thr); out = {loss, sum(mask)} */
void orc_l1_masked(const float* a, const float* b, const uint8_t* mask, const float* mu, const float* mv,
                   float thr, float* out, long long n) {
  double s = 0, m = 0;
  for (long long i = 0; i < n; ++i) {
    float mk = 1.f;
    if (mask) mk = mask[i] ? 1.f : 0.f;
    else if (mu) mk = (mu[i] > thr || mv[i] > thr) ? 1.f : 0.f;
    s += (double)(fabsf(a[i] - b[i]) * mk);
    m += (double)mk;
  }
  if (!mask && !mu) m = (double)n;
  out[0] = m == 0.0 ? 0.f : (float)(s / m);
  out[1] = (float)m;
}

/* ===================================================================================
 * Network layers (NCHW / NCDHW planar fp32, as the reference keeps them)
 *   orc_conv          nn.Conv2d / nn.Conv3d (zero padding, stride)     models/networks.py:983..1214,
 *                                                                       vxm networks.py:1515,1077
 *   orc_instnorm      nn.InstanceNorm2d(affine=False), eps 1e-5        models/networks.py:125
 *   orc_pad_reflect   nn.ReflectionPad2d                               models/networks.py:982,1022,1193
 *   orc_blur_down     Downsample (reflect pad 1, [1,2,1]^2/16, s2)     models/networks.py:37-60
 *   orc_blur_up       Upsample (replicate pad, conv_transpose 4x4, crop) models/networks.py:73-93
 *   orc_upsample_nn   nn.Upsample(scale_factor=2, 'nearest')           vxm networks.py:62,100
 *   orc_patchnce      PatchNCELoss.forward                             models/patchnce.py:14-55
 * =================================================================================== */

/* x (N,Cin,*I) w (Cout,Cin,*K) b (Cout, nullable) -> y (N,Cout,*O); nd 2 or 3 (2-D: leading dims 1) */
void orc_conv(const float* x, const float* w, const float* b, float* y, int N, int Cin, int Cout, int nd,
              const int* I_, const int* K_, int stride, int pad) {
  int I[3] = {1, 1, 1}, K[3] = {1, 1, 1}, O[3] = {1, 1, 1}, P[3] = {0, 0, 0};
  for (int a = 0; a < nd; ++a) {
    I[a + 3 - nd] = I_[a]; K[a + 3 - nd] = K_[a]; P[a + 3 - nd] = pad;
    O[a + 3 - nd] = (I_[a] + 2 * pad - K_[a]) / stride + 1;
  }
  const long long ivox = (long long)I[0] * I[1] * I[2], ovox = (long long)O[0] * O[1] * O[2];
  const long long kvox = (long long)K[0] * K[1] * K[2];
  for (int n = 0; n < N; ++n)
    for (int co = 0; co < Cout; ++co)
      for (int oz = 0; oz < O[0]; ++oz)
        for (int oy = 0; oy < O[1]; ++oy)
          for (int ox = 0; ox < O[2]; ++ox) {
            double acc = 0.0; /* wide accumulator: the reference's blocked fp32 sums are order-dependent */
            for (int ci = 0; ci < Cin; ++ci)
              for (int kz = 0; kz < K[0]; ++kz) {
                const int iz = oz * stride - P[0] + kz;
                if (iz < 0 || iz >= I[0]) continue;
                for (int ky = 0; ky < K[1]; ++ky) {
                  const int iy = oy * stride - P[1] + ky;
                  if (iy < 0 || iy >= I[1]) continue;
                  for (int kx = 0; kx < K[2]; ++kx) {
                    const int ix = ox * stride - P[2] + kx;
                    if (ix < 0 || ix >= I[2]) continue;
                    acc += (double)x[((long long)n * Cin + ci) * ivox + ((long long)iz * I[1] + iy) * I[2] + ix] *
                           (double)w[((long long)co * Cin + ci) * kvox + ((long long)kz * K[1] + ky) * K[2] + kx];
                  }
                }
              }
            if (b) acc += (double)b[co];
            y[((long long)n * Cout + co) * ovox + ((long long)oz * O[1] + oy) * O[2] + ox] = (float)acc;
          }
}

/* per (n,c) plane of HW elements: (x - mean) / sqrt(var_biased + eps) */
void orc_instnorm(const float* x, float* y, int NC, long long HW, float eps) {
  for (int p = 0; p < NC; ++p) {
    const float* xp = x + (long long)p * HW;
    double s = 0, ss = 0;
    for (long long i = 0; i < HW; ++i) s += xp[i];
    const double m = s / (double)HW;
    for (long long i = 0; i < HW; ++i) ss += ((double)xp[i] - m) * ((double)xp[i] - m);
    const double r = 1.0 / sqrt(ss / (double)HW + (double)eps);
    for (long long i = 0; i < HW; ++i) y[(long long)p * HW + i] = (float)(((double)xp[i] - m) * r);
  }
}

static int reflect_i(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * (n - 1) - i; return i; }
static int clamp_i(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

void orc_pad_reflect(const float* x, float* y, int NC, int H, int W, int p) {
  const int HP = H + 2 * p, WP = W + 2 * p;
  for (int c = 0; c < NC; ++c)
    for (int h = 0; h < HP; ++h)
      for (int w = 0; w < WP; ++w)
        y[((long long)c * HP + h) * WP + w] = x[((long long)c * H + reflect_i(h - p, H)) * W + reflect_i(w - p, W)];
}

/* F.conv2d(reflect_pad(x,1), [1,2,1]x[1,2,1]/16, stride 2, groups=C) */
void orc_blur_down(const float* x, float* y, int NC, int H, int W) {
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  static const float f[3] = {1.f, 2.f, 1.f};
  for (int c = 0; c < NC; ++c)
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow) {
        float acc = 0.f;
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b)
            acc += x[((long long)c * H + reflect_i(2 * oh - 1 + a, H)) * W + reflect_i(2 * ow - 1 + b, W)] *
                   (f[a] * f[b] / 16.f);
        y[((long long)c * OH + oh) * OW + ow] = acc;
      }
}

/* conv_transpose2d(replicate_pad(x,1), [1,3,3,1]x[1,3,3,1]/64*4, stride 2, padding 2)[1:,1:][:-1,:-1],
 * restated literally: scatter every padded input sample through the 4x4 filter, then crop. */
void orc_blur_up(const float* x, float* y, int NC, int H, int W) {
  static const float f[4] = {1.f, 3.f, 3.f, 1.f};
  const int PH = H + 2, PW = W + 2;
  const int TH = (PH - 1) * 2 - 4 + 4, TW = (PW - 1) * 2 - 4 + 4; /* conv_transpose output */
  const int OH = 2 * H, OW = 2 * W;
  float* tmp = (float*)malloc(sizeof(float) * (size_t)TH * TW);
  for (int c = 0; c < NC; ++c) {
    memset(tmp, 0, sizeof(float) * (size_t)TH * TW);
    for (int i = 0; i < PH; ++i)
      for (int j = 0; j < PW; ++j) {
        const float v = x[((long long)c * H + clamp_i(i - 1, H)) * W + clamp_i(j - 1, W)];
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) {
            const int oi = 2 * i - 2 + a, oj = 2 * j - 2 + b;
            if (oi < 0 || oi >= TH || oj < 0 || oj >= TW) continue;
            tmp[(long long)oi * TW + oj] += v * (f[a] * f[b] / 64.f * 4.f);
          }
      }
    for (int oh = 0; oh < OH; ++oh)
      for (int ow = 0; ow < OW; ++ow) y[((long long)c * OH + oh) * OW + ow] = tmp[(long long)(oh + 1) * TW + ow + 1];
  }
  free(tmp);
}

/* nearest x2: x (NC,*S) -> y (NC,*2S), nd 2 or 3 */
void orc_upsample_nn(const float* x, float* y, int NC, int nd, const int* S_) {
  int S[3] = {1, 1, 1};
  for (int a = 0; a < nd; ++a) S[a + 3 - nd] = S_[a];
  const int O0 = nd == 3 ? 2 * S[0] : 1, O1 = 2 * S[1], O2 = 2 * S[2];
  for (int c = 0; c < NC; ++c)
    for (int z = 0; z < O0; ++z)
      for (int yy = 0; yy < O1; ++yy)
        for (int xx = 0; xx < O2; ++xx)
          y[(((long long)c * O0 + z) * O1 + yy) * O2 + xx] =
              x[(((long long)c * S[0] + (nd == 3 ? z / 2 : 0)) * S[1] + yy / 2) * S[2] + xx / 2];
}

/* q,k (B*P, D): loss[r] = logsumexp([q.k_pos, q.k_j (j != i; -10 at j == i)] / T) - q.k_pos / T */
void orc_patchnce(const float* q, const float* k, float* loss, int B, int P, int D, float T) {
  float* row = (float*)malloc(sizeof(float) * (size_t)(P + 1));
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < P; ++i) {
      const float* qi = q + ((long long)b * P + i) * D;
      for (int j = 0; j < P; ++j) {
        const float* kj = k + ((long long)b * P + j) * D;
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc += qi[d] * kj[d];
        row[j + 1] = acc;
      }
      row[0] = row[i + 1];
      row[i + 1] = -10.0f;
      float mx = -INFINITY;
      for (int j = 0; j <= P; ++j) { row[j] = row[j] / T; if (row[j] > mx) mx = row[j]; }
      double s = 0;
      for (int j = 0; j <= P; ++j) s += exp((double)row[j] - mx);
      loss[(long long)b * P + i] = (float)((double)mx + log(s) - (double)row[0]);
    }
  free(row);
}
