// "Thin" 2-D convolutions: the 7x7 stem (Cin = 1 -> 64, models/networks.py:983) and head (64 -> Cout = 1,
// :1023) of ResnetGenerator, forward, data gradient and weight gradient.  Their GEMM view has N = 1 or
// K = 49: no tensor-core tile fits, and they are bound by the one multi-channel tensor they
// stream (0.54 GB at batch 32), so they run as direct convolutions on the CUDA cores with shared-memory
// halo tiles and register reuse instead of as a degenerate implicit GEMM.
//
// Three kernels cover the six cases (s = the single-channel image, v / out = the C-channel tensor):
//   expand : out[n,oh,ow,c] = act(b[c] + sum_tap s[n,oh+r-ph,ow+q-pw] * W[tap][c])     stem fwd, head dgrad
//   reduce : out[n,oh,ow]   = act(b + sum_tap sum_c v[n,oh+r-ph,ow+q-pw,c] * W[tap][c]) head fwd, stem dgrad
//   wgrad  : dW[tap][c]    += sum_p s[p+tap-(ph,pw)] * v[p][c]   (db[c] += sum_p v[p][c]) stem / head wgrad
// Data gradients are the same correlations with flipped taps and p' = K-1-p; the head's weight gradient
// iterates over the input domain with s = dy.  Sources are zero outside their bounds.
#include "common.cuh"
#include "umma.cuh"
#include "dfmir_b200.h"
#include <stdlib.h>

namespace thin {

struct ThinP {
  int N, C, OH, OW;      // iteration domain (output pixels; for wgrad: the domain of v)
  int SH, SW;            // spatial size of the tap-shifted source
  int ph, pw, flip, act;
  long long s_n, s_h, s_w;   // strides of the shifted source (single channel for expand / wgrad; v for reduce, c stride 1)
  long long o_n, o_h, o_w;   // strides of out (expand / reduce) or of v (wgrad), c stride 1
  int tiles_h, tiles_w;
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == DFMIR_ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  if (act == DFMIR_ACT_TANH) return tanhf(v);
  if (act == DFMIR_ACT_RELU) return v > 0.f ? v : 0.f;
  return v;
}

// ------------------------------------------------------------------ expand: 1 -> C channels
// block = 32 x 8 output pixels, one pixel per thread, CT output channels in registers.
template <int K, int CT>
__global__ void __launch_bounds__(256)
thin_expand_kernel(const float* __restrict__ s, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out, const ThinP p) {
  constexpr int TW = 32, TH = 8, SWD = TW + K - 1, SHT = TH + K - 1, TAPS = K * K;
  __shared__ float st[SHT][SWD + 1];
  __shared__ __align__(16) float wt[TAPS][CT];
  int tile = blockIdx.x;
  const int tw_i = tile % p.tiles_w; tile /= p.tiles_w;
  const int th_i = tile % p.tiles_h; const int n = tile / p.tiles_h;
  const int h0 = th_i * TH, w0 = tw_i * TW, c0 = blockIdx.y * CT;
  const int t = threadIdx.x, tx = t & 31, ty = t >> 5;
  const float* sb = s + (long long)n * p.s_n;
  for (int e = t; e < SHT * SWD; e += 256) {
    const int r = e / SWD, c = e - r * SWD;
    const int ih = h0 + r - p.ph, iw = w0 + c - p.pw;
    st[r][c] = (ih >= 0 && ih < p.SH && iw >= 0 && iw < p.SW) ? __ldg(sb + ih * p.s_h + iw * p.s_w) : 0.f;
  }
  for (int e = t; e < TAPS * CT; e += 256) {
    const int tp = e / CT, c = e - tp * CT;
    wt[tp][c] = (c0 + c < p.C) ? __ldg(w + (long long)(p.flip ? TAPS - 1 - tp : tp) * p.C + c0 + c) : 0.f;
  }
  __syncthreads();
  float acc[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) acc[c] = (bias && c0 + c < p.C) ? __ldg(bias + c0 + c) : 0.f;
#pragma unroll 1
  for (int r = 0; r < K; ++r) {
#pragma unroll
    for (int q = 0; q < K; ++q) {
      const float sv = st[ty + r][tx + q];
      const float4* wr = reinterpret_cast<const float4*>(wt[r * K + q]);
#pragma unroll
      for (int c4 = 0; c4 < CT / 4; ++c4) {
        const float4 wv = wr[c4];
        acc[4 * c4] = fmaf(sv, wv.x, acc[4 * c4]);
        acc[4 * c4 + 1] = fmaf(sv, wv.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(sv, wv.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(sv, wv.w, acc[4 * c4 + 3]);
      }
    }
  }
  const int oh = h0 + ty, ow = w0 + tx;
  if (oh >= p.OH || ow >= p.OW) return;
  float* op = out + (long long)n * p.o_n + (long long)oh * p.o_h + (long long)ow * p.o_w + c0;
  if (c0 + CT <= p.C && (((uintptr_t)op) & 15) == 0) {
#pragma unroll
    for (int c4 = 0; c4 < CT / 4; ++c4)
      reinterpret_cast<float4*>(op)[c4] = make_float4(act_apply(acc[4 * c4], p.act), act_apply(acc[4 * c4 + 1], p.act),
                                                      act_apply(acc[4 * c4 + 2], p.act), act_apply(acc[4 * c4 + 3], p.act));
  } else {
#pragma unroll
    for (int c = 0; c < CT; ++c)
      if (c0 + c < p.C) op[c] = act_apply(acc[c], p.act);
  }
}


// ------------------------------------------------------------------ expand on the tensor cores (C = 64)
// The direct kernel above is FMA-bound: 49 x 64 = 3136 multiply-adds per pixel (0.45 ms for 32 images of 256^2 at
// ~40 % of the fp32 peak) to write 0.54 GB.  As a GEMM it is M = pixels, N = 64, K = 49: the CTA builds the im2col
// operand of a 16 x 8 pixel tile itself - 128 rows x 56 (49 + 7 zero) taps from a 22 x 14 staged patch of the single
// source channel, written in the K-major SWIZZLE_128B layout the tcgen05 descriptor expects - and the weights once.
// Arithmetic stays fp32-class: both operands are split hi = x & ~0x1fff (what the tensor core keeps of an fp32
// operand), lo = x - hi (exact), and three products hi*hi + hi*lo + lo*hi are accumulated (the dropped lo*lo term is
// 2^-22 relative) - 21 MMAs of 128 x 64 x 8 per tile, a few hundred clocks against the ~1400 the tile's 32 KB of output
// take at this SM's share of HBM.  Two CTAs per SM overlap each other's build / MMA / store phases.
namespace ut = umma;

constexpr int XK = 56;                       // taps padded to whole K = 8 slices
constexpr int XA_CHUNK = 128 * 128;          // one 32-tap chunk of the 128-row operand
constexpr int XB_CHUNK = 64 * 128;           // one 32-tap chunk of the 64-channel weight operand
constexpr int X_SMEM = 4 * XA_CHUNK + 4 * XB_CHUNK + 22 * 16 * 4 + 64 + 1024;

__device__ __forceinline__ uint32_t sw128_off(int row, int kk) {      // byte offset of element (row, kk < 32) in a chunk
  return (uint32_t)(row * 128 + ((((kk >> 2) ^ (row & 7)) << 4) | ((kk & 3) << 2)));
}

__global__ void __launch_bounds__(256, 2)
thin_expand_umma_kernel(const float* __restrict__ s, const float* __restrict__ w, const float* __restrict__ bias,
                        float* __restrict__ out, const ThinP p, int ntiles) {
  constexpr int K = 7, TAPS = 49, TH = 16, TW = 8, PH = TH + K - 1, PW = TW + K - 1;   // 22 x 14 patch
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sAhi = smem;                       // [2 chunks][128 rows][128 B]
  uint8_t* sAlo = sAhi + 2 * XA_CHUNK;
  uint8_t* sBhi = sAlo + 2 * XA_CHUNK;        // [2 chunks][64 rows][128 B]
  uint8_t* sBlo = sBhi + 2 * XB_CHUNK;
  float* patch = (float*)(sBlo + 2 * XB_CHUNK);            // [22][16]
  uint64_t* bar = (uint64_t*)(patch + PH * 16);
  uint32_t* tmem_slot = (uint32_t*)(bar + 1);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  if (t == 0) { ut::mbar_init(bar, 1); ut::fence_barrier_init(); }
  if (warp == 0) ut::tmem_alloc(tmem_slot, 64u);
  // weights: B[n = channel][k = tap], hi / lo, zero beyond the 49 taps and p.C channels
  for (int e = t; e < 64 * 64; e += 256) {
    const int n = e >> 6, k = e & 63;
    float v = 0.f;
    if (k < TAPS && n < p.C) v = __ldg(w + (long long)(p.flip ? TAPS - 1 - k : k) * p.C + n);
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const uint32_t off = (uint32_t)((k >> 5) * XB_CHUNK) + sw128_off(n, k & 31);
    *(float*)(sBhi + off) = hi;
    *(float*)(sBlo + off) = v - hi;
  }
  ut::tc_fence_before();
  __syncthreads();
  ut::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = ut::instr_desc_tf32(128, 64);
  const int quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane, th = row >> 3, tw = row & 7;
  float bv[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) bv[i] = (bias && half * 32 + i < p.C) ? __ldg(bias + half * 32 + i) : 0.f;
  uint32_t phase = 0;
  // the source patch of the NEXT tile is fetched into registers while this tile's MMAs run (its global-load latency
  // was the longest serial piece of a tile)
  constexpr int PER_T = (PH * PW + 255) / 256;
  float pre[PER_T];
  auto fetch = [&](int tile) {
    int r = tile;
    const int tw_i = r % p.tiles_w; r /= p.tiles_w;
    const int th_i = r % p.tiles_h; const int n = r / p.tiles_h;
    const float* sb = s + (long long)n * p.s_n;
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
      const int e = t + 256 * j;
      const int pr = e / PW, pc = e - pr * PW;
      const int ih = th_i * TH + pr - p.ph, iw = tw_i * TW + pc - p.pw;
      pre[j] = (e < PH * PW && tile < ntiles && ih >= 0 && ih < p.SH && iw >= 0 && iw < p.SW) ? __ldg(sb + ih * p.s_h + iw * p.s_w) : 0.f;
    }
  };
  fetch(blockIdx.x);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int r = tile;
    const int tw_i = r % p.tiles_w; r /= p.tiles_w;
    const int th_i = r % p.tiles_h; const int n = r / p.tiles_h;
    const int h0 = th_i * TH, w0 = tw_i * TW;
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
      const int e = t + 256 * j;
      if (e < PH * PW) patch[(e / PW) * 16 + e % PW] = pre[j];
    }
    __syncthreads();
    // im2col rows: warp -> pixels warp, warp + 8, ...; lane = tap within the 32-tap chunk (conflict-free swizzled stores)
    for (int it = 0; it < 32; ++it) {
      const int px = warp + 8 * (it >> 1), c = it & 1;
      const int k = 32 * c + lane;
      float v = 0.f;
      if (k < TAPS) v = patch[((px >> 3) + k / K) * 16 + (px & 7) + k % K];
      const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      const uint32_t off = (uint32_t)(c * XA_CHUNK) + sw128_off(px, lane);
      *(float*)(sAhi + off) = hi;
      *(float*)(sAlo + off) = v - hi;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (warp == 0) {         // converged warp, one elected lane issues (no per-MMA ELECT / BRA.U.ANY wrapper)
      ut::tc_fence_after();
      const uint64_t ahi = ut::smem_desc_sw128(ut::smem_u32(sAhi), 16, 1024), alo = ut::smem_desc_sw128(ut::smem_u32(sAlo), 16, 1024);
      const uint64_t bhi = ut::smem_desc_sw128(ut::smem_u32(sBhi), 16, 1024), blo = ut::smem_desc_sw128(ut::smem_u32(sBlo), 16, 1024);
#pragma unroll
      for (int seg = 0; seg < 3; ++seg) {
        const uint64_t ad = seg == 2 ? alo : ahi, bd = seg == 1 ? blo : bhi;
#pragma unroll
        for (int ks = 0; ks < XK / 8; ++ks) {
          const int c = ks >> 2, k = ks & 3;
          const uint64_t a = ad + (uint64_t)(c * (XA_CHUNK / 16) + 2 * k), b = bd + (uint64_t)(c * (XB_CHUNK / 16) + 2 * k);
          if ((seg | ks) != 0) ut::umma_tf32_elect<true>(tmem_base, a, b, idesc);
          else ut::umma_tf32_elect<false>(tmem_base, a, b, idesc);
        }
      }
      ut::umma_commit_elect(bar);
    }
    fetch(tile + (int)gridDim.x);
    ut::mbar_wait(bar, phase);
    phase ^= 1;
    ut::tc_fence_after();
    {
      // epilogue through shared memory (the operand tiles are free once the MMAs have retired): a thread holds 32
      // channels of ONE pixel, 256 bytes apart from its neighbour's - stored directly that is 32 half-used sectors per
      // instruction (lg_throttle-bound: 18 % of HBM); staged, every warp instruction writes 512 contiguous bytes
      float v[32];
      ut::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 32), v);
      float* stage = reinterpret_cast<float*>(sAhi);             // [128 pixels][68 floats]: rows padded against bank conflicts
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(stage + row * 68 + half * 32 + 4 * i) =
            make_float4(act_apply(v[4 * i] + bv[4 * i], p.act), act_apply(v[4 * i + 1] + bv[4 * i + 1], p.act),
                        act_apply(v[4 * i + 2] + bv[4 * i + 2], p.act), act_apply(v[4 * i + 3] + bv[4 * i + 3], p.act));
      ut::tc_fence_before();
      __syncthreads();
      const bool vec = p.C == 64 && (p.o_w & 3) == 0 && (p.o_h & 3) == 0 && (p.o_n & 3) == 0 && (((uintptr_t)out) & 15) == 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = t + 256 * j;                             // float4 index in the tile: pixel * 16 + channel quad
        const int px = idx >> 4, c4 = idx & 15;
        const int oh = h0 + (px >> 3), ow = w0 + (px & 7);
        if (oh >= p.OH || ow >= p.OW) continue;
        const float4 val = *reinterpret_cast<const float4*>(stage + px * 68 + 4 * c4);
        float* op = out + (long long)n * p.o_n + (long long)oh * p.o_h + (long long)ow * p.o_w + 4 * c4;
        if (vec) *reinterpret_cast<float4*>(op) = val;
        else {
          if (4 * c4 < p.C) op[0] = val.x;
          if (4 * c4 + 1 < p.C) op[1] = val.y;
          if (4 * c4 + 2 < p.C) op[2] = val.z;
          if (4 * c4 + 3 < p.C) op[3] = val.w;
        }
      }
    }
    ut::tc_fence_before();
    __syncthreads();         // the accumulator, the operand tiles and the patch are reused by the next tile
  }
  if (warp == 0) ut::tmem_dealloc(tmem_base, 64u);
}

// expand through the tensor-core kernel when it applies (7 x 7, 17..64 channels); returns false to fall back
static bool launch_expand_umma(const float* s, const float* w, const float* bias, float* out, ThinP p, cudaStream_t st) {
  static const int on = getenv("DFMIR_THIN_TC") ? atoi(getenv("DFMIR_THIN_TC")) : 1;
  if (!on || p.C <= 16 || p.C > 64) return false;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(thin_expand_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X_SMEM) != cudaSuccess) return false;
    attr = true;
  }
  p.tiles_h = (p.OH + 15) / 16; p.tiles_w = (p.OW + 7) / 8;
  const long long ntiles = (long long)p.N * p.tiles_h * p.tiles_w;
  if (ntiles <= 0 || ntiles > 0x7fffffff) return false;
  long long grid = 2LL * dfmir_num_sms();
  if (grid > ntiles) grid = ntiles;
  thin_expand_umma_kernel<<<(unsigned)grid, 256, X_SMEM, st>>>(s, w, bias, out, p, (int)ntiles);
  return true;
}

// ------------------------------------------------------------------ reduce: C -> 1 channel
// block = 32 x 32 output pixels, 4 warps; lane = column, each thread 8 vertically adjacent pixels, so a
// column of 8+K-1 source float4s (4 channels) is loaded once per tap column and reused for 8 x K taps.
template <int K>
__global__ void __launch_bounds__(128)
thin_reduce_kernel(const float* __restrict__ v, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out, const ThinP p) {
  constexpr int TW = 32, TH = 32, PH = 8, HW_ = TW + K - 1, HH = TH + K - 1, CC = 8, TAPS = K * K;
  __shared__ float4 tile[CC / 4][HH][HW_];
  __shared__ float4 wsm[TAPS][CC / 4];
  int tl = blockIdx.x;
  const int tw_i = tl % p.tiles_w; tl /= p.tiles_w;
  const int th_i = tl % p.tiles_h; const int n = tl / p.tiles_h;
  const int h0 = th_i * TH, w0 = tw_i * TW;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float* vb = v + (long long)n * p.s_n;
  float acc[PH];
#pragma unroll
  for (int j = 0; j < PH; ++j) acc[j] = 0.f;
  for (int c0 = 0; c0 < p.C; c0 += CC) {
    __syncthreads();
    for (int e = t; e < HH * HW_ * (CC / 4); e += 128) {
      const int c4 = e % (CC / 4), px = e / (CC / 4);
      const int row = px / HW_, col = px - row * HW_;
      const int ih = h0 + row - p.ph, iw = w0 + col - p.pw;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ih >= 0 && ih < p.SH && iw >= 0 && iw < p.SW && c0 + c4 * 4 < p.C)
        val = __ldg(reinterpret_cast<const float4*>(vb + ih * p.s_h + iw * p.s_w + c0 + c4 * 4));
      tile[c4][row][col] = val;
    }
    for (int e = t; e < TAPS * (CC / 4); e += 128) {
      const int tp = e / (CC / 4), c4 = e - tp * (CC / 4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + c4 * 4 < p.C) val = __ldg(reinterpret_cast<const float4*>(w + (long long)(p.flip ? TAPS - 1 - tp : tp) * p.C + c0 + c4 * 4));
      wsm[tp][c4] = val;
    }
    __syncthreads();
#pragma unroll
    for (int c4 = 0; c4 < CC / 4; ++c4) {
#pragma unroll 1
      for (int q = 0; q < K; ++q) {
        float4 xs[PH + K - 1];
#pragma unroll
        for (int i = 0; i < PH + K - 1; ++i) xs[i] = tile[c4][warp * PH + i][lane + q];
#pragma unroll
        for (int r = 0; r < K; ++r) {
          const float4 wv = wsm[r * K + q][c4];
#pragma unroll
          for (int j = 0; j < PH; ++j) {
            const float4 x = xs[j + r];
            acc[j] = fmaf(x.x, wv.x, acc[j]);
            acc[j] = fmaf(x.y, wv.y, acc[j]);
            acc[j] = fmaf(x.z, wv.z, acc[j]);
            acc[j] = fmaf(x.w, wv.w, acc[j]);
          }
        }
      }
    }
  }
  const float b = bias ? __ldg(bias) : 0.f;
  const int ow = w0 + lane;
  if (ow >= p.OW) return;
#pragma unroll
  for (int j = 0; j < PH; ++j) {
    const int oh = h0 + warp * PH + j;
    if (oh < p.OH) out[(long long)n * p.o_n + (long long)oh * p.o_h + (long long)ow * p.o_w] = act_apply(acc[j] + b, p.act);
  }
}

// ------------------------------------------------------------------ weight gradient on the tensor cores (C = 64)
// dW[tap][c] = sum_p s[p + tap] v[p][c] is a GEMM with the PIXELS as the reduction: M = taps (49 of the 128 operand
// rows; row 49 is all ones, so the bias gradient sum_p v[p][c] falls out of the same product), N = 64 channels,
// K = the 64 pixels of an 8 x 8 tile.  A = the im2col patch TRANSPOSED (row = tap, 32 pixels per 128-byte row, K-major
// SWIZZLE_128B), built by the CTA from a 14 x 14 staged patch; B = the v tile exactly as it lies in memory (row = pixel,
// 2 column groups of 32 channels: MN-major, written by TMA in SWIZZLE_128B_ATOM_32B).  3xTF32 as in the expand kernel:
// both operands split hi / lo in shared memory (the split of the TMA-written tile is elementwise, so layout-blind).
// The accumulator stays in TMEM over all the tiles of the CTA; one atomic pass at the end.
constexpr int WA_CHUNK = 64 * 128;            // 64 tap rows x 32 pixels (the MMA reads 128 rows: the upper 64 alias what follows)
constexpr int WB_GROUP = 64 * 128;            // 64 pixels x 32 channels
constexpr int W_SMEM = 4 * WA_CHUNK + 4 * WB_GROUP + 14 * 16 * 4 + 64 + 1024;

__global__ void __launch_bounds__(256, 3)
thin_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmV, const float* __restrict__ s, float* __restrict__ dw,
                       float* __restrict__ db, const ThinP p, int ntiles) {
  constexpr int K = 7, TAPS = 49, TH = 8, TW = 8, PH = TH + K - 1, PW = TW + K - 1;    // 14 x 14 patch
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sAhi = smem;                       // [2 chunks][64 rows][128 B]
  uint8_t* sAlo = sAhi + 2 * WA_CHUNK;
  uint8_t* sBhi = sAlo + 2 * WA_CHUNK;        // [2 channel groups][64 pixel rows][128 B], written by TMA
  uint8_t* sBlo = sBhi + 2 * WB_GROUP;
  float* patch = (float*)(sBlo + 2 * WB_GROUP);             // [14][16]
  uint64_t* tma_bar = (uint64_t*)(patch + PH * 16);
  uint64_t* mma_bar = tma_bar + 1;
  uint32_t* tmem_slot = (uint32_t*)(mma_bar + 1);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  if (t == 0) { ut::prefetch_tmap(&tmV); ut::mbar_init(tma_bar, 1); ut::mbar_init(mma_bar, 1); ut::fence_barrier_init(); }
  if (warp == 0) ut::tmem_alloc(tmem_slot, 64u);
  // constant operand rows: 49 = ones (bias gradient), 50..63 = zero, in both chunks; lo parts zero
  for (int e = t; e < 2 * 15 * 32; e += 256) {
    const int c = e / (15 * 32), rr = (e / 32) % 15, kk = e & 31;
    const uint32_t off = (uint32_t)(c * WA_CHUNK) + sw128_off(TAPS + rr, kk);
    *(float*)(sAhi + off) = rr == 0 ? 1.f : 0.f;
    *(float*)(sAlo + off) = 0.f;
  }
  ut::tc_fence_before();
  __syncthreads();
  ut::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = ut::instr_desc_tf32(128, 64, 0, 1);        // A K-major, B MN-major

  constexpr int PER_T = (PH * PW + 255) / 256;
  float pre[PER_T];
  auto coords = [&](int tile, int& n, int& h0, int& w0) {
    int r = tile;
    const int tw_i = r % p.tiles_w; r /= p.tiles_w;
    const int th_i = r % p.tiles_h; n = r / p.tiles_h;
    h0 = th_i * TH; w0 = tw_i * TW;
  };
  auto fetch = [&](int tile) {
    int n, h0, w0;
    coords(tile, n, h0, w0);
    const float* sb = s + (long long)n * p.s_n;
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
      const int e = t + 256 * j;
      const int pr = e / PW, pc = e - pr * PW;
      const int ih = h0 + pr - p.ph, iw = w0 + pc - p.pw;
      pre[j] = (e < PH * PW && tile < ntiles && ih >= 0 && ih < p.SH && iw >= 0 && iw < p.SW) ? __ldg(sb + ih * p.s_h + iw * p.s_w) : 0.f;
    }
  };
  fetch(blockIdx.x);
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    int n, h0, w0;
    coords(tile, n, h0, w0);
    if (it > 0) { ut::mbar_wait(mma_bar, (it - 1) & 1); ut::tc_fence_after(); }     // the previous tile's MMAs have read the operands
    if (t == 0) {
      ut::mbar_expect_tx(tma_bar, (uint32_t)(2 * WB_GROUP));
      ut::tma_load_4d(sBhi, &tmV, tma_bar, 0, w0, h0, n);
      ut::tma_load_4d(sBhi + WB_GROUP, &tmV, tma_bar, 32, w0, h0, n);
    }
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
      const int e = t + 256 * j;
      if (e < PH * PW) patch[(e / PW) * 16 + e % PW] = pre[j];
    }
    __syncthreads();
    fetch(tile + (int)gridDim.x);
    // A rows: (tap, 32-pixel chunk) items over the warps; lane = pixel within the chunk
    for (int item = warp; item < 2 * TAPS; item += 8) {
      const int tap = item >> 1, c = item & 1;
      const int px = 32 * c + lane;
      const float v = patch[((px >> 3) + tap / K) * 16 + (px & 7) + tap % K];
      const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      const uint32_t off = (uint32_t)(c * WA_CHUNK) + sw128_off(tap, lane);
      *(float*)(sAhi + off) = hi;
      *(float*)(sAlo + off) = v - hi;
    }
    ut::mbar_wait(tma_bar, it & 1);
    // split the v tile in place (elementwise: the swizzled position of an element does not matter)
    for (int e = t; e < 2 * WB_GROUP / 16; e += 256) {
      float4 v = *reinterpret_cast<float4*>(sBhi + 16 * e);
      float4 hi;
      hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
      hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
      *reinterpret_cast<float4*>(sBhi + 16 * e) = hi;
      *reinterpret_cast<float4*>(sBlo + 16 * e) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      ut::tc_fence_after();
      const uint64_t ahi = ut::smem_desc_sw128(ut::smem_u32(sAhi), 16, 1024), alo = ut::smem_desc_sw128(ut::smem_u32(sAlo), 16, 1024);
      // MN-major operand: 32-channel column groups WB_GROUP apart, 4-row (pixel) groups 512 bytes apart
      const uint64_t bhi = ut::smem_desc(ut::smem_u32(sBhi), WB_GROUP, 512, ut::LAYOUT_SW128_BASE32B);
      const uint64_t blo = ut::smem_desc(ut::smem_u32(sBlo), WB_GROUP, 512, ut::LAYOUT_SW128_BASE32B);
#pragma unroll
      for (int seg = 0; seg < 3; ++seg) {
        const uint64_t ad = seg == 2 ? alo : ahi, bd = seg == 1 ? blo : bhi;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {              // 8 pixels per K slice
          const uint64_t a = ad + (uint64_t)((ks >> 2) * (WA_CHUNK / 16) + 2 * (ks & 3)), b = bd + (uint64_t)(8 * 8 * ks);
          if ((seg | ks) != 0) ut::umma_tf32_elect<true>(tmem_base, a, b, idesc);
          else ut::umma_tf32_elect_rt(tmem_base, a, b, idesc, (uint32_t)(it != 0));
        }
      }
      ut::umma_commit_elect(mma_bar);
    }
  }
  // epilogue: TMEM lane = tap (49 = the ones row), column = channel
  if (it > 0) { ut::mbar_wait(mma_bar, (it - 1) & 1); ut::tc_fence_after(); }
  if (warp < 2 && it > 0) {
    const int rowi = warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float v[32];
      ut::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (rowi < TAPS) {
        float* dst = dw + (long long)(p.flip ? TAPS - 1 - rowi : rowi) * p.C + c0;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < p.C) atomicAdd(dst + i, v[i]);
      } else if (rowi == TAPS && db) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < p.C) atomicAdd(db + c0 + i, v[i]);
      }
    }
  }
  ut::tc_fence_before();
  __syncthreads();
  if (warp == 0) ut::tmem_dealloc(tmem_base, 64u);
}

static bool launch_wgrad_umma(const float* s, const float* v, float* dw, float* db, ThinP p, cudaStream_t st) {
  static const int on = getenv("DFMIR_THIN_TC") ? atoi(getenv("DFMIR_THIN_TC")) : 1;
  if (!on || p.C != 64 || (((uintptr_t)v) & 15) || (p.o_n & 3) || (p.o_h & 3) || (p.o_w & 3)) return false;
  CUtensorMap tm;
  const long long as[4] = {p.o_n, p.o_h, p.o_w, 1};
  if (ut::encode_act_map(&tm, v, as, p.C, p.OW, p.OH, p.N, 8, 8, "dfmir_conv_wgrad(thin)", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) != DFMIR_OK) return false;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(thin_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM) != cudaSuccess) return false;
    attr = true;
  }
  p.tiles_h = (p.OH + 7) / 8; p.tiles_w = (p.OW + 7) / 8;
  const long long ntiles = (long long)p.N * p.tiles_h * p.tiles_w;
  if (ntiles <= 0 || ntiles > 0x7fffffff) return false;
  long long grid = 3LL * dfmir_num_sms();
  if (grid > ntiles) grid = ntiles;
  thin_wgrad_umma_kernel<<<(unsigned)grid, 256, W_SMEM, st>>>(tm, s, dw, db, p, (int)ntiles);
  return true;
}

// ------------------------------------------------------------------ weight gradient
// thread = (channel c, pixel group g); a unit is 8 rows x 32 columns of v's domain, group g owns 8
// columns.  Per row the thread keeps 8 v values and, per tap row, 8+K-1 source values in registers
// (broadcast shared-memory loads), and accumulates all K*K taps of its channel in registers.
template <int K>
__global__ void __launch_bounds__(256)
thin_wgrad_kernel(const float* __restrict__ s, const float* __restrict__ v, float* __restrict__ dw, float* __restrict__ db,
                  const ThinP p, int units) {
  constexpr int TW = 32, TH = 8, SWD = TW + K - 1, SHT = TH + K - 1, TAPS = K * K, CT = 64, G = 4, PW = TW / G;
  __shared__ __align__(16) float st[SHT][SWD + 2];
  __shared__ float red[TAPS + 1][CT];
  const int t = threadIdx.x, cl = t % CT, g = t / CT;
  const int c = blockIdx.y * CT + cl;
  const bool c_ok = c < p.C;
  float acc[TAPS];
#pragma unroll
  for (int i = 0; i < TAPS; ++i) acc[i] = 0.f;
  float bacc = 0.f;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    int tl = u;
    const int tw_i = tl % p.tiles_w; tl /= p.tiles_w;
    const int th_i = tl % p.tiles_h; const int n = tl / p.tiles_h;
    const int h0 = th_i * TH, w0 = tw_i * TW;
    const float* sb = s + (long long)n * p.s_n;
    __syncthreads();
    for (int e = t; e < SHT * SWD; e += 256) {
      const int r = e / SWD, cc = e - r * SWD;
      const int ih = h0 + r - p.ph, iw = w0 + cc - p.pw;
      st[r][cc] = (ih >= 0 && ih < p.SH && iw >= 0 && iw < p.SW) ? __ldg(sb + ih * p.s_h + iw * p.s_w) : 0.f;
    }
    __syncthreads();
    const float* vb = v + (long long)n * p.o_n + c;
#pragma unroll 1
    for (int row = 0; row < TH; ++row) {
      const int oh = h0 + row;
      if (oh >= p.OH) break;
      float vv[PW];
#pragma unroll
      for (int j = 0; j < PW; ++j) {
        const int ow = w0 + g * PW + j;
        vv[j] = (c_ok && ow < p.OW) ? __ldg(vb + (long long)oh * p.o_h + (long long)ow * p.o_w) : 0.f;
        bacc += vv[j];
      }
#pragma unroll
      for (int r = 0; r < K; ++r) {
        float sr[PW + K - 1];
#pragma unroll
        for (int i = 0; i < PW + K - 1; ++i) sr[i] = st[row + r][g * PW + i];
#pragma unroll
        for (int q = 0; q < K; ++q)
#pragma unroll
          for (int j = 0; j < PW; ++j) acc[r * K + q] = fmaf(sr[j + q], vv[j], acc[r * K + q]);
      }
    }
  }
  // combine the G pixel groups in shared memory, then one atomic per (tap, channel) per block
  for (int gg = 0; gg < G; ++gg) {
    __syncthreads();
    if (g == gg) {
#pragma unroll
      for (int i = 0; i < TAPS; ++i) red[i][cl] = (gg == 0 ? 0.f : red[i][cl]) + acc[i];
      red[TAPS][cl] = (gg == 0 ? 0.f : red[TAPS][cl]) + bacc;
    }
  }
  __syncthreads();
  for (int e = t; e < TAPS * CT; e += 256) {
    const int tp = e / CT, cc = e - tp * CT;
    const int ch = blockIdx.y * CT + cc;
    if (ch < p.C) atomicAdd(dw + (long long)(p.flip ? TAPS - 1 - tp : tp) * p.C + ch, red[tp][cc]);
  }
  if (db && t < CT && blockIdx.y * CT + t < p.C) atomicAdd(db + blockIdx.y * CT + t, red[TAPS][t]);
}

// out[0] += sum of x[0..n)
__global__ void __launch_bounds__(256)
sum_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
  __shared__ float sm[32];
  float a = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a += x[i];
  a = block_sum(a, sm);
  if (threadIdx.x == 0) atomicAdd(out, a);
}

}  // namespace thin

using thin::ThinP;

static bool thin_geom_ok(const dfmir_conv_desc* d) {
  return d->nd == 2 && d->stride == 1 && d->kernel[0] == 7 && d->kernel[1] == 7 && (d->Cin == 1 || d->Cout == 1) &&
         !(d->Cin == 1 && d->Cout == 1);
}
static bool vec4_ok(const float* ptr, const long long* st, int C) {   // strides {n,h,w,c}
  return st[3] == 1 && C % 4 == 0 && (((uintptr_t)ptr) & 15) == 0 && st[0] % 4 == 0 && st[1] % 4 == 0 && st[2] % 4 == 0;
}

// Returns 1 when the call was handled (rc holds the status), 0 when the generic kernel must run.
int dfmir_thin_fwd(const float* x, const float* w, const float* bias, float* y, const dfmir_conv_desc* d, cudaStream_t st,
                   int* rc) {
  if (!thin_geom_ok(d)) return 0;
  ThinP p{};
  p.N = d->N; p.OH = d->out_shape[0]; p.OW = d->out_shape[1]; p.SH = d->in_shape[0]; p.SW = d->in_shape[1];
  p.ph = d->pad[0]; p.pw = d->pad[1]; p.flip = 0; p.act = d->act;
  p.s_n = d->x_strides[0]; p.s_h = d->x_strides[1]; p.s_w = d->x_strides[2];
  p.o_n = d->y_strides[0]; p.o_h = d->y_strides[1]; p.o_w = d->y_strides[2];
  if (d->Cin == 1) {
    if (d->y_strides[3] != 1) return 0;
    p.C = d->Cout; p.tiles_h = (p.OH + 7) / 8; p.tiles_w = (p.OW + 31) / 32;
    const unsigned gx = (unsigned)(p.N * p.tiles_h * p.tiles_w);
    // 32 output channels per thread (two CTAs per pixel tile for the 64-channel stem): 3 CTAs resident per SM
    // instead of 1 with 64 accumulators per thread
    if (thin::launch_expand_umma(x, w, bias, y, p, st)) {}
    else if (p.C > 16) thin::thin_expand_kernel<7, 32><<<dim3(gx, (p.C + 31) / 32), 256, 0, st>>>(x, w, bias, y, p);
    else thin::thin_expand_kernel<7, 16><<<dim3(gx, 1), 256, 0, st>>>(x, w, bias, y, p);
  } else {
    if (!vec4_ok(x, d->x_strides, d->Cin) || (((uintptr_t)w) & 15)) return 0;
    p.C = d->Cin; p.tiles_h = (p.OH + 31) / 32; p.tiles_w = (p.OW + 31) / 32;
    thin::thin_reduce_kernel<7><<<(unsigned)(p.N * p.tiles_h * p.tiles_w), 128, 0, st>>>(x, w, bias, y, p);
  }
  cudaError_t e = cudaGetLastError();
  dfmir_count_launch();
  if (e != cudaSuccess) { dfmir_set_error("dfmir_conv_fwd(thin): launch failed: %s", cudaGetErrorString(e)); *rc = DFMIR_ERR_CUDA; }
  else *rc = DFMIR_OK;
  return 1;
}

// wt: [tap][Cout][Cin] (the dgrad layout of the generic path)
int dfmir_thin_dgrad(const float* dy, const float* wt, float* dx, const dfmir_conv_desc* d, cudaStream_t st, int* rc) {
  if (!thin_geom_ok(d)) return 0;
  ThinP p{};
  p.N = d->N; p.OH = d->in_shape[0]; p.OW = d->in_shape[1]; p.SH = d->out_shape[0]; p.SW = d->out_shape[1];
  p.ph = 6 - d->pad[0]; p.pw = 6 - d->pad[1]; p.flip = 1; p.act = DFMIR_ACT_NONE;
  p.s_n = d->y_strides[0]; p.s_h = d->y_strides[1]; p.s_w = d->y_strides[2];
  p.o_n = d->x_strides[0]; p.o_h = d->x_strides[1]; p.o_w = d->x_strides[2];
  if (d->Cout == 1) {        // head: dy has one channel, dx has Cin channels
    if (d->x_strides[3] != 1) return 0;
    p.C = d->Cin; p.tiles_h = (p.OH + 7) / 8; p.tiles_w = (p.OW + 31) / 32;
    const unsigned gx = (unsigned)(p.N * p.tiles_h * p.tiles_w);
    if (thin::launch_expand_umma(dy, wt, nullptr, dx, p, st)) {}
    else if (p.C > 16) thin::thin_expand_kernel<7, 32><<<dim3(gx, (p.C + 31) / 32), 256, 0, st>>>(dy, wt, nullptr, dx, p);
    else thin::thin_expand_kernel<7, 16><<<dim3(gx, 1), 256, 0, st>>>(dy, wt, nullptr, dx, p);
  } else {                   // stem: dy has Cout channels, dx has one
    if (!vec4_ok(dy, d->y_strides, d->Cout) || (((uintptr_t)wt) & 15)) return 0;
    p.C = d->Cout; p.tiles_h = (p.OH + 31) / 32; p.tiles_w = (p.OW + 31) / 32;
    thin::thin_reduce_kernel<7><<<(unsigned)(p.N * p.tiles_h * p.tiles_w), 128, 0, st>>>(dy, wt, nullptr, dx, p);
  }
  cudaError_t e = cudaGetLastError();
  dfmir_count_launch();
  if (e != cudaSuccess) { dfmir_set_error("dfmir_conv_dgrad(thin): launch failed: %s", cudaGetErrorString(e)); *rc = DFMIR_ERR_CUDA; }
  else *rc = DFMIR_OK;
  return 1;
}

int dfmir_thin_wgrad(const float* x, const float* dy, float* dw, float* db, const dfmir_conv_desc* d, cudaStream_t st,
                     int* rc) {
  if (!thin_geom_ok(d)) return 0;
  ThinP p{};
  p.N = d->N; p.act = 0;
  const float *s, *v;
  if (d->Cin == 1) {   // stem: s = x shifted by +tap - pad, v = dy over the output domain
    if (d->y_strides[3] != 1) return 0;
    s = x; v = dy; p.C = d->Cout; p.flip = 0;
    p.OH = d->out_shape[0]; p.OW = d->out_shape[1]; p.SH = d->in_shape[0]; p.SW = d->in_shape[1];
    p.ph = d->pad[0]; p.pw = d->pad[1];
    p.s_n = d->x_strides[0]; p.s_h = d->x_strides[1]; p.s_w = d->x_strides[2];
    p.o_n = d->y_strides[0]; p.o_h = d->y_strides[1]; p.o_w = d->y_strides[2];
  } else {             // head: s = dy shifted by -tap + pad (flipped taps), v = x over the input domain
    if (d->x_strides[3] != 1) return 0;
    s = dy; v = x; p.C = d->Cin; p.flip = 1;
    p.OH = d->in_shape[0]; p.OW = d->in_shape[1]; p.SH = d->out_shape[0]; p.SW = d->out_shape[1];
    p.ph = 6 - d->pad[0]; p.pw = 6 - d->pad[1];
    p.s_n = d->y_strides[0]; p.s_h = d->y_strides[1]; p.s_w = d->y_strides[2];
    p.o_n = d->x_strides[0]; p.o_h = d->x_strides[1]; p.o_w = d->x_strides[2];
  }
  p.tiles_h = (p.OH + 7) / 8; p.tiles_w = (p.OW + 31) / 32;
  const int units = p.N * p.tiles_h * p.tiles_w;
  const int ctiles = (p.C + 63) / 64;
  int gx = 2 * dfmir_num_sms() / ctiles;
  if (gx > units) gx = units;
  if (gx < 1) gx = 1;
  if (!thin::launch_wgrad_umma(s, v, dw, d->Cin == 1 ? db : nullptr, p, st))
    thin::thin_wgrad_kernel<7><<<dim3((unsigned)gx, (unsigned)ctiles), 256, 0, st>>>(s, v, dw, d->Cin == 1 ? db : nullptr, p, units);
  dfmir_count_launch();
  if (d->Cout == 1 && db) {   // bias gradient of the single output channel: sum of dy (dense (N,OH,OW,1))
    const long long n = (long long)d->N * d->out_shape[0] * d->out_shape[1];
    const bool dense = d->y_strides[2] == 1 && d->y_strides[1] == d->out_shape[1] &&
                       d->y_strides[0] == (long long)d->out_shape[0] * d->out_shape[1];
    if (!dense) { dfmir_set_error("dfmir_conv_wgrad(thin): bias gradient needs a dense dy"); *rc = DFMIR_ERR_ARG; return 1; }
    thin::sum_kernel<<<dfmir_num_sms(), 256, 0, st>>>(dy, db, n);
    dfmir_count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { dfmir_set_error("dfmir_conv_wgrad(thin): launch failed: %s", cudaGetErrorString(e)); *rc = DFMIR_ERR_CUDA; }
  else *rc = DFMIR_OK;
  return 1;
}
