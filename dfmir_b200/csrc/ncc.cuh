// Windowed box-sum machinery of the local-NCC kernels, shared by losses.cu and fused_reg.cu.
// (see losses.cu for the description of the algorithm)
#pragma once
#include "common.cuh"

namespace nccdev {

constexpr int TX = 32, TY = 16, NT = TX * TY;  // outputs per slice per CTA == threads
constexpr int SEG = 8;                          // consecutive outputs one thread builds in the x / y passes

struct BoxGeom {
  int B, D, H, W;      // 2-D volumes use D = 1
  int win, wz;         // in-plane window, z window (1 for 2-D)
  int zchunk;          // output slices per CTA along z
  int nzc;             // chunks per volume
};

template <int WIN>
struct BoxDims {
  static constexpr int R = WIN / 2;
  static constexpr int IY = TY + 2 * R, IX = TX + 2 * R;
  static constexpr int XIN = ((SEG + WIN - 1 + 3) / 4) * 4;        // floats one x-pass item reads (whole float4s)
  static constexpr int NEED = (TX - SEG + XIN) > IX ? (TX - SEG + XIN) : IX;
  // row stride of the staged slice: 4 mod 8 floats, so that the 8 lanes of one LDS.128 phase (4 segments of a row,
  // then the next row) fall on 8 different 16-byte bank groups
  static constexpr int IXP = ((NEED + 3) / 8) * 8 + 4;
  static constexpr int SXP = TX + 4;                               // same rule for the STS.128 of the x sums
};

// Sums of WIN consecutive values for N consecutive windows, a[0 .. N + WIN - 2] -> o[0 .. N - 1], by doubling
// (pairs, quads, octets shared between neighbouring windows): 4 - 6 additions per window instead of WIN - 1, every
// window summed afresh in one fixed order (no running sum that could drift).
template <int WIN, int N>
__device__ __forceinline__ void win_sums(const float* a, float* o) {
  static_assert(WIN == 3 || WIN == 5 || WIN == 7 || WIN == 9 || WIN == 11, "window");
  constexpr int L = N + WIN - 1;
  float a2[L - 1];
#pragma unroll
  for (int i = 0; i < L - 1; ++i) a2[i] = a[i] + a[i + 1];
  if constexpr (WIN == 3) {
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = a2[i] + a[i + 2];
  } else {
    float a4[L - 3];
#pragma unroll
    for (int i = 0; i < L - 3; ++i) a4[i] = a2[i] + a2[i + 2];
    if constexpr (WIN == 5) {
#pragma unroll
      for (int i = 0; i < N; ++i) o[i] = a4[i] + a[i + 4];
    } else if constexpr (WIN == 7) {
#pragma unroll
      for (int i = 0; i < N; ++i) o[i] = a4[i] + (a2[i + 4] + a[i + 6]);
    } else {
      float a8[L - 7];
#pragma unroll
      for (int i = 0; i < L - 7; ++i) a8[i] = a4[i] + a4[i + 4];
      if constexpr (WIN == 9) {
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = a8[i] + a[i + 8];
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) o[i] = a8[i] + (a2[i + 8] + a[i + 10]);
      }
    }
  }
}

// z window: the last WIN slice sums of this thread's output column live in registers; slot S is overwritten with the
// newest slice and the window is summed oldest slice first as a fixed pairwise tree.  S is a compile-time constant
// (the caller switches on the slot), which is what keeps ring[][] in registers.
template <int NQ, int WIN, int S>
__device__ __forceinline__ void ring_push_sum(float (&ring)[WIN][NQ], const float* v, float* sums) {
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    ring[S][k] = v[k];
    float t[WIN];
#pragma unroll
    for (int j = 0; j < WIN; ++j) t[j] = ring[(S + 1 + j) % WIN][k];
#pragma unroll
    for (int w = 1; w < WIN; w *= 2)
#pragma unroll
      for (int j = 0; j + w < WIN; j += 2 * w) t[j] += t[j + w];
    sums[k] = t[0];
  }
}

// March a (TY x TX) column through z computing window sums of NQ per-voxel quantities.
//   Loader::load(b, z, y, x, inb, q[NQ])  : per input voxel quantities (zero outside the volume)
//   Consumer::consume(b, z, y, x, sums[NQ]) : called for every in-volume output voxel
// Per slice: (1) the tile with its halo is staged in shared memory (the next slice is already in flight in
// registers), (2) x pass: each item loads SEG + WIN - 1 staged values of one row with LDS.128 and builds SEG window
// sums by doubling, (3) y pass: the same down the columns, (4) z pass in registers.  Three barriers per slice.
template <int NQ, int WIN, class Loader, class Consumer>
__device__ __forceinline__ void box_march(const BoxGeom& g, Loader& ld, Consumer& cs, float* smem, int tile_x, int tile_y,
                                          int tile_z) {
  using BD = BoxDims<WIN>;
  constexpr int R = BD::R, IX = BD::IX, IY = BD::IY, IXP = BD::IXP, SXP = BD::SXP, XIN = BD::XIN;
  float* sIn = smem;                         // [NQ][IY][IXP]
  float* sX = sIn + NQ * IY * IXP;           // [NQ][IY][SXP]
  float* sY = sX + NQ * IY * SXP;            // [NQ][TY][TX]
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int x0 = tile_x * TX, y0 = tile_y * TY;
  const int b = tile_z / g.nzc, zc = tile_z % g.nzc;
  const int zo0 = zc * g.zchunk;
  const int zo1 = min(g.D, zo0 + g.zchunk);
  const int rz = g.wz / 2;
  const bool own = y0 + ty < g.H && x0 + tx < g.W;

  constexpr int PER_T = (IY * IX + NT - 1) / NT;
  float pre[PER_T][NQ];
  auto prefetch = [&](int z) {
    const bool zin = z >= 0 && z < g.D;
#pragma unroll
    for (int e = 0; e < PER_T; ++e) {
      const int i = tid + e * NT;
      if (zin && i < IY * IX) {
        const int ly = i / IX, lx = i - ly * IX;
        const int gy = y0 - R + ly, gx = x0 - R + lx;
        const bool inb = gy >= 0 && gy < g.H && gx >= 0 && gx < g.W;
        ld.load(b, z, gy, gx, inb, pre[e]);
      }
    }
  };
  float ring[WIN][NQ];
#pragma unroll
  for (int j = 0; j < WIN; ++j)
#pragma unroll
    for (int k = 0; k < NQ; ++k) ring[j][k] = 0.f;

  prefetch(zo0 - rz);
  int slot = 0;
  for (int z = zo0 - rz; z < zo1 + rz; ++z) {
    const bool zin = z >= 0 && z < g.D;       // uniform over the CTA
    float v[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) v[k] = 0.f;
    if (zin) {
#pragma unroll
      for (int e = 0; e < PER_T; ++e) {
        const int i = tid + e * NT;
        if (i < IY * IX) {
          const int ly = i / IX, lx = i - ly * IX;
#pragma unroll
          for (int k = 0; k < NQ; ++k) sIn[(k * IY + ly) * IXP + lx] = pre[e][k];
        }
      }
      __syncthreads();
      if (z + 1 < zo1 + rz) prefetch(z + 1);
      // x pass
      for (int i = tid; i < NQ * IY * (TX / SEG); i += NT) {
        const int seg = i % (TX / SEG), rowq = i / (TX / SEG);      // rowq = q * IY + row
        float a[XIN], o[SEG];
        const float4* src = reinterpret_cast<const float4*>(sIn + rowq * IXP + seg * SEG);
#pragma unroll
        for (int j = 0; j < XIN / 4; ++j) {
          const float4 t = src[j];
          a[4 * j] = t.x; a[4 * j + 1] = t.y; a[4 * j + 2] = t.z; a[4 * j + 3] = t.w;
        }
        win_sums<WIN, SEG>(a, o);
        float4* dst = reinterpret_cast<float4*>(sX + rowq * SXP + seg * SEG);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
      __syncthreads();
      // y pass
      for (int i = tid; i < NQ * (TY / SEG) * TX; i += NT) {
        const int x = i % TX, hq = i / TX;
        const int half = hq % (TY / SEG), q = hq / (TY / SEG);
        float a[SEG + WIN - 1], o[SEG];
        const float* src = sX + (q * IY + half * SEG) * SXP + x;
#pragma unroll
        for (int j = 0; j < SEG + WIN - 1; ++j) a[j] = src[j * SXP];
        win_sums<WIN, SEG>(a, o);
        float* dst = sY + (q * TY + half * SEG) * TX + x;
#pragma unroll
        for (int j = 0; j < SEG; ++j) dst[j * TX] = o[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NQ; ++k) v[k] = sY[(k * TY + ty) * TX + tx];
    } else if (z + 1 < zo1 + rz) {
      prefetch(z + 1);
    }
    float sums[NQ];
    if (g.wz == 1) {
#pragma unroll
      for (int k = 0; k < NQ; ++k) sums[k] = v[k];
    } else {
      switch (slot) {
#define DFMIR_RING_CASE(S) case S: if constexpr (S < WIN) ring_push_sum<NQ, WIN, (S < WIN ? S : 0)>(ring, v, sums); break;
        DFMIR_RING_CASE(0) DFMIR_RING_CASE(1) DFMIR_RING_CASE(2) DFMIR_RING_CASE(3) DFMIR_RING_CASE(4) DFMIR_RING_CASE(5)
        DFMIR_RING_CASE(6) DFMIR_RING_CASE(7) DFMIR_RING_CASE(8) DFMIR_RING_CASE(9) DFMIR_RING_CASE(10)
#undef DFMIR_RING_CASE
        default: break;
      }
      if (++slot == WIN) slot = 0;
    }
    const int zo = z - rz;
    if (own && zo >= zo0 && zo < zo1) cs.consume(b, zo, y0 + ty, x0 + tx, sums);
  }
  __syncthreads();       // the next tile of a persistent caller restages sIn
}

template <int NQ, int WIN>
constexpr size_t box_smem_bytes(int /*wz*/) {
  using BD = BoxDims<WIN>;
  return sizeof(float) * ((size_t)NQ * BD::IY * BD::IXP + (size_t)NQ * BD::IY * BD::SXP + (size_t)NQ * TY * TX);
}

// cc and the partials of the reference formula (util/losses.py:199-207, :241), fp32 op for op.
struct CcTerms { float cross, ivar, jvar, uI, uJ, denom, cc; };
__device__ __forceinline__ CcTerms cc_terms(const float* s, float wsz, float eps) {
  // s = {I_sum, J_sum, I2_sum, J2_sum, IJ_sum}
  CcTerms t;
  t.uI = __fdiv_rn(s[0], wsz);
  t.uJ = __fdiv_rn(s[1], wsz);
  t.cross = __fadd_rn(__fsub_rn(__fsub_rn(s[4], __fmul_rn(t.uJ, s[0])), __fmul_rn(t.uI, s[1])),
                      __fmul_rn(__fmul_rn(t.uI, t.uJ), wsz));
  t.ivar = __fadd_rn(__fsub_rn(s[2], __fmul_rn(__fmul_rn(2.f, t.uI), s[0])),
                     __fmul_rn(__fmul_rn(t.uI, t.uI), wsz));
  t.jvar = __fadd_rn(__fsub_rn(s[3], __fmul_rn(__fmul_rn(2.f, t.uJ), s[1])),
                     __fmul_rn(__fmul_rn(t.uJ, t.uJ), wsz));
  t.denom = __fadd_rn(__fmul_rn(t.ivar, t.jvar), eps);
  t.cc = __fdiv_rn(__fmul_rn(t.cross, t.cross), t.denom);
  return t;
}

struct IJLoader {
  const float* I; const float* J; long long vol, hw; int W;
  __device__ __forceinline__ void load(int b, int z, int y, int x, bool inb, float* q) const {
    float i = 0.f, j = 0.f;
    if (inb) {
      const long long o = b * vol + z * hw + (long long)y * W + x;
      i = I[o]; j = J[o];
    }
    q[0] = i; q[1] = j; q[2] = i * i; q[3] = j * j; q[4] = i * j;
  }
};

struct CcReduce {
  const float* mask; long long vol, hw; int W; float wsz, eps;
  double acc_cc, acc_m;
  __device__ __forceinline__ void consume(int b, int z, int y, int x, const float* s) {
    const CcTerms t = cc_terms(s, wsz, eps);
    if (mask) {
      const float m = mask[b * vol + z * hw + (long long)y * W + x];
      acc_cc += (double)(t.cc * m); acc_m += (double)m;
    } else {
      acc_cc += (double)t.cc;
    }
  }
};


// host-side geometry helpers
inline int make_box(BoxGeom& g, int B, int nd, const int* shape, int win) {
  if (nd < 2 || nd > 3 || B < 1) return -1;
  g.B = B;
  g.D = nd == 3 ? shape[0] : 1;
  g.H = shape[nd - 2]; g.W = shape[nd - 1];
  if (g.D <= 0 || g.H <= 0 || g.W <= 0) return -1;
  g.win = win; g.wz = nd == 3 ? win : 1;
  // z chunks per tile column: the split with the fewest slice marches on the critical path, counting the 2 * (win / 2)
  // halo slices every chunk re-reads and the number of waves of CTAs (one resident CTA per SM)
  const long long tiles = (long long)dfmir_ceil_div(g.W, TX) * dfmir_ceil_div(g.H, TY) * B;
  int nzc = 1;
  if (nd == 3) {
    const long long sms = dfmir_num_sms();
    long long best = -1;
    for (int c = 1; c <= 16 && g.D / c >= win; ++c) {
      const long long waves = (tiles * c + sms - 1) / sms;
      const long long cost = waves * (dfmir_ceil_div(g.D, c) + 2 * (win / 2));
      if (best < 0 || cost < best) { best = cost; nzc = c; }
    }
  }
  g.zchunk = dfmir_ceil_div(g.D, nzc);
  g.nzc = dfmir_ceil_div(g.D, g.zchunk);
  return 0;
}

inline dim3 box_grid(const BoxGeom& g) {
  return dim3(dfmir_ceil_div(g.W, TX), dfmir_ceil_div(g.H, TY), g.B * g.nzc);
}

}  // namespace nccdev
