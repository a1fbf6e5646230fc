// Windowed box-sum machinery of the local-NCC kernels, shared by losses.cu and fused_reg.cu.
// (see losses.cu for the description of the algorithm)
#pragma once
#include "common.cuh"

namespace nccdev {

constexpr int TX = 32, TY = 8, NT = TX * TY;  // outputs per slice per CTA == threads

struct BoxGeom {
  int B, D, H, W;      // 2-D volumes use D = 1
  int win, wz;         // in-plane window, z window (1 for 2-D)
  int zchunk;          // output slices per CTA along z
  int nzc;             // chunks per volume
};

// March a (TY x TX) column through z computing window sums of NQ per-voxel quantities.
//   Loader::load(b, z, y, x, inb, q[NQ])  : per input voxel quantities (zero outside the volume)
//   Consumer::consume(b, z, y, x, sums[NQ]) : called for every in-volume output voxel
template <int NQ, int WIN, class Loader, class Consumer>
__device__ __forceinline__ void box_march(const BoxGeom& g, Loader& ld, Consumer& cs, float* smem, int tile_x, int tile_y,
                                          int tile_z) {
  constexpr int R = WIN / 2;
  constexpr int IX = TX + 2 * R, IY = TY + 2 * R;
  float* sIn = smem;                         // [NQ][IY][IX]
  float* sX = sIn + NQ * IY * IX;            // [NQ][IY][TX]
  float* ring = sX + NQ * IY * TX;           // [wz][NQ][NT]
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int x0 = tile_x * TX, y0 = tile_y * TY;
  const int b = tile_z / g.nzc, zc = tile_z % g.nzc;
  const int zo0 = zc * g.zchunk;
  const int zo1 = min(g.D, zo0 + g.zchunk);
  const int rz = g.wz / 2;

  // software pipeline: the halo tile of slice z+1 is fetched into registers while slice z is summed
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
  prefetch(zo0 - rz);
  for (int z = zo0 - rz; z < zo1 + rz; ++z) {
    const bool zin = z >= 0 && z < g.D;
    const int slot = ((z % g.wz) + g.wz) % g.wz;
    if (zin) {
#pragma unroll
      for (int e = 0; e < PER_T; ++e) {
        const int i = tid + e * NT;
        if (i < IY * IX) {
          const int ly = i / IX, lx = i - ly * IX;
#pragma unroll
          for (int k = 0; k < NQ; ++k) sIn[(k * IY + ly) * IX + lx] = pre[e][k];
        }
      }
    }
    __syncthreads();
    if (z + 1 < zo1 + rz) prefetch(z + 1);
    if (zin) {
      for (int i = tid; i < IY * TX; i += NT) {
        const int ly = i / TX, lx = i - ly * TX;
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
          const float* row = sIn + (k * IY + ly) * IX + lx;
          float s = row[0];
#pragma unroll
          for (int j = 1; j < WIN; ++j) s += row[j];
          sX[(k * IY + ly) * TX + lx] = s;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        const float* col = sX + (k * IY + ty) * TX + tx;
        float s = col[0];
#pragma unroll
        for (int j = 1; j < WIN; ++j) s += col[j * TX];
        ring[(slot * NQ + k) * NT + tid] = s;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NQ; ++k) ring[(slot * NQ + k) * NT + tid] = 0.f;
    }
    // ring slots are private to the thread (indexed by tid): no barrier needed for them, but sIn /
    // sX are reused by the next slice.
    const int zo = z - rz;
    if (zo >= zo0 && zo < zo1) {
      // sum the ring oldest slice first (z-rz .. z+rz ascending).  One modulo per slice, not per term:
      // the slot of slice zo-rz, then +1 with wrap-around (90 integer divisions per voxel otherwise).
      float sums[NQ];
#pragma unroll
      for (int k = 0; k < NQ; ++k) sums[k] = 0.f;
      int sl = (((zo - rz) % g.wz) + g.wz) % g.wz;
      for (int j = 0; j < g.wz; ++j) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) sums[k] += ring[(sl * NQ + k) * NT + tid];
        if (++sl == g.wz) sl = 0;
      }
      const int oy = y0 + ty, ox = x0 + tx;
      if (oy < g.H && ox < g.W) cs.consume(b, zo, oy, ox, sums);
    }
    __syncthreads();
  }
}

template <int NQ, int WIN>
constexpr size_t box_smem_bytes(int wz) {
  return sizeof(float) * ((size_t)NQ * (TY + 2 * (WIN / 2)) * (TX + 2 * (WIN / 2)) +
                          (size_t)NQ * (TY + 2 * (WIN / 2)) * TX + (size_t)wz * NQ * NT);
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
  // split z so that the grid has >= ~4 CTAs per SM, but keep chunks >= 2*win slices deep
  const long long tiles = (long long)dfmir_ceil_div(g.W, TX) * dfmir_ceil_div(g.H, TY) * B;
  int nzc = 1;
  if (nd == 3) {
    const long long want = 4LL * dfmir_num_sms();
    while (tiles * nzc < want && g.D / (nzc * 2) >= 2 * win) nzc *= 2;
  }
  g.zchunk = dfmir_ceil_div(g.D, nzc);
  g.nzc = dfmir_ceil_div(g.D, g.zchunk);
  return 0;
}

inline dim3 box_grid(const BoxGeom& g) {
  return dim3(dfmir_ceil_div(g.W, TX), dfmir_ceil_div(g.H, TY), g.B * g.nzc);
}

}  // namespace nccdev
