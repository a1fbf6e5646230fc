// Device-side sampling arithmetic shared by the warp / VecInt / fused kernels.
//
// Restates, operation for operation in IEEE fp32 (no FMA contraction, no fast division), the
// coordinate pipeline of the reference:
//   models/voxelmorph/torchvoxelmorph/layers.py:32-37   new_locs = grid + flow ; 2*(x/(S-1) - 0.5)
//   ATen grid_sampler_unnormalize(align_corners=True)    ((n + 1) / 2) * (S - 1)
// so that floor()/nearbyint() of the un-normalised coordinate picks the same voxel as the
// reference (hard part H1).
#pragma once
#include "common.cuh"

#define DFMIR_INTERP_LINEAR 0
#define DFMIR_INTERP_NEAREST 1
#define DFMIR_COORD_IEEE_DIV 0  // CPU ATen: true division by (S-1)
#define DFMIR_COORD_RCP_MUL 1   // CUDA ATen: multiply by fp32 reciprocal of (S-1)

template <int COORD_MODE>
__device__ __forceinline__ float dfmir_unnorm_coord(int i, float f, int S) {
  const float sm1 = (float)(S - 1);
  const float loc = __fadd_rn((float)i, f);
  float q;
  if (COORD_MODE == DFMIR_COORD_IEEE_DIV)
    q = __fdiv_rn(loc, sm1);
  else
    q = __fmul_rn(loc, __fdiv_rn(1.0f, sm1));
  const float n = __fmul_rn(2.0f, __fsub_rn(q, 0.5f));
  // (n + 1) / 2: multiplying by 0.5 is the same correctly-rounded value as dividing by 2, at a tenth of the cost
  return __fmul_rn(__fmul_rn(__fadd_rn(n, 1.0f), 0.5f), sm1);
}

// float -> int that is safe for NaN / huge values (maps them far out of bounds).
__device__ __forceinline__ int dfmir_safe_int(float v) {
  if (!(v > -1.0e9f)) return -1000000000;  // also catches NaN
  if (v > 1.0e9f) return 1000000000;
  return (int)v;
}

// Per-voxel sampling site: base corner + 1-D weights per dimension.
template <int ND>
struct SampleSite {
  int i0[ND];     // floor(ix) per dim (ij order: dim 0 = slowest spatial axis)
  float w0[ND];   // weight of corner i0   : (i0 + 1) - ix
  float w1[ND];   // weight of corner i0+1 : ix - i0
};

template <int ND, int COORD_MODE>
__device__ __forceinline__ void dfmir_make_site(SampleSite<ND>& s, const int* pos, const float* f,
                                                const int* S) {
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const float ix = dfmir_unnorm_coord<COORD_MODE>(pos[d], f[d], S[d]);
    const float fl = floorf(ix);
    s.i0[d] = dfmir_safe_int(fl);
    s.w1[d] = ix - fl;
    s.w0[d] = (fl + 1.0f) - ix;
  }
}

template <int ND, int COORD_MODE>
__device__ __forceinline__ void dfmir_nearest_index(int* idx, const int* pos, const float* f,
                                                    const int* S) {
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const float ix = dfmir_unnorm_coord<COORD_MODE>(pos[d], f[d], S[d]);
    idx[d] = dfmir_safe_int(nearbyintf(ix));  // round-half-even, as ATen's std::nearbyint
  }
}

// Weight of corner c: product taken x first, then y, then z (the order ATen multiplies in).
template <int ND>
__device__ __forceinline__ float dfmir_corner_weight(const SampleSite<ND>& s, int c) {
  float w = 1.f;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    const float wd = ((c >> (ND - 1 - d)) & 1) ? s.w1[d] : s.w0[d];
    w = (d == ND - 1) ? wd : w * wd;
  }
  return w;
}

// Gather-interpolate one channel plane (contiguous, row-major over S) at a site. Zero padding.
template <int ND>
__device__ __forceinline__ float dfmir_sample(const float* __restrict__ plane, const SampleSite<ND>& s,
                                              const int* S) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) {
    long long off = 0;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      // bit (ND-1-d) of c selects the upper corner of dim d, so the fastest axis toggles first
      const int id = s.i0[d] + ((c >> (ND - 1 - d)) & 1);
      ok = ok && (id >= 0) && (id < S[d]);
      off = off * S[d] + id;
    }
    if (ok) acc += __ldg(plane + off) * dfmir_corner_weight<ND>(s, c);
  }
  return acc;
}

// The same gather with the corner offsets / weights / bounds computed ONCE per site and shared by every plane
// sampled there (the nd components of a field, the channels of an image): same products, same order of additions
// as dfmir_sample, so results are bit-identical.  Offsets are plane-relative and fit 32 bits for < 2^31 voxels.
template <int ND>
struct CornerSet {
  int off[1 << ND];
  float w[1 << ND];
  bool ok[1 << ND];
};

template <int ND>
__device__ __forceinline__ void dfmir_corners(CornerSet<ND>& cs, const SampleSite<ND>& s, const int* S) {
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) {
    int off = 0;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const int id = s.i0[d] + ((c >> (ND - 1 - d)) & 1);
      ok = ok && (id >= 0) && (id < S[d]);
      off = off * S[d] + id;
    }
    cs.off[c] = off; cs.ok[c] = ok; cs.w[c] = dfmir_corner_weight<ND>(s, c);
  }
}

template <int ND>
__device__ __forceinline__ float dfmir_sample_corners(const float* __restrict__ plane, const CornerSet<ND>& cs) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c)
    if (cs.ok[c]) acc += __ldg(plane + cs.off[c]) * cs.w[c];
  return acc;
}
