// Linear resampling arithmetic (ATen upsample_{bi,tri}linear, align_corners=True) shared by resize.cu and
// fused_reg.cu, so that the fused launch reproduces the stand-alone kernel bit for bit.
#pragma once
#include "common.cuh"

namespace resizedev {

struct RGeom {
  int BC;          // batch * channels (planes)
  int I[3], O[3];  // input / output spatial sizes, unused dims = 1
  float sc[3];     // (in-1)/(out-1) or 0
  long long nin, nout;
};

template <int ND, typename I>
__device__ __forceinline__ void setup(const RGeom& g, I v, int* i0, int* i1, float* l0, float* l1) {
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    const int o = (int)(v % g.O[d]);
    v /= g.O[d];
    const float s = g.sc[d] * (float)o;
    int a = (int)s;
    if (a > g.I[d] - 1) a = g.I[d] - 1;
    i0[d] = a;
    i1[d] = a + (a < g.I[d] - 1 ? 1 : 0);
    l1[d] = s - (float)a;
    l0[d] = 1.0f - l1[d];
  }
}


// the same from the output position itself (callers that already hold it: no division)
template <int ND>
__device__ __forceinline__ void setup_pos(const RGeom& g, const int* pos, int* i0, int* i1, float* l0, float* l1) {
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const float s = g.sc[d] * (float)pos[d];
    int a = (int)s;
    if (a > g.I[d] - 1) a = g.I[d] - 1;
    i0[d] = a;
    i1[d] = a + (a < g.I[d] - 1 ? 1 : 0);
    l1[d] = s - (float)a;
    l0[d] = 1.0f - l1[d];
  }
}

// value of one plane at a prepared site: pre_mul is applied to every input sample (layers.py:91-94)
template <int ND>
__device__ __forceinline__ float interp_site(const float* __restrict__ xp, const RGeom& g, const int* i0, const int* i1,
                                             const float* l0, const float* l1, float pre_mul) {
  float r;
  if (ND == 1) {
    r = l0[0] * (pre_mul * xp[i0[0]]) + l1[0] * (pre_mul * xp[i1[0]]);
  } else if (ND == 2) {
    const long long r0 = (long long)i0[0] * g.I[1], r1 = (long long)i1[0] * g.I[1];
    r = l0[0] * (l0[1] * (pre_mul * xp[r0 + i0[1]]) + l1[1] * (pre_mul * xp[r0 + i1[1]])) +
        l1[0] * (l0[1] * (pre_mul * xp[r1 + i0[1]]) + l1[1] * (pre_mul * xp[r1 + i1[1]]));
  } else {
    const long long hw = (long long)g.I[1] * g.I[2];
    const long long z0 = i0[0] * hw, z1 = i1[0] * hw;
    const long long y0 = (long long)i0[1] * g.I[2], y1 = (long long)i1[1] * g.I[2];
    const int a = i0[2], b = i1[2];
    r = l0[0] * (l0[1] * (l0[2] * (pre_mul * xp[z0 + y0 + a]) + l1[2] * (pre_mul * xp[z0 + y0 + b])) +
                 l1[1] * (l0[2] * (pre_mul * xp[z0 + y1 + a]) + l1[2] * (pre_mul * xp[z0 + y1 + b]))) +
        l1[0] * (l0[1] * (l0[2] * (pre_mul * xp[z1 + y0 + a]) + l1[2] * (pre_mul * xp[z1 + y0 + b])) +
                 l1[1] * (l0[2] * (pre_mul * xp[z1 + y1 + a]) + l1[2] * (pre_mul * xp[z1 + y1 + b])));
  }
  return r;
}

// value of output voxel v of one plane
template <int ND, typename I>
__device__ __forceinline__ float interp(const float* __restrict__ xp, const RGeom& g, I v, float pre_mul) {
  int i0[ND], i1[ND]; float l0[ND], l1[ND];
  setup<ND, I>(g, v, i0, i1, l0, l1);
  return interp_site<ND>(xp, g, i0, i1, l0, l1, pre_mul);
}

}  // namespace resizedev
