// ResizeTransform: linear (bi/tri-linear) resampling with align_corners=True plus the
// vector-field rescale, one launch.  Replaces models/voxelmorph/torchvoxelmorph/layers.py:85-97
// (F.interpolate + scalar multiply).  Arithmetic follows ATen's upsample_{bi,tri}linear with
// align_corners=True: scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = (int)src,
// i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1.
#include "common.cuh"
#include "resize.cuh"
#include "dfmir_b200.h"

namespace {
using namespace resizedev;

// I: int when BC * nout < 2^31 (64-bit div/mod per element made this kernel integer-ALU bound), else long long
template <int ND, typename I>
__global__ void __launch_bounds__(256)
resize_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, RGeom g, float pre_mul, float post_mul) {
  const I nout = (I)g.nout;
  const I total = (I)g.BC * nout;
  for (I it = (I)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (I)gridDim.x * blockDim.x) {
    const int p = (int)(it / nout);
    const I v = it - (I)p * nout;
    const float r = interp<ND, I>(x + (long long)p * g.nin, g, v, pre_mul);
    y[it] = post_mul * r;
  }
}

// Adjoint: scatter gy * (pre_mul*post_mul) * weights into gx (pre-zeroed) with atomics.
template <int ND, typename I>
__global__ void __launch_bounds__(256)
resize_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, RGeom g, float mul) {
  const I nout = (I)g.nout;
  const I total = (I)g.BC * nout;
  for (I it = (I)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (I)gridDim.x * blockDim.x) {
    const int p = (int)(it / nout);
    const I v = it - (I)p * nout;
    int i0[ND], i1[ND]; float l0[ND], l1[ND];
    setup<ND, I>(g, v, i0, i1, l0, l1);
    float* gp = gx + (long long)p * g.nin;
    const float go = gy[it] * mul;
#pragma unroll
    for (int c = 0; c < (1 << ND); ++c) {
      long long off = 0; float w = 1.f;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const int hi = (c >> (ND - 1 - d)) & 1;
        off = off * g.I[d] + (hi ? i1[d] : i0[d]);
        w *= hi ? l1[d] : l0[d];
      }
      if (w != 0.f) atomicAdd(gp + off, w * go);
    }
  }
}

int make_rgeom(RGeom& g, int BC, int nd, const int* in_shape, const int* out_shape) {
  if (nd < 1 || nd > 3 || BC < 0) return -1;
  g.BC = BC; g.nin = 1; g.nout = 1;
  for (int d = 0; d < 3; ++d) {
    g.I[d] = d < nd ? in_shape[d] : 1;
    g.O[d] = d < nd ? out_shape[d] : 1;
    if (g.I[d] <= 0 || g.O[d] <= 0) return -1;
    g.sc[d] = g.O[d] > 1 ? (float)(g.I[d] - 1) / (float)(g.O[d] - 1) : 0.f;
    g.nin *= g.I[d]; g.nout *= g.O[d];
  }
  return 0;
}

inline int grid_for(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace

extern "C" int dfmir_resize_linear_fwd(const float* x, float* y, int BC, int nd, const int* in_shape,
                                       const int* out_shape, float pre_mul, float post_mul, void* stream) {
  RGeom g;
  DFMIR_CHECK_ARG(make_rgeom(g, BC, nd, in_shape, out_shape) == 0, "dfmir_resize_linear_fwd: bad geometry");
  DFMIR_CHECK_ARG(x && y, "dfmir_resize_linear_fwd: null pointer");
  const long long items = (long long)BC * g.nout;
  if (items == 0) return DFMIR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (items < (1LL << 31)) {
    if (nd == 1) resize_fwd_kernel<1, int><<<grid_for(items), 256, 0, st>>>(x, y, g, pre_mul, post_mul);
    else if (nd == 2) resize_fwd_kernel<2, int><<<grid_for(items), 256, 0, st>>>(x, y, g, pre_mul, post_mul);
    else resize_fwd_kernel<3, int><<<grid_for(items), 256, 0, st>>>(x, y, g, pre_mul, post_mul);
  } else {
    if (nd == 1) resize_fwd_kernel<1, long long><<<grid_for(items), 256, 0, st>>>(x, y, g, pre_mul, post_mul);
    else if (nd == 2) resize_fwd_kernel<2, long long><<<grid_for(items), 256, 0, st>>>(x, y, g, pre_mul, post_mul);
    else resize_fwd_kernel<3, long long><<<grid_for(items), 256, 0, st>>>(x, y, g, pre_mul, post_mul);
  }
  DFMIR_CHECK_LAUNCH("dfmir_resize_linear_fwd");
  return DFMIR_OK;
}

extern "C" int dfmir_resize_linear_bwd(const float* gy, float* gx, int BC, int nd, const int* in_shape,
                                       const int* out_shape, float pre_mul, float post_mul, void* stream) {
  RGeom g;
  DFMIR_CHECK_ARG(make_rgeom(g, BC, nd, in_shape, out_shape) == 0, "dfmir_resize_linear_bwd: bad geometry");
  DFMIR_CHECK_ARG(gy && gx, "dfmir_resize_linear_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DFMIR_CUDA(cudaMemsetAsync(gx, 0, sizeof(float) * (long long)BC * g.nin, st));
  const long long items = (long long)BC * g.nout;
  if (items == 0) return DFMIR_OK;
  const float mul = pre_mul * post_mul;
  if (items < (1LL << 31)) {
    if (nd == 1) resize_bwd_kernel<1, int><<<grid_for(items), 256, 0, st>>>(gy, gx, g, mul);
    else if (nd == 2) resize_bwd_kernel<2, int><<<grid_for(items), 256, 0, st>>>(gy, gx, g, mul);
    else resize_bwd_kernel<3, int><<<grid_for(items), 256, 0, st>>>(gy, gx, g, mul);
  } else {
    if (nd == 1) resize_bwd_kernel<1, long long><<<grid_for(items), 256, 0, st>>>(gy, gx, g, mul);
    else if (nd == 2) resize_bwd_kernel<2, long long><<<grid_for(items), 256, 0, st>>>(gy, gx, g, mul);
    else resize_bwd_kernel<3, long long><<<grid_for(items), 256, 0, st>>>(gy, gx, g, mul);
  }
  DFMIR_CHECK_LAUNCH("dfmir_resize_linear_bwd");
  return DFMIR_OK;
}
