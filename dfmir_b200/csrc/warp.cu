// K3 / K4: SpatialTransformer warp (forward, backward) and VecInt scaling-and-squaring steps.
//
// Replaces models/voxelmorph/torchvoxelmorph/layers.py:30-48 (SpatialTransformer.forward: grid add,
// per-axis normalise, permute, channel flip, F.grid_sample) and :64-68 (VecInt.forward) of the
// reference.  One launch per warp: the sampling coordinates are computed in registers, no grid
// tensor is ever materialised.  Layout: src (B,C,*S), flow (B,nd,*S), out (B,C,*S), fp32, planar
// (the reference's NCHW / NCDHW), so flow reads / out writes are 128-bit coalesced along x.
#include "warp.cuh"
#include "dfmir_b200.h"

namespace {

struct Geom {
  int B, C;
  int S[3];          // spatial sizes, ij order; unused trailing dims = 1
  long long nvox;    // prod(S)
};

template <int ND, typename I>
__device__ __forceinline__ void unravel(I v, const int* S, int* pos) {
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    pos[d] = (int)(v % S[d]);
    v /= S[d];
  }
}

// ---------------------------------------------------------------- forward, linear / nearest
// VEC consecutive x positions per thread; flow is read and out written as float4 when VEC == 4.
// I: index type of the element loops — int when B * nvox < 2^31 (64-bit div/mod per voxel would dominate
// these memory-bound kernels), long long otherwise.
template <int ND, int COORD_MODE, int VEC, bool NEAREST, typename I>
__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float* __restrict__ src, const float* __restrict__ flow, float* __restrict__ out,
                int32_t* __restrict__ idx_out, Geom g) {
  const I items_per_b = (I)(g.nvox / VEC);
  const I total = (I)g.B * items_per_b;
  for (I it = (I)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (I)gridDim.x * blockDim.x) {
    const int b = (int)(it / items_per_b);
    const I v0 = (it - (I)b * items_per_b) * VEC;
    int pos[ND];
    unravel<ND, I>(v0, g.S, pos);

    float f[ND][VEC];
    const float* fb = flow + (long long)b * ND * g.nvox + v0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      if (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(fb + (long long)d * g.nvox);
        f[d][0] = t.x; f[d][1 % VEC] = t.y; f[d][2 % VEC] = t.z; f[d][3 % VEC] = t.w;
      } else {
        f[d][0] = fb[(long long)d * g.nvox];
      }
    }

    const float* sb = src + (long long)b * g.C * g.nvox;
    float* ob = out + (long long)b * g.C * g.nvox + v0;

    if (NEAREST) {
      int ni[VEC][ND];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        int p[ND]; float fj[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { p[d] = pos[d]; fj[d] = f[d][j]; }
        p[ND - 1] += j;
        dfmir_nearest_index<ND, COORD_MODE>(ni[j], p, fj, g.S);
        if (idx_out) {
#pragma unroll
          for (int d = 0; d < ND; ++d)
            idx_out[((long long)b * ND + d) * g.nvox + v0 + j] = ni[j][d];
        }
      }
      for (int c = 0; c < g.C; ++c) {
        float r[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          long long off = 0; bool ok = true;
#pragma unroll
          for (int d = 0; d < ND; ++d) {
            ok = ok && ni[j][d] >= 0 && ni[j][d] < g.S[d];
            off = off * g.S[d] + ni[j][d];
          }
          r[j] = ok ? __ldg(sb + (long long)c * g.nvox + off) : 0.f;
        }
        if (VEC == 4)
          *reinterpret_cast<float4*>(ob + (long long)c * g.nvox) = make_float4(r[0], r[1 % VEC], r[2 % VEC], r[3 % VEC]);
        else
          ob[(long long)c * g.nvox] = r[0];
      }
    } else {
      SampleSite<ND> site[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        int p[ND]; float fj[ND];
#pragma unroll
        for (int d = 0; d < ND; ++d) { p[d] = pos[d]; fj[d] = f[d][j]; }
        p[ND - 1] += j;
        dfmir_make_site<ND, COORD_MODE>(site[j], p, fj, g.S);
        if (idx_out) {
#pragma unroll
          for (int d = 0; d < ND; ++d)
            idx_out[((long long)b * ND + d) * g.nvox + v0 + j] = site[j].i0[d];
        }
      }
      for (int c = 0; c < g.C; ++c) {
        const float* plane = sb + (long long)c * g.nvox;
        float r[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) r[j] = dfmir_sample<ND>(plane, site[j], g.S);
        if (VEC == 4)
          *reinterpret_cast<float4*>(ob + (long long)c * g.nvox) = make_float4(r[0], r[1 % VEC], r[2 % VEC], r[3 % VEC]);
        else
          ob[(long long)c * g.nvox] = r[0];
      }
    }
  }
}

// ---------------------------------------------------------------- backward (linear only)
// d_src is accumulated with fp32 atomics (caller zero-fills); d_flow is written directly.
// d(ix)/d(flow) = 1 analytically (normalise . unnormalise), so d_flow[d] = sum_c d(out_c)/d(ix_d) * g_c.
template <int ND, int COORD_MODE, typename I>
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ src,
                const float* __restrict__ flow, float* __restrict__ d_src, float* __restrict__ d_flow,
                Geom g) {
  const I nvox = (I)g.nvox;
  const I total = (I)g.B * nvox;
  for (I it = (I)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (I)gridDim.x * blockDim.x) {
    const int b = (int)(it / nvox);
    const I v = it - (I)b * nvox;
    int pos[ND]; float f[ND];
    unravel<ND, I>(v, g.S, pos);
#pragma unroll
    for (int d = 0; d < ND; ++d) f[d] = flow[((long long)b * ND + d) * g.nvox + v];
    SampleSite<ND> s;
    dfmir_make_site<ND, COORD_MODE>(s, pos, f, g.S);

    float gf[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) gf[d] = 0.f;

    // corner offsets / validity once, reused over channels
    long long off[1 << ND]; bool ok[1 << ND]; float w[1 << ND];
#pragma unroll
    for (int c = 0; c < (1 << ND); ++c) {
      long long o = 0; bool k = true;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const int id = s.i0[d] + ((c >> (ND - 1 - d)) & 1);
        k = k && id >= 0 && id < g.S[d];
        o = o * g.S[d] + id;
      }
      off[c] = o; ok[c] = k; w[c] = dfmir_corner_weight<ND>(s, c);
    }

    for (int ch = 0; ch < g.C; ++ch) {
      const long long pl = ((long long)b * g.C + ch) * g.nvox;
      const float go = gout[pl + v];
#pragma unroll
      for (int c = 0; c < (1 << ND); ++c) {
        if (!ok[c]) continue;
        if (d_src) atomicAdd(d_src + pl + off[c], w[c] * go);
        if (d_flow) {
          const float val = __ldg(src + pl + off[c]) * go;
#pragma unroll
          for (int d = 0; d < ND; ++d) {
            // derivative of the product of 1-D weights wrt ix_d: +/- product of the other dims
            float o = 1.f;
#pragma unroll
            for (int e = 0; e < ND; ++e)
              if (e != d) o *= ((c >> (ND - 1 - e)) & 1) ? s.w1[e] : s.w0[e];
            gf[d] += (((c >> (ND - 1 - d)) & 1) ? o : -o) * val;
          }
        }
      }
    }
    if (d_flow) {
#pragma unroll
      for (int d = 0; d < ND; ++d) d_flow[((long long)b * ND + d) * g.nvox + v] = gf[d];
    }
  }
}

// ---------------------------------------------------------------- VecInt step
// out[bv] = s*in[bi] + warp(s*in[bi], s*in[bi]),  bi = bv % B_in, s = bv < B_in ? scale_lo : scale_hi.
// First step: in = raw velocity (B), scale_lo = 2^-n, scale_hi = -2^-n (bidirectional: the
// negated flow of vxm/networks.py:1125 is integrated in the same launch as virtual batches
// B..2B-1). Later steps: B_in = Bv, scales 1. Scaling by +-2^-n commutes exactly with fp32
// rounding, so sampling the unscaled field and scaling is bit-identical to layers.py:65.
template <int ND, int COORD_MODE, typename I>
__global__ void __launch_bounds__(256)
vecint_step_kernel(const float* __restrict__ in, float* __restrict__ out, int B_in, int Bv,
                   float scale_lo, float scale_hi, Geom g) {
  const I nvox = (I)g.nvox;
  const I total = (I)Bv * nvox;
  for (I it = (I)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (I)gridDim.x * blockDim.x) {
    const int bv = (int)(it / nvox);
    const I v = it - (I)bv * nvox;
    const int bi = bv % B_in;
    const float sc = bv < B_in ? scale_lo : scale_hi;
    const float* ib = in + (long long)bi * ND * g.nvox;
    int pos[ND]; float f[ND];
    unravel<ND, I>(v, g.S, pos);
#pragma unroll
    for (int d = 0; d < ND; ++d) f[d] = ib[(long long)d * g.nvox + v] * sc;
    SampleSite<ND> s;
    dfmir_make_site<ND, COORD_MODE>(s, pos, f, g.S);
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const float smp = dfmir_sample<ND>(ib + (long long)d * g.nvox, s, g.S) * sc;
      out[((long long)bv * ND + d) * g.nvox + v] = __fadd_rn(f[d], smp);
    }
  }
}

// Backward of one step. g_in (pre-zeroed, shape (B_in, nd, S)) receives, via atomics,
//   sc * [ g_out (identity term) + d_flow term ] at the voxel itself and sc * w * g_out at the corners.
template <int ND, int COORD_MODE, typename I>
__global__ void __launch_bounds__(256)
vecint_step_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ in,
                       float* __restrict__ g_in, int B_in, int Bv, float scale_lo, float scale_hi,
                       Geom g) {
  const I nvox = (I)g.nvox;
  const I total = (I)Bv * nvox;
  for (I it = (I)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (I)gridDim.x * blockDim.x) {
    const int bv = (int)(it / nvox);
    const I v = it - (I)bv * nvox;
    const int bi = bv % B_in;
    const float sc = bv < B_in ? scale_lo : scale_hi;
    const float* ib = in + (long long)bi * ND * g.nvox;
    float* gb = g_in + (long long)bi * ND * g.nvox;
    int pos[ND]; float f[ND], go[ND], gf[ND];
    unravel<ND, I>(v, g.S, pos);
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      f[d] = ib[(long long)d * g.nvox + v] * sc;
      go[d] = gout[((long long)bv * ND + d) * g.nvox + v];
      gf[d] = go[d];
    }
    SampleSite<ND> s;
    dfmir_make_site<ND, COORD_MODE>(s, pos, f, g.S);
#pragma unroll
    for (int c = 0; c < (1 << ND); ++c) {
      long long o = 0; bool k = true;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const int id = s.i0[d] + ((c >> (ND - 1 - d)) & 1);
        k = k && id >= 0 && id < g.S[d];
        o = o * g.S[d] + id;
      }
      if (!k) continue;
      const float w = dfmir_corner_weight<ND>(s, c);
#pragma unroll
      for (int ch = 0; ch < ND; ++ch) {
        atomicAdd(gb + (long long)ch * g.nvox + o, sc * w * go[ch]);
        const float val = __ldg(ib + (long long)ch * g.nvox + o) * sc * go[ch];
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          float ow = 1.f;
#pragma unroll
          for (int e = 0; e < ND; ++e)
            if (e != d) ow *= ((c >> (ND - 1 - e)) & 1) ? s.w1[e] : s.w0[e];
          gf[d] += (((c >> (ND - 1 - d)) & 1) ? ow : -ow) * val;
        }
      }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) atomicAdd(gb + (long long)d * g.nvox + v, sc * gf[d]);
  }
}

int make_geom(Geom& g, int B, int C, int nd, const int* shape) {
  if (nd < 1 || nd > 3 || B < 0 || C < 0) return -1;
  g.B = B; g.C = C; g.nvox = 1;
  for (int d = 0; d < 3; ++d) {
    g.S[d] = d < nd ? shape[d] : 1;
    if (g.S[d] <= 0) return -1;
    g.nvox *= g.S[d];
  }
  return 0;
}

inline int grid_for(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)dfmir_num_sms() * 16;  // grid-stride; multiple of the SM count
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <int ND, int CM>
int launch_fwd(const float* src, const float* flow, float* out, int32_t* idx, const Geom& g, int interp,
               cudaStream_t st) {
  const bool vec4 = (g.S[ND - 1] % 4 == 0) && (((uintptr_t)flow | (uintptr_t)out) % 16 == 0);
  const long long items = (long long)g.B * g.nvox / (vec4 ? 4 : 1);
  if (items == 0) return DFMIR_OK;
  const int grid = grid_for(items, 256);
  const bool narrow = (long long)g.B * g.nvox * (g.C > ND ? g.C : ND) < (1LL << 31);
#define WARP_LAUNCH(V, NEAR)                                                                              \
  do {                                                                                                    \
    if (narrow) warp_fwd_kernel<ND, CM, V, NEAR, int><<<grid, 256, 0, st>>>(src, flow, out, idx, g);      \
    else warp_fwd_kernel<ND, CM, V, NEAR, long long><<<grid, 256, 0, st>>>(src, flow, out, idx, g);       \
  } while (0)
  if (interp == DFMIR_INTERP_NEAREST) {
    if (vec4) WARP_LAUNCH(4, true); else WARP_LAUNCH(1, true);
  } else {
    if (vec4) WARP_LAUNCH(4, false); else WARP_LAUNCH(1, false);
  }
#undef WARP_LAUNCH
  DFMIR_CHECK_LAUNCH("dfmir_warp_fwd");
  return DFMIR_OK;
}

}  // namespace

#define DISPATCH_ND_CM(nd, cm, ...)                                         \
  if (cm == DFMIR_COORD_IEEE_DIV) {                                          \
    if (nd == 1) { constexpr int ND = 1, CM = DFMIR_COORD_IEEE_DIV; __VA_ARGS__; }  \
    else if (nd == 2) { constexpr int ND = 2, CM = DFMIR_COORD_IEEE_DIV; __VA_ARGS__; } \
    else { constexpr int ND = 3, CM = DFMIR_COORD_IEEE_DIV; __VA_ARGS__; }          \
  } else {                                                                   \
    if (nd == 1) { constexpr int ND = 1, CM = DFMIR_COORD_RCP_MUL; __VA_ARGS__; }   \
    else if (nd == 2) { constexpr int ND = 2, CM = DFMIR_COORD_RCP_MUL; __VA_ARGS__; } \
    else { constexpr int ND = 3, CM = DFMIR_COORD_RCP_MUL; __VA_ARGS__; }           \
  }

extern "C" int dfmir_warp_fwd(const float* src, const float* flow, float* out, int32_t* idx_out, int B,
                              int C, int nd, const int* shape, int interp, int coord_mode,
                              void* stream) {
  Geom g;
  DFMIR_CHECK_ARG(make_geom(g, B, C, nd, shape) == 0, "dfmir_warp_fwd: bad geometry (nd=%d)", nd);
  DFMIR_CHECK_ARG(interp == DFMIR_INTERP_LINEAR || interp == DFMIR_INTERP_NEAREST,
                  "dfmir_warp_fwd: interp must be 0 (linear) or 1 (nearest)");
  DFMIR_CHECK_ARG(coord_mode == 0 || coord_mode == 1, "dfmir_warp_fwd: bad coord_mode");
  DFMIR_CHECK_ARG(src && flow && out, "dfmir_warp_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_ND_CM(nd, coord_mode, return (launch_fwd<ND, CM>(src, flow, out, idx_out, g, interp, st)));
  return DFMIR_OK;
}

extern "C" int dfmir_warp_bwd(const float* grad_out, const float* src, const float* flow, float* d_src,
                              float* d_flow, int B, int C, int nd, const int* shape, int coord_mode,
                              void* stream) {
  Geom g;
  DFMIR_CHECK_ARG(make_geom(g, B, C, nd, shape) == 0, "dfmir_warp_bwd: bad geometry (nd=%d)", nd);
  DFMIR_CHECK_ARG(coord_mode == 0 || coord_mode == 1, "dfmir_warp_bwd: bad coord_mode");
  DFMIR_CHECK_ARG(grad_out && src && flow, "dfmir_warp_bwd: null pointer");
  if (!d_src && !d_flow) return DFMIR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const long long items = (long long)g.B * g.nvox;
  if (items == 0) return DFMIR_OK;
  const int grid = grid_for(items, 256);
  const bool narrow = (long long)g.B * g.nvox * (g.C > nd ? g.C : nd) < (1LL << 31);
  if (narrow) {
    DISPATCH_ND_CM(nd, coord_mode,
                   (warp_bwd_kernel<ND, CM, int><<<grid, 256, 0, st>>>(grad_out, src, flow, d_src, d_flow, g)));
  } else {
    DISPATCH_ND_CM(nd, coord_mode,
                   (warp_bwd_kernel<ND, CM, long long><<<grid, 256, 0, st>>>(grad_out, src, flow, d_src, d_flow, g)));
  }
  DFMIR_CHECK_LAUNCH("dfmir_warp_bwd");
  return DFMIR_OK;
}

// steps: (n_slabs, Bv, nd, *S) where n_slabs = keep_all ? nsteps : 2 (ping-pong); the integrated
// field is slab (keep_all ? nsteps-1 : (nsteps-1) & 1).  Bv = bidir ? 2B : B.
extern "C" int dfmir_vecint_fwd(const float* vel, float* steps, int B, int nd, const int* shape,
                                int nsteps, int bidir, int keep_all, int coord_mode, void* stream) {
  Geom g;
  DFMIR_CHECK_ARG(make_geom(g, B, nd, nd, shape) == 0, "dfmir_vecint_fwd: bad geometry (nd=%d)", nd);
  DFMIR_CHECK_ARG(nsteps >= 1 && nsteps < 31, "dfmir_vecint_fwd: nsteps must be in [1,30], got %d", nsteps);
  DFMIR_CHECK_ARG(vel && steps, "dfmir_vecint_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int Bv = bidir ? 2 * B : B;
  const long long slab = (long long)Bv * nd * g.nvox;
  const float sc = 1.0f / (float)(1 << nsteps);
  const long long items = (long long)Bv * g.nvox;
  if (items == 0) return DFMIR_OK;
  const int grid = grid_for(items, 256);
  for (int k = 0; k < nsteps; ++k) {
    const float* in = k == 0 ? vel : steps + (long long)(keep_all ? k - 1 : (k - 1) & 1) * slab;
    float* out = steps + (long long)(keep_all ? k : k & 1) * slab;
    const int B_in = k == 0 ? B : Bv;
    const float lo = k == 0 ? sc : 1.f, hi = k == 0 ? -sc : 1.f;
    if (slab < (1LL << 31)) {
      DISPATCH_ND_CM(nd, coord_mode,
                     (vecint_step_kernel<ND, CM, int><<<grid, 256, 0, st>>>(in, out, B_in, Bv, lo, hi, g)));
    } else {
      DISPATCH_ND_CM(nd, coord_mode,
                     (vecint_step_kernel<ND, CM, long long><<<grid, 256, 0, st>>>(in, out, B_in, Bv, lo, hi, g)));
    }
    DFMIR_CHECK_LAUNCH("dfmir_vecint_fwd");
  }
  return DFMIR_OK;
}

// grad_out: (Bv, nd, *S) gradient wrt the integrated field(s); steps: the keep_all buffer of the
// forward; work: 2 slabs of scratch (Bv, nd, *S); d_vel: (B, nd, *S), overwritten.
extern "C" int dfmir_vecint_bwd(const float* grad_out, const float* vel, const float* steps, float* work,
                                float* d_vel, int B, int nd, const int* shape, int nsteps, int bidir,
                                int coord_mode, void* stream) {
  Geom g;
  DFMIR_CHECK_ARG(make_geom(g, B, nd, nd, shape) == 0, "dfmir_vecint_bwd: bad geometry (nd=%d)", nd);
  DFMIR_CHECK_ARG(nsteps >= 1 && nsteps < 31, "dfmir_vecint_bwd: nsteps must be in [1,30], got %d", nsteps);
  DFMIR_CHECK_ARG(grad_out && vel && steps && work && d_vel, "dfmir_vecint_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int Bv = bidir ? 2 * B : B;
  const long long slab = (long long)Bv * nd * g.nvox;
  const float sc = 1.0f / (float)(1 << nsteps);
  const long long items = (long long)Bv * g.nvox;
  if (items == 0) return DFMIR_OK;
  const int grid = grid_for(items, 256);
  const float* gcur = grad_out;
  for (int k = nsteps - 1; k >= 0; --k) {
    const float* in = k == 0 ? vel : steps + (long long)(k - 1) * slab;
    float* gin = k == 0 ? d_vel : work + (long long)(k & 1) * slab;
    const int B_in = k == 0 ? B : Bv;
    const float lo = k == 0 ? sc : 1.f, hi = k == 0 ? -sc : 1.f;
    DFMIR_CUDA(cudaMemsetAsync(gin, 0, sizeof(float) * (k == 0 ? (long long)B * nd * g.nvox : slab), st));
    if (slab < (1LL << 31)) {
      DISPATCH_ND_CM(nd, coord_mode,
                     (vecint_step_bwd_kernel<ND, CM, int><<<grid, 256, 0, st>>>(gcur, in, gin, B_in, Bv, lo, hi, g)));
    } else {
      DISPATCH_ND_CM(nd, coord_mode,
                     (vecint_step_bwd_kernel<ND, CM, long long><<<grid, 256, 0, st>>>(gcur, in, gin, B_in, Bv, lo, hi, g)));
    }
    DFMIR_CHECK_LAUNCH("dfmir_vecint_bwd");
    gcur = gin;
  }
  return DFMIR_OK;
}
