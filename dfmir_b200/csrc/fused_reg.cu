// K4 + K5 fused: ONE cooperative launch for the registration tail of VxmDense + its losses,
//
//   integrate (VecInt: nsteps scaling-and-squaring steps at half resolution, layers.py:64-68)
//   -> ResizeTransform fullsize (x2 linear upsampling of the field, rescaled, layers.py:91-94)
//   -> SpatialTransformer warp of the moving image (layers.py:30-48)
//   -> local NCC(warped, fixed) (util/losses.py:183-261) + Grad / smoothness of the field (util/losses.py:92-130)
//
// The reference runs this as ~130 kernels (14 grid_samples, each with its coordinate-prep elementwise
// passes, 2 interpolations, 5 dense box-filter convolutions, slices / abs / means).  Here a persistent
// grid (occupancy x 148 CTAs, all co-resident) walks the phases with grid.sync() in between:
//   phase 1..n  one squaring step each; the half-resolution fields (<= 7.4 MB) ping-pong through L2
//   phase n+1   per full-resolution voxel: interpolate the field (flow_full, an output), build the
//               sampling site, gather the moving image (warped, an output) — no grid tensor, no
//               separate resize pass
//   phase n+2   NCC tiles (shared-memory halo box sums, see ncc.cuh) + forward differences of the field,
//               reduced to per-CTA fp64 partials; after a last grid.sync() CTA 0 finalises both scalars
// Every phase calls the same device functions as the stand-alone kernels (warp.cuh, resize.cuh, ncc.cuh),
// so flow_full / warped are bit-identical to the unfused path and the losses agree to reduction order.
// The squaring steps are all kept (steps buffer) because the backward pass needs them.
#include <cooperative_groups.h>
#include "warp.cuh"
#include "resize.cuh"
#include "ncc.cuh"
#include "dfmir_b200.h"

namespace cg = cooperative_groups;

namespace {
using namespace nccdev;

struct FusedP {
  int B, C;                 // batch, channels of the moving image
  int Sh[3], Sf[3];         // half / full resolution spatial sizes (ij order, unused trailing dims = 1)
  long long nh, nf;         // voxels per half / full volume
  int nsteps;
  resizedev::RGeom rg;      // half -> full
  float pre_mul;            // ResizeTransform factor (2)
  BoxGeom box;              // NCC geometry at full resolution
  int tiles_x, tiles_y, tiles_z;
  float eps;
  int ncc_reduction, grad_penalty;
  float grad_mult;
  int G3[3];                // full-resolution dims right-aligned for the Grad loss (2-D: {1,H,W})
  DfmirFastDiv dh[3], df[3], dnh, dnf, dg1, dg2;   // exact division by Sh[d], Sf[d], nh, nf, G3[1], G3[2]
};

// element loops run on 32-bit indices (the host entry rejects volumes with B * nd * nvox >= 2^31)
template <int ND>
__device__ __forceinline__ void unravel(int v, const DfmirFastDiv* S, int* pos) {
  uint32_t u = (uint32_t)v;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) { uint32_t r; u = S[d].divmod(u, r); pos[d] = (int)r; }
}

template <int ND, int WIN, int CM>
__global__ void __launch_bounds__(NT)
fused_reg_kernel(const float* __restrict__ vel, const float* __restrict__ moving, const float* __restrict__ fixed,
                 float* __restrict__ steps, float* __restrict__ flow_full, float* __restrict__ warped,
                 double* __restrict__ partials, float* __restrict__ out, const FusedP p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double sred[5][NT / 32];
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gthreads = gridDim.x * blockDim.x;
  const int nh = (int)p.nh, nf = (int)p.nf;

  // ---- phases 1..nsteps: scaling and squaring (same arithmetic as vecint_step_kernel)
  const long long slab = (long long)p.B * ND * p.nh;
  const float sc0 = 1.0f / (float)(1 << p.nsteps);
  for (int k = 0; k < p.nsteps; ++k) {
    const float* in = k == 0 ? vel : steps + (long long)(k - 1) * slab;
    float* o = steps + (long long)k * slab;
    const float sc = k == 0 ? sc0 : 1.f;
    // two voxels per trip (all gathers of both issued before either result is stored): these phases are bound by
    // load latency at 16 warps per SM, not by bandwidth
    for (int it0 = gtid; it0 < p.B * nh; it0 += 2 * gthreads) {
      float res[2][ND]; long long oo[2]; bool live[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int it = it0 + u * gthreads;
        live[u] = it < p.B * nh;
        if (!live[u]) continue;
        uint32_t vr;
        const int b = (int)p.dnh.divmod((uint32_t)it, vr);
        const int v = (int)vr;
        const float* ib = in + (long long)b * ND * p.nh;
        int pos[ND]; float f[ND];
        unravel<ND>(v, p.dh, pos);
#pragma unroll
        for (int d = 0; d < ND; ++d) f[d] = ib[(long long)d * p.nh + v] * sc;
        SampleSite<ND> s;
        dfmir_make_site<ND, CM>(s, pos, f, p.Sh);
        CornerSet<ND> cs;
        dfmir_corners<ND>(cs, s, p.Sh);
#pragma unroll
        for (int d = 0; d < ND; ++d) {
          const float smp = dfmir_sample_corners<ND>(ib + (long long)d * p.nh, cs) * sc;
          res[u][d] = __fadd_rn(f[d], smp);
        }
        oo[u] = (long long)b * ND * p.nh + v;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (live[u])
#pragma unroll
          for (int d = 0; d < ND; ++d) o[oo[u] + (long long)d * p.nh] = res[u][d];
    }
    grid.sync();
  }

  // ---- phase nsteps+1: upsample the integrated field and warp the moving image
  {
    const float* field = steps + (long long)(p.nsteps - 1) * slab;
    // (single-channel images take the two-voxel path; more channels keep one voxel per trip)
    for (int it0 = gtid; it0 < p.B * nf; it0 += 2 * gthreads) {
      float f[2][ND], wv[2]; int bb[2], vv[2]; bool live[2];
      CornerSet<ND> cs[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int it = it0 + u * gthreads;
        live[u] = it < p.B * nf;
        if (!live[u]) continue;
        uint32_t vr;
        bb[u] = (int)p.dnf.divmod((uint32_t)it, vr);
        vv[u] = (int)vr;
        int pos[ND];
        unravel<ND>(vv[u], p.df, pos);
        int i0[ND], i1[ND]; float l0[ND], l1[ND];
        resizedev::setup_pos<ND>(p.rg, pos, i0, i1, l0, l1);        // one interpolation site for the nd components
#pragma unroll
        for (int d = 0; d < ND; ++d)
          f[u][d] = resizedev::interp_site<ND>(field + ((long long)bb[u] * ND + d) * p.nh, p.rg, i0, i1, l0, l1, p.pre_mul);   // post_mul = 1
        SampleSite<ND> s;
        dfmir_make_site<ND, CM>(s, pos, f[u], p.Sf);
        dfmir_corners<ND>(cs[u], s, p.Sf);
        wv[u] = dfmir_sample_corners<ND>(moving + (long long)bb[u] * p.C * p.nf, cs[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!live[u]) continue;
#pragma unroll
        for (int d = 0; d < ND; ++d) flow_full[((long long)bb[u] * ND + d) * p.nf + vv[u]] = f[u][d];
        warped[(long long)bb[u] * p.C * p.nf + vv[u]] = wv[u];
        for (int c = 1; c < p.C; ++c)
          warped[((long long)bb[u] * p.C + c) * p.nf + vv[u]] = dfmir_sample_corners<ND>(moving + ((long long)bb[u] * p.C + c) * p.nf, cs[u]);
      }
    }
    grid.sync();
  }

  // ---- phase nsteps+2: NCC(warped, fixed) tiles and Grad(flow_full)
  const long long hw = (long long)p.box.H * p.box.W, vol = hw * p.box.D;
  IJLoader ld{warped, fixed, vol, hw, p.box.W};
  const float wsz = (float)p.box.win * (float)p.box.win * (float)p.box.wz;
  CcReduce cs{nullptr, vol, hw, p.box.W, wsz, p.eps, 0.0, 0.0};
  const int ntiles = p.tiles_x * p.tiles_y * p.tiles_z;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tx = t % p.tiles_x, ty = (t / p.tiles_x) % p.tiles_y, tz = t / (p.tiles_x * p.tiles_y);
    box_march<5, WIN>(p.box, ld, cs, smem, tx, ty, tz);
  }
  float acc[3] = {0.f, 0.f, 0.f};
  {
    const int st1 = p.G3[2], st0 = p.G3[1] * p.G3[2];
    const int total = p.B * ND * nf;
    auto pen = [&](float a, float c) { const float d = fabsf(a - c); return p.grad_penalty == 2 ? d * d : d; };
    if ((st1 & 3) == 0) {
      // four voxels along x per thread: the centre, the y + 1 and z + 1 rows as 128-bit loads (rows are 16-byte
      // aligned: W % 4 == 0), one scalar for the x neighbour of the fourth voxel
      const float4* f4 = reinterpret_cast<const float4*>(flow_full);
      for (int i4 = gtid; i4 < total / 4; i4 += gthreads) {
        const int it = 4 * i4;
        uint32_t vr, pxr, pyr;
        p.dnf.divmod((uint32_t)it, vr);
        const uint32_t row = p.dg2.divmod(vr, pxr);
        const int pz = (int)p.dg1.divmod(row, pyr);
        const int px = (int)pxr, py = (int)pyr;
        const bool hy = py + 1 < p.G3[1], hz = pz + 1 < p.G3[0], hx = px + 4 < p.G3[2];
        const float4 c = f4[i4];
        float4 ny = c, nz = c;
        float nx = c.w;
        if (hy) ny = f4[i4 + st1 / 4];
        if (hz) nz = f4[i4 + st0 / 4];
        if (hx) nx = flow_full[it + 4];
        acc[2] += pen(c.y, c.x); acc[2] += pen(c.z, c.y); acc[2] += pen(c.w, c.z);
        if (hx) acc[2] += pen(nx, c.w);
        if (hy) { acc[1] += pen(ny.x, c.x); acc[1] += pen(ny.y, c.y); acc[1] += pen(ny.z, c.z); acc[1] += pen(ny.w, c.w); }
        if (hz) { acc[0] += pen(nz.x, c.x); acc[0] += pen(nz.y, c.y); acc[0] += pen(nz.z, c.z); acc[0] += pen(nz.w, c.w); }
      }
    } else {
      for (int it = gtid; it < total; it += gthreads) {
        uint32_t vr, pxr, pyr;
        p.dnf.divmod((uint32_t)it, vr);
        const uint32_t row = p.dg2.divmod(vr, pxr);
        const int pz = (int)p.dg1.divmod(row, pyr);
        const int px = (int)pxr, py = (int)pyr;
        const float c = flow_full[it];
        if (px + 1 < p.G3[2]) acc[2] += pen(flow_full[it + 1], c);
        if (py + 1 < p.G3[1]) acc[1] += pen(flow_full[it + st1], c);
        if (pz + 1 < p.G3[0]) acc[0] += pen(flow_full[it + st0], c);
      }
    }
  }
  {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double r[5] = {cs.acc_cc, cs.acc_m, (double)acc[0], (double)acc[1], (double)acc[2]};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double s = warp_sum_d(r[k]);
      if (lane == 0) sred[k][w] = s;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      double s = 0;
      for (int i = 0; i < NT / 32; ++i) s += sred[threadIdx.x][i];
      partials[5 * (long long)blockIdx.x + threadIdx.x] = s;
    }
  }
  grid.sync();

  // ---- finalise (CTA 0): out = {ncc loss, sum cc, voxel count, grad loss}
  if (blockIdx.x == 0) {
    double r[5] = {0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
#pragma unroll
      for (int k = 0; k < 5; ++k) r[k] += partials[5 * (long long)i + k];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double s = warp_sum_d(r[k]);
      if (lane == 0) sred[k][w] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t[5] = {0, 0, 0, 0, 0};
      for (int i = 0; i < NT / 32; ++i)
#pragma unroll
        for (int k = 0; k < 5; ++k) t[k] += sred[k][i];
      const double count = (double)p.B * (double)vol;
      out[0] = p.ncc_reduction == 0 ? -sqrtf((float)(t[0] / count)) : -(float)(t[0] / count);
      out[1] = (float)t[0]; out[2] = (float)count;
      float d = 0.f;
      const long long planes = (long long)p.B * ND;
      for (int k = 2; k >= 3 - ND; --k) {
        const double cnt = (double)planes * (double)(p.nf / p.G3[k]) * (double)(p.G3[k] - 1);
        d += (float)(t[2 + k] / cnt);
      }
      out[3] = d / (float)ND * p.grad_mult;
    }
  }
}

template <int ND, int WIN, int CM>
int launch_fused(const float* vel, const float* moving, const float* fixed, float* steps, float* flow_full, float* warped,
                 double* partials, size_t partial_slots, float* out, FusedP p, cudaStream_t st, int* grid_out) {
  auto kern = fused_reg_kernel<ND, WIN, CM>;
  const size_t sm = box_smem_bytes<5, WIN>(p.box.wz);
  DFMIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int per_sm = 0;
  DFMIR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, sm));
  if (per_sm < 1) { dfmir_set_error("dfmir_fused_reg_fwd: kernel does not fit on an SM"); return DFMIR_ERR_CUDA; }
  if (per_sm > 4) per_sm = 4;
  int grid = per_sm * dfmir_num_sms();          // all CTAs co-resident: required by grid.sync()
  if ((size_t)grid > partial_slots) grid = (int)partial_slots;
  if (grid_out) *grid_out = grid;
  void* args[] = {(void*)&vel, (void*)&moving, (void*)&fixed, (void*)&steps, (void*)&flow_full, (void*)&warped,
                  (void*)&partials, (void*)&out, (void*)&p};
  DFMIR_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(NT), args, sm, st));
  dfmir_count_launch();
  return DFMIR_OK;
}

constexpr size_t PARTIAL_SLOTS = 148 * 4;

}  // namespace

extern "C" size_t dfmir_fused_reg_workspace_bytes(void) { return sizeof(double) * 5 * PARTIAL_SLOTS + 256; }

// vel (B,nd,*half) -> steps (nsteps,B,nd,*half) [all squaring steps], flow_full (B,nd,*full), warped (B,C,*full),
// out[4] = {ncc loss, sum cc, voxel count, grad loss}.  full = 2 * half per axis.
extern "C" int dfmir_fused_reg_fwd(const float* vel, const float* moving, const float* fixed, float* steps,
                                   float* flow_full, float* warped, float* out, void* ws, size_t ws_bytes, int B, int C,
                                   int nd, const int* half_shape, int nsteps, int win, float eps, int ncc_reduction,
                                   int grad_penalty, float grad_mult, int coord_mode, void* stream) {
  const char* who = "dfmir_fused_reg_fwd";
  DFMIR_CHECK_ARG(vel && moving && fixed && steps && flow_full && warped && out && ws, "%s: null pointer", who);
  DFMIR_CHECK_ARG(nd == 2 || nd == 3, "%s: nd must be 2 or 3", who);
  DFMIR_CHECK_ARG(B >= 1 && C >= 1 && nsteps >= 1 && nsteps < 31, "%s: bad sizes (B=%d C=%d nsteps=%d)", who, B, C, nsteps);
  DFMIR_CHECK_ARG(grad_penalty == 1 || grad_penalty == 2, "%s: penalty must be 1 (l1) or 2 (l2)", who);
  DFMIR_CHECK_ARG(coord_mode == 0 || coord_mode == 1, "%s: bad coord_mode", who);
  DFMIR_CHECK_ARG(ws_bytes >= dfmir_fused_reg_workspace_bytes(), "%s: workspace too small", who);
  FusedP p{};
  p.B = B; p.C = C; p.nsteps = nsteps; p.nh = 1; p.nf = 1;
  int full_shape[3] = {1, 1, 1};
  for (int d = 0; d < 3; ++d) {
    p.Sh[d] = d < nd ? half_shape[d] : 1;
    p.Sf[d] = d < nd ? 2 * half_shape[d] : 1;
    DFMIR_CHECK_ARG(p.Sh[d] >= 1, "%s: bad shape", who);
    if (d < nd) full_shape[d] = p.Sf[d];
    p.nh *= p.Sh[d]; p.nf *= p.Sf[d];
  }
  DFMIR_CHECK_ARG((long long)B * nd * p.nf < (1LL << 31) && (long long)B * C * p.nf < (1LL << 31),
                  "%s: volume too large for the 32-bit element loops (B * nd * voxels must stay below 2^31)", who);
  p.rg.BC = B * nd; p.rg.nin = p.nh; p.rg.nout = p.nf;
  for (int d = 0; d < 3; ++d) {
    p.rg.I[d] = p.Sh[d]; p.rg.O[d] = p.Sf[d];
    p.rg.sc[d] = p.rg.O[d] > 1 ? (float)(p.rg.I[d] - 1) / (float)(p.rg.O[d] - 1) : 0.f;
  }
  p.pre_mul = 2.0f;
  for (int d = 0; d < 3; ++d) { p.dh[d] = dfmir_fastdiv((uint32_t)p.Sh[d]); p.df[d] = dfmir_fastdiv((uint32_t)p.Sf[d]); }
  p.dnh = dfmir_fastdiv((uint32_t)p.nh); p.dnf = dfmir_fastdiv((uint32_t)p.nf);
  DFMIR_CHECK_ARG(make_box(p.box, B, nd, full_shape, win) == 0, "%s: bad NCC geometry", who);
  const dim3 bg = box_grid(p.box);
  p.tiles_x = bg.x; p.tiles_y = bg.y; p.tiles_z = bg.z;
  p.eps = eps; p.ncc_reduction = ncc_reduction; p.grad_penalty = grad_penalty; p.grad_mult = grad_mult;
  for (int d = 0; d < 3; ++d) { const int src = d - (3 - nd); p.G3[d] = src >= 0 ? full_shape[src] : 1; }
  p.dg1 = dfmir_fastdiv((uint32_t)p.G3[1]); p.dg2 = dfmir_fastdiv((uint32_t)p.G3[2]);
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = (double*)ws;
#define FUSED_CASE(NDv, WINv)                                                                                        \
  return coord_mode == 0                                                                                             \
             ? launch_fused<NDv, WINv, 0>(vel, moving, fixed, steps, flow_full, warped, partials, PARTIAL_SLOTS, out, p, st, nullptr) \
             : launch_fused<NDv, WINv, 1>(vel, moving, fixed, steps, flow_full, warped, partials, PARTIAL_SLOTS, out, p, st, nullptr)
  if (nd == 2) {
    switch (win) {
      case 5: FUSED_CASE(2, 5);
      case 7: FUSED_CASE(2, 7);
      case 9: FUSED_CASE(2, 9);
      default: break;
    }
  } else {
    switch (win) {
      case 5: FUSED_CASE(3, 5);
      case 7: FUSED_CASE(3, 7);
      case 9: FUSED_CASE(3, 9);
      default: break;
    }
  }
#undef FUSED_CASE
  dfmir_set_error("%s: window %d not supported (5, 7, 9)", who, win);
  return DFMIR_ERR_UNSUPPORTED;
}
