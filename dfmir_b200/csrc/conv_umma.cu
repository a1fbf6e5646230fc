// K1 / K2 (tensor-core path): 2-D and 3-D stride-1 convolution as an implicit GEMM on the 5th-generation tensor
// cores — tcgen05.mma (kind::tf32, fp32 accumulators in TMEM), operands staged in shared memory by
// TMA, persistent CTAs (one per SM) that each walk a list of output tiles.
//
// Replaces the cuDNN implicit-GEMM engines behind nn.Conv2d in ResnetGenerator
// (models/networks.py:995,1016 and the 18 ResnetBlock convs :1201,1214) and behind the stride-1
// nn.Conv2d / nn.Conv3d of VoxelMorph's U-Net (vxm networks.py:1515,1077).  Any output channel count
// (tile widths 16..128, missing channels are zero-filled weight rows), reduction-side channel counts
// that are multiples of 4 (a partial 32-channel chunk is zero-filled by the TMA unit).  The reference's
// own CUDA path runs these in TF32 (cuDNN default); this kernel does the same arithmetic class:
// fp32 storage, TF32 operands (the tensor core truncates fp32 to 10 mantissa bits), fp32 accumulate.
//
// GEMM view per work item: MT sub-tiles of 128 pixels x BN output channels,
//     D_j[128][BN] = sum over (tap, 32-channel chunk) A_j,tap[128][32] * W_tap[BN][32]^T ,  j < MT
//   A_j,tap : the (TH x TW) pixel rectangle of sub-tile j shifted by the tap, fetched from the
//           channels-last activation by ONE tiled TMA box load {32 ch, TW, TH, 1}: rows land as
//           128-byte, 128B-swizzled K-major rows — the canonical UMMA operand layout; taps that
//           reach outside the image are zero-filled by the TMA unit (zero padding for free).
//           Reflection padding is materialised by the producer kernel (norm_resample.cu), so the
//           ResnetBlock convs run here with pad 0 on the padded buffer.
//   W_tap : weights pre-arranged [tap][Cout][Cin] (Cin contiguous), box {32, BN, 1}; ONE weight
//           tile per pipeline stage feeds all MT sub-tiles, which is what keeps the operand
//           stream (L2 -> shared memory, the binding resource for 4-byte TF32 operands) below
//           the tensor pipe's appetite: (MT*16 + BN/8) KB per MT*4 MMAs.
// TMEM: MT*BN accumulator columns per stage, AS stages (512 columns in all): with AS = 2 the
// epilogue of one work item overlaps the MMAs of the next.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..9 = epilogue (TMEM -> registers -> bias/activation -> global, one pixel row per thread;
// two warps share each TMEM lane quarter and split the sub-tiles).
// The data gradient is the same kernel run on dy with taps flipped and pad' = k-1-pad.
#include <type_traits>
#include "umma.cuh"
#include "dfmir_b200.h"
#include <stdlib.h>

namespace {
using namespace umma;

constexpr int BM = 128;          // pixels per sub-tile (UMMA M)
constexpr int KCH = 32;          // tf32 elements per 128-byte swizzled row
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction
constexpr int A_BYTES = BM * 128;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;

struct UmmaP {
  int N, D, H, W;         // output sample count and spatial size (2-D: D = 1)
  int Cin, Cout;          // K-side / N-side channel counts of this product
  int KD, KH, KW, pad_d, pad_h, pad_w;
  int TD, TH, TW, tiles_d, tiles_h, tiles_w;
  int ptiles;             // N * tiles_d * tiles_h * tiles_w sub-tiles of 128 voxels
  int flip;               // 1: use tap (taps-1-t) of the weight tensor (data gradient)
  int act;
  int per_sample;         // 1: the "tap" coordinate of the weight map is the sample index (batched A[b] * B[b]^T)
  int accum;              // 1: y += result (CTA-pair kernel only: the residual branch's gradient is already in y)
  long long ys[5];        // output element strides n, d, h, w, c
  float2* stat_rows;      // nullable: per-(32-voxel row group, channel) {sum, sum of squares} of the stored result, the
                          // InstanceNorm statistics of the layer that follows (halo / CTA-pair kernels), [row][Cout]
};

// Column sums of a 32 x 32 register tile held one row per lane: after five exchange steps lane L holds the sum over
// the 32 lanes of element L (31 shuffles instead of 160 for 32 separate warp reductions).  v is destroyed.
template <int K>
__device__ __forceinline__ void col_sum_step(float* v, int lane) {
  const bool up = (lane & K) != 0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float send = up ? v[j] : v[j + K];
    const float keep = up ? v[j + K] : v[j];
    v[j] = keep + __shfl_xor_sync(0xffffffffu, send, K);
  }
}
__device__ __forceinline__ float col_sum32(float* v, int lane) {
  col_sum_step<16>(v, lane); col_sum_step<8>(v, lane); col_sum_step<4>(v, lane); col_sum_step<2>(v, lane); col_sum_step<1>(v, lane);
  return v[0];
}
// InstanceNorm statistics by-product of an epilogue: v = this lane's 32 stored channel values (garbage if !valid)
__device__ __forceinline__ void emit_stat_row(float* v, bool valid, int lane, float2* row, int c_first, int Cout) {
  float sq[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { v[i] = valid ? v[i] : 0.f; sq[i] = v[i] * v[i]; }
  const float s1 = col_sum32(v, lane), s2 = col_sum32(sq, lane);
  if (c_first + lane < Cout) row[c_first + lane] = make_float2(s1, s2);
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == DFMIR_ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  if (act == DFMIR_ACT_TANH) return tanhf(v);
  if (act == DFMIR_ACT_RELU) return v > 0.f ? v : 0.f;
  return v;
}

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
template <int BN, int MT, int STAGES, int AS>
struct SmemLayout {
  static_assert(MT * BN * AS <= 512, "TMEM has 512 columns");
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int NBARS = 2 * STAGES + 2 * AS;
  static constexpr int TOTAL = BAR_OFF + NBARS * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int MT, int STAGES, int AS>
__global__ void __launch_bounds__(THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const float* __restrict__ bias, float* __restrict__ y, const UmmaP p) {
  using L = SmemLayout<BN, MT, STAGES, AS>;
  constexpr int CW = BN < 32 ? BN : 32;       // accumulator columns per TMEM load
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B: 1024-byte aligned
  uint64_t* full = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + AS;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + AS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int pgroups = (p.ptiles + MT - 1) / MT;
  const int items = pgroups * n_tiles;
  const int cchunks = (p.Cin + KCH - 1) / KCH;   // a partial last chunk is zero-filled by the TMA unit
  const int taps = p.KD * p.KH * p.KW;
  const int num_kb = taps * cchunks;
  const int tiles_hw = p.tiles_h * p.tiles_w;
  const int tiles_per_img = p.tiles_d * tiles_hw;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int a = 0; a < AS; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      uint32_t it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int nt = item % n_tiles, pg = item / n_tiles;
        int sn[MT], sd[MT], sh[MT], sw[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int pt = pg * MT + j;
          if (pt < p.ptiles) {
            const int n = pt / tiles_per_img; int rem = pt - n * tiles_per_img;
            const int td_i = rem / tiles_hw; rem -= td_i * tiles_hw;
            const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
            sn[j] = n; sd[j] = td_i * p.TD; sh[j] = th_i * p.TH; sw[j] = tw_i * p.TW;
          } else { sn[j] = p.N; sd[j] = 0; sh[j] = 0; sw[j] = 0; }      // beyond the batch: the TMA unit zero-fills
        }
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty + s, ph ^ 1);
          const int tap = kb / cchunks, cc = kb - tap * cchunks;
          const int q = tap % p.KW; const int t2 = tap / p.KW;
          const int r = t2 % p.KH, kd = t2 / p.KH;
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          mbar_expect_tx(full + s, (uint32_t)L::STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < MT; ++j)
            tma_load_5d(sa + j * A_BYTES, &tmA, full + s, cc * KCH, sw[j] + q - p.pad_w, sh[j] + r - p.pad_h,
                        sd[j] + kd - p.pad_d, sn[j]);
          tma_load_3d(sa + MT * A_BYTES, &tmB, full + s, cc * KCH, nt * BN,
                      p.per_sample ? sn[0] : (p.flip ? taps - 1 - tap : tap));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      constexpr uint32_t idesc = instr_desc_tf32(BM, BN);
      uint32_t it = 0, ti = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
        const uint32_t as = ti % AS, aph = (ti / AS) & 1;
        mbar_wait(tmem_empty + as, aph ^ 1);        // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t acc = tmem_base + as * (MT * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
          const uint64_t bdesc = smem_desc_sw128(sa + MT * A_BYTES, 16, 1024);
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            const uint64_t adesc = smem_desc_sw128(sa + j * A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < KCH / UMMA_K; ++k) {
              // advance both operands by 32 bytes (8 tf32) inside the swizzled row: +2 in 16-byte units
              umma_tf32(acc + j * BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            }
          }
          umma_commit(empty + s);            // frees the smem stage when these MMAs retire
        }
        umma_commit(tmem_full + as);         // accumulators of this work item complete
      }
    }
  } else {
    // ---------------- epilogue: warp e reads TMEM lanes 32*(e%4) .. +31 of sub-tiles e/4, e/4 + 2, ...
    const int e = warp - 2;
    const int quarter = warp & 3;            // hardware rule: a warp may access TMEM lanes 32*(warp%4)..+31
    const int jfirst = e >> 2;
    const int row = quarter * 32 + lane;
    const int thw = p.TH * p.TW;
    const int td = row / thw, th = (row - td * thw) / p.TW, tw = row % p.TW;
    const bool vec_ok = p.ys[4] == 1 && (p.Cout & 3) == 0;
    uint32_t ti = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
      const uint32_t as = ti % AS, aph = (ti / AS) & 1;
      const int nt = item % n_tiles, pg = item / n_tiles;
      const int n0 = nt * BN;
      mbar_wait(tmem_full + as, aph);
      tc_fence_after();
#pragma unroll 1
      for (int j = jfirst; j < MT; j += EPI_WARPS / 4) {
        const int pt = pg * MT + j;
        const int n = pt / tiles_per_img; int rem = pt - n * tiles_per_img;
        const int td_i = rem / tiles_hw; rem -= td_i * tiles_hw;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        const int od = td_i * p.TD + td, oh = th_i * p.TH + th, ow = tw_i * p.TW + tw;
        const bool valid = pt < p.ptiles && od < p.D && oh < p.H && ow < p.W;
        float* yp = y + (long long)n * p.ys[0] + (long long)od * p.ys[1] + (long long)oh * p.ys[2] + (long long)ow * p.ys[3];
        const uint32_t acc = tmem_base + as * (MT * BN) + j * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CW) {
          float v[CW];
          if (CW == 32) tmem_ld_32x32(acc + (uint32_t)c0, v); else tmem_ld_32x16(acc + (uint32_t)c0, v);
          if (valid && n0 + c0 < p.Cout) {
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              float t = v[i];
              if (bias && n0 + c0 + i < p.Cout) t += __ldg(bias + n0 + c0 + i);
              v[i] = act_apply(t, p.act);
            }
            if (vec_ok && n0 + c0 + CW <= p.Cout) {
              float4* dst = reinterpret_cast<float4*>(yp + n0 + c0);
#pragma unroll
              for (int i = 0; i < CW / 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < CW; ++i)
                if (n0 + c0 + i < p.Cout) yp[(long long)(n0 + c0 + i) * p.ys[4]] = v[i];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// ---------------------------------------------------------------- halo variant
// Same product, different operand delivery: per 32-channel chunk ONE activation tile WITH ITS HALO is
// loaded (box {32, 8*SW+KW-1, 16*SH+KH-1, SD+KD-1, 1}, rows in (d,h,w) order), and every tap reads its
// shifted window of that tile through the UMMA descriptor: start address = the window's first row, stride
// between 8-row groups = the halo row pitch.  SWIZZLE_128B is a function of the absolute shared-memory
// address, so a window may start at any 128-byte row and use any 128-byte-multiple group stride (measured:
// tools/umma_shift_probe.cu).  A sub-tile is 16 (h) x 8 (w) voxels at one depth: its sixteen 8-row groups
// are exactly (8*SW+KW-1) rows apart.  Activation traffic per output voxel drops from taps x 128 B to
// ~1.3-2 x 128 B per chunk (9x / 27x fewer L2 -> shared-memory bytes in 2-D / 3-D); the weights stream as
// before, one {32, BN} box per (tap, chunk) through their own deeper pipeline.
// compile-time loop: f(integral_constant<int, I>) for I = 0 .. N-1 (#pragma unroll leaves loops around mbarrier waits rolled)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

template <int BN, int SD, int SH, int SW, int BST>
struct HaloCfg {
  static constexpr int MT = SD * SH * SW;
  static constexpr int AS = (2 * MT * BN <= 512) ? 2 : 1;
  static constexpr int B_BYTES = BN * 128;
};

struct HaloP {
  UmmaP u;
  int HD, HH, HW;          // halo tile extents
  int a_bytes;             // bytes of one activation stage (rows * 128, rounded up to 1024)
};

// KDT > 0: the kernel is 3 x 3 with depth KDT (1 or 3), known at compile time: the tap loops of the TMA and MMA
// threads unroll completely and every descriptor is the stage's base descriptor plus a constant, which keeps the
// single MMA-issuing thread at ~6 instructions per tcgen05.mma (with run-time kernel extents it spent ~270
// instructions per tap and the tensor pipe waited for it: 59 % active on the ResnetBlock conv).  KDT = 0: any extents.
template <int BN, int SD, int SH, int SW, int BST, int KDT>
__global__ void __launch_bounds__(THREADS, 1)
conv_umma_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const float* __restrict__ bias, float* __restrict__ y, const HaloP hp) {
  using C = HaloCfg<BN, SD, SH, SW, BST>;
  constexpr int MT = C::MT, AS = C::AS;
  constexpr int CW = BN < 32 ? BN : 32;
  const UmmaP& p = hp.u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                              // 2 activation stages
  uint8_t* sB = smem + 2 * hp.a_bytes;             // BST weight stages
  uint64_t* a_full = (uint64_t*)(sB + BST * C::B_BYTES);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + BST;
  uint64_t* tmem_full = b_empty + BST;
  uint64_t* tmem_empty = tmem_full + AS;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + AS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int tiles_hw = p.tiles_h * p.tiles_w;
  const int tiles_per_img = p.tiles_d * tiles_hw;
  const int items = p.ptiles * n_tiles;            // ptiles = CTA tiles here (each MT sub-tiles)
  const int cchunks = (p.Cin + KCH - 1) / KCH;
  const int taps = p.KD * p.KH * p.KW;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < 2; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < BST; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
    for (int a = 0; a < AS; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      uint32_t ia = 0, ib = 0;
      int sbp = 0; uint32_t bphp = 0;              // weight stage / phase counters of the unrolled variant
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int nt = item % n_tiles, pt = item / n_tiles;
        const int n = pt / tiles_per_img; int rem = pt - n * tiles_per_img;
        const int td_i = rem / tiles_hw; rem -= td_i * tiles_hw;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        const int d0 = td_i * SD, h0 = th_i * 16 * SH, w0 = tw_i * 8 * SW;
        for (int cc = 0; cc < cchunks; ++cc, ++ia) {
          const int sa = ia & 1;
          mbar_wait(a_empty + sa, ((ia >> 1) & 1) ^ 1);
          mbar_expect_tx(a_full + sa, (uint32_t)(hp.HD * hp.HH * hp.HW * 128));
          tma_load_5d(sA + sa * hp.a_bytes, &tmA, a_full + sa, cc * KCH, w0 - p.pad_w, h0 - p.pad_h, d0 - p.pad_d, n);
          if constexpr (KDT > 0) {
            for (int tap = 0; tap < KDT * 9; ++tap) {
              mbar_wait(b_empty + sbp, bphp ^ 1);
              mbar_expect_tx(b_full + sbp, (uint32_t)C::B_BYTES);
              tma_load_3d(sB + sbp * C::B_BYTES, &tmB, b_full + sbp, cc * KCH, nt * BN, p.flip ? KDT * 9 - 1 - tap : tap);
              if (++sbp == BST) { sbp = 0; bphp ^= 1; }
            }
          } else {
            for (int tap = 0; tap < taps; ++tap, ++ib) {
              const int sb = ib % BST;
              mbar_wait(b_empty + sb, ((ib / BST) & 1) ^ 1);
              mbar_expect_tx(b_full + sb, (uint32_t)C::B_BYTES);
              tma_load_3d(sB + sb * C::B_BYTES, &tmB, b_full + sb, cc * KCH, nt * BN, p.flip ? taps - 1 - tap : tap);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer
      constexpr uint32_t idesc = instr_desc_tf32(BM, BN);
      const uint32_t sbo = (uint32_t)hp.HW * 128u;             // next h row of the halo tile
      const uint64_t bdesc0 = smem_desc_sw128(smem_u32(sB), 16, 1024);
      int sbm = 0; uint32_t bphm = 0;              // weight stage / phase counters of the unrolled variant
      uint32_t ia = 0, ib = 0, ti = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
        const uint32_t as = ti % AS, aph = (ti / AS) & 1;
        mbar_wait(tmem_empty + as, aph ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + as * (MT * BN);
        for (int cc = 0; cc < cchunks; ++cc, ++ia) {
          const int sa = ia & 1;
          mbar_wait(a_full + sa, (ia >> 1) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + sa * hp.a_bytes);
          // a partial last chunk (Cin = 16, 36, ...) holds zero-filled channels: issue only the K slices with data
          const int rem_c = p.Cin - cc * KCH;
          const int ksteps = rem_c >= KCH ? KCH / UMMA_K : (rem_c + UMMA_K - 1) / UMMA_K;
          if constexpr (KDT > 0) {
            constexpr int HH = 16 * SH + 2, HW = 8 * SW + 2;
            const uint64_t adesc0 = smem_desc_sw128(a_base, 16, HW * 128);     // 8-row groups one halo row apart
            auto issue = [&](auto full_c) {
              constexpr bool FULL = decltype(full_c)::value;          // all four K slices of the chunk hold data
              static_for<0, KDT * 9>([&](auto tap_c) {
                constexpr int tap = decltype(tap_c)::value;
                constexpr int kd = tap / 9, r = (tap / 3) % 3, q = tap % 3;
                mbar_wait(b_full + sbm, bphm);
                tc_fence_after();
                const uint64_t bdesc = bdesc0 + (uint64_t)(sbm * (C::B_BYTES / 16));
#pragma unroll
                for (int j = 0; j < MT; ++j) {
                  const int sw = j % SW, sh = (j / SW) % SH, sd = j / (SW * SH);
                  const int row = ((sd + kd) * HH + 16 * sh + r) * HW + 8 * sw + q;      // compile-time constant
                  const uint64_t adesc = adesc0 + (uint64_t)(8 * row);
#pragma unroll
                  for (int k = 0; k < KCH / UMMA_K; ++k)
                    if (FULL || k < ksteps)
                      umma_tf32(acc + j * BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (tap | k) != 0 ? 1u : (uint32_t)(cc != 0));
                }
                umma_commit(b_empty + sbm);
                if (++sbm == BST) { sbm = 0; bphm ^= 1; }
              });
            };
            if (ksteps == KCH / UMMA_K) issue(std::true_type{}); else issue(std::false_type{});
            umma_commit(a_empty + sa);
            continue;
          }
          int tap = 0;
          for (int kd = 0; kd < p.KD; ++kd)
            for (int r = 0; r < p.KH; ++r)
              for (int q = 0; q < p.KW; ++q, ++tap, ++ib) {
                const int sb = ib % BST;
                mbar_wait(b_full + sb, (ib / BST) & 1);
                tc_fence_after();
                const uint64_t bdesc = smem_desc_sw128(smem_u32(sB + sb * C::B_BYTES), 16, 1024);
#pragma unroll
                for (int j = 0; j < MT; ++j) {
                  const int sw = j % SW, sh = (j / SW) % SH, sd = j / (SW * SH);
                  const uint32_t row = (uint32_t)(((sd + kd) * hp.HH + 16 * sh + r) * hp.HW + 8 * sw + q);
                  const uint64_t adesc = smem_desc_sw128(a_base + row * 128u, 16, sbo);
#pragma unroll
                  for (int k = 0; k < KCH / UMMA_K; ++k)
                    if (k < ksteps) umma_tf32(acc + j * BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (cc | tap | k) != 0);
                }
                umma_commit(b_empty + sb);
              }
          umma_commit(a_empty + sa);           // every tap of this chunk has been issued: frees the halo tile when they retire
        }
        umma_commit(tmem_full + as);
      }
    }
  } else {
    // ---------------- epilogue
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int jfirst = e >> 2;
    const int row = quarter * 32 + lane;
    const int th = row >> 3, tw = row & 7;
    const bool vec_ok = p.ys[4] == 1 && (p.Cout & 3) == 0;
    uint32_t ti = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++ti) {
      const uint32_t as = ti % AS, aph = (ti / AS) & 1;
      const int nt = item % n_tiles, pt = item / n_tiles;
      const int n0 = nt * BN;
      const int n = pt / tiles_per_img; int rem = pt - n * tiles_per_img;
      const int td_i = rem / tiles_hw; rem -= td_i * tiles_hw;
      const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
      mbar_wait(tmem_full + as, aph);
      tc_fence_after();
#pragma unroll 1
      for (int j = jfirst; j < MT; j += EPI_WARPS / 4) {
        const int sw = j % SW, sh = (j / SW) % SH, sd = j / (SW * SH);
        const int od = td_i * SD + sd, oh = th_i * 16 * SH + 16 * sh + th, ow = tw_i * 8 * SW + 8 * sw + tw;
        const bool valid = od < p.D && oh < p.H && ow < p.W;
        float* yp = y + (long long)n * p.ys[0] + (long long)od * p.ys[1] + (long long)oh * p.ys[2] + (long long)ow * p.ys[3];
        const uint32_t acc = tmem_base + as * (MT * BN) + j * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CW) {
          float v[CW];
          if (CW == 32) tmem_ld_32x32(acc + (uint32_t)c0, v); else tmem_ld_32x16(acc + (uint32_t)c0, v);
          if (valid && n0 + c0 < p.Cout) {
#pragma unroll
            for (int i = 0; i < CW; ++i) {
              float t = v[i];
              if (bias && n0 + c0 + i < p.Cout) t += __ldg(bias + n0 + c0 + i);
              v[i] = act_apply(t, p.act);
            }
            if (vec_ok && n0 + c0 + CW <= p.Cout) {
              float4* dst = reinterpret_cast<float4*>(yp + n0 + c0);
#pragma unroll
              for (int i = 0; i < CW / 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < CW; ++i)
                if (n0 + c0 + i < p.Cout) yp[(long long)(n0 + c0 + i) * p.ys[4]] = v[i];
            }
          }
          if constexpr (CW == 32) {
            if (p.stat_rows)         // uniform per launch; every lane takes part in the exchange
              emit_stat_row(v, valid, lane, p.stat_rows + ((long long)(pt * MT + j) * 4 + quarter) * p.Cout, n0 + c0, p.Cout);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// ---------------------------------------------------------------- depth-march variant: 3 x 3 x 3, thin layers (BN <= 64)
// The halo kernel above is bound by the tensor core's shared-memory operand reads, ~74 bytes / clock measured: a
// 128 x BN x 8 TF32 MMA reads 4 KB of activations however narrow BN is, and a 3 x 3 x 3 layer re-reads every activation
// row 27 times - 82 clocks per output voxel on the 36 -> 16 channel full-resolution layer of the registration U-Net
// (7 % of the tensor pipe, 11 % of HBM).  Here a CTA owns a 16 x 8 (h, w) column of voxels and MARCHES along depth:
// for every INPUT slice u one halo tile {32 ch, 10, 18} is loaded, and each of its nine in-plane taps is multiplied by
// the weights of all THREE depth taps at once - B = [kd][BN] rows, one MMA of N = 3 * BN - because slice u feeds output
// slice u + pad - kd through depth tap kd.  The three column groups of that MMA are the accumulators of three
// consecutive output slices: TMEM is a ring of 512 / BN slots, slot(step) moves down by one per input slice, so an
// output slice is overwritten by its kd = 0 product, accumulated by kd = 1 and kd = 2 of the next two steps, and then
// complete: the epilogue drains one slot per step while the MMAs of the following steps run.  Activation reads by the
// tensor core drop 3x (27 -> 9 per row), the halo tile of a slice is fetched once per column instead of once per
// SD + 2 slices, and N grows from BN to 3 * BN.
// tcgen05.mma with the accumulate flag known at compile time (no predicate set-up in the issuing thread, whose
// instruction rate is what bounds these thin layers: ~7 clocks per instruction for a lone warp)
template <bool ACC>
__device__ __forceinline__ void umma_tf32_c(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}

// RES: the whole weight tensor of the layer (<= RES_CHUNKS chunks x 9 in-plane taps x [3][BN] rows) stays in shared memory
// for the life of the CTA instead of streaming through BST stages once per input slice - on the 36 -> 16 layer the
// streamed weights were 70 % of the L2 -> shared-memory traffic (108 of 154 KB per slice, 5 TB/s over the chip).
template <int BN, bool RES = false>
struct DmCfg {
  static constexpr int R = 512 / BN;                      // accumulator slots
  static constexpr int RES_CHUNKS = BN == 16 ? 2 : 1;     // 110.6 KB of weights (BN = 16, 32), 162 KB at BN = 48
  static constexpr int AST = (RES && BN == 48) ? 2 : 4;   // activation stages (one input slice x 32 channels each)
  static constexpr int BST = BN <= 32 ? 9 : 3;            // weight stages (one in-plane tap x 3 depth taps each); divides 9, so
                                                          // that the stage of a tap is a compile-time constant
  static constexpr int NB = 4;                            // steps the epilogue may lag behind the MMA issuer
  static constexpr int HH = 18, HW = 10;
  static constexpr int A_ST = (HH * HW * 128 + 1023) / 1024 * 1024;
  static constexpr int B_ST = 3 * BN * 128;
  static constexpr int NBARS = 2 * AST + 2 * BST + 2 * NB;
  static constexpr int B_BYTES = RES ? RES_CHUNKS * 9 * B_ST : BST * B_ST;
  static constexpr int SMEM = AST * A_ST + B_BYTES + NBARS * 8 + 16 + 1024;
  static_assert(R >= NB + 4, "the ring must hold the slots in flight");
  static_assert(!RES || BN <= 48, "resident weights: 16 / 32 / 48-channel tiles");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

struct DmP {
  UmmaP u;
  int zchunk, nzc;         // output slices per work item, items per (h, w) column
};

template <int BN, bool RES>
__global__ void __launch_bounds__(THREADS, 1)
conv_umma_dmarch_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const float* __restrict__ bias, float* __restrict__ y, const DmP dp) {
  using C = DmCfg<BN, RES>;
  constexpr int R = C::R, AST = C::AST, BST = C::BST, NB = C::NB;
  const UmmaP& p = dp.u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + AST * C::A_ST;
  uint64_t* a_full = (uint64_t*)(sB + C::B_BYTES);
  uint64_t* a_empty = a_full + AST;
  uint64_t* b_full = a_empty + AST;
  uint64_t* b_empty = b_full + BST;
  uint64_t* acc_full = b_empty + BST;
  uint64_t* acc_empty = acc_full + NB;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + NB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int tiles_hw = p.tiles_h * p.tiles_w;
  const int items = p.N * tiles_hw * dp.nzc * n_tiles;
  const int cchunks = (p.Cin + KCH - 1) / KCH;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < AST; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < BST; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> output-channel tile, depth chunk, (h, w) column, image
  auto decode = [&](int item, int& nt, int& n, int& h0, int& w0, int& z0, int& z1) {
    nt = item % n_tiles; int rest = item / n_tiles;
    const int zc = rest % dp.nzc; rest /= dp.nzc;
    const int tw_i = rest % p.tiles_w; rest /= p.tiles_w;
    const int th_i = rest % p.tiles_h; n = rest / p.tiles_h;
    h0 = th_i * 16; w0 = tw_i * 8;
    z0 = zc * dp.zchunk; z1 = min(p.D, z0 + dp.zchunk);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      uint32_t ia = 0, ib = 0;
      if constexpr (RES) {          // every weight tile once: stage (cc * 9 + rq), one barrier for all of them
        mbar_expect_tx(b_full, (uint32_t)(cchunks * 9 * C::B_ST));
        for (int cc = 0; cc < cchunks; ++cc)
          for (int tap = 0; tap < 27; ++tap)
            tma_load_3d(sB + (cc * 9 + tap % 9) * C::B_ST + (tap / 9) * BN * 128, &tmB, b_full, cc * KCH, 0, p.flip ? 26 - tap : tap);
      }
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int nt, n, h0, w0, z0, z1;
        decode(item, nt, n, h0, w0, z0, z1);
        for (int u = z0 - p.pad_d; u <= z1 + 1 - p.pad_d; ++u)
          for (int cc = 0; cc < cchunks; ++cc, ++ia) {
            const int sa = ia % AST;
            mbar_wait(a_empty + sa, ((ia / AST) & 1) ^ 1);
            mbar_expect_tx(a_full + sa, (uint32_t)(C::HH * C::HW * 128));
            tma_load_5d(sA + sa * C::A_ST, &tmA, a_full + sa, cc * KCH, w0 - p.pad_w, h0 - p.pad_h, u, n);
            if constexpr (RES) continue;
            // tap rq always uses weight stage rq % BST; a stage is used 9 / BST (odd) times per chunk, so the phase of its
            // j-th use in chunk number ib is (ib + j) & 1
#pragma unroll
            for (int rq = 0; rq < 9; ++rq) {
              const int sb = rq % BST;
              mbar_wait(b_empty + sb, ((ib + rq / BST) & 1) ^ 1);
              mbar_expect_tx(b_full + sb, (uint32_t)C::B_ST);
#pragma unroll
              for (int kd = 0; kd < 3; ++kd) {
                const int tap = kd * 9 + rq;
                tma_load_3d(sB + sb * C::B_ST + kd * BN * 128, &tmB, b_full + sb, cc * KCH, nt * BN, p.flip ? 26 - tap : tap);
              }
            }
            ++ib;
          }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issuer: the whole warp walks the loops (converged), one elected lane issues
      constexpr uint32_t idesc1 = instr_desc_tf32(BM, BN), idesc2 = instr_desc_tf32(BM, 2 * BN), idesc3 = instr_desc_tf32(BM, 3 * BN);
      constexpr uint64_t GROUP = (uint64_t)(BN * 128 / 16);          // one depth tap's weight rows, in descriptor units
      const uint64_t bdesc0 = smem_desc_sw128(smem_u32(sB), 16, 1024);
      // One 32-channel chunk of one input slice: nine in-plane taps, KS K-slices each.  Everything that shapes the
      // instruction stream is a template argument - KS, whether the three-slot window wraps around the ring (WRAP 1:
      // slots R-2, R-1 | 0; WRAP 2: R-1 | 0, 1), whether this is the step's first chunk (its very first product
      // overwrites the newest slot) - so that the issuing thread runs ~7 instructions per tcgen05.mma.
      auto issue_chunk = [&](auto ks_c, auto wrap_c, auto first_c, uint32_t bph, uint64_t adesc0, int s0, uint64_t bdesc_cc) {
        constexpr int KS = decltype(ks_c)::value, WRAP = decltype(wrap_c)::value;
        constexpr bool FIRST = decltype(first_c)::value;
        const uint32_t c0 = tmem_base + (uint32_t)(s0 * BN);
        static_for<0, 9>([&](auto rq_c) {
          constexpr int rq = decltype(rq_c)::value;
          constexpr int sb = RES ? rq : rq % BST;
          constexpr int row = (rq / 3) * C::HW + rq % 3;
          if constexpr (!RES) {
            mbar_wait(b_full + sb, (bph + rq / BST) & 1);
            tc_fence_after();
          }
          const uint64_t bdesc = bdesc_cc + (uint64_t)(sb * (C::B_ST / 16));
          const uint64_t adesc = adesc0 + (uint64_t)(8 * row);
          static_for<0, KS>([&](auto k_c) {
            constexpr int k = decltype(k_c)::value;
            const uint64_t ak = adesc + (uint64_t)(2 * k), bk = bdesc + (uint64_t)(2 * k);
            if constexpr (FIRST && rq == 0 && k == 0) {
              umma_tf32_elect<false>(c0, ak, bk, idesc1);
              umma_tf32_elect<true>(tmem_base + (uint32_t)(((s0 + 1) % R) * BN), ak, bk + GROUP, idesc1);
              umma_tf32_elect<true>(tmem_base + (uint32_t)(((s0 + 2) % R) * BN), ak, bk + 2 * GROUP, idesc1);
            } else if constexpr (WRAP == 0) {
              umma_tf32_elect<true>(c0, ak, bk, idesc3);
            } else if constexpr (WRAP == 1) {
              umma_tf32_elect<true>(c0, ak, bk, idesc2);
              umma_tf32_elect<true>(tmem_base, ak, bk + 2 * GROUP, idesc1);
            } else {
              umma_tf32_elect<true>(c0, ak, bk, idesc1);
              umma_tf32_elect<true>(tmem_base, ak, bk + GROUP, idesc2);
            }
          });
          if constexpr (!RES) umma_commit_elect(b_empty + sb);
        });
      };
      auto issue_ks = [&](auto wrap_c, auto first_c, int ksteps, uint32_t bph, uint64_t adesc0, int s0, uint64_t bd) {
        switch (ksteps) {
          case 1: issue_chunk(std::integral_constant<int, 1>{}, wrap_c, first_c, bph, adesc0, s0, bd); break;
          case 2: issue_chunk(std::integral_constant<int, 2>{}, wrap_c, first_c, bph, adesc0, s0, bd); break;
          case 3: issue_chunk(std::integral_constant<int, 3>{}, wrap_c, first_c, bph, adesc0, s0, bd); break;
          default: issue_chunk(std::integral_constant<int, 4>{}, wrap_c, first_c, bph, adesc0, s0, bd); break;
        }
      };
      auto issue_wrap = [&](auto first_c, int wrap, int ksteps, uint32_t bph, uint64_t adesc0, int s0, uint64_t bd) {
        if (wrap == 0) issue_ks(std::integral_constant<int, 0>{}, first_c, ksteps, bph, adesc0, s0, bd);
        else if (wrap == 1) issue_ks(std::integral_constant<int, 1>{}, first_c, ksteps, bph, adesc0, s0, bd);
        else issue_ks(std::integral_constant<int, 2>{}, first_c, ksteps, bph, adesc0, s0, bd);
      };
      if constexpr (RES) { mbar_wait(b_full, 0); tc_fence_after(); }
      uint32_t ia = 0, ib = 0, st = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int nt, n, h0, w0, z0, z1;
        decode(item, nt, n, h0, w0, z0, z1);
        const int nsteps = z1 - z0 + 2;
        for (int j = 0; j < nsteps; ++j, ++st) {
          const uint32_t bi = st % NB;
          mbar_wait(acc_empty + bi, ((st / NB) & 1) ^ 1);        // the epilogue has drained step st - NB
          tc_fence_after();
          const int s0 = (R - (int)(st % R)) % R;                // slot of the output slice first touched by this step
          const int wrap = s0 + 2 < R ? 0 : (s0 + 2 == R ? 1 : 2);
          for (int cc = 0; cc < cchunks; ++cc, ++ia, ++ib) {
            const int sa = ia % AST;
            mbar_wait(a_full + sa, (ia / AST) & 1);
            tc_fence_after();
            const uint64_t adesc0 = smem_desc_sw128(smem_u32(sA + sa * C::A_ST), 16, C::HW * 128);
            const int rem_c = p.Cin - cc * KCH;
            const int ksteps = rem_c >= KCH ? KCH / UMMA_K : (rem_c + UMMA_K - 1) / UMMA_K;
            const uint64_t bd = RES ? bdesc0 + (uint64_t)(cc * 9 * (C::B_ST / 16)) : bdesc0;
            if (cc == 0) issue_wrap(std::true_type{}, wrap, ksteps, ib & 1, adesc0, s0, bd);
            else issue_wrap(std::false_type{}, wrap, ksteps, ib & 1, adesc0, s0, bd);
            umma_commit_elect(a_empty + sa);
          }
          umma_commit_elect(acc_full + bi);        // depth tap 2 of this step completed slot s0 + 2
        }
      }
    }
  } else {
    // ---------------- epilogue: warps 2..5 take the even steps, 6..9 the odd ones; one voxel row per thread
    const int set = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int th = row >> 3, tw = row & 7;
    const bool vec_ok = p.ys[4] == 1 && (p.Cout & 3) == 0;
    constexpr int CW = BN < 32 ? BN : 32;
    uint32_t st = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int nt, n, h0, w0, z0, z1;
      decode(item, nt, n, h0, w0, z0, z1);
      const int n0 = nt * BN;
      const int nsteps = z1 - z0 + 2;
      const int oh = h0 + th, ow = w0 + tw;
      const bool inb = oh < p.H && ow < p.W;
      float* ybase = y + (long long)n * p.ys[0] + (long long)oh * p.ys[2] + (long long)ow * p.ys[3];
      for (int j = 0; j < nsteps; ++j, ++st) {
        if ((int)(st & 1) != set) continue;
        const uint32_t bi = st % NB;
        mbar_wait(acc_full + bi, (st / NB) & 1);
        tc_fence_after();
        if (j >= 2) {
          const int od = z0 + j - 2;
          const int slot = ((R - (int)(st % R)) % R + 2) % R;
          const uint32_t acc = tmem_base + (uint32_t)(slot * BN) + ((uint32_t)(quarter * 32) << 16);
          float* yp = ybase + (long long)od * p.ys[1];
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += CW) {
            float v[CW];
            if (CW == 32) tmem_ld_32x32(acc + (uint32_t)c0, v); else tmem_ld_32x16(acc + (uint32_t)c0, v);
            if (inb && n0 + c0 < p.Cout) {
#pragma unroll
              for (int i = 0; i < CW; ++i) {
                float t = v[i];
                if (bias && n0 + c0 + i < p.Cout) t += __ldg(bias + n0 + c0 + i);
                v[i] = act_apply(t, p.act);
              }
              if (vec_ok && n0 + c0 + CW <= p.Cout) {
                float4* dst = reinterpret_cast<float4*>(yp + n0 + c0);
#pragma unroll
                for (int i = 0; i < CW / 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              } else {
#pragma unroll
                for (int i = 0; i < CW; ++i)
                  if (n0 + c0 + i < p.Cout) yp[(long long)(n0 + c0 + i) * p.ys[4]] = v[i];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + bi);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512u);
}

// ---------------------------------------------------------------- CTA-pair variant (cta_group::2), 3x3, 256-channel tiles
// The single-CTA kernels above are bound by the shared-memory operand reads of the MMA (a TF32 128 x 128 x 8 MMA reads
// 8 KB for 64 cycles of math, a 128 x 256 x 8 one 12 KB for 128: 59 % / 75 % tensor-pipe activity measured).  Here
// two CTAs of a cluster (two SMs of a TPC) run ONE 256 x 256 x 8 MMA per K slice: each SM holds its own 128-voxel
// halo tile (the A half) and HALF of the weight tile (128 of the 256 output channels); the tensor cores of both SMs
// read A locally and B from both shared memories, so every SM reads 8 KB per 128 cycles of math and streams half
// the weight bytes from L2.  Structure as conv_umma_halo_kernel (16 x 8 voxel tile per CTA, TMEM double buffering):
//   * both CTAs issue their own TMA loads (cta_group::2 form: the transaction bytes are signalled on the LEADER's
//     barrier, whose expect_tx covers both halves);
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 and commits with a multicast arrive, which frees the
//     stage in both CTAs and hands the accumulators (each CTA's own 128 TMEM lanes) to both epilogues;
//   * the epilogue warps of both CTAs release the accumulator stage on the leader's barrier (remote arrive).
constexpr int PAIR_BST = 8;                // weight stages: this CTA's BN / 2 output channels x 32 tf32 (16 KB at BN = 256)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the same barrier in the leader CTA (rank 0) of the pair, as a shared::cluster address
__device__ __forceinline__ uint32_t leader_addr(const void* local) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(0));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
conv_umma_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const float* __restrict__ bias, float* __restrict__ y, const HaloP hp) {
  constexpr int AS = 2, BST = PAIR_BST;
  constexpr int PAIR_B_BYTES = (BN / 2) * 128;
  constexpr int HH = 18, HW = 10;                  // 16 x 8 voxel tile + halo of a 3 x 3 kernel
  const UmmaP& p = hp.u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                              // 2 activation stages
  uint8_t* sB = smem + 2 * hp.a_bytes;             // BST weight stages (this CTA's 128 output channels)
  uint64_t* a_full = (uint64_t*)(sB + BST * PAIR_B_BYTES);
  uint64_t* a_empty = a_full + 2;
  uint64_t* b_full = a_empty + 2;
  uint64_t* b_empty = b_full + BST;
  uint64_t* tmem_full = b_empty + BST;
  uint64_t* tmem_empty = tmem_full + AS;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + AS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const int tiles_hw = p.tiles_h * p.tiles_w;
  const int pairs = (p.ptiles + 1) / 2;            // ptiles = 16 x 8 voxel tiles; CTA `rank` of a pair takes tile 2 * pair + rank
  const int items = pairs * n_tiles;
  const int cchunks = (p.Cin + KCH - 1) / KCH;
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB);
    for (int s = 0; s < 2; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < BST; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
    for (int a = 0; a < AS; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer (both CTAs): own halo tile, own half of the weight tile
      uint32_t ia = 0;
      int sb = 0; uint32_t bph = 0;
      const uint32_t a_bytes_box = (uint32_t)(HH * HW * 128);
      for (int item = cluster_id; item < items; item += nclusters) {
        const int nt = item % n_tiles, pt = 2 * (item / n_tiles) + (int)rank;
        const int n = pt / tiles_hw; const int rem = pt - n * tiles_hw;      // pt >= ptiles: n >= N, the box is zero-filled
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        const int h0 = th_i * 16, w0 = tw_i * 8;
        for (int cc = 0; cc < cchunks; ++cc, ++ia) {
          const int sa = ia & 1;
          mbar_wait(a_empty + sa, ((ia >> 1) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(a_full + sa, 2 * a_bytes_box);
          tma_load_5d_pair(sA + sa * hp.a_bytes, &tmA, leader_addr(a_full + sa), cc * KCH, w0 - p.pad_w, h0 - p.pad_h, 0, n);
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(b_empty + sb, bph ^ 1);
            if (rank == 0) mbar_expect_tx(b_full + sb, 2 * PAIR_B_BYTES);
            tma_load_3d_pair(sB + sb * PAIR_B_BYTES, &tmB, leader_addr(b_full + sb), cc * KCH, nt * BN + (int)rank * (BN / 2), p.flip ? 8 - tap : tap);
            if (++sb == BST) { sb = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ---------------- MMA issuer (leader CTA only): M = 256 (128 voxels of each CTA), N = 256
      constexpr uint32_t idesc = instr_desc_tf32(256, BN);
      const uint64_t bdesc0 = smem_desc_sw128(smem_u32(sB), 16, 1024);
      int sb = 0; uint32_t bph = 0;
      uint32_t ia = 0, ti = 0;
      for (int item = cluster_id; item < items; item += nclusters, ++ti) {
        const uint32_t as = ti % AS, aph = (ti / AS) & 1;
        mbar_wait(tmem_empty + as, aph ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + as * BN;
        for (int cc = 0; cc < cchunks; ++cc, ++ia) {
          const int sa = ia & 1;
          mbar_wait(a_full + sa, (ia >> 1) & 1);
          tc_fence_after();
          const uint64_t adesc0 = smem_desc_sw128(smem_u32(sA + sa * hp.a_bytes), 16, HW * 128);
          static_for<0, 9>([&](auto tap_c) {
            constexpr int tap = decltype(tap_c)::value;
            constexpr int row = (tap / 3) * HW + tap % 3;
            mbar_wait(b_full + sb, bph);
            tc_fence_after();
            const uint64_t bdesc = bdesc0 + (uint64_t)(sb * (PAIR_B_BYTES / 16));
            const uint64_t adesc = adesc0 + (uint64_t)(8 * row);
#pragma unroll
            for (int k = 0; k < KCH / UMMA_K; ++k)
              umma_tf32_pair(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (tap | k) != 0 ? 1u : (uint32_t)(cc != 0));
            umma_commit_pair(b_empty + sb);
            if (++sb == BST) { sb = 0; bph ^= 1; }
          });
          umma_commit_pair(a_empty + sa);
        }
        umma_commit_pair(tmem_full + as);
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): warp e reads TMEM lanes 32 * (e % 4) .. + 31, columns (BN / 2) * (e / 4) .. + BN / 2 - 1
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int chalf = e >> 2;
    const int row = quarter * 32 + lane;
    const int th = row >> 3, tw = row & 7;
    const bool vec_ok = p.ys[4] == 1 && (p.Cout & 3) == 0;
    uint32_t ti = 0;
    for (int item = cluster_id; item < items; item += nclusters, ++ti) {
      const uint32_t as = ti % AS, aph = (ti / AS) & 1;
      const int nt = item % n_tiles, pt = 2 * (item / n_tiles) + (int)rank;
      const int n0 = nt * BN;
      const int n = pt / tiles_hw; const int rem = pt - n * tiles_hw;
      const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
      mbar_wait(tmem_full + as, aph);
      tc_fence_after();
      const int oh = th_i * 16 + th, ow = tw_i * 8 + tw;
      const bool valid = n < p.N && oh < p.H && ow < p.W;
      float* yp = y + (long long)n * p.ys[0] + (long long)oh * p.ys[2] + (long long)ow * p.ys[3];
      const uint32_t acc = tmem_base + as * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
        float v[32];
        tmem_ld_32x32(acc + (uint32_t)c0, v);
        if (valid && n0 + c0 < p.Cout) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float t = v[i];
            if (bias && n0 + c0 + i < p.Cout) t += __ldg(bias + n0 + c0 + i);
            v[i] = act_apply(t, p.act);
          }
          if (vec_ok && n0 + c0 + 32 <= p.Cout) {
            float4* dst = reinterpret_cast<float4*>(yp + n0 + c0);
            if (p.accum) {
              float4 o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = dst[i];
#pragma unroll
              for (int i = 0; i < 8; ++i) { v[4 * i] += o[i].x; v[4 * i + 1] += o[i].y; v[4 * i + 2] += o[i].z; v[4 * i + 3] += o[i].w; }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (n0 + c0 + i < p.Cout) {
                float* q = yp + (long long)(n0 + c0 + i) * p.ys[4];
                *q = p.accum ? *q + v[i] : v[i];
              }
          }
        }
        if (p.stat_rows && n < p.N)
          emit_stat_row(v, valid, lane, p.stat_rows + ((long long)pt * 4 + quarter) * p.Cout, n0 + c0, p.Cout);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_addr(tmem_empty + as));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                  // the leader's MMAs read the peer's shared memory until the end
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512u);
}

// ---------------------------------------------------------------- host side
// InstanceNorm-statistics by-product of a forward launch (dfmir_conv_umma_fwd_stats): `query` only reports how many
// statistic rows per image the kernel variant chosen for this shape writes (0: that variant has no such epilogue)
struct StatArg { float2* buf; bool query; int rows_per_image; };

// strides arrive as {n, spatial[nd], c}
struct Strides5 { long long n, d, h, w, c; };
Strides5 spread(const long long* s, int nd) {
  Strides5 r;
  r.n = s[0]; r.c = s[nd + 1];
  if (nd == 3) { r.d = s[1]; r.h = s[2]; r.w = s[3]; } else { r.h = s[1]; r.w = s[2]; r.d = 0; }
  return r;
}

bool operand_ok(const long long* s, int nd) {     // TMA source: unit channel stride, 16-byte multiples elsewhere
  if (s[nd + 1] != 1) return false;
  for (int i = 0; i <= nd; ++i) if (s[i] & 3) return false;
  return true;
}

int umma_shape_ok(const dfmir_conv_desc* d, int dgrad) {
  if (!d || (d->nd != 2 && d->nd != 3) || d->stride != 1) return 0;
  const int Cin = dgrad ? d->Cout : d->Cin, Cout = dgrad ? d->Cin : d->Cout;
  // K side: rows of the weight matrix must be 16-byte multiples (Cin % 4; K slices that hold only zero-filled
  // channels are not issued); N side: any count (padded to the tile by zero-filled weight rows)
  if (Cin % 4 || Cin < 4 || Cout < 1) return 0;
  const long long* is = dgrad ? d->y_strides : d->x_strides;
  if (!operand_ok(is, d->nd)) return 0;
  for (int a = 0; a < d->nd; ++a) {
    const int pd = dgrad ? d->kernel[a] - 1 - d->pad[a] : d->pad[a];
    if (pd < 0) return 0;
  }
  return 1;
}

int fill_umma(UmmaP& p, const dfmir_conv_desc* d, int dgrad, const char* who) {
  if (!umma_shape_ok(d, dgrad)) {
    dfmir_set_error("%s: the tensor-core path covers 2-D / 3-D stride-1 convolutions whose reduction-side channel count is a "
                    "multiple of 4, on channels-last operands with 16-byte aligned strides", who);
    return DFMIR_ERR_UNSUPPORTED;
  }
  const int nd = d->nd, sh = 3 - nd;
  p.N = d->N; p.Cin = dgrad ? d->Cout : d->Cin; p.Cout = dgrad ? d->Cin : d->Cout;
  int K[3] = {1, 1, 1}, P[3] = {0, 0, 0}, O[3] = {1, 1, 1};
  const int* osh = dgrad ? d->in_shape : d->out_shape;
  for (int a = 0; a < nd; ++a) {
    K[a + sh] = d->kernel[a];
    P[a + sh] = dgrad ? d->kernel[a] - 1 - d->pad[a] : d->pad[a];
    O[a + sh] = osh[a];
  }
  p.KD = K[0]; p.KH = K[1]; p.KW = K[2]; p.pad_d = P[0]; p.pad_h = P[1]; p.pad_w = P[2];
  p.D = O[0]; p.H = O[1]; p.W = O[2];
  p.flip = dgrad; p.act = dgrad ? DFMIR_ACT_NONE : d->act; p.per_sample = 0; p.accum = 0; p.stat_rows = nullptr;
  const Strides5 os = spread(dgrad ? d->x_strides : d->y_strides, nd);
  p.ys[0] = os.n; p.ys[1] = os.d; p.ys[2] = os.h; p.ys[3] = os.w; p.ys[4] = os.c;
  // tile box TD x TH x TW = 128 voxels (powers of two, TW >= 8): the shape that wastes the fewest voxels on
  // partial tiles (66x66 data-gradient outputs: 16x8 covers 76 % vs 52 % for 64x2); ties go to the widest
  // tile (longest contiguous runs per TMA box row)
  long long best = -1;
  for (int TW = 128; TW >= 8; TW >>= 1)
    for (int TH = BM / TW; TH >= 1; TH >>= 1) {
      const int TD = BM / (TW * TH);
      if (TD > 1 && nd == 2) continue;
      const long long vol = (long long)((p.W + TW - 1) / TW) * TW * ((p.H + TH - 1) / TH) * TH * ((p.D + TD - 1) / TD) * TD;
      if (best < 0 || vol < best) { best = vol; p.TW = TW; p.TH = TH; p.TD = TD; }
    }
  p.tiles_w = (p.W + p.TW - 1) / p.TW;
  p.tiles_h = (p.H + p.TH - 1) / p.TH;
  p.tiles_d = (p.D + p.TD - 1) / p.TD;
  p.ptiles = p.N * p.tiles_d * p.tiles_h * p.tiles_w;
  return DFMIR_OK;
}

template <int BN, int MT, int STAGES, int AS>
int launch_umma(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, float* y, const UmmaP& p, cudaStream_t st,
                const char* who) {
  using L = SmemLayout<BN, MT, STAGES, AS>;
  DFMIR_CUDA(cudaFuncSetAttribute(conv_umma_kernel<BN, MT, STAGES, AS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  const int items = ((p.ptiles + MT - 1) / MT) * ((p.Cout + BN - 1) / BN);
  if (items == 0) return DFMIR_OK;
  // persistent CTAs, one per SM; with fewer items than SMs every item gets its own CTA
  int grid = dfmir_num_sms();
  if (grid > items) grid = items;
  // balance: the same number of rounds on every CTA that runs (e.g. 256 items -> 128 CTAs x 2)
  const int rounds = (items + grid - 1) / grid;
  grid = (items + rounds - 1) / rounds;
  conv_umma_kernel<BN, MT, STAGES, AS><<<grid, THREADS, L::TOTAL, st>>>(tmA, tmB, bias, y, p);
  DFMIR_CHECK_LAUNCH(who);
  return DFMIR_OK;
}

template <int BN, int SD, int SH, int SW, int BST, int KDT>
int launch_halo_k(const float* act, const Strides5& as, int ID, int IH, int IW, const CUtensorMap& tmB, const float* bias, float* y,
                UmmaP p, cudaStream_t st, const char* who, StatArg* stat) {
  using C = HaloCfg<BN, SD, SH, SW, BST>;
  HaloP hp;
  hp.HD = SD + p.KD - 1; hp.HH = 16 * SH + p.KH - 1; hp.HW = 8 * SW + p.KW - 1;
  hp.a_bytes = (hp.HD * hp.HH * hp.HW * 128 + 1023) / 1024 * 1024;
  const int nbars = 4 + 2 * BST + 2 * C::AS;
  const size_t smem = 2 * (size_t)hp.a_bytes + (size_t)BST * C::B_BYTES + nbars * 8 + 16 + 1024;
  if (smem > 227 * 1024 || hp.HW > 256 || hp.HH > 256 || hp.HD > 256) return DFMIR_ERR_UNSUPPORTED;
  p.tiles_d = (p.D + SD - 1) / SD; p.tiles_h = (p.H + 16 * SH - 1) / (16 * SH); p.tiles_w = (p.W + 8 * SW - 1) / (8 * SW);
  p.ptiles = p.N * p.tiles_d * p.tiles_h * p.tiles_w;
  if (stat) {
    stat->rows_per_image = BN >= 32 ? p.tiles_d * p.tiles_h * p.tiles_w * C::MT * 4 : 0;
    if (stat->query) return DFMIR_OK;
    if (stat->buf && stat->rows_per_image == 0) { dfmir_set_error("%s: no statistics epilogue for 16-channel tiles", who); return DFMIR_ERR_UNSUPPORTED; }
    p.stat_rows = stat->buf;
  }
  hp.u = p;
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  CUtensorMap tmA;
  const long long sd = ID > 1 ? as.d : as.h * IH;
  cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)ID, (cuuint64_t)p.N};
  cuuint64_t strides[4] = {(cuuint64_t)as.w * 4, (cuuint64_t)as.h * 4, (cuuint64_t)sd * 4, (cuuint64_t)as.n * 4};
  cuuint32_t box[5] = {KCH, (cuuint32_t)hp.HW, (cuuint32_t)hp.HH, (cuuint32_t)hp.HD, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(halo tile) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  DFMIR_CUDA(cudaFuncSetAttribute(conv_umma_halo_kernel<BN, SD, SH, SW, BST, KDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = p.ptiles * ((p.Cout + BN - 1) / BN);
  if (items == 0) return DFMIR_OK;
  int grid = dfmir_num_sms();
  if (grid > items) grid = items;
  const int rounds = (items + grid - 1) / grid;
  grid = (items + rounds - 1) / rounds;
  conv_umma_halo_kernel<BN, SD, SH, SW, BST, KDT><<<grid, THREADS, smem, st>>>(tmA, tmB, bias, y, hp);
  DFMIR_CHECK_LAUNCH(who);
  return DFMIR_OK;
}

// 3 x 3 (x 3) kernels take the instantiation with compile-time taps; 2-D tile configurations have SD = 1
template <int BN, int SD, int SH, int SW, int BST>
int launch_halo(const float* act, const Strides5& as, int ID, int IH, int IW, const CUtensorMap& tmB, const float* bias, float* y,
                UmmaP p, cudaStream_t st, const char* who, StatArg* stat) {
  static const int fixed = getenv("DFMIR_UMMA_FIXED_TAPS") ? atoi(getenv("DFMIR_UMMA_FIXED_TAPS")) : 1;
  constexpr int KDT = SD > 1 ? 3 : 1;
  if (fixed && p.KH == 3 && p.KW == 3 && p.KD == KDT)
    return launch_halo_k<BN, SD, SH, SW, BST, KDT>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
  return launch_halo_k<BN, SD, SH, SW, BST, 0>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
}

// CTA-pair kernel: 2-D 3 x 3 convolutions with Cout a multiple of BN = 256 or 128 (ResnetBlock and down / up-sampling
// convs of the generator, forward and data gradient)
template <int BN>
int launch_pair(const float* act, const Strides5& as, int IH, int IW, const float* w, const float* bias, float* y, UmmaP p,
                cudaStream_t st, const char* who, StatArg* stat) {
  HaloP hp;
  hp.HD = 1; hp.HH = 18; hp.HW = 10;
  hp.a_bytes = (hp.HH * hp.HW * 128 + 1023) / 1024 * 1024;
  const int nbars = 4 + 2 * PAIR_BST + 4;
  constexpr int PAIR_B_BYTES = (BN / 2) * 128;
  const size_t smem = 2 * (size_t)hp.a_bytes + (size_t)PAIR_BST * PAIR_B_BYTES + nbars * 8 + 16 + 1024;
  p.tiles_d = 1; p.tiles_h = (p.H + 15) / 16; p.tiles_w = (p.W + 7) / 8;
  p.ptiles = p.N * p.tiles_h * p.tiles_w;
  if (stat) {
    stat->rows_per_image = p.tiles_h * p.tiles_w * 4;
    if (stat->query) return DFMIR_OK;
    p.stat_rows = stat->buf;
  }
  hp.u = p;
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)IW, (cuuint64_t)IH, 1, (cuuint64_t)p.N};
    cuuint64_t strides[4] = {(cuuint64_t)as.w * 4, (cuuint64_t)as.h * 4, (cuuint64_t)as.h * IH * 4, (cuuint64_t)as.n * 4};
    cuuint32_t box[5] = {KCH, (cuuint32_t)hp.HW, (cuuint32_t)hp.HH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(halo tile) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, 9};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Cin * p.Cout * 4};
    cuuint32_t box[3] = {KCH, BN / 2, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(weights) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  }
  DFMIR_CUDA(cudaFuncSetAttribute(conv_umma_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int items = ((p.ptiles + 1) / 2) * ((p.Cout + BN - 1) / BN);
  if (items == 0) return DFMIR_OK;
  int clusters = dfmir_num_sms() / 2;
  if (clusters > items) clusters = items;
  const int rounds = (items + clusters - 1) / clusters;
  clusters = (items + rounds - 1) / rounds;
  conv_umma_pair_kernel<BN><<<2 * clusters, THREADS, smem, st>>>(tmA, tmB, bias, y, hp);
  DFMIR_CHECK_LAUNCH(who);
  return DFMIR_OK;
}

// depth-march kernel: 3-D 3 x 3 x 3, output-channel tiles of 16 / 32 / 64
template <int BN, bool RES>
int launch_dmarch(const float* act, const Strides5& as, int ID, int IH, int IW, const float* w, const float* bias, float* y, UmmaP p,
                  cudaStream_t st, const char* who) {
  using C = DmCfg<BN, RES>;
  static_assert(C::SMEM <= 227 * 1024, "shared memory");
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  p.tiles_h = (p.H + 15) / 16; p.tiles_w = (p.W + 7) / 8; p.tiles_d = 1;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  const long long cols = (long long)p.N * p.tiles_h * p.tiles_w * n_tiles;
  // depth chunks per column: fewest slice steps on the critical path, counting the two extra input slices of every
  // chunk and the rounds of the persistent grid
  DmP dp;
  {
    const long long sms = dfmir_num_sms();
    long long best = -1; int nzc = 1;
    for (int c = 1; c <= 32 && p.D / c >= 4; ++c) {
      const long long rounds = (cols * c + sms - 1) / sms;
      const long long cost = rounds * ((p.D + c - 1) / c + 2);
      if (best < 0 || cost < best) { best = cost; nzc = c; }
    }
    dp.zchunk = (p.D + nzc - 1) / nzc;
    dp.nzc = (p.D + dp.zchunk - 1) / dp.zchunk;
  }
  p.ptiles = (int)(cols / n_tiles) * dp.nzc;
  dp.u = p;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)ID, (cuuint64_t)p.N};
    cuuint64_t strides[4] = {(cuuint64_t)as.w * 4, (cuuint64_t)as.h * 4, (cuuint64_t)as.d * 4, (cuuint64_t)as.n * 4};
    cuuint32_t box[5] = {KCH, (cuuint32_t)C::HW, (cuuint32_t)C::HH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(slice tile) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, 27};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Cin * p.Cout * 4};
    cuuint32_t box[3] = {KCH, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(weights) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  }
  DFMIR_CUDA(cudaFuncSetAttribute(conv_umma_dmarch_kernel<BN, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  const long long items = cols * dp.nzc;
  if (items == 0) return DFMIR_OK;
  int grid = dfmir_num_sms();
  if (grid > items) grid = (int)items;
  const int rounds = (int)((items + grid - 1) / grid);
  grid = (int)((items + rounds - 1) / rounds);
  conv_umma_dmarch_kernel<BN, RES><<<grid, THREADS, C::SMEM, st>>>(tmA, tmB, bias, y, dp);
  DFMIR_CHECK_LAUNCH(who);
  return DFMIR_OK;
}

// act: source activation (channels-last, c stride 1) with element strides `as` and spatial size (ID, IH, IW)
int run_umma(const float* act, const Strides5& as, int ID, int IH, int IW, const float* w, const float* bias, float* y,
             const UmmaP& p, cudaStream_t st, const char* who, StatArg* stat = nullptr) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { dfmir_set_error("%s: cuTensorMapEncodeTiled not available from the driver", who); return DFMIR_ERR_CUDA; }
  if (((uintptr_t)act & 15) || ((uintptr_t)w & 15)) { dfmir_set_error("%s: TMA needs 16-byte aligned base pointers", who); return DFMIR_ERR_ARG; }
  int BN = p.Cout <= 16 ? 16 : (p.Cout <= 32 ? 32 : (p.Cout <= 64 ? 64 : 128));
  // 16x16-voxel CTA tiles waste a third of the work on a 66x66 output (the data gradient of the ResnetBlock
  // convs): there, 16x8 tiles with all 256 channels per CTA measured 542 vs 441 TFLOP/s (batch 32)
  const double eff16 = (double)p.H * p.W / ((double)((p.H + 15) / 16 * 16) * ((p.W + 15) / 16 * 16));
  const double eff8 = (double)p.H * p.W / ((double)((p.H + 15) / 16 * 16) * ((p.W + 7) / 8 * 8));
  const bool narrow256 = p.Cout == 256 && ID == 1 && p.KD * p.KH * p.KW > 1 && eff16 < 0.72 && eff8 > eff16 * 1.08;
  if (narrow256) BN = 256;
  static const int pair = getenv("DFMIR_UMMA_PAIR") ? atoi(getenv("DFMIR_UMMA_PAIR")) : 1;
  if (p.accum && !(pair && ID == 1 && p.KD == 1 && p.KH == 3 && p.KW == 3 && p.Cin % KCH == 0 && p.ys[4] == 1 && p.Cout % 256 == 0)) {
    dfmir_set_error("%s: accumulation into the output is implemented by the CTA-pair kernel (2-D 3x3, 256-channel tiles)", who);
    return DFMIR_ERR_UNSUPPORTED;
  }
  if (pair && ID == 1 && p.KD == 1 && p.KH == 3 && p.KW == 3 && p.Cin % KCH == 0 && !p.per_sample && p.ys[4] == 1) {
    if (p.Cout % 256 == 0) return launch_pair<256>(act, as, IH, IW, w, bias, y, p, st, who, stat);
    // DFMIR_UMMA_PAIR=2 also pairs the 128-channel layers: measured slower than the single-CTA 2 x 128 tiles (5.7 vs 5.2 ms / step)
    if (p.Cout % 128 == 0 && pair > 1) return launch_pair<128>(act, as, IH, IW, w, bias, y, p, st, who, stat);
  }
  static const int dmarch = getenv("DFMIR_UMMA_DMARCH") ? atoi(getenv("DFMIR_UMMA_DMARCH")) : 1;
  if (dmarch && ID > 1 && p.KD == 3 && p.KH == 3 && p.KW == 3 && BN <= 64 && !p.per_sample && !p.accum && !stat && p.D >= 8 &&
      (long long)p.N * p.D * p.H * p.W >= 8192) {
    static const int res = getenv("DFMIR_UMMA_DMARCH_RES") ? atoi(getenv("DFMIR_UMMA_DMARCH_RES")) : 1;
    const int cch = (p.Cin + KCH - 1) / KCH;
    // 33..48 output channels (the data gradient of the 36-channel concat layer) with one input chunk: 48-column tiles keep
    // the whole weight tensor resident (as 64-column tiles they streamed 216 KB of weights per input slice: 46 % L2 throughput)
    if (BN == 64 && p.Cout <= 48 && res && cch == 1) return launch_dmarch<48, true>(act, as, ID, IH, IW, w, bias, y, p, st, who);
    if (BN == 64) return launch_dmarch<64, false>(act, as, ID, IH, IW, w, bias, y, p, st, who);
    if (BN == 32) return res && cch <= DmCfg<32, true>::RES_CHUNKS ? launch_dmarch<32, true>(act, as, ID, IH, IW, w, bias, y, p, st, who)
                                                                   : launch_dmarch<32, false>(act, as, ID, IH, IW, w, bias, y, p, st, who);
    return res && cch <= DmCfg<16, true>::RES_CHUNKS ? launch_dmarch<16, true>(act, as, ID, IH, IW, w, bias, y, p, st, who)
                                                     : launch_dmarch<16, false>(act, as, ID, IH, IW, w, bias, y, p, st, who);
  }
  CUtensorMap tmA, tmB;
  {
    const long long sd = ID > 1 ? as.d : as.h * IH;      // 2-D: a depth axis of extent 1 (its stride is never used)
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)ID, (cuuint64_t)p.N};
    cuuint64_t strides[4] = {(cuuint64_t)as.w * 4, (cuuint64_t)as.h * 4, (cuuint64_t)sd * 4, (cuuint64_t)as.n * 4};
    cuuint32_t box[5] = {KCH, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TD, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(activation) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  }
  {
    const int taps = p.per_sample ? p.N : p.KD * p.KH * p.KW;
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Cin * p.Cout * 4};
    cuuint32_t box[3] = {KCH, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(weights) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  }
  // halo variant (one activation tile per chunk serves every tap): kernels larger than 1x1
  static const int halo = getenv("DFMIR_UMMA_HALO") ? atoi(getenv("DFMIR_UMMA_HALO")) : 1;
  if (halo && narrow256) {
    int rc = launch_halo<256, 1, 1, 1, 4>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
    if (rc != DFMIR_ERR_UNSUPPORTED) return rc;
  }
  if (halo && p.KD * p.KH * p.KW > 1 && BN <= 128) {
    int rc = DFMIR_ERR_UNSUPPORTED;
    if (ID > 1) {
      if (BN == 128) rc = launch_halo<128, 2, 1, 1, 2>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
      else if (BN == 64) rc = launch_halo<64, 2, 1, 1, 4>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
      else if (BN == 32) rc = launch_halo<32, 2, 1, 1, 6>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
      else rc = launch_halo<16, 2, 1, 1, 8>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
    } else {
      if (BN == 128) rc = launch_halo<128, 1, 1, 2, 6>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
      else if (BN == 64) rc = launch_halo<64, 1, 2, 2, 4>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
      else if (BN == 32) rc = launch_halo<32, 1, 2, 2, 6>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
      else rc = launch_halo<16, 1, 2, 2, 8>(act, as, ID, IH, IW, tmB, bias, y, p, st, who, stat);
    }
    if (rc != DFMIR_ERR_UNSUPPORTED) return rc;
  }
  if (stat) {         // the plain variant has no statistics epilogue
    stat->rows_per_image = 0;
    if (stat->query) return DFMIR_OK;
    if (stat->buf) { dfmir_set_error("%s: no statistics epilogue for this shape", who); return DFMIR_ERR_UNSUPPORTED; }
  }
  // plain variant (1x1 kernels, DFMIR_UMMA_HALO=0).  128-channel tiles x 2 sub-tiles: 48 KB stages x 4 and
  // double-buffered accumulators beat one 256-wide tile (64 KB x 3, no epilogue overlap): 552 vs 430 TFLOP/s on the
  // ResnetBlock conv at batch 32
  if (BN == 256) return launch_umma<256, 2, 3, 1>(tmA, tmB, bias, y, p, st, who);
  if (BN == 128) return launch_umma<128, 2, 4, 2>(tmA, tmB, bias, y, p, st, who);
  if (BN == 64) return launch_umma<64, 4, 3, 2>(tmA, tmB, bias, y, p, st, who);
  if (BN == 32) return launch_umma<32, 4, 3, 2>(tmA, tmB, bias, y, p, st, who);
  return launch_umma<16, 4, 3, 2>(tmA, tmB, bias, y, p, st, who);
}

}  // namespace

extern "C" int dfmir_conv_umma_supported(const dfmir_conv_desc* d, int dgrad) { return umma_shape_ok(d, dgrad); }

// Forward on the tensor cores.  w: [tap][Cout][Cin] (Cin contiguous).
extern "C" int dfmir_conv_umma_fwd(const float* x, const float* w, const float* bias, float* y,
                                   const dfmir_conv_desc* d, void* stream) {
  UmmaP p;
  int rc = fill_umma(p, d, 0, "dfmir_conv_umma_fwd");
  if (rc) return rc;
  DFMIR_CHECK_ARG(x && w && y, "dfmir_conv_umma_fwd: null pointer");
  const int nd = d->nd;
  return run_umma(x, spread(d->x_strides, nd), nd == 3 ? d->in_shape[0] : 1, d->in_shape[nd - 2], d->in_shape[nd - 1], w, bias, y, p,
                  (cudaStream_t)stream, "dfmir_conv_umma_fwd");
}

// Forward + the InstanceNorm statistics of its result as a by-product of the epilogue (models/networks.py:984,996,1020,
// 1201-1215: every generator convolution is followed by InstanceNorm2d): while a 32-voxel x 32-channel block of the result
// is in registers, its per-channel sum and sum of squares go to stat_rows[(image-major row)][Cout] as float2, which
// dfmir_instnorm_fwd_rows reduces instead of a pass over y.  dfmir_conv_umma_stat_rows: rows per image the kernel
// chosen for this shape writes, 0 if it has no such epilogue (then call dfmir_conv_umma_fwd and dfmir_instnorm_fwd).
extern "C" int dfmir_conv_umma_stat_rows(const dfmir_conv_desc* d) {
  UmmaP p;
  if (!umma_shape_ok(d, 0) || fill_umma(p, d, 0, "dfmir_conv_umma_stat_rows")) return 0;
  const int nd = d->nd;
  if (d->y_strides[nd + 1] != 1 || d->Cout % 4) return 0;
  StatArg stat{nullptr, true, 0};
  // 16-byte aligned dummy pointers: only the tile geometry is evaluated
  if (run_umma((const float*)16, spread(d->x_strides, nd), nd == 3 ? d->in_shape[0] : 1, d->in_shape[nd - 2], d->in_shape[nd - 1],
               (const float*)16, nullptr, (float*)16, p, nullptr, "dfmir_conv_umma_stat_rows", &stat)) return 0;
  return stat.rows_per_image;
}

extern "C" int dfmir_conv_umma_fwd_stats(const float* x, const float* w, const float* bias, float* y, const dfmir_conv_desc* d,
                                         float* stat_rows, void* stream) {
  UmmaP p;
  int rc = fill_umma(p, d, 0, "dfmir_conv_umma_fwd_stats");
  if (rc) return rc;
  DFMIR_CHECK_ARG(x && w && y && stat_rows, "dfmir_conv_umma_fwd_stats: null pointer");
  DFMIR_CHECK_ARG(((uintptr_t)stat_rows & 7) == 0, "dfmir_conv_umma_fwd_stats: stat_rows must be 8-byte aligned");
  const int nd = d->nd;
  StatArg stat{(float2*)stat_rows, false, 0};
  return run_umma(x, spread(d->x_strides, nd), nd == 3 ? d->in_shape[0] : 1, d->in_shape[nd - 2], d->in_shape[nd - 1], w, bias, y, p,
                  (cudaStream_t)stream, "dfmir_conv_umma_fwd_stats", &stat);
}

// Data gradient on the tensor cores: dx = conv(dy, flipped taps, pad' = k-1-pad).
// w: the FORWARD layout of the fp32 path, [tap][Cin][Cout] (Cout contiguous): K-major for this product.
extern "C" int dfmir_conv_umma_dgrad(const float* dy, const float* w, float* dx, const dfmir_conv_desc* d,
                                     void* stream) {
  UmmaP p;
  int rc = fill_umma(p, d, 1, "dfmir_conv_umma_dgrad");
  if (rc) return rc;
  DFMIR_CHECK_ARG(dy && w && dx, "dfmir_conv_umma_dgrad: null pointer");
  const int nd = d->nd;
  return run_umma(dy, spread(d->y_strides, nd), nd == 3 ? d->out_shape[0] : 1, d->out_shape[nd - 2], d->out_shape[nd - 1], w, nullptr,
                  dx, p, (cudaStream_t)stream, "dfmir_conv_umma_dgrad");
}

// Same product ADDED to dx (dx += conv(dy, flipped taps)): the ResnetBlock input receives the residual branch's
// gradient first (dfmir_instnorm_bwd writes it, models/networks.py:1218-1221 out = x + conv_block(x)) and the
// convolution branch's data gradient lands on top of it, instead of a separate add over both tensors.
// DFMIR_ERR_UNSUPPORTED (nothing launched) unless the CTA-pair kernel covers the shape; dfmir_conv_umma_dgrad_acc_supported tells.
extern "C" int dfmir_conv_umma_dgrad_acc_supported(const dfmir_conv_desc* d) {
  static const int pair = getenv("DFMIR_UMMA_PAIR") ? atoi(getenv("DFMIR_UMMA_PAIR")) : 1;
  if (!pair || !umma_shape_ok(d, 1) || d->nd != 2 || d->kernel[0] != 3 || d->kernel[1] != 3) return 0;
  return d->Cin % 256 == 0 && d->Cout % KCH == 0 && d->x_strides[3] == 1;
}

extern "C" int dfmir_conv_umma_dgrad_acc(const float* dy, const float* w, float* dx, const dfmir_conv_desc* d, void* stream) {
  UmmaP p;
  int rc = fill_umma(p, d, 1, "dfmir_conv_umma_dgrad_acc");
  if (rc) return rc;
  DFMIR_CHECK_ARG(dy && w && dx, "dfmir_conv_umma_dgrad_acc: null pointer");
  p.accum = 1;
  return run_umma(dy, spread(d->y_strides, 2), 1, d->out_shape[0], d->out_shape[1], w, nullptr, dx, p, (cudaStream_t)stream,
                  "dfmir_conv_umma_dgrad_acc");
}

// Batched C[b] (M x N) = A[b] (M x K) * B[b]^T (N x K), all row-major and dense, on the same tcgen05 kernel:
// the M rows of a sample are the "pixels", K the input channels, and the weight map's tap coordinate selects
// the sample.  Used for the PatchNCE logits S = Q K^T (models/patchnce.py:20,42) and their backward product.
// TF32 operands: callers that need fp32-class accuracy pass the 3-way split operands of dfmir_tf32_split3.
extern "C" int dfmir_bmm_nt_umma(const float* A, const float* B, float* C, int batch, int M, int N, int K, void* stream) {
  const char* who = "dfmir_bmm_nt_umma";
  DFMIR_CHECK_ARG(A && B && C, "%s: null pointer", who);
  DFMIR_CHECK_ARG(batch >= 1 && M >= 128 && M % 128 == 0 && N >= 1 && K >= 16 && K % 4 == 0,
                  "%s: needs M %% 128 == 0, K %% 4 == 0, K >= 16 (got batch=%d M=%d N=%d K=%d)", who, batch, M, N, K);
  UmmaP p{};
  p.N = batch; p.D = 1; p.H = 1; p.W = M; p.Cin = K; p.Cout = N;
  p.KD = p.KH = p.KW = 1; p.pad_d = p.pad_h = p.pad_w = 0;
  p.TD = 1; p.TH = 1; p.TW = 128; p.tiles_d = 1; p.tiles_h = 1; p.tiles_w = M / 128;
  p.ptiles = batch * p.tiles_w; p.flip = 0; p.act = DFMIR_ACT_NONE; p.per_sample = 1; p.accum = 0; p.stat_rows = nullptr;
  p.ys[0] = (long long)M * N; p.ys[1] = 0; p.ys[2] = 0; p.ys[3] = N; p.ys[4] = 1;
  const int mt = N <= 64 ? 4 : 2;      // sub-tiles per work item of the tile configuration run_umma picks
  DFMIR_CHECK_ARG(p.tiles_w % mt == 0, "%s: M = %d must be a multiple of %d so that a work item stays inside one sample", who, M, 128 * mt);
  Strides5 as; as.n = (long long)M * K; as.d = 0; as.h = (long long)M * K; as.w = K; as.c = 1;
  return run_umma(A, as, 1, 1, M, B, nullptr, C, p, (cudaStream_t)stream, who);
}

namespace {
// out (rows, 3*D): a-style [hi | hi | lo], b-style [hi | lo | hi]; hi = x truncated to TF32 (what the tensor core
// keeps), lo = x - hi (exact).  sum over the 3D columns of a-style * b-style = hi*hi + hi*lo + lo*hi.
__global__ void __launch_bounds__(256)
tf32_split3_kernel(const float* __restrict__ x, float* __restrict__ out, long long total, int D, int b_style) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D; const int c = (int)(i - r * D);
    const float v = x[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const float lo = v - hi;
    float* o = out + r * 3 * D + c;
    o[0] = hi; o[D] = b_style ? lo : hi; o[2 * D] = b_style ? hi : lo;
  }
}
// the same split stacked along the rows, for products that reduce over the rows (weight gradients): out (3, rows, D),
// a-style [hi ; hi ; lo], b-style [hi ; lo ; hi]
__global__ void __launch_bounds__(256)
tf32_split3_rows_kernel(const float* __restrict__ x, float* __restrict__ out, long long total, int b_style) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const float lo = v - hi;
    out[i] = hi; out[total + i] = b_style ? lo : hi; out[2 * total + i] = b_style ? hi : lo;
  }
}
}  // namespace

extern "C" int dfmir_tf32_split3_rows(const float* x, float* out, long long rows, int D, int b_style, void* stream) {
  DFMIR_CHECK_ARG(x && out && rows >= 0 && D > 0, "dfmir_tf32_split3_rows: bad argument");
  const long long total = rows * D;
  if (total == 0) return DFMIR_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  tf32_split3_rows_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(x, out, total, b_style);
  DFMIR_CHECK_LAUNCH("dfmir_tf32_split3_rows");
  return DFMIR_OK;
}

extern "C" int dfmir_tf32_split3(const float* x, float* out, long long rows, int D, int b_style, void* stream) {
  DFMIR_CHECK_ARG(x && out && rows >= 0 && D > 0, "dfmir_tf32_split3: bad argument");
  const long long total = rows * D;
  if (total == 0) return DFMIR_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)dfmir_num_sms() * 16;
  tf32_split3_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(x, out, total, D, b_style);
  DFMIR_CHECK_LAUNCH("dfmir_tf32_split3");
  return DFMIR_OK;
}
