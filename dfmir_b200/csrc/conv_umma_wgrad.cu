// K1 / K2 weight gradient on the 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in
// TMEM, operands staged by TMA), for the stride-1 convolutions of ResnetGenerator (models/networks.py:995,
// 1016,1201,1214) and of VoxelMorph's U-Net in 2-D and 3-D (vxm networks.py:1515) — the backward of the
// cuDNN wgrad engines.  Channel counts that do not fill a tile are padded by the TMA unit's zero fill
// (M side to 128 channels, N side to 32 / 64 / 128 / 256).
//
//   dW[tap][ci][co] = sum over output pixels p of  x[p + tap - pad][ci] * dy[p][co]
//
// GEMM view per tap: D[M][N] = A[M][K] * B[N][K]^T with K = output pixels.  Both operands are
// channels-last activations, i.e. the channel axis (M or N) is the contiguous one: they are fed to
// the tensor core as MN-major operands.  One TMA box {32 channels, TW, TH, 1} lands 32 pixels as 32
// rows of 128 bytes in the one layout tcgen05 accepts for MN-major 32-bit operands
// (SWIZZLE_128B_BASE32B = TMA's SWIZZLE_128B_ATOM_32B): 4-pixel groups 512 bytes apart (SBO),
// 32-channel column groups PIX*128 bytes apart (LBO).  The x box is shifted by the
// tap; pixels outside the image are zero-filled by the TMA unit (zero padding), and partial tiles
// contribute zero because the dy box is zero-filled there.
//
// Work split: grid = (taps, M-tiles * N-tiles, S).  A CTA owns one 128 x BN tile of dW for one tap
// and walks the pixel chunks s, s+S, s+2S, ... (so the CTAs that run concurrently read neighbouring
// chunks and share them through L2: the 9 taps and all channel tiles of a chunk hit the same lines),
// keeps the accumulator in TMEM for its whole pixel range, and finishes with one red.global.add pass
// into dW (S-way split-K; dW is zero-filled by the caller as for the fp32 path).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 =
// epilogue (TMEM -> registers -> red.global.add.f32).
#include "umma.cuh"
#include "dfmir_b200.h"

namespace {
using namespace umma;

constexpr int PIX = 32;                 // voxels (K) per pipeline stage
constexpr int CHUNK_BYTES = PIX * 128;  // one 32-channel column group of a stage
constexpr int UMMA_K = 8;               // tf32

struct WgradP {
  int N, D, H, W;           // dy: samples and spatial size (output of the forward conv); 2-D: D = 1
  int Cin, Cout;
  int KH, KW, pad_d, pad_h, pad_w;
  int TD, TH, TW, tiles_d, tiles_h, tiles_w;   // voxel chunk box (TD*TH*TW = 32) and chunks per axis
  int nchunks;
  int x_is_m;               // 1: M = input channels (x), N = output channels (dy); 0: swapped
  int m_tiles, n_tiles;
};

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int BN, int STAGES>
struct WgLayout {
  static constexpr int A_BYTES = 4 * CHUNK_BYTES;          // 128 channels
  static constexpr int B_BYTES = (BN / 32) * CHUNK_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192)
conv_wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                       float* __restrict__ dw, const WgradP p) {
  using L = WgLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = blockIdx.x;
  const int mt = blockIdx.y / p.n_tiles, nt = blockIdx.y - mt * p.n_tiles;
  const int split = blockIdx.z, S = gridDim.z;
  const int q = tap % p.KW; const int t2 = tap / p.KW;
  const int r = t2 % p.KH, kd = t2 / p.KH;
  const int dd = kd - p.pad_d, dh = r - p.pad_h, dwv = q - p.pad_w;      // x voxel = dy voxel + (dd, dh, dw)
  const int m0 = mt * 128, n0 = nt * BN;
  const int iters = (p.nchunks - split + S - 1) / S;   // host guarantees S <= nchunks
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmG);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mapM = p.x_is_m ? &tmX : &tmG;
      const CUtensorMap* mapN = p.x_is_m ? &tmG : &tmX;
      const int mdd = p.x_is_m ? dd : 0, mdh = p.x_is_m ? dh : 0, mdw = p.x_is_m ? dwv : 0;
      const int ndd = p.x_is_m ? 0 : dd, ndh = p.x_is_m ? 0 : dh, ndw = p.x_is_m ? 0 : dwv;
      const int tiles_hw = p.tiles_h * p.tiles_w;
      const int tiles_img = p.tiles_d * tiles_hw;
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const int c = split + it * S;
        const int n = c / tiles_img; int rem = c - n * tiles_img;
        const int td_i = rem / tiles_hw; rem -= td_i * tiles_hw;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        const int d0 = td_i * p.TD, h0 = th_i * p.TH, w0 = tw_i * p.TW;
        mbar_wait(empty + s, ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_expect_tx(full + s, (uint32_t)L::STAGE_BYTES);
        // channel groups beyond the tensor's channel count are zero-filled by the TMA unit (M padded to 128)
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_5d(sa + g * CHUNK_BYTES, mapM, full + s, m0 + g * 32, w0 + mdw, h0 + mdh, d0 + mdd, n);
#pragma unroll
        for (int g = 0; g < BN / 32; ++g) tma_load_5d(sb + g * CHUNK_BYTES, mapN, full + s, n0 + g * 32, w0 + ndw, h0 + ndh, d0 + ndd, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_tf32(128, BN, 1, 1);   // both operands MN-major
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t adesc = smem_desc(sa, CHUNK_BYTES, 512, LAYOUT_SW128_BASE32B);
        const uint64_t bdesc = smem_desc(sa + L::A_BYTES, CHUNK_BYTES, 512, LAYOUT_SW128_BASE32B);
#pragma unroll
        for (int k = 0; k < PIX / UMMA_K; ++k) {
          // next 8 voxels: +1024 bytes = +64 in 16-byte address units
          umma_tf32(tmem_base, adesc + (uint64_t)(64 * k), bdesc + (uint64_t)(64 * k), idesc, (it | k) != 0);
        }
        umma_commit(empty + s);
      }
      umma_commit(tmem_full);
    }
  } else {
    // epilogue: thread = accumulator row (M index); 32 consecutive N columns per TMEM load
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int CM = p.x_is_m ? p.Cin : p.Cout, CN = p.x_is_m ? p.Cout : p.Cin;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* dwt = dw + (long long)tap * p.Cin * p.Cout;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      if (m0 + row >= CM || n0 + c0 >= CN) continue;
      if (p.x_is_m) {
        float* dst = dwt + (long long)(m0 + row) * p.Cout + n0 + c0;     // row = ci, columns = co (contiguous)
        if (n0 + c0 + 32 <= CN && (p.Cout & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) red_add_v4(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < CN) atomicAdd(dst + j, v[j]);
        }
      } else {
        float* dst = dwt + (long long)(n0 + c0) * p.Cout + m0 + row;     // row = co (coalesced over lanes), columns = ci
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n0 + c0 + j < CN) atomicAdd(dst + (long long)j * p.Cout, v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------- CTA-pair variant (cta_group::2) of the per-tap kernel
// The 256 x 256-channel weight gradient above is bound by L2 -> shared-memory traffic: each of the two M tiles of
// a tap re-reads the whole dy chunk (48 KB per 512 cycles of MMA per SM, 12.5 TB/s over the chip).  Here the two
// M tiles of a tap form a CTA pair that runs ONE 256 x 256 x 8 MMA per K slice: CTA r stages x channels
// [128 r, 128 r + 128) (its half of M) and dy channels [128 r, 128 r + 128) (its half of N), 32 KB per stage
// instead of 48, and the accumulator rows of its x channels stay in its own TMEM.  Barrier protocol as in
// conv_umma_pair_kernel (conv_umma.cu): TMA bytes of both CTAs are counted on the leader's barrier, the leader
// issues the MMAs and commits with a multicast arrive.  grid = (2 * taps, M pairs, S), cluster (2, 1, 1).
constexpr int WP_STAGES = 6;
constexpr int WP_STAGE_BYTES = 8 * CHUNK_BYTES;     // 4 x-channel groups + 4 dy-channel groups of 32

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t leader_addr(const void* local) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(0));
  return r;
}
__device__ __forceinline__ void tma_load_5d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192)
conv_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                       float* __restrict__ dw, const WgradP p) {
  constexpr int STAGES = WP_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + STAGES * WP_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tap = blockIdx.x >> 1;
  const int split = blockIdx.z, S = gridDim.z;
  const int q = tap % p.KW; const int t2 = tap / p.KW;
  const int r = t2 % p.KH, kd = t2 / p.KH;
  const int dd = kd - p.pad_d, dh = r - p.pad_h, dwv = q - p.pad_w;
  const int m0 = blockIdx.y * 256 + (int)rank * 128;          // this CTA's x channels (rows of its accumulator)
  const int nh = (int)rank * 128;                             // this CTA's half of the 256 dy channels
  const int iters = (p.nchunks - split + S - 1) / S;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmG);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int tiles_hw = p.tiles_h * p.tiles_w;
      const int tiles_img = p.tiles_d * tiles_hw;
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const int c = split + it * S;
        const int n = c / tiles_img; int rem = c - n * tiles_img;
        const int td_i = rem / tiles_hw; rem -= td_i * tiles_hw;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        const int d0 = td_i * p.TD, h0 = th_i * p.TH, w0 = tw_i * p.TW;
        mbar_wait(empty + s, ph ^ 1);
        uint8_t* sa = smem + s * WP_STAGE_BYTES;
        uint8_t* sb = sa + 4 * CHUNK_BYTES;
        if (rank == 0) mbar_expect_tx(full + s, 2u * WP_STAGE_BYTES);
        const uint32_t bar = leader_addr(full + s);
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_5d_pair(sa + g * CHUNK_BYTES, &tmX, bar, m0 + g * 32, w0 + dwv, h0 + dh, d0 + dd, n);
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_5d_pair(sb + g * CHUNK_BYTES, &tmG, bar, nh + g * 32, w0, h0, d0, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = instr_desc_tf32(256, 256, 1, 1);   // both operands MN-major
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * WP_STAGE_BYTES);
        const uint64_t adesc = smem_desc(sa, CHUNK_BYTES, 512, LAYOUT_SW128_BASE32B);
        const uint64_t bdesc = smem_desc(sa + 4 * CHUNK_BYTES, CHUNK_BYTES, 512, LAYOUT_SW128_BASE32B);
#pragma unroll
        for (int k = 0; k < PIX / UMMA_K; ++k)
          umma_tf32_pair(tmem_base, adesc + (uint64_t)(64 * k), bdesc + (uint64_t)(64 * k), idesc, (uint32_t)((it | k) != 0));
        umma_commit_pair(empty + s);
      }
      umma_commit_pair(tmem_full);
    }
  } else {
    // epilogue (both CTAs): thread = accumulator row = x channel m0 + row; 256 dy-channel columns
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* dst = dw + ((long long)tap * p.Cin + m0 + row) * p.Cout;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) red_add_v4(dst + c0 + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// ---------------------------------------------------------------- halo variant (few-channel 3x3 / 3x3x3 layers)
// The per-tap kernel above re-reads x and dy once per tap (27 times in 3-D) and pads the M side to 128
// channels: right for the 256-channel ResnetBlock convs, 5 ms per launch for VoxelMorph-3D's full-resolution
// 36 -> 32 layer.  Here a CTA owns ALL taps of a 32-input-channel chunk and walks tiles of TH x 32 output
// voxels:
//   * x arrives ONCE per tile as a box with its halo, {32 ch, 35, TH + 2, KD, 1} (zero-filled outside the volume =
//     the convolution's zero padding), dy as {32 ch, 32, TH, 1, 1};
//   * an MN-major operand may start at any 128-byte row (= voxel) of the TMA-written tile and its four
//     32-channel column groups may be ONE row apart (leading byte offset 128; tools/umma_mn_shift_probe.cu), so
//     the M = 128 operand of one MMA is [kw = 0..3] x [32 input channels]: three taps along w per instruction
//     (the fourth column group is computed and dropped) instead of one tap with 96 zero-padded rows;
//   * K = 8 consecutive voxels along w; an accumulator per (kd, kh): KD*3 accumulators of BN columns in TMEM,
//     resident over the CTA's whole tile range, then one red.global.add pass (split-K over CTAs, dW zero-filled
//     by the caller as for the other kernels).
// grid = (S, input-channel chunks, output-channel tiles); 192 threads: warp 0 TMA, warp 1 MMA, warps 2..5 epilogue.
template <int BN, int KD>
struct WgHalo {
  static constexpr int TW = 32, TH = 4, KH = 3, KW = 3;
  static constexpr int HW = TW + 3, HH = TH + KH - 1, HD = KD;
  static constexpr int ROWS = HW * HH * HD;
  static constexpr int X_BYTES = (ROWS * 128 + 1023) / 1024 * 1024;
  static constexpr int GA = (BN + 31) / 32;                 // 32-channel column groups of the dy operand
  static constexpr int GA_BYTES = TH * TW * 128;
  static constexpr int G_BYTES = GA * GA_BYTES;
  static constexpr int STAGE_BYTES = X_BYTES + G_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES >= 4 ? 4 : (200 * 1024) / STAGE_BYTES;
  static_assert(STAGES >= 2, "two stages must fit");
  static constexpr int ACCS = KD * KH;
  static constexpr int COLS = ACCS * BN;
  static constexpr uint32_t TMEM_COLS = COLS <= 32 ? 32 : (COLS <= 64 ? 64 : (COLS <= 128 ? 128 : (COLS <= 256 ? 256 : 512)));
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
  static_assert(COLS <= 512, "TMEM has 512 columns");
};

struct WgHaloP {
  int N, D, H, W, Cin, Cout;
  int pad_d, pad_h, pad_w;
  int tiles_h, tiles_w, ntiles;
};

template <int BN, int KD>
__global__ void __launch_bounds__(192)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                       float* __restrict__ dw, const WgHaloP p) {
  using L = WgHalo<BN, KD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* empty = full + L::STAGES;
  uint64_t* tmem_full = empty + L::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, S = gridDim.x;
  const int c0 = blockIdx.y * 32, n0 = blockIdx.z * BN;
  const int iters = (p.ntiles - split + S - 1) / S;      // host guarantees S <= ntiles

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmG);
    for (int s = 0; s < L::STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int tiles_hw = p.tiles_h * p.tiles_w;
      const int tiles_img = p.D * tiles_hw;
      for (int it = 0; it < iters; ++it) {
        const int s = it % L::STAGES;
        const uint32_t ph = (it / L::STAGES) & 1;
        const int t = split + it * S;
        const int n = t / tiles_img; int rem = t - n * tiles_img;
        const int d0 = rem / tiles_hw; rem -= d0 * tiles_hw;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        const int h0 = th_i * L::TH, w0 = tw_i * L::TW;
        mbar_wait(empty + s, ph ^ 1);
        uint8_t* sx = smem + s * L::STAGE_BYTES;
        mbar_expect_tx(full + s, (uint32_t)(L::ROWS * 128 + L::G_BYTES));
        tma_load_5d(sx, &tmX, full + s, c0, w0 - p.pad_w, h0 - p.pad_h, d0 - p.pad_d, n);
#pragma unroll
        for (int g = 0; g < L::GA; ++g) tma_load_5d(sx + L::X_BYTES + g * L::GA_BYTES, &tmG, full + s, n0 + 32 * g, w0, h0, d0, n);
      }
    }
  } else if (warp == 1) {
    {     // converged warp, elected issue (umma.cuh): no ELECT / BRA.U.ANY loop around every tcgen05.mma
      constexpr uint32_t idesc = instr_desc_tf32(128, BN, 1, 1);   // both operands MN-major
      for (int it = 0; it < iters; ++it) {
        const int s = it % L::STAGES;
        const uint32_t ph = (it / L::STAGES) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t sx = smem_u32(smem + s * L::STAGE_BYTES);
        // column groups of the x operand one voxel (128 bytes) apart; 4-voxel K groups 512 bytes apart
        const uint64_t ad0 = smem_desc(sx, 128, 512, LAYOUT_SW128_BASE32B);
        const uint64_t bd0 = smem_desc(sx + L::X_BYTES, L::GA_BYTES, 512, LAYOUT_SW128_BASE32B);     // dy column groups GA_BYTES apart
#pragma unroll
        for (int a = 0; a < L::ACCS; ++a) {
          const int kd = a / L::KH, kh = a % L::KH;
#pragma unroll
          for (int hy = 0; hy < L::TH; ++hy) {
#pragma unroll
            for (int ks = 0; ks < L::TW / UMMA_K; ++ks) {
              const int row_x = (kd * L::HH + kh + hy) * L::HW + UMMA_K * ks;     // first voxel of this K slice, tap kw = 0
              const int row_g = hy * L::TW + UMMA_K * ks;
              if ((hy | ks) != 0)
                umma_tf32_elect<true>(tmem_base + (uint32_t)(a * BN), ad0 + (uint64_t)(8 * row_x), bd0 + (uint64_t)(8 * row_g), idesc);
              else
                umma_tf32_elect_rt(tmem_base + (uint32_t)(a * BN), ad0 + (uint64_t)(8 * row_x), bd0 + (uint64_t)(8 * row_g), idesc,
                                   (uint32_t)(it != 0));
            }
          }
        }
        umma_commit_elect(empty + s);
      }
      umma_commit_elect(tmem_full);
    }
  } else {
    // epilogue: TMEM lane = kw * 32 + (input channel - c0); BN output-channel columns per accumulator
    const int kw = warp & 3;
    const int ci = c0 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const bool live = kw < L::KW && ci < p.Cin;
    const bool vec = (p.Cout & 3) == 0;
#pragma unroll 1
    for (int a = 0; a < L::ACCS; ++a) {
      float* dst = dw + ((long long)(a * L::KW + kw) * p.Cin + ci) * p.Cout + n0;
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 16) {
        float v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(kw * 32) << 16) + (uint32_t)(a * BN + cc), v);
        if (!live || n0 + cc >= p.Cout) continue;
        if (vec && n0 + cc + 16 <= p.Cout) {
#pragma unroll
          for (int j = 0; j < 4; ++j) red_add_v4(dst + cc + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + cc + j < p.Cout) atomicAdd(dst + cc + j, v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, L::TMEM_COLS);
}

// db[c] += sum over voxels of dy[voxel][c]  (dy channels-last, voxel stride Cs; this launch covers C <= 256 channels).
// thread = (channel, voxel lane); 4 independent accumulators keep 4 loads in flight per thread.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, long long pixels, int C, int Cs, long long per_block) {
  __shared__ float part[256];
  const int t = threadIdx.x;
  const int lanes = 256 / C;
  const int c = t % C, pl = t / C;
  const long long p0 = (long long)blockIdx.x * per_block;
  const long long p1 = min(pixels, p0 + per_block);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (pl < lanes) {
    long long px = p0 + pl;
    for (; px + 3LL * lanes < p1; px += 4LL * lanes) {
      a0 += __ldg(dy + px * Cs + c);
      a1 += __ldg(dy + (px + lanes) * Cs + c);
      a2 += __ldg(dy + (px + 2LL * lanes) * Cs + c);
      a3 += __ldg(dy + (px + 3LL * lanes) * Cs + c);
    }
    for (; px < p1; px += lanes) a0 += __ldg(dy + px * Cs + c);
  }
  float acc = (a0 + a1) + (a2 + a3);
  part[t] = acc;
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < lanes; ++l) acc += part[l * C + c];
    atomicAdd(db + c, acc);
  }
}

inline int ceil_bn(int c) { return c <= 32 ? 32 : (c <= 64 ? 64 : (c <= 128 ? 128 : 256)); }

// operand roles: the M side is padded to multiples of 128 channels, the N side to its tile width; pick the
// assignment with the smaller padded product (ties: x on the M side, whose epilogue is vectorised)
inline int pick_x_is_m(int Cin, int Cout) {
  const long long cx = (long long)((Cin + 127) / 128 * 128) * ((Cout + ceil_bn(Cout) - 1) / ceil_bn(Cout) * ceil_bn(Cout));
  const long long cg = (long long)((Cout + 127) / 128 * 128) * ((Cin + ceil_bn(Cin) - 1) / ceil_bn(Cin) * ceil_bn(Cin));
  return cx <= cg;
}

// The halo variant covers 3x3 / 3x3x3 stride-1 layers; it pays off while the per-tap kernel would pad most of
// its 128-row operand (few input channels) and the output channels fit one or two 32-column tiles.
bool wgrad_halo_fits(const dfmir_conv_desc* d) {
  static const int off = getenv("DFMIR_WGRAD_HALO") ? atoi(getenv("DFMIR_WGRAD_HALO")) == 0 : 0;
  if (off) return false;
  for (int a = 0; a < d->nd; ++a) if (d->kernel[a] != 3 || d->pad[a] < 0 || d->pad[a] > 2) return false;
  // output channels per CTA: 9 accumulators x 32 columns in 3-D (two tiles up to 64), 3 x 128 in 2-D
  // 2-D: up to 256 input channels through 32-channel chunks (grid.y): the 256 -> 128 up-sampling convolution ran 0.92 ms
  // on the per-tap kernel, 0.33 ms here (two 128-column tiles for 256 output channels measured no faster than the per-tap kernel)
  if (d->nd == 2) return d->Cin <= 256 && d->Cout <= 128 && d->out_shape[1] >= 24;
  return d->Cin <= 128 && d->Cout <= 64 && d->out_shape[2] >= 24;
}

int wgrad_supported(const dfmir_conv_desc* d) {
  if (!d || (d->nd != 2 && d->nd != 3) || d->stride != 1) return 0;
  if (d->Cin % 4 || d->Cout % 4 || d->Cin < 4 || d->Cout < 4) return 0;     // rows of both operands: 16-byte multiples
  const int nd = d->nd;
  if (d->x_strides[nd + 1] != 1 || d->y_strides[nd + 1] != 1) return 0;
  for (int i = 0; i <= nd; ++i) if ((d->x_strides[i] & 3) || (d->y_strides[i] & 3)) return 0;
  // the per-tap kernel pads the M side to 128 channels: below 16 channels only the halo variant is worth running
  if ((d->Cin < 16 || d->Cout < 16) && !wgrad_halo_fits(d)) return 0;
  return 1;
}

template <int BN>
int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmG, float* dw, const WgradP& p, int taps, int S, cudaStream_t st) {
  constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  using L = WgLayout<BN, STAGES>;
  DFMIR_CUDA(cudaFuncSetAttribute(conv_wgrad_umma_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  dim3 grid((unsigned)taps, (unsigned)(p.m_tiles * p.n_tiles), (unsigned)S);
  conv_wgrad_umma_kernel<BN, STAGES><<<grid, 192, L::TOTAL, st>>>(tmX, tmG, dw, p);
  DFMIR_CHECK_LAUNCH("dfmir_conv_umma_wgrad");
  return DFMIR_OK;
}

// 5-D tiled map over a channels-last activation, box {32 ch, TW, TH, TD, 1}, MN-major TF32 swizzle
int encode_map5(CUtensorMap* tm, const float* act, const long long* st, int nd, int C, const int* shape, int N, int TW, int TH,
                int TD, const char* who) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { dfmir_set_error("%s: cuTensorMapEncodeTiled not available from the driver", who); return DFMIR_ERR_CUDA; }
  if ((uintptr_t)act & 15) { dfmir_set_error("%s: TMA needs a 16-byte aligned base pointer", who); return DFMIR_ERR_ARG; }
  const int D = nd == 3 ? shape[0] : 1, H = shape[nd - 2], W = shape[nd - 1];
  const long long sn = st[0], sw = st[nd], sh = st[nd - 1], sd = nd == 3 ? st[1] : sh * H;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)sw * 4, (cuuint64_t)sh * 4, (cuuint64_t)sd * 4, (cuuint64_t)sn * 4};
  cuuint32_t box[5] = {32, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TD, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  return DFMIR_OK;
}

int bias_grad(const float* dy, float* db, const dfmir_conv_desc* d, cudaStream_t st, const char* who) {
  const int nd = d->nd;
  long long pixels = d->N, dense = d->Cout;
  bool contiguous = d->y_strides[nd + 1] == 1;
  for (int a = nd - 1; a >= 0; --a) { contiguous = contiguous && d->y_strides[1 + a] == dense; dense *= d->out_shape[a]; pixels *= d->out_shape[a]; }
  contiguous = contiguous && d->y_strides[0] == dense;
  DFMIR_CHECK_ARG(contiguous, "%s: bias gradient needs a contiguous dy", who);
  int blocks = 16 * dfmir_num_sms();
  long long per_block = (pixels + blocks - 1) / blocks;
  if (per_block < 32) per_block = 32;
  blocks = (int)((pixels + per_block - 1) / per_block);
  for (int c0 = 0; c0 < d->Cout; c0 += 256) {      // 256 channels per launch
    const int C = d->Cout - c0 < 256 ? d->Cout - c0 : 256;
    bias_grad_kernel<<<blocks, 256, 0, st>>>(dy + c0, db + c0, pixels, C, d->Cout, per_block);
    DFMIR_CHECK_LAUNCH(who);
  }
  return DFMIR_OK;
}

template <int BN, int KD>
int launch_wgrad_halo(const float* x, const float* dy, float* dw, const dfmir_conv_desc* d, cudaStream_t st, const char* who) {
  using L = WgHalo<BN, KD>;
  const int nd = d->nd;
  WgHaloP p;
  p.N = d->N; p.D = nd == 3 ? d->out_shape[0] : 1; p.H = d->out_shape[nd - 2]; p.W = d->out_shape[nd - 1];
  p.Cin = d->Cin; p.Cout = d->Cout;
  p.pad_d = nd == 3 ? d->pad[0] : 0; p.pad_h = d->pad[nd - 2]; p.pad_w = d->pad[nd - 1];
  p.tiles_h = (p.H + L::TH - 1) / L::TH; p.tiles_w = (p.W + L::TW - 1) / L::TW;
  p.ntiles = p.N * p.D * p.tiles_h * p.tiles_w;
  if (p.ntiles == 0) return DFMIR_OK;
  const int chunks = (d->Cin + 31) / 32, n_tiles = (d->Cout + BN - 1) / BN;
  int S = dfmir_num_sms() / (chunks * n_tiles);
  if (S < 1) S = 1;
  if (S > p.ntiles) S = p.ntiles;
  CUtensorMap tmX, tmG;
  int rc = encode_map5(&tmX, x, d->x_strides, nd, d->Cin, d->in_shape, d->N, L::HW, L::HH, L::HD, who);
  if (rc) return rc;
  rc = encode_map5(&tmG, dy, d->y_strides, nd, d->Cout, d->out_shape, d->N, L::TW, L::TH, 1, who);
  if (rc) return rc;
  DFMIR_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel<BN, KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  dim3 grid((unsigned)S, (unsigned)chunks, (unsigned)n_tiles);
  conv_wgrad_halo_kernel<BN, KD><<<grid, 192, L::TOTAL, st>>>(tmX, tmG, dw, p);
  DFMIR_CHECK_LAUNCH(who);
  return DFMIR_OK;
}

}  // namespace

extern "C" int dfmir_conv_umma_wgrad_supported(const dfmir_conv_desc* d) { return wgrad_supported(d); }

// Weight / bias gradient on the tensor cores.  dw [tap][Cin][Cout] and db [Cout] (nullable) are
// ACCUMULATED into (zero-fill them first).  x and dy channels-last with unit channel stride.
extern "C" int dfmir_conv_umma_wgrad(const float* x, const float* dy, float* dw, float* db, const dfmir_conv_desc* d,
                                     void* stream) {
  const char* who = "dfmir_conv_umma_wgrad";
  if (!wgrad_supported(d)) {
    dfmir_set_error("%s: needs a 2-D / 3-D stride-1 convolution with channel counts that are multiples of 4 (>= 16 unless 3x3 / 3x3x3) on "
                    "channels-last operands with 16-byte aligned strides", who);
    return DFMIR_ERR_UNSUPPORTED;
  }
  DFMIR_CHECK_ARG(x && dy && dw, "%s: null pointer", who);
  cudaStream_t st = (cudaStream_t)stream;
  const int nd = d->nd;
  if (wgrad_halo_fits(d)) {
    int rc;
    if (nd == 3) rc = d->Cout <= 16 ? launch_wgrad_halo<16, 3>(x, dy, dw, d, st, who) : launch_wgrad_halo<32, 3>(x, dy, dw, d, st, who);
    else if (d->Cout <= 16) rc = launch_wgrad_halo<16, 1>(x, dy, dw, d, st, who);
    else if (d->Cout <= 32) rc = launch_wgrad_halo<32, 1>(x, dy, dw, d, st, who);
    else if (d->Cout <= 64) rc = launch_wgrad_halo<64, 1>(x, dy, dw, d, st, who);
    else rc = launch_wgrad_halo<128, 1>(x, dy, dw, d, st, who);
    if (rc) return rc;
    return db ? bias_grad(dy, db, d, st, who) : DFMIR_OK;
  }
  WgradP p;
  p.N = d->N; p.D = nd == 3 ? d->out_shape[0] : 1; p.H = d->out_shape[nd - 2]; p.W = d->out_shape[nd - 1];
  p.Cin = d->Cin; p.Cout = d->Cout;
  p.KH = d->kernel[nd - 2]; p.KW = d->kernel[nd - 1];
  p.pad_d = nd == 3 ? d->pad[0] : 0; p.pad_h = d->pad[nd - 2]; p.pad_w = d->pad[nd - 1];
  int TW = PIX;
  while (TW > p.W && TW > 8) TW >>= 1;
  p.TW = TW; p.TH = PIX / TW; p.TD = 1;
  if (p.TH > p.H && p.D > 1) { p.TD = p.TH; p.TH = 1; }     // narrow 3-D volumes: take the rest of the chunk along depth
  p.tiles_w = (p.W + p.TW - 1) / p.TW;
  p.tiles_h = (p.H + p.TH - 1) / p.TH;
  p.tiles_d = (p.D + p.TD - 1) / p.TD;
  p.nchunks = p.N * p.tiles_d * p.tiles_h * p.tiles_w;
  if (p.nchunks == 0) return DFMIR_OK;
  p.x_is_m = pick_x_is_m(d->Cin, d->Cout);
  const int CM = p.x_is_m ? d->Cin : d->Cout, CN = p.x_is_m ? d->Cout : d->Cin;
  const int BN = ceil_bn(CN);
  p.m_tiles = (CM + 127) / 128; p.n_tiles = (CN + BN - 1) / BN;
  int taps = 1;
  for (int a = 0; a < nd; ++a) taps *= d->kernel[a];
  const int tiles = taps * p.m_tiles * p.n_tiles;
  int S = dfmir_num_sms() / tiles;
  if (S < 1) S = 1;
  if (S > p.nchunks) S = p.nchunks;

  CUtensorMap tmX, tmG;
  int rc = encode_map5(&tmX, x, d->x_strides, nd, d->Cin, d->in_shape, d->N, p.TW, p.TH, p.TD, who);
  if (rc) return rc;
  rc = encode_map5(&tmG, dy, d->y_strides, nd, d->Cout, d->out_shape, d->N, p.TW, p.TH, p.TD, who);
  if (rc) return rc;
  static const int pair = getenv("DFMIR_WGRAD_PAIR") ? atoi(getenv("DFMIR_WGRAD_PAIR")) : 1;
  if (pair && p.x_is_m && d->Cout == 256 && d->Cin % 256 == 0) {
    const int smem = WP_STAGES * WP_STAGE_BYTES + (2 * WP_STAGES + 1) * 8 + 16 + 1024;
    DFMIR_CUDA(cudaFuncSetAttribute(conv_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int Sp = dfmir_num_sms() / (2 * taps * (d->Cin / 256));
    if (Sp < 1) Sp = 1;
    if (Sp > p.nchunks) Sp = p.nchunks;
    dim3 grid((unsigned)(2 * taps), (unsigned)(d->Cin / 256), (unsigned)Sp);
    conv_wgrad_pair_kernel<<<grid, 192, smem, st>>>(tmX, tmG, dw, p);
    DFMIR_CHECK_LAUNCH(who);
    rc = DFMIR_OK;
  } else if (BN == 256) rc = launch_wgrad<256>(tmX, tmG, dw, p, taps, S, st);
  else if (BN == 128) rc = launch_wgrad<128>(tmX, tmG, dw, p, taps, S, st);
  else if (BN == 64) rc = launch_wgrad<64>(tmX, tmG, dw, p, taps, S, st);
  else rc = launch_wgrad<32>(tmX, tmG, dw, p, taps, S, st);
  if (rc) return rc;
  if (db) return bias_grad(dy, db, d, st, who);
  return DFMIR_OK;
}
