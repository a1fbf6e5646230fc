// K1 weight gradient on the 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in
// TMEM, operands staged by TMA), for the 2-D stride-1 convolutions of ResnetGenerator
// (models/networks.py:995,1016,1201,1214 — the backward of the cuDNN wgrad engines).
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

constexpr int PIX = 32;                 // pixels (K) per pipeline stage
constexpr int CHUNK_BYTES = PIX * 128;  // one 32-channel column group of a stage
constexpr int UMMA_K = 8;               // tf32

struct WgradP {
  int N, H, W;              // dy: samples and spatial size (output of the forward conv)
  int Cin, Cout;
  int KW, pad_h, pad_w;
  int TW, TH, tiles_w, tiles_h;
  int nchunks;
  int x_is_m;               // 1: M = input channels (x), N = output channels (dy); 0: swapped
  int m_tiles, n_tiles;
};

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
  const int r = tap / p.KW, q = tap - r * p.KW;
  const int dh = r - p.pad_h, dwv = q - p.pad_w;      // x pixel = dy pixel + (dh, dw)
  const int m0 = mt * 128, n0 = nt * BN;
  const int iters = (p.nchunks - split + S - 1) / S;   // host guarantees S <= nchunks

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmG);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mapM = p.x_is_m ? &tmX : &tmG;
      const CUtensorMap* mapN = p.x_is_m ? &tmG : &tmX;
      const int mdh = p.x_is_m ? dh : 0, mdw = p.x_is_m ? dwv : 0;
      const int ndh = p.x_is_m ? 0 : dh, ndw = p.x_is_m ? 0 : dwv;
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        int c = split + it * S;
        const int tw_i = c % p.tiles_w; c /= p.tiles_w;
        const int th_i = c % p.tiles_h; const int n = c / p.tiles_h;
        const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
        mbar_wait(empty + s, ph ^ 1);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_expect_tx(full + s, (uint32_t)L::STAGE_BYTES);
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_4d(sa + g * CHUNK_BYTES, mapM, full + s, m0 + g * 32, w0 + mdw, h0 + mdh, n);
#pragma unroll
        for (int g = 0; g < BN / 32; ++g) tma_load_4d(sb + g * CHUNK_BYTES, mapN, full + s, n0 + g * 32, w0 + ndw, h0 + ndh, n);
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
          // next 8 pixels: +1024 bytes = +64 in 16-byte address units
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
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* dwt = dw + (long long)tap * p.Cin * p.Cout;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      if (p.x_is_m) {
        float* dst = dwt + (long long)(m0 + row) * p.Cout + n0 + c0;     // row = ci, columns = co (contiguous)
#pragma unroll
        for (int j = 0; j < 8; ++j) red_add_v4(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        float* dst = dwt + (long long)(n0 + c0) * p.Cout + m0 + row;     // row = co (coalesced over lanes), columns = ci
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + (long long)j * p.Cout, v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)BN);
}

// db[c] += sum over pixels of dy[pixel][c]  (dy channels-last, contiguous pixels x C; C in {64,128,256}).
// thread = (channel, pixel lane); 4 independent accumulators keep 4 loads in flight per thread.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ db, long long pixels, int C, long long per_block) {
  __shared__ float part[256];
  const int t = threadIdx.x;
  const int lanes = 256 / C;
  const int c = t % C, pl = t / C;
  const long long p0 = (long long)blockIdx.x * per_block;
  const long long p1 = min(pixels, p0 + per_block);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  long long px = p0 + pl;
  for (; px + 3LL * lanes < p1; px += 4LL * lanes) {
    a0 += __ldg(dy + px * C + c);
    a1 += __ldg(dy + (px + lanes) * C + c);
    a2 += __ldg(dy + (px + 2LL * lanes) * C + c);
    a3 += __ldg(dy + (px + 3LL * lanes) * C + c);
  }
  for (; px < p1; px += lanes) a0 += __ldg(dy + px * C + c);
  float acc = (a0 + a1) + (a2 + a3);
  part[t] = acc;
  __syncthreads();
  if (pl == 0) {
    for (int l = 1; l < lanes; ++l) acc += part[l * C + c];
    atomicAdd(db + c, acc);
  }
}

int wgrad_supported(const dfmir_conv_desc* d) {
  if (!d || d->nd != 2 || d->stride != 1) return 0;
  const int Cin = d->Cin, Cout = d->Cout;
  const bool ok_x_m = (Cin % 128 == 0) && (Cout == 64 || Cout == 128 || Cout % 256 == 0);
  const bool ok_g_m = (Cout % 128 == 0) && (Cin == 64 || Cin == 128 || Cin % 256 == 0);
  if (!ok_x_m && !ok_g_m) return 0;
  const long long* xs = d->x_strides; const long long* ys = d->y_strides;
  if (xs[3] != 1 || ys[3] != 1) return 0;
  for (int i = 0; i < 3; ++i) if ((xs[i] & 3) || (ys[i] & 3)) return 0;
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

}  // namespace

extern "C" int dfmir_conv_umma_wgrad_supported(const dfmir_conv_desc* d) { return wgrad_supported(d); }

// Weight / bias gradient on the tensor cores.  dw [tap][Cin][Cout] and db [Cout] (nullable) are
// ACCUMULATED into (zero-fill them first).  dy must be channels-last with unit channel stride.
extern "C" int dfmir_conv_umma_wgrad(const float* x, const float* dy, float* dw, float* db, const dfmir_conv_desc* d,
                                     void* stream) {
  const char* who = "dfmir_conv_umma_wgrad";
  if (!wgrad_supported(d)) {
    dfmir_set_error("%s: needs a 2-D stride-1 convolution with one channel count a multiple of 128 and the other in "
                    "{64,128,256k}, channels-last operands", who);
    return DFMIR_ERR_UNSUPPORTED;
  }
  DFMIR_CHECK_ARG(x && dy && dw, "%s: null pointer", who);
  cudaStream_t st = (cudaStream_t)stream;
  WgradP p;
  p.N = d->N; p.H = d->out_shape[0]; p.W = d->out_shape[1];
  p.Cin = d->Cin; p.Cout = d->Cout;
  p.KW = d->kernel[1]; p.pad_h = d->pad[0]; p.pad_w = d->pad[1];
  int TW = PIX;
  while (TW > p.W && TW > 8) TW >>= 1;
  p.TW = TW; p.TH = PIX / TW;
  p.tiles_w = (p.W + p.TW - 1) / p.TW;
  p.tiles_h = (p.H + p.TH - 1) / p.TH;
  p.nchunks = p.N * p.tiles_h * p.tiles_w;
  if (p.nchunks == 0) return DFMIR_OK;
  // operand roles: prefer x on the M side (vectorised epilogue)
  p.x_is_m = (d->Cin % 128 == 0) && (d->Cout == 64 || d->Cout == 128 || d->Cout % 256 == 0);
  const int CM = p.x_is_m ? d->Cin : d->Cout, CN = p.x_is_m ? d->Cout : d->Cin;
  const int BN = CN >= 256 ? 256 : CN;
  p.m_tiles = CM / 128; p.n_tiles = CN / BN;
  const int taps = d->kernel[0] * d->kernel[1];
  const int tiles = taps * p.m_tiles * p.n_tiles;
  int S = dfmir_num_sms() / tiles;
  if (S < 1) S = 1;
  if (S > p.nchunks) S = p.nchunks;

  CUtensorMap tmX, tmG;
  int rc = encode_act_map(&tmX, x, d->x_strides, d->Cin, d->in_shape[1], d->in_shape[0], d->N, p.TW, p.TH, who,
                          CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = encode_act_map(&tmG, dy, d->y_strides, d->Cout, d->out_shape[1], d->out_shape[0], d->N, p.TW, p.TH, who,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  if (BN == 256) rc = launch_wgrad<256>(tmX, tmG, dw, p, taps, S, st);
  else if (BN == 128) rc = launch_wgrad<128>(tmX, tmG, dw, p, taps, S, st);
  else rc = launch_wgrad<64>(tmX, tmG, dw, p, taps, S, st);
  if (rc) return rc;
  if (db) {
    const long long pixels = (long long)d->N * p.H * p.W;
    // dy contiguous (N,H,W,Cout) is required for the bias reduction
    DFMIR_CHECK_ARG(d->y_strides[2] == d->Cout && d->y_strides[1] == (long long)p.W * d->Cout &&
                    d->y_strides[0] == (long long)p.H * p.W * d->Cout, "%s: bias gradient needs a contiguous dy", who);
    DFMIR_CHECK_ARG(d->Cout == 64 || d->Cout == 128 || d->Cout == 256, "%s: bias gradient covers Cout in {64,128,256}", who);
    int blocks = 16 * dfmir_num_sms();
    long long per_block = (pixels + blocks - 1) / blocks;
    if (per_block < 32) per_block = 32;
    blocks = (int)((pixels + per_block - 1) / per_block);
    bias_grad_kernel<<<blocks, 256, 0, st>>>(dy, db, pixels, d->Cout, per_block);
    DFMIR_CHECK_LAUNCH(who);
  }
  return DFMIR_OK;
}
