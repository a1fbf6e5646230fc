// K1 (tensor-core path): 2-D stride-1 convolution as an implicit GEMM on the 5th-generation tensor
// cores — tcgen05.mma (kind::tf32, fp32 accumulators in TMEM), operands staged in shared memory by
// TMA, persistent CTAs (one per SM) that each walk a list of output tiles.
//
// Replaces the cuDNN implicit-GEMM engines behind nn.Conv2d in ResnetGenerator
// (models/networks.py:995,1016 and the 18 ResnetBlock convs :1201,1214) for the layers whose
// channel counts fill a tensor-core tile (Cin % 32 == 0, Cout in {64,128,256}).  The reference's
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
  int N, H, W;            // output sample count and spatial size
  int Cin, Cout;
  int KH, KW, pad_h, pad_w;
  int TW, TH, tiles_w, tiles_h;
  int ptiles;             // N * tiles_h * tiles_w sub-tiles of 128 pixels
  int flip;               // 1: use tap (KH*KW-1-t) of the weight tensor (data gradient)
  int act;
  long long ys[4];        // output element strides n, h, w, c
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == DFMIR_ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  if (act == DFMIR_ACT_TANH) return tanhf(v);
  if (act == DFMIR_ACT_RELU) return v > 0.f ? v : 0.f;
  return v;
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
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B: 1024-byte aligned
  uint64_t* full = (uint64_t*)(smem + L::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + AS;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + AS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.Cout / BN;
  const int pgroups = (p.ptiles + MT - 1) / MT;
  const int items = pgroups * n_tiles;
  const int cchunks = p.Cin / KCH;
  const int taps = p.KH * p.KW;
  const int num_kb = taps * cchunks;
  const int tiles_per_img = p.tiles_h * p.tiles_w;

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
        int sn[MT], sh[MT], sw[MT];
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int pt = pg * MT + j;
          if (pt < p.ptiles) {
            const int n = pt / tiles_per_img, rem = pt - n * tiles_per_img;
            const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
            sn[j] = n; sh[j] = th_i * p.TH; sw[j] = tw_i * p.TW;
          } else { sn[j] = p.N; sh[j] = 0; sw[j] = 0; }      // beyond the batch: the TMA unit zero-fills
        }
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty + s, ph ^ 1);
          const int tap = kb / cchunks, cc = kb - tap * cchunks;
          const int r = tap / p.KW, q = tap - r * p.KW;
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          mbar_expect_tx(full + s, (uint32_t)L::STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < MT; ++j)
            tma_load_4d(sa + j * A_BYTES, &tmA, full + s, cc * KCH, sw[j] + q - p.pad_w, sh[j] + r - p.pad_h, sn[j]);
          tma_load_3d(sa + MT * A_BYTES, &tmB, full + s, cc * KCH, nt * BN, p.flip ? taps - 1 - tap : tap);
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
    const int th = row / p.TW, tw = row - th * p.TW;
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
        const int n = pt / tiles_per_img, rem = pt - n * tiles_per_img;
        const int th_i = rem / p.tiles_w, tw_i = rem - th_i * p.tiles_w;
        const int oh = th_i * p.TH + th, ow = tw_i * p.TW + tw;
        const bool valid = pt < p.ptiles && oh < p.H && ow < p.W;
        float* yp = y + (long long)n * p.ys[0] + (long long)oh * p.ys[1] + (long long)ow * p.ys[2];
        const uint32_t acc = tmem_base + as * (MT * BN) + j * BN + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          float v[32];
          tmem_ld_32x32(acc + (uint32_t)c0, v);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float t = v[i];
              if (bias) t += __ldg(bias + n0 + c0 + i);
              v[i] = act_apply(t, p.act);
            }
            if (p.ys[3] == 1) {
              float4* dst = reinterpret_cast<float4*>(yp + n0 + c0);
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) yp[(long long)(n0 + c0 + i) * p.ys[3]] = v[i];
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

// ---------------------------------------------------------------- host side
int fill_umma(UmmaP& p, const dfmir_conv_desc* d, int dgrad, const char* who) {
  if (!d) { dfmir_set_error("%s: null descriptor", who); return DFMIR_ERR_ARG; }
  if (d->nd != 2 || d->stride != 1) { dfmir_set_error("%s: tensor-core path covers 2-D stride-1 convolutions", who); return DFMIR_ERR_UNSUPPORTED; }
  const int Cin = dgrad ? d->Cout : d->Cin, Cout = dgrad ? d->Cin : d->Cout;
  if (Cin % KCH || !(Cout == 64 || Cout == 128 || Cout == 256)) {
    dfmir_set_error("%s: needs Cin %% 32 == 0 and Cout in {64,128,256} (got %d -> %d)", who, Cin, Cout);
    return DFMIR_ERR_UNSUPPORTED;
  }
  p.N = d->N; p.Cin = Cin; p.Cout = Cout;
  p.KH = d->kernel[0]; p.KW = d->kernel[1];
  const int* osh = dgrad ? d->in_shape : d->out_shape;
  p.H = osh[0]; p.W = osh[1];
  p.pad_h = dgrad ? d->kernel[0] - 1 - d->pad[0] : d->pad[0];
  p.pad_w = dgrad ? d->kernel[1] - 1 - d->pad[1] : d->pad[1];
  p.flip = dgrad; p.act = dgrad ? DFMIR_ACT_NONE : d->act;
  const long long* os = dgrad ? d->x_strides : d->y_strides;
  for (int i = 0; i < 4; ++i) p.ys[i] = os[i];
  // tile rectangle TH x TW = 128 pixels (powers of two, TW >= 8): the shape that wastes the fewest
  // pixels on partial tiles (66x66 data-gradient outputs: 16x8 covers 76 % vs 52 % for 64x2); ties go to
  // the widest tile (longest contiguous runs per TMA box row)
  long long best = -1;
  for (int TW = 128; TW >= 8; TW >>= 1) {
    const int TH = BM / TW;
    const long long area = (long long)((p.W + TW - 1) / TW) * TW * ((p.H + TH - 1) / TH) * TH;
    if (best < 0 || area < best) { best = area; p.TW = TW; p.TH = TH; }
  }
  p.tiles_w = (p.W + p.TW - 1) / p.TW;
  p.tiles_h = (p.H + p.TH - 1) / p.TH;
  p.ptiles = p.N * p.tiles_h * p.tiles_w;
  return DFMIR_OK;
}

template <int BN, int MT, int STAGES, int AS>
int launch_umma(const CUtensorMap& tmA, const CUtensorMap& tmB, const float* bias, float* y, const UmmaP& p, cudaStream_t st,
                const char* who) {
  using L = SmemLayout<BN, MT, STAGES, AS>;
  DFMIR_CUDA(cudaFuncSetAttribute(conv_umma_kernel<BN, MT, STAGES, AS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  const int items = ((p.ptiles + MT - 1) / MT) * (p.Cout / BN);
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

// act: source activation (channels-last, element strides {n,h,w,c}, c stride 1) of spatial size (IH, IW)
int run_umma(const float* act, const long long* as, int IH, int IW, const float* w, const float* bias, float* y,
             const UmmaP& p, cudaStream_t st, const char* who) {
  static const int cfg = getenv("DFMIR_UMMA_CFG") ? atoi(getenv("DFMIR_UMMA_CFG")) : 0;   // tuning experiments
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { dfmir_set_error("%s: cuTensorMapEncodeTiled not available from the driver", who); return DFMIR_ERR_CUDA; }
  if (as[3] != 1 || ((uintptr_t)act & 15) || (as[0] & 3) || (as[1] & 3) || (as[2] & 3) || ((uintptr_t)w & 15)) {
    dfmir_set_error("%s: TMA needs unit channel stride and 16-byte aligned rows", who); return DFMIR_ERR_ARG;
  }
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)p.N};
    cuuint64_t strides[3] = {(cuuint64_t)as[2] * 4, (cuuint64_t)as[1] * 4, (cuuint64_t)as[0] * 4};
    cuuint32_t box[4] = {KCH, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)act, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(activation) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  }
  {
    const int taps = p.KH * p.KW;
    const int BN = (p.Cout == 256 && cfg != 1 && cfg != 5) ? 128 : (p.Cout > 256 ? 256 : p.Cout);
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Cin * p.Cout * 4};
    cuuint32_t box[3] = {KCH, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { dfmir_set_error("%s: cuTensorMapEncodeTiled(weights) failed (%d)", who, (int)r); return DFMIR_ERR_CUDA; }
  }
  if (p.Cout == 256) {
    // two 128-channel halves per pixel group: 48 KB stages x 4 and double-buffered accumulators beat one
    // 256-wide tile (64 KB x 3, no epilogue overlap): 552 vs 430 TFLOP/s on the ResnetBlock conv at batch 32
    if (cfg == 1) return launch_umma<256, 1, 4, 2>(tmA, tmB, bias, y, p, st, who);
    if (cfg == 5) return launch_umma<256, 2, 3, 1>(tmA, tmB, bias, y, p, st, who);
    if (cfg == 3) return launch_umma<128, 4, 2, 1>(tmA, tmB, bias, y, p, st, who);
    if (cfg == 4) return launch_umma<128, 1, 6, 2>(tmA, tmB, bias, y, p, st, who);
    return launch_umma<128, 2, 4, 2>(tmA, tmB, bias, y, p, st, who);
  }
  if (p.Cout == 128) return launch_umma<128, 2, 4, 2>(tmA, tmB, bias, y, p, st, who);
  return launch_umma<64, 4, 3, 2>(tmA, tmB, bias, y, p, st, who);
}

}  // namespace

extern "C" int dfmir_conv_umma_supported(const dfmir_conv_desc* d, int dgrad) {
  if (!d || d->nd != 2 || d->stride != 1) return 0;
  const int Cin = dgrad ? d->Cout : d->Cin, Cout = dgrad ? d->Cin : d->Cout;
  if (Cin % KCH || !(Cout == 64 || Cout == 128 || Cout == 256)) return 0;
  const long long* is = dgrad ? d->y_strides : d->x_strides;
  const long long* os = dgrad ? d->x_strides : d->y_strides;
  if (is[3] != 1 || (is[0] & 3) || (is[1] & 3) || (is[2] & 3)) return 0;
  if (os[3] == 1 && ((os[0] & 3) || (os[1] & 3) || (os[2] & 3))) return 0;
  return 1;
}

// Forward on the tensor cores.  w: [tap][Cout][Cin] (Cin contiguous).
extern "C" int dfmir_conv_umma_fwd(const float* x, const float* w, const float* bias, float* y,
                                   const dfmir_conv_desc* d, void* stream) {
  UmmaP p;
  int rc = fill_umma(p, d, 0, "dfmir_conv_umma_fwd");
  if (rc) return rc;
  DFMIR_CHECK_ARG(x && w && y, "dfmir_conv_umma_fwd: null pointer");
  return run_umma(x, d->x_strides, d->in_shape[0], d->in_shape[1], w, bias, y, p, (cudaStream_t)stream, "dfmir_conv_umma_fwd");
}

// Data gradient on the tensor cores: dx = conv(dy, flipped taps, pad' = k-1-pad).
// w: the FORWARD layout of the fp32 path, [tap][Cin][Cout] (Cout contiguous): K-major for this product.
extern "C" int dfmir_conv_umma_dgrad(const float* dy, const float* w, float* dx, const dfmir_conv_desc* d,
                                     void* stream) {
  UmmaP p;
  int rc = fill_umma(p, d, 1, "dfmir_conv_umma_dgrad");
  if (rc) return rc;
  DFMIR_CHECK_ARG(dy && w && dx, "dfmir_conv_umma_dgrad: null pointer");
  return run_umma(dy, d->y_strides, d->out_shape[0], d->out_shape[1], w, nullptr, dx, p, (cudaStream_t)stream, "dfmir_conv_umma_dgrad");
}
