// Development probe for the halo weight-gradient kernel: an MN-major (SWIZZLE_128B_BASE32B) tf32 operand read
// through a descriptor whose start address is shifted by whole 128-byte rows (= K indices, voxels) into a larger
// TMA-written tile, and whose leading byte offset is ONE row (128 bytes), so that the four 32-channel column
// groups of an M = 128 operand are four consecutive voxel shifts of the same tile (three taps along w + one
// unused).  Expected: D[g*32 + c][n] = sum_p A[shift + g + p][c] * B[p][n].
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../dfmir_b200/csrc/umma.cuh"
using namespace umma;

constexpr int ROWS = 64, PIXK = 32, N = 32;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                             float* out, int shift, int lbo_rows) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + ROWS * 128;
  uint64_t* bar = (uint64_t*)(sB + PIXK * 128); uint64_t* done = bar + 1; uint32_t* slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(slot, 32);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * 128 + PIXK * 128);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sA)), "l"(&tmA), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sB)), "l"(&tmB), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint64_t ad = smem_desc(smem_u32(sA) + shift * 128, lbo_rows * 128, 512, LAYOUT_SW128_BASE32B);
    const uint64_t bd = smem_desc(smem_u32(sB), PIXK * 128, 512, LAYOUT_SW128_BASE32B);
    for (int k = 0; k < PIXK / 8; ++k) umma_tf32(tm, ad + 64 * k, bd + 64 * k, instr_desc_tf32(128, N, 1, 1), k != 0);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  float v[32];
  tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16), v);
  for (int i = 0; i < 32; ++i) out[threadIdx.x * 32 + i] = v[i];
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 32);
}

int main() {
  std::vector<float> A(ROWS * 32), B(PIXK * 32);
  for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 32; ++c) A[r * 32 + c] = (float)(((r * 7 + c * 3) % 31) - 15);
  for (int p = 0; p < PIXK; ++p) for (int n = 0; n < 32; ++n) B[p * 32 + n] = (float)(((p * 5 + n) % 13) - 6);
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  auto enc = get_encode();
  CUtensorMap tA, tB;
  { cuuint64_t d[2] = {32, ROWS}, s[1] = {128}; cuuint32_t b[2] = {32, ROWS}, e[2] = {1, 1};
    CUresult r = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encA %d\n", r); return 1; } }
  { cuuint64_t d[2] = {32, PIXK}, s[1] = {128}; cuuint32_t b[2] = {32, PIXK}, e[2] = {1, 1};
    CUresult r = enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encB %d\n", r); return 1; } }
  const int smem = ROWS * 128 + PIXK * 128 + 64 + 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * 32);
  for (int lbo_rows : {1, 2, 8})
    for (int shift = 0; shift < 12; ++shift) {
      probe<<<1, 128, smem>>>(tA, tB, dO, shift, lbo_rows);
      cudaError_t e = cudaDeviceSynchronize();
      if (e) { printf("lbo %d shift %d: %s\n", lbo_rows, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, badg[4] = {0, 0, 0, 0};
      for (int m = 0; m < 128; ++m) {
        const int g = m / 32, c = m % 32;
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int p = 0; p < PIXK; ++p) {
            const int row = shift + g * lbo_rows + p;
            ref += (row < ROWS ? A[row * 32 + c] : 0.f) * B[p * 32 + n];
          }
          if (O[m * 32 + n] != ref) { ++bad; ++badg[g]; }
        }
      }
      printf("lbo_rows %d shift %2d : %s (%d wrong of 4096; per column group %d %d %d %d)\n", lbo_rows, shift, bad ? "MISMATCH" : "ok", bad,
             badg[0], badg[1], badg[2], badg[3]);
    }
  return 0;
}
