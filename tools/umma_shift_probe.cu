// Development probe: can a K-major SWIZZLE_128B UMMA operand start at a 128-byte row offset that is NOT a
// multiple of 1024 bytes (shifted window into a larger TMA-written tile)?  Tests row shifts s = 0..9 with the
// descriptor's base_offset field = 0 and = (start >> 7) & 7, and two stride-byte-offset values.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../dfmir_b200/csrc/umma.cuh"
using namespace umma;

constexpr int ROWS = 192, KC = 32, N = 16;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                             float* out, int shift, int mode, int sbo_rows) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + ROWS * 128;
  uint64_t* bar = (uint64_t*)(sB + 4096); uint64_t* done = bar + 1; uint32_t* slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(slot, 32);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * 128 + N * 128);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sA)), "l"(&tmA), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sB)), "l"(&tmB), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA) + shift * 128;
    uint64_t ad = smem_desc_sw128(a0, 16, sbo_rows * 128);
    if (mode == 1) ad |= (uint64_t)((a0 >> 7) & 7) << 49;
    const uint64_t bd = smem_desc_sw128(smem_u32(sB), 16, 1024);
    for (int k = 0; k < 4; ++k) umma_tf32(tm, ad + 2 * k, bd + 2 * k, instr_desc_tf32(128, N), k != 0);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  float v[16];
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(tm + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
  for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = v[i];
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 32);
}

int main() {
  std::vector<float> A(ROWS * KC), B(N * KC);
  for (int r = 0; r < ROWS; ++r) for (int k = 0; k < KC; ++k) A[r * KC + k] = (float)(((r * 7 + k * 3) % 31) - 15);
  for (int n = 0; n < N; ++n) for (int k = 0; k < KC; ++k) B[n * KC + k] = (float)(((n * 5 + k) % 13) - 6);
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * 16 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  auto enc = get_encode();
  CUtensorMap tA, tB;
  { cuuint64_t d[2] = {KC, ROWS}, s[1] = {KC * 4}; cuuint32_t b[2] = {KC, ROWS}, e[2] = {1, 1};
    CUresult r = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encA %d\n", r); return 1; } }
  { cuuint64_t d[2] = {KC, N}, s[1] = {KC * 4}; cuuint32_t b[2] = {KC, N}, e[2] = {1, 1};
    CUresult r = enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("encB %d\n", r); return 1; } }
  const int smem = ROWS * 128 + 4096 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * 16);
  // sbo_rows = 8: consecutive 8-row groups (plain shifted window); sbo_rows = 10: groups 10 rows apart (halo pitch TW+2 = 10)
  for (int sbo_rows : {8, 10})
    for (int mode = 0; mode < 2; ++mode)
      for (int shift = 0; shift < 12; ++shift) {
        probe<<<1, 128, smem>>>(tA, tB, dO, shift, mode, sbo_rows);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("sbo %d mode %d shift %d: %s\n", sbo_rows, mode, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int src = shift + (m / 8) * sbo_rows + (m % 8);
          for (int n = 0; n < N; ++n) {
            float ref = 0; for (int k = 0; k < KC; ++k) ref += A[src * KC + k] * B[n * KC + k];
            if (O[m * 16 + n] != ref) ++bad;
          }
        }
        printf("sbo_rows %2d base_offset_mode %d shift %2d : %s (%d wrong of 2048)\n", sbo_rows, mode, shift, bad ? "MISMATCH" : "ok", bad);
      }
  return 0;
}
