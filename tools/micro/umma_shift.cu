// Hardware question for the halo-reuse conv design: a K-major SWIZZLE_128B operand tile is written by TMA at a
// 1024-byte aligned address; can tcgen05.mma read 128 rows starting at row `s` (start address = base + s*128 B,
// not 1024-aligned) and get rows s..s+127 of the matrix?  Variants: base_offset field 0, or (addr>>7)&7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_shift umma_shift.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepspeaker_pytorch_b200/csrc/dsk_ptx.cuh"
using namespace dsk;

constexpr int ROWS_A = 192, N = 64, K = 64;

__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int shift,
             int use_base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                   // 192 rows x 128 B
  uint8_t* sb = smem + ROWS_A * 128;    // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + N * 128);
  uint64_t* done = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(tptr, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, (ROWS_A + N) * 128);
    tma_load_3d(sa, &tmA, bar, 0, 0, 0);
    tma_load_3d(sb, &tmB, bar, 0, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sa) + shift * 128;
    uint64_t da = umma_desc_sw128(a_addr);
    if (use_base_offset) da |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t db = umma_desc_sw128(smem_u32(sb));
    constexpr uint32_t idesc = umma_idesc_f16(128, N, false);
    for (int k = 0; k < K / 16; ++k) umma_f16(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v0[32], v1[32];
  const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  tmem_ld_32x32(taddr, v0); tmem_ld_32x32(taddr + 32, v1); tmem_ld_wait();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 32; ++c) { out[row * N + c] = __uint_as_float(v0[c]); out[row * N + 32 + c] = __uint_as_float(v1[c]); }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  std::vector<__half> A(ROWS_A * K), B(N * K);
  std::vector<float> Af(ROWS_A * K), Bf(N * K);
  srand(1);
  for (int i = 0; i < ROWS_A * K; ++i) { float v = (rand() % 17 - 8) / 8.0f; A[i] = __float2half(v); Af[i] = v; }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.0f; B[i] = __float2half(v); Bf[i] = v; }
  __half *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
  CUtensorMap tmA, tmB;
  cuuint32_t es[3] = {1, 1, 1};
  { cuuint64_t d[3] = {K, ROWS_A, 1}, s[2] = {K * 2, (cuuint64_t)ROWS_A * K * 2}; cuuint32_t b[3] = {64, ROWS_A, 1};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dA, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encA failed\n"); return 1; } }
  { cuuint64_t d[3] = {K, N, 1}, s[2] = {K * 2, (cuuint64_t)N * K * 2}; cuuint32_t b[3] = {64, N, 1};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dB, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encB failed\n"); return 1; } }
  const int smem = (ROWS_A + N) * 128 + 1024 + 64;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * N);
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int shift : {0, 1, 2, 3, 5, 7, 8, 9, 17, 33, 34, 63}) {
      cudaMemset(dO, 0, 128 * N * 4);
      shift_kernel<<<1, 128, smem>>>(tmA, tmB, dO, shift, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d: CUDA error %s\n", shift, ubo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; int bad_rows = 0;
      for (int m = 0; m < 128; ++m) {
        double rowerr = 0;
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)Af[(m + shift) * K + k] * Bf[n * K + k];
          rowerr = fmax(rowerr, fabs(ref - O[m * N + n]));
        }
        if (rowerr > 1e-3) ++bad_rows;
        maxerr = fmax(maxerr, rowerr);
      }
      printf("shift %2d base_offset_field %d: max_err %.4f bad_rows %d/128 %s\n", shift, ubo, maxerr, bad_rows,
             bad_rows ? "MISMATCH" : "OK");
    }
  return 0;
}
