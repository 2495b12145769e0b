// Round-2 groundwork (NOT YET RUN ON HARDWARE): can an MN-major SWIZZLE_128B operand be read at an arbitrary ROW offset?
// For the weight gradient on the zero-padded layout, K = positions and a position is one 128-byte row of 64 channels, so
// one X halo tile (128 + 2W + 4 rows) could serve several filter taps through descriptors whose start address is shifted
// by `shift` rows (as the K-major A operand of the forward halo kernel: tools/micro/umma_shift.cu) — 3 passes over G / X
// with 4 tap accumulators in TMEM instead of 9.  Here: B has 192 rows, D[m][n] = sum_k At[k][m] * Bt[k + shift][n].
// Original question (umma_mnmajor.cu): both operands MN-major (the reduction index K = pixel rows,
// 128-byte rows of 64 channels as TMA writes an NHWC box with SWIZZLE_128B).
//   D[m][n] = sum_k At[k][m] * Bt[k][n],  K = 128 rows, M = 128 (2 atoms of 64 channels), N = 128 (2 atoms).
// Descriptor: start = atom0 + kstep*16 rows*128 B, LBO = atom stride (16 KB), SBO = 1024 B (8-row groups).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepspeaker_pytorch_b200/csrc/dsk_ptx.cuh"
using namespace dsk;
constexpr int KR = 128, M = 128, N = 128, KB = 192;  // B has KB rows so that a shifted window of KR rows exists
constexpr int kBAtom = KB * 128;                     // bytes of one 64-channel atom of B

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
mn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int shift) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // 2 atoms x (128 rows x 128 B)
  uint8_t* sb = smem + 2 * 16384;     // 2 atoms of KB rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * 16384 + 2 * kBAtom);
  uint64_t* done = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(tptr, 128); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 2 * 16384 + 2 * kBAtom);
    tma_load_3d(sa, &tmA, bar, 0, 0, 0);
    tma_load_3d(sa + 16384, &tmA, bar, 64, 0, 0);
    tma_load_3d(sb, &tmB, bar, 0, 0, 0);
    tma_load_3d(sb + kBAtom, &tmB, bar, 64, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    // idesc: f32 accum, f16 operands, a_major = b_major = MN (bits 15, 16)
    const uint32_t idesc = umma_idesc_f16(M, N, false) | (1u << 15) | (1u << 16);
    for (int k = 0; k < KR / 16; ++k) {  // LBO = atom stride, SBO = 1024 B (the variant umma_mnmajor.cu found correct)
      const uint64_t da = desc_mn_sw128(smem_u32(sa) + k * 2048, 16384, 1024);
      const uint64_t db = desc_mn_sw128(smem_u32(sb) + shift * 128 + k * 2048, kBAtom, 1024);
      umma_f16(tmem, da, db, idesc, k > 0);
    }
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int j = 0; j < N / 32; ++j) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + j * 32, v);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) out[row * N + j * 32 + c] = __uint_as_float(v[c]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

int main() {
  std::vector<__half> A(KR * M), B(KB * N);
  std::vector<float> Af(KR * M), Bf(KB * N);
  srand(2);
  for (int i = 0; i < KR * M; ++i) { float v = (rand() % 17 - 8) / 8.0f; A[i] = __float2half(v); Af[i] = v; }
  for (int i = 0; i < KB * N; ++i) { float v = (rand() % 13 - 6) / 4.0f; B[i] = __float2half(v); Bf[i] = v; }
  __half *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
  CUtensorMap tmA, tmB; cuuint32_t es[3] = {1, 1, 1};
  { cuuint64_t d[3] = {M, KR, 1}, s[2] = {M * 2, (cuuint64_t)KR * M * 2}; cuuint32_t b[3] = {64, KR, 1};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dA, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encA failed\n"); return 1; } }
  { cuuint64_t d[3] = {N, KB, 1}, s[2] = {N * 2, (cuuint64_t)KB * N * 2}; cuuint32_t b[3] = {64, KB, 1};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dB, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encB failed\n"); return 1; } }
  const int smem = 2 * 16384 + 2 * kBAtom + 1024 + 64;
  cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(M * N);
  for (int shift : {0, 1, 2, 3, 7, 8, 9, 17, 18, 19, 33, 34, 35, 64}) {
    cudaMemset(dO, 0, M * N * 4);
    mn_kernel<<<1, 128, smem>>>(tmA, tmB, dO, shift);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < KR; ++k) ref += (double)Af[k * M + m] * Bf[(k + shift) * N + n];
        const double err = fabs(ref - O[m * N + n]);
        if (err > 1e-2) ++bad;
        maxerr = fmax(maxerr, err);
      }
    printf("MN-major B read at row shift %2d: max_err %.4f bad %d/%d %s\n", shift, maxerr, bad, M * N, bad ? "MISMATCH" : "OK");
  }
  return 0;
}
