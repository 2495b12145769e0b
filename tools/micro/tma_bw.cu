// Microbenchmark: how fast can 148 persistent CTAs pull 16 KB TMA boxes out of L2 into shared memory?
// Modes: 0 = every CTA streams its own region (A-operand like), 1 = all CTAs read the same 1 MB (weights like),
//        2 = each CTA re-reads its own 64 KB 9 times (per-tap reload pattern).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu -lcuda ; run: ./tma_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../deepspeaker_pytorch_b200/csrc/dsk_ptx.cuh"
using namespace dsk;

template <int STAGES>
__global__ void __launch_bounds__(128, 1) bw_kernel(const __grid_constant__ CUtensorMap tm, int iters, int mode,
                                                    int rows_total, int box_rows) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * box_rows * 128);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const int nboxes = rows_total / box_rows;
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      int box;
      if (mode == 0) box = (blockIdx.x * iters + it) % nboxes;
      else if (mode == 1) box = it % 64;
      else box = (blockIdx.x * 4 + (it / 9) % 4 + (it / 36) * 592) % nboxes;
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], box_rows * 128);
      tma_load_3d(smem + stage * box_rows * 128, &tm, &full[stage], 0, box * box_rows, 0);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }
}

int main() {
  const int rows_total = 1 << 20;  // 1M rows x 128 B = 128 MB
  void* buf; cudaMalloc(&buf, (size_t)rows_total * 128); cudaMemset(buf, 1, (size_t)rows_total * 128);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
  for (int box_rows : {128, 256, 64, 32}) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {64, (cuuint64_t)rows_total, 1}; cuuint64_t str[2] = {128, (cuuint64_t)rows_total * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d\n", (int)r); return 1; }
    for (int mode = 0; mode < 3; ++mode) {
      for (int stages : {4, 8}) {
        const int iters = 2000 * 128 / box_rows;
        const int smem = stages * box_rows * 128 + 1024 + 256;
        auto k = stages == 4 ? bw_kernel<4> : bw_kernel<8>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          k<<<148, 128, smem>>>(tm, iters, mode, rows_total, box_rows);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = 148.0 * iters * box_rows * 128;
        printf("box %3d rows (%5d B) mode %d stages %d: %.3f ms  %.2f TB/s  (%.1f B/clk/SM @1.9GHz)  err=%s\n", box_rows,
               box_rows * 128, mode, stages, ms, bytes / ms / 1e9, bytes / ms / 1e6 / 148 / 1.9e3 * 1e0,
               cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
