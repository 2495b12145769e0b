// Round-2 groundwork: tcgen05.mma.cta_group::2 (CTA pair, M = 256 across two SMs) — correctness of the operand split and
// the MMA rate.  In SS mode a single CTA is bound by its 128 B/clk shared-memory port (umma_rate.cu: 48 / 64 cycles per
// M128 x K16 MMA at N = 64 / 128).  In a pair each CTA supplies its own 128 rows of A and HALF of B (N/2 rows), so the
// per-CTA operand bytes per MMA drop from 4 KB + N*32 B to 4 KB + N*16 B: expected 40 cycles at N = 64, tensor-bound
// 64 / 128 cycles at N = 128 / 256.
//   * cluster of 2 CTAs; each CTA writes its A half (128 x 64, K-major SWIZZLE_128B) and its B half (N/2 x 64) into its
//     own shared memory with plain stores (no TMA needed for the question asked here)
//   * one warp per CTA allocates TMEM with cta_group::2; the leader (rank 0) issues the MMAs; completion is multicast
//     to an mbarrier at the same offset in both CTAs; each CTA reads its 128 x N accumulator half back
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_pair umma_pair.cu
// NOT YET RUN ON HARDWARE (written at the end of round 1 without GPU budget): first thing to run in round 2.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepspeaker_pytorch_b200/csrc/dsk_ptx.cuh"
using namespace dsk;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on `bar` in both CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// A: [256][64] halfs, B: [N][64] halfs (row-major, K contiguous); D: [256][N] fp32; cyc[0..1]: issue / complete cycles
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, long long* cyc, int iters) {
  __shared__ __align__(1024) uint8_t sA[128 * 128];
  __shared__ __align__(1024) uint8_t sB[(N / 2) * 128];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  // swizzled K-major rows: element (m, k) at m*128 + ((k/8) ^ (m&7))*16 + (k%8)*2
  {
    const uint4* src = reinterpret_cast<const uint4*>(A + (static_cast<size_t>(rank) * 128 + tid) * 64);
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sA + tid * 128 + ((c ^ (tid & 7)) << 4)) = src[c];
    if (tid < N / 2) {
      const uint4* sb = reinterpret_cast<const uint4*>(B + (static_cast<size_t>(rank) * (N / 2) + tid) * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sB + tid * 128 + ((c ^ (tid & 7)) << 4)) = sb[c];
    }
  }
  if (tid == 0) {
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(&tptr, N < 32 ? 32 : N);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' operands written, barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem = tptr;
  if (rank == 0 && warp == 0) {
    // instruction descriptor: M = 256 (field M>>4 = 16), N, fp16 operands, fp32 accumulate
    constexpr uint32_t idesc = umma_idesc_f16(256, N, false);
    const uint64_t da = umma_desc_sw128(smem_u32(sA));
    const uint64_t db = umma_desc_sw128(smem_u32(sB));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_pair(tmem, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit_pair(&done);
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (tid == 0) {
      cyc[0] = t1 - t0;
      cyc[1] = t2 - t0;
    }
  }
  mbar_wait(&done, 0);  // the multicast commit arrives in both CTAs
  tc_fence_after();
  // each CTA holds rows [128*rank, 128*rank+128) of D, all N columns
  const int row = rank * 128 + tid;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) D[static_cast<size_t>(row) * N + c0 + c] = __uint_as_float(v[c]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs done with TMEM before it is released
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, N < 32 ? 32 : N);
  }
}

template <int N>
void run(int grid_pairs) {
  std::vector<__half> A(256 * 64), B(N * 64);
  std::vector<float> Af(256 * 64), Bf(N * 64);
  srand(7);
  for (size_t i = 0; i < A.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; A[i] = __float2half(v); Af[i] = v; }
  for (size_t i = 0; i < B.size(); ++i) { float v = (rand() % 13 - 6) / 4.0f; B[i] = __float2half(v); Bf[i] = v; }
  __half *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 256 * N * 4); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  for (int iters : {1, 2000}) {
    cudaMemset(dD, 0, 256 * N * 4);
    pair_kernel<N><<<2 * grid_pairs, 128>>>(dA, dB, dD, dC, iters);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: CUDA error %s\n", N, cudaGetErrorString(e)); exit(1); }
    long long C[2];
    cudaMemcpy(C, dC, 16, cudaMemcpyDeviceToHost);
    if (iters == 1) {
      std::vector<float> D(256 * N);
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < 256; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)Af[m * 64 + k] * Bf[n * 64 + k];
          const double er = fabs(ref - D[m * N + n]);
          if (er > 1e-3) ++bad;
          maxerr = fmax(maxerr, er);
        }
      printf("N %3d pairs %3d: D = A B^T max_err %.4f bad %d/%d %s\n", N, grid_pairs, maxerr, bad, 256 * N, bad ? "MISMATCH" : "OK");
    } else {
      printf("N %3d pairs %3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (tensor-bound %d, single-CTA port bound %d)\n", N,
             grid_pairs, C[0] / (iters * 4.0), C[1] / (iters * 4.0), N / 2, (4096 + N * 32) / 128);
    }
    fflush(stdout);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
}

int main() {
  for (int pairs : {1, 74}) {
    run<64>(pairs);
    run<128>(pairs);
    run<256>(pairs);
  }
  return 0;
}
