// Round-2 groundwork, part 2: the complete CTA-pair (cta_group::2) GEMM pipeline the halo conv kernel will adopt —
//   * both CTAs of a cluster run a TMA producer: each loads ITS 128 rows of A and ITS half of B per K block into its own
//     shared memory, but the transaction bytes are signalled on the LEADER's full barrier (cp.async.bulk.tensor
//     .cta_group::2 with the barrier address' CTA-rank bit cleared); the peer also arrives there remotely (mapa)
//   * the leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256) and releases ring slots / publishes
//     accumulators with multicast commits that arrive on the barrier at the same offset in both CTAs
//   * both CTAs run an epilogue on their own 128 x N accumulator half (two TMEM stages); the peer hands accumulators back
//     by a remote arrive on the leader's tmem_empty barrier
// D[512 x N] = A[512 x K] B[N x K]^T for two 256-row tiles, K = 64 * KB; checked against a CPU GEMM, and cycles per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_pair_pipe umma_pair_pipe.cu -lcuda
// NOT YET RUN ON HARDWARE (written at the end of round 1 without GPU budget).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deepspeaker_pytorch_b200/csrc/dsk_ptx.cuh"
using namespace dsk;

constexpr int kStages = 3, kTiles = 2, kThreads = 192;  // warp 0 producer, 1 MMA, 2..5 epilogue (TMEM lanes by warp % 4)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
               : "memory");
}
// 2-D TMA load into this CTA's shared memory; completion bytes go to the barrier at `bar_cluster_addr` (any CTA of the pair)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
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
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

template <int N>
struct Smem {
  static constexpr int kA = 128 * 128, kB = (N / 2) * 128, kStage = kA + kB;
  static constexpr int kTotal = kStages * kStage + 1024 + 256;
};

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pipe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D, int KB,
            long long* cyc) {
  using S = Smem<N>;
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStage);
  uint64_t* full = bars;                 // [kStages] used in the leader only: 2 producer arrivals + both CTAs' bytes
  uint64_t* empty = full + kStages;      // [kStages] in each CTA: multicast commit from the leader
  uint64_t* tfull = empty + kStages;     // [2]       in each CTA: multicast commit from the leader
  uint64_t* tempty = tfull + 2;          // [2]       used in the leader only: 4 epilogue warps x 2 CTAs
  uint32_t* tptr = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;      // all pairs compute the same problem (rate measurement across the chip)
  (void)pair;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tptr, 2 * N);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem = *tptr;

  if (warp == 0) {
    // ---- producer (both CTAs) ----
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < kTiles; ++t) {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one_sync()) {
          const uint32_t lead_full = mapa_u32(&full[stage], 0);
          uint8_t* sa = smem + stage * S::kStage;
          if (rank == 0) mbar_arrive_expect_tx_cluster(lead_full, 2 * S::kStage);  // both CTAs' bytes land on this barrier
          else mbar_arrive_remote(lead_full);
          tma_load_2d_pair(sa, &tmA, lead_full, kb * 64, t * 256 + rank * 128);
          tma_load_2d_pair(sa + S::kA, &tmB, lead_full, kb * 64, rank * (N / 2));
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ---- MMA issuer (leader only) ----
    constexpr uint32_t idesc = umma_idesc_f16(256, N, false);
    int stage = 0;
    uint32_t phase = 0;
    long long t0 = 0;
    for (int t = 0; t < kTiles; ++t) {
      mbar_wait(&tempty[t & 1], ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem + (t & 1) * N;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (t == 0 && kb == 0) t0 = clock64();
        if (elect_one_sync()) {
          const uint64_t da = umma_desc_sw128(smem_u32(smem + stage * S::kStage));
          const uint64_t db = umma_desc_sw128(smem_u32(smem + stage * S::kStage + S::kA));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_pair(d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty[stage]);
          if (kb == KB - 1) umma_commit_pair(&tfull[t & 1]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    mbar_wait(&tfull[(kTiles - 1) & 1], ((kTiles - 1) >> 1) & 1);  // all MMAs done
    if (lane == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
  } else if (warp >= 2) {
    // ---- epilogue (both CTAs): rows [128*rank, +128) of each 256-row tile ----
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    for (int t = 0; t < kTiles; ++t) {
      mbar_wait(&tfull[t & 1], (t >> 1) & 1);
      tc_fence_after();
      const int row = t * 256 + rank * 128 + q * 32 + lane;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + (t & 1) * N + c0, v);
        tmem_ld_wait();
        if (blockIdx.x < 2)
          for (int c = 0; c < 32; ++c) D[static_cast<size_t>(row) * N + c0 + c] = __uint_as_float(v[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_u32(&tempty[t & 1], 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 2 * N);
  }
}

template <int N>
int run(int pairs, int KB) {
  const int M = 256 * kTiles, K = 64 * KB;
  std::vector<__half> A(static_cast<size_t>(M) * K), B(static_cast<size_t>(N) * K);
  std::vector<float> Af(A.size()), Bf(B.size());
  srand(11);
  for (size_t i = 0; i < A.size(); ++i) { float v = (rand() % 9 - 4) / 8.0f; A[i] = __float2half(v); Af[i] = v; }
  for (size_t i = 0; i < B.size(); ++i) { float v = (rand() % 7 - 3) / 4.0f; B[i] = __float2half(v); Bf[i] = v; }
  __half *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, static_cast<size_t>(M) * N * 4); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, static_cast<size_t>(M) * N * 4);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
  CUtensorMap tmA, tmB;
  cuuint32_t es[2] = {1, 1};
  { cuuint64_t d[2] = {(cuuint64_t)K, (cuuint64_t)M}, s[1] = {(cuuint64_t)K * 2}; cuuint32_t b[2] = {64, 128};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encA failed\n"); return 1; } }
  { cuuint64_t d[2] = {(cuuint64_t)K, (cuuint64_t)N}, s[1] = {(cuuint64_t)K * 2}; cuuint32_t b[2] = {64, (cuuint32_t)(N / 2)};
    if (enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encB failed\n"); return 1; } }
  cudaFuncSetAttribute(pipe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<N>::kTotal);
  pipe_kernel<N><<<2 * pairs, kThreads, Smem<N>::kTotal>>>(tmA, tmB, dD, KB, dC);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N %d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
  long long C[2];
  cudaMemcpy(C, dC, 16, cudaMemcpyDeviceToHost);
  std::vector<float> Dh(static_cast<size_t>(M) * N);
  cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; long bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)Af[(size_t)m * K + k] * Bf[(size_t)n * K + k];
      const double er = fabs(ref - Dh[(size_t)m * N + n]);
      if (er > 1e-2) ++bad;
      maxerr = fmax(maxerr, er);
    }
  printf("N %3d pairs %3d KB %3d: max_err %.4f bad %ld/%d %s | %.1f cycles per M256 MMA (tensor-bound %d)\n", N, pairs, KB, maxerr,
         bad, M * N, bad ? "MISMATCH" : "OK", (double)C[0] / (kTiles * KB * 4.0), N / 2);
  fflush(stdout);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad ? 1 : 0;
}

int main() {
  int rc = 0;
  for (int pairs : {1, 74}) {
    rc |= run<64>(pairs, 16);
    rc |= run<128>(pairs, 16);
    rc |= run<256>(pairs, 16);
  }
  return rc;
}
