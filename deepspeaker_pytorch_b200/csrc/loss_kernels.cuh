// Distance / triplet-loss / selection kernels (fp32).
// The summation orders are part of the contract: oracle/dsk_oracle.c restates them step by step so
// that the selection indices are bit-exact between GPU and oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsk_ptx.cuh"

namespace dsk {

// Canonical row reduction used by every distance in this file:
//   lane l accumulates fmaf(d,d,acc) over j = l, l+32, l+64, ... ; then xor-butterfly 16,8,4,2,1.
__device__ __forceinline__ float row_sqdist(const float* __restrict__ a, const float* __restrict__ b, int D,
                                            int lane) {
  float acc = 0.f;
  for (int j = lane; j < D; j += 32) {
    const float d = a[j] - b[j];
    acc = fmaf(d, d, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// PairwiseDistance(2).forward, /root/reference/model.py:13-18.  One warp per row.
__global__ void pairwise_distance_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int B, int D,
                                         float eps, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const float s = row_sqdist(x1 + static_cast<long>(row) * D, x2 + static_cast<long>(row) * D, D, lane);
  if (lane == 0) out[row] = sqrtf(s + eps);
}

// grad_x1 = grad_out * (x1-x2)/dist ; grad_x2 = -grad_x1.
__global__ void pairwise_distance_bwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                             const float* __restrict__ dist, const float* __restrict__ go, int B,
                                             int D, float* __restrict__ g1, float* __restrict__ g2) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long>(B) * D) return;
  const int row = i / D;
  const float g = go[row] * (x1[i] - x2[i]) / dist[row];
  if (g1) g1[i] = g;
  if (g2) g2[i] = -g;
}

// TripletMarginLoss.forward, /root/reference/model.py:27-33: d_p, d_n per row (one warp per row).
__global__ void triplet_dist_kernel(const float* __restrict__ a, const float* __restrict__ p,
                                    const float* __restrict__ n, int B, int D, float eps, float* __restrict__ d_p,
                                    float* __restrict__ d_n) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const long off = static_cast<long>(row) * D;
  const float sp = row_sqdist(a + off, p + off, D, lane);
  const float sn = row_sqdist(a + off, n + off, D, lane);
  if (lane == 0) {
    d_p[row] = sqrtf(sp + eps);
    d_n[row] = sqrtf(sn + eps);
  }
}

// loss = mean_i clamp(margin + d_p - d_n, min=0).  Single block, fixed reduction order.
__global__ void hinge_mean_kernel(const float* __restrict__ d_p, const float* __restrict__ d_n, int B, float margin,
                                  float* __restrict__ loss) {
  __shared__ float red[1024];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += fmaxf((margin + d_p[i]) - d_n[i], 0.f);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] / static_cast<float>(B);
}

// d loss / d{a,p,n}.  torch.clamp(min=0) passes the gradient where its input >= 0.
__global__ void triplet_loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ p,
                                        const float* __restrict__ n, const float* __restrict__ d_p,
                                        const float* __restrict__ d_n, const float* __restrict__ grad_loss, int B,
                                        int D, float margin, float* __restrict__ ga, float* __restrict__ gp,
                                        float* __restrict__ gn) {
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long>(B) * D) return;
  const int row = i / D;
  const float dp = d_p[row], dn = d_n[row];
  const float active = ((margin + dp) - dn >= 0.f) ? 1.f : 0.f;
  const float g = active * grad_loss[0] / static_cast<float>(B);
  const float up = g * (a[i] - p[i]) / dp;   // d loss / d a via d_p
  const float un = -g * (a[i] - n[i]) / dn;  // d loss / d a via -d_n
  ga[i] = up + un;
  gp[i] = -up;
  gn[i] = -un;
}

// idx = ascending { i : d_n[i] - d_p[i] < margin }  == np.where(mask == 1), train_triplet.py:251-262.
// Single block; ordered compaction with warp ballots + a block scan per 1024-element chunk.
__global__ void margin_select_kernel(const float* __restrict__ d_p, const float* __restrict__ d_n, int B,
                                     float margin, int64_t* __restrict__ idx, int32_t* __restrict__ count) {
  __shared__ int warp_cnt[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < B; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const bool sel = (i < B) && ((d_n[i] - d_p[i]) < margin);
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (sel) idx[off + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < nwarps; ++w) t += warp_cnt[w];
      base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) count[0] = base;
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                   const int32_t* __restrict__ count, int64_t row_elems, float* __restrict__ out) {
  const int j = blockIdx.x;
  if (j >= count[0]) return;
  const float* s = src + idx[j] * row_elems;
  float* o = out + static_cast<int64_t>(j) * row_elems;
  for (int64_t e = threadIdx.x; e < row_elems; e += blockDim.x) o[e] = s[e];
}

// ---------------------------------------------------------------------------------------------
// All-pairs distances (fp32, direct differences, sequential-in-d fmaf order) into a dense N x N
// matrix, 64x64 tile per block of 256 threads (4x4 outputs per thread).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
allpairs_sqdist_kernel(const float* __restrict__ E, int N, int D, float* __restrict__ S) {
  constexpr int TM = 64, TK = 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TM + 4];
  const int i0 = blockIdx.y * TM, j0 = blockIdx.x * TM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int k0 = 0; k0 < D; k0 += TK) {
    for (int t = threadIdx.x; t < TM * TK; t += 256) {
      const int r = t / TK, k = t % TK;
      As[k][r] = (i0 + r < N && k0 + k < D) ? E[static_cast<long>(i0 + r) * D + k0 + k] : 0.f;
      Bs[k][r] = (j0 + r < N && k0 + k < D) ? E[static_cast<long>(j0 + r) * D + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) av[r] = As[k][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) bv[c] = Bs[k][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float d = av[r] - bv[c];
          acc[r][c] = fmaf(d, d, acc[r][c]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ty * 4 + r, j = j0 + tx * 4 + c;
      if (i < N && j < N) S[static_cast<long>(i) * N + j] = acc[r][c];
    }
}

// Per row: the k smallest sqrt(S+eps) among columns with a different label, ties -> lower column.
// One warp per row; each pass extracts the lexicographic (value, index) minimum.
__global__ void topk_rows_kernel(const float* __restrict__ S, const int64_t* __restrict__ labels, int N, float eps,
                                 int k, int64_t* __restrict__ idx, float* __restrict__ val) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int lane = threadIdx.x & 31;
  const float* s = S + static_cast<long>(row) * N;
  const int64_t my_label = labels[row];
  float last_v = -1.f;
  int last_j = -1;
  for (int t = 0; t < k; ++t) {
    float bv = __int_as_float(0x7f800000);  // +inf
    int bj = 0x7fffffff;
    for (int j = lane; j < N; j += 32) {
      if (labels[j] == my_label) continue;
      const float v = sqrtf(s[j] + eps);
      // strictly after (last_v, last_j) in lexicographic order
      const bool after = (v > last_v) || (v == last_v && j > last_j);
      if (after && (v < bv || (v == bv && j < bj))) {
        bv = v;
        bj = j;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ov < bv || (ov == bv && oj < bj)) {
        bv = ov;
        bj = oj;
      }
    }
    if (lane == 0) {
      idx[static_cast<long>(row) * k + t] = (bj == 0x7fffffff) ? -1 : bj;
      val[static_cast<long>(row) * k + t] = bv;
    }
    last_v = bv;
    last_j = bj;
  }
}

}  // namespace dsk

// =================================================================================================
// Tensor-core all-pairs path: fp16 Gram on tcgen05 (conv_umma_kernel used as a plain GEMM, fp32 output), candidate
// selection from the approximate distances, EXACT fp32 refinement of the candidates in the canonical order of
// allpairs_sqdist_kernel (sequential-in-d fmaf), so the result is bit-identical to the exact path / the oracle.
// =================================================================================================
namespace dsk {

// E fp32 [N][D] -> 16-bit [Npad][D] (rows >= N zero) + squared norms of the ROUNDED rows. One block per row.
template <bool BF16>
__global__ void allpairs_prep_kernel(const float* __restrict__ E, int N, int D, uint16_t* __restrict__ E16,
                                     float* __restrict__ norms) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  float s = 0.f;
  for (int t = threadIdx.x; t < D; t += blockDim.x) {
    uint16_t h = 0;
    if (row < N) {
      h = to16<BF16>(E[static_cast<long>(row) * D + t]);
      const float f = from16<BF16>(h);
      s = fmaf(f, f, s);
    }
    E16[static_cast<long>(row) * D + t] = h;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    norms[row] = t;
  }
}

constexpr int kApCand = 16;  // candidates refined exactly per row (must be >= k; host enforces k <= 8)

__device__ __forceinline__ float exact_dist_seq(const float* __restrict__ a, const float* __restrict__ b, int D,
                                                float eps) {
  float acc = 0.f;
  for (int t = 0; t < D; ++t) {
    const float d = a[t] - b[t];
    acc = fmaf(d, d, acc);
  }
  return sqrtf(acc + eps);
}

// One warp per row.  G: fp32 Gram of the rounded rows [Npad][Npad]; approximate squared distance
// a_ij = n_i + n_j - 2 G_ij (exact squared distance of the ROUNDED rows up to fp32 accumulation).
// Exactness argument: rounding row e to 16 bit moves it by at most u*||e|| (u = unit roundoff), so for every pair
//   | a_ij - ||e_i - e_j||^2 | <= 2 d r + r^2 + slack,   r = u (||e_i|| + max_j ||e_j||),  d = ||e_i - e_j||.
// If the worst kept candidate's a exceeds (k-th exact distance)^2 by more than that bound, no discarded column can
// beat the k-th result and the refined top-k is the exact answer; otherwise the warp scans the whole row exactly.
__global__ void __launch_bounds__(256)
allpairs_select_refine_kernel(const float* __restrict__ E, const float* __restrict__ G, const float* __restrict__ norms,
                              const int64_t* __restrict__ labels, int N, int Npad, int D, float eps, int k, float u,
                              int64_t* __restrict__ idx, float* __restrict__ val) {
  __shared__ int cand_j[8][kApCand];
  __shared__ float cand_a[8][kApCand];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + w;
  if (row >= N) return;
  const float* g = G + static_cast<long>(row) * Npad;
  const float ni = norms[row];
  const int64_t my_label = labels[row];
  const float INF = __int_as_float(0x7f800000);
  // ---- candidates: the kApCand smallest approximate distances, extracted in lexicographic (a, j) order.
  // Rows of up to 1024 columns keep their 32 values per lane in registers; longer rows re-read G (L1/L2 hits).
  constexpr int kRegCols = 32;
  const bool in_regs = N <= 32 * kRegCols;
  float areg[kRegCols];
  if (in_regs) {
#pragma unroll
    for (int q = 0; q < kRegCols; ++q) {
      const int j = lane + 32 * q;
      areg[q] = (j < N && labels[j] != my_label) ? (ni + norms[j]) - 2.0f * g[j] : INF;
    }
  }
  float last_a = -INF;
  int last_j = -1;
  for (int t = 0; t < kApCand; ++t) {
    float ba = INF;
    int bj = 0x7fffffff;
    if (in_regs) {
#pragma unroll
      for (int q = 0; q < kRegCols; ++q) {
        const int j = lane + 32 * q;
        const float a = areg[q];
        const bool after = (a > last_a) || (a == last_a && j > last_j);
        if (a < INF && after && (a < ba || (a == ba && j < bj))) {
          ba = a;
          bj = j;
        }
      }
    } else {
      for (int j = lane; j < N; j += 32) {
        if (labels[j] == my_label) continue;
        const float a = (ni + norms[j]) - 2.0f * g[j];
        const bool after = (a > last_a) || (a == last_a && j > last_j);
        if (after && (a < ba || (a == ba && j < bj))) {
          ba = a;
          bj = j;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float oa = __shfl_xor_sync(0xffffffffu, ba, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (oa < ba || (oa == ba && oj < bj)) {
        ba = oa;
        bj = oj;
      }
    }
    if (lane == 0) {
      cand_j[w][t] = bj;
      cand_a[w][t] = ba;
    }
    last_a = ba;
    last_j = bj;
  }
  __syncwarp();
  // ---- exact refinement of the candidates (lane t < kApCand owns candidate t)
  float v = INF;
  int j = 0x7fffffff;
  if (lane < kApCand && cand_j[w][lane] != 0x7fffffff) {
    j = cand_j[w][lane];
    v = exact_dist_seq(E + static_cast<long>(row) * D, E + static_cast<long>(j) * D, D, eps);
  }
  // k-th smallest exact value among the candidates
  float kth = INF;
  {
    float lv = -1.f;
    int lj = -1;
    for (int t = 0; t < k; ++t) {
      float bv = INF;
      int bj = 0x7fffffff;
      const bool after = (v > lv) || (v == lv && j > lj);
      if (after) {
        bv = v;
        bj = j;
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ov < bv || (ov == bv && oj < bj)) {
          bv = ov;
          bj = oj;
        }
      }
      lv = bv;
      lj = bj;
      kth = bv;
    }
  }
  // safe iff every column NOT among the candidates has approximate squared distance >= kth^2 + bound
  const float worst_a = cand_a[w][kApCand - 1];
  float nmax = 0.f;
  for (int jj = lane; jj < N; jj += 32) nmax = fmaxf(nmax, norms[jj]);
  for (int o = 16; o > 0; o >>= 1) nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
  const float r = u * (sqrtf(ni) + sqrtf(nmax)) * 1.01f;
  const float dmax = sqrtf(fmaxf(worst_a, 0.f)) + r + 1e-3f;
  const float tol = 2.f * dmax * r + r * r + 1e-5f * (ni + nmax) + 1e-4f;
  const bool all_candidates = cand_j[w][kApCand - 1] == 0x7fffffff;  // fewer valid columns than candidates
  const bool safe = all_candidates || (worst_a >= (kth * kth - eps) + tol);
  if (!safe) {
    // ---- exact fallback: scan the whole row (rare); same arithmetic as the exact path
    float lv = -1.f;
    int lj = -1;
    for (int t = 0; t < k; ++t) {
      float bv = INF;
      int bj = 0x7fffffff;
      for (int jj = lane; jj < N; jj += 32) {
        if (labels[jj] == my_label) continue;
        const float vv = exact_dist_seq(E + static_cast<long>(row) * D, E + static_cast<long>(jj) * D, D, eps);
        const bool after = (vv > lv) || (vv == lv && jj > lj);
        if (after && (vv < bv || (vv == bv && jj < bj))) {
          bv = vv;
          bj = jj;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (ov < bv || (ov == bv && oj < bj)) {
          bv = ov;
          bj = oj;
        }
      }
      if (lane == 0) {
        idx[static_cast<long>(row) * k + t] = (bj == 0x7fffffff) ? -1 : bj;
        val[static_cast<long>(row) * k + t] = bv;
      }
      lv = bv;
      lj = bj;
    }
    return;
  }
  // ---- final top-k among the exactly refined candidates, (value, index) lexicographic
  float lv = -1.f;
  int lj = -1;
  for (int t = 0; t < k; ++t) {
    float bv = INF;
    int bj = 0x7fffffff;
    const bool after = (v > lv) || (v == lv && j > lj);
    if (after) {
      bv = v;
      bj = j;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ov < bv || (ov == bv && oj < bj)) {
        bv = ov;
        bj = oj;
      }
    }
    if (lane == 0) {
      idx[static_cast<long>(row) * k + t] = (bj == 0x7fffffff) ? -1 : bj;
      val[static_cast<long>(row) * k + t] = bv;
    }
    lv = bv;
    lj = bj;
  }
}

}  // namespace dsk
