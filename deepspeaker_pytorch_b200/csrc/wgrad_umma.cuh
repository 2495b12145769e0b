// Weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces cuDNN's convolution-backward-filter reached by autograd for every nn.Conv2d of the
// ResCNN (/root/reference/model.py:58,61,98,102,106; backward triggered at train_triplet.py:223).
//
//   dW[tap][co][ci] = sum over pixels  G[pix][co] * X[pix shifted by tap][ci]
//
// The reduction dimension is the pixel index, so both operands are read from channel-major
// ("transposed") 16-bit copies written by the BatchNorm kernels: GT [co][n][h][w], XT [ci][n](plane)[h][w].
// One K-step = 64 pixels: a TMA box {kw, kh, 1, kn, rows} lands in shared memory as rows x 128 bytes,
// i.e. the same K-major SWIZZLE_128B operand tile the forward conv uses; the tap shift and the zero
// padding again come from TMA coordinates / out-of-bounds fill.  Work item = (tap, co tile of 128,
// ci tile of N_TILE, K split); partial tiles are reduced with fp32 atomics into dW.
#pragma once
#include "conv_umma.cuh"

namespace dsk {

struct WgradParams {
  int taps, co_tiles, ci_tiles, ksplit;
  int cout, cin;            // real channel counts (rows beyond are TMA zero fill and are not written)
  int chunks_w, chunks_h, chunks_n;  // K-chunk grid; chunk = box {kw, kh, kn}
  int kw, kh, kn;
  int8_t tap_dw[kMaxTaps];
  int8_t tap_dh[kMaxTaps];
  int8_t tap_plane[kMaxTaps];
  float* dw;                // fp32 [tap][cout][cin], pre-zeroed
};

template <int N_TILE>
struct WgradSmem {
  static constexpr int kStages = (N_TILE == 64) ? 6 : (N_TILE == 128 ? 5 : 4);
  static constexpr int kBTileBytes = N_TILE * 128;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kTotal = kStages * kStageBytes + 256 + 1024;
};

template <int N_TILE, bool BF16>
__global__ void __launch_bounds__(256, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  using S = WgradSmem<N_TILE>;
  constexpr int kStages = S::kStages;
  constexpr int kTmemCols = 2 * N_TILE;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kATileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_chunks = p.chunks_w * p.chunks_h * p.chunks_n;
  const int num_items = p.taps * p.co_tiles * p.ci_tiles * p.ksplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // item -> (tap, co tile, ci tile, K range).  K split fastest: CTAs running together share the weight tile's
  // operands' neighbourhood in L2 and finish a (tap, co, ci) tile at about the same time.
  auto decode = [&](int item, int& tap, int& co0, int& ci0, int& k_begin, int& k_end) {
    const int ks = item % p.ksplit;
    int r = item / p.ksplit;
    const int cit = r % p.ci_tiles;
    r /= p.ci_tiles;
    const int cot = r % p.co_tiles;
    tap = r / p.co_tiles;
    co0 = cot * kTileM;
    ci0 = cit * N_TILE;
    const int per = (total_chunks + p.ksplit - 1) / p.ksplit;
    k_begin = ks * per;
    k_end = k_begin + per < total_chunks ? k_begin + per : total_chunks;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tap, co0, ci0, kb, ke;
        decode(item, tap, co0, ci0, kb, ke);
        const int dw = p.tap_dw[tap], dh = p.tap_dh[tap], plane = p.tap_plane[tap];
        for (int k = kb; k < ke; ++k) {
          const int cw = k % p.chunks_w;
          const int r = k / p.chunks_w;
          const int chh = r % p.chunks_h;
          const int cn = r / p.chunks_h;
          const int w0 = cw * p.kw, h0 = chh * p.kh, n0 = cn * p.kn;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_5d(smem_a + stage * kATileBytes, &tmG, &full_bar[stage], w0, h0, 0, n0, co0);
          tma_load_5d(smem_b + stage * S::kBTileBytes, &tmX, &full_bar[stage], w0 + dw, h0 + dh, plane, n0, ci0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, N_TILE, BF16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tap, co0, ci0, kb, ke;
        decode(item, tap, co0, ci0, kb, ke);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * N_TILE;
        for (int k = kb; k < ke; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * kATileBytes));
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * S::kBTileBytes));
#pragma unroll
          for (int kk = 0; kk < kKStep / 16; ++kk)
            umma_f16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (k > kb || kk > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int tap, co0, ci0, kb, ke;
      decode(item, tap, co0, ci0, kb, ke);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int co = co0 + row;
      float* dst = p.dw + (static_cast<long>(tap) * p.cout + co) * p.cin + ci0;
      const bool live = (co < p.cout) && (ke > kb);
#pragma unroll 1
      for (int j = 0; j < N_TILE / 32; ++j) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * N_TILE + j * 32, v);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (ci0 + j * 32 + e < p.cin) atomicAdd(dst + j * 32 + e, __uint_as_float(v[e]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace dsk
