// Weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces cuDNN's convolution-backward-filter reached by autograd for every nn.Conv2d of the
// ResCNN (/root/reference/model.py:58,61,98,102,106; backward triggered at train_triplet.py:223).
//
//   dW[tap][co][ci] = sum over pixels  G[pix][co] * X[pix shifted by tap][ci]
//
// The reduction index is the pixel, and both tensors are NHWC (channels contiguous), so both operands are
// "MN-major" for the tensor core: a TMA box of 128 pixels x 64 channels lands in shared memory as 128 rows of
// 128 bytes (SWIZZLE_128B) and is consumed as one 64-wide operand atom (descriptor: leading byte offset = atom
// stride, stride byte offset = 1024 B between 8-pixel groups, major bits = MN).  The tap shift and the zero padding
// are TMA coordinates on the outer (w, h) dims exactly as in the forward conv, so no transposed copies exist.
//
// Work item = (tap, 128 output channels, N_TILE input channels, K split).  Every K split writes its partial tile with
// plain stores into its OWN slice dw[ks][tap][cout][cin]; unpack_wgrad_kernel then adds the slices in fixed order, so
// the weight gradient is bit-reproducible from run to run (no atomics, no pre-zeroing).  Layers with 64 output
// channels use the swapped form (M = two taps x 64 input channels, N = 64 output channels) so that the MMA still
// has M = 128.
#pragma once
#include "conv_umma.cuh"

namespace dsk {

constexpr int kAtomBytes = 128 * 128;  // 128 pixels x 64 channels x 2 B

struct WgradParams {
  int swapped;              // 0: A = G (2 atoms of 64 co), B = X;  1: A = X for two taps, B = G (cout == 64)
  int taps, co_tiles, ci_tiles, ksplit;
  int cout, cin;
  int chunks_w, chunks_h, chunks_n;  // K-chunk grid; chunk = pixel box {wt, hb, nb} of 128 pixels
  int wt, hb, nb;
  int16_t tap_c[kMaxTaps];  // X view: channel offset (parity column), w/h offsets, parity row
  int8_t tap_dw[kMaxTaps];
  int8_t tap_ph[kMaxTaps];
  int8_t tap_dh[kMaxTaps];
  float* dw;                // fp32 [ksplit][tap][cout][cin]: one slice per K split, every element written exactly once
  long slice_elems;         // taps * cout * cin
};

template <int N_TILE>
struct WgradSmem {
  static constexpr int kBAtoms = N_TILE / 64;
  static constexpr int kStageBytes = (2 + kBAtoms) * kAtomBytes;
  static constexpr int kStages = (N_TILE == 64) ? 4 : 3;
  static constexpr int kTotal = kStages * kStageBytes + 256 + 1024;
};

// MN-major SWIZZLE_128B operand descriptor (see header comment)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(kAtomBytes >> 4) << 16;  // leading byte offset: next 64-channel atom
  d |= static_cast<uint64_t>(1024 >> 4) << 32;        // stride byte offset: next group of 8 pixel rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int N_TILE, bool BF16>
__global__ void __launch_bounds__(256, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  using S = WgradSmem<N_TILE>;
  constexpr int kStages = S::kStages;
  constexpr int kBAtoms = S::kBAtoms;
  constexpr int kTmemCols = 2 * N_TILE;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_chunks = p.chunks_w * p.chunks_h * p.chunks_n;
  const int tap_units = p.swapped ? (p.taps + 1) / 2 : p.taps;
  const int num_items = tap_units * p.co_tiles * p.ci_tiles * p.ksplit;

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

  // item -> (tap unit, co tile, ci tile, K range); K split fastest
  auto decode = [&](int item, int& tu, int& co0, int& ci0, int& k_begin, int& k_end) -> int {
    const int ks = item % p.ksplit;
    int r = item / p.ksplit;
    const int cit = r % p.ci_tiles;
    r /= p.ci_tiles;
    const int cot = r % p.co_tiles;
    tu = r / p.co_tiles;
    co0 = cot * kTileM;
    ci0 = cit * N_TILE;
    const int per = (total_chunks + p.ksplit - 1) / p.ksplit;
    k_begin = ks * per;
    k_end = k_begin + per < total_chunks ? k_begin + per : total_chunks;
    return ks;
  };

  if (warp == 0) {
    // TMA producer: warp-converged loop, one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int tu, co0, ci0, kb, ke;
      decode(item, tu, co0, ci0, kb, ke);
      for (int k = kb; k < ke; ++k) {
        const int cw = k % p.chunks_w;
        const int r = k / p.chunks_w;
        const int chh = r % p.chunks_h;
        const int cn = r / p.chunks_h;
        const int w0 = cw * p.wt, h0 = chh * p.hb, n0 = cn * p.nb;
        uint8_t* sa = smem + stage * S::kStageBytes;
        uint8_t* sb = sa + 2 * kAtomBytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          if (!p.swapped) {
            tma_load_5d(sa, &tmG, &full_bar[stage], co0, w0, 0, h0, n0);
            tma_load_5d(sa + kAtomBytes, &tmG, &full_bar[stage], co0 + 64, w0, 0, h0, n0);
#pragma unroll
            for (int j = 0; j < kBAtoms; ++j)
              tma_load_5d(sb + j * kAtomBytes, &tmX, &full_bar[stage], p.tap_c[tu] + ci0 + 64 * j, w0 + p.tap_dw[tu],
                          p.tap_ph[tu], h0 + p.tap_dh[tu], n0);
          } else {
            const int t0 = 2 * tu, t1 = (2 * tu + 1 < p.taps) ? 2 * tu + 1 : 2 * tu;
            tma_load_5d(sa, &tmX, &full_bar[stage], p.tap_c[t0], w0 + p.tap_dw[t0], p.tap_ph[t0], h0 + p.tap_dh[t0], n0);
            tma_load_5d(sa + kAtomBytes, &tmX, &full_bar[stage], p.tap_c[t1], w0 + p.tap_dw[t1], p.tap_ph[t1],
                        h0 + p.tap_dh[t1], n0);
            tma_load_5d(sb, &tmG, &full_bar[stage], 0, w0, 0, h0, n0);
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: fp32 accumulate, both operands MN-major (bits 15 and 16)
    constexpr uint32_t idesc = umma_idesc_f16(kTileM, N_TILE, BF16) | (1u << 15) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int tu, co0, ci0, kb, ke;
      decode(item, tu, co0, ci0, kb, ke);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * N_TILE;
      for (int k = kb; k < ke; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t da = umma_desc_mn_sw128(smem_u32(smem + stage * S::kStageBytes));
          const uint64_t db = umma_desc_mn_sw128(smem_u32(smem + stage * S::kStageBytes + 2 * kAtomBytes));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // 128 pixels = 8 x K16; one K16 step = 16 rows x 128 B = +128 in the addr field
            umma_f16(d_tmem, da + 128 * kk, db + 128 * kk, idesc, (k > kb || kk > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (k == ke - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (ke <= kb) {  // empty K range (split rounding): still hand an (ignored) accumulator to the epilogue
        if (elect_one_sync()) umma_commit(&tmem_full[acc]);
        __syncwarp();
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int tu, co0, ci0, kb, ke;
      const int ks = decode(item, tu, co0, ci0, kb, ke);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      // destination of D[row][col] inside this K split's slice; an empty K range (split rounding) stores zeros
      float* dst = p.dw + static_cast<long>(ks) * p.slice_elems;
      const bool empty = ke <= kb;
      bool live = true;
      if (!p.swapped) {
        const int co = co0 + row;
        live = co < p.cout;
        dst += (static_cast<long>(tu) * p.cout + co) * p.cin + ci0;
      } else {
        const int tap = 2 * tu + (row >> 6);
        live = tap < p.taps;
        dst += static_cast<long>(tap) * p.cout * p.cin + (row & 63);  // [tap][co = col][ci = row % 64]
      }
#pragma unroll 1
      for (int j = 0; j < N_TILE / 32; ++j) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * N_TILE + j * 32, v);
        tmem_ld_wait();
        if (empty) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
        if (live) {
          if (!p.swapped) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              *reinterpret_cast<uint4*>(dst + j * 32 + e * 4) = make_uint4(v[e * 4], v[e * 4 + 1], v[e * 4 + 2], v[e * 4 + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) dst[static_cast<long>(j * 32 + e) * p.cin] = __uint_as_float(v[e]);
          }
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
