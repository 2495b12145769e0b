// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces the cuDNN calls the reference reaches through nn.Conv2d for the 3x3 s1 p1 convs inside
// BasicBlock (/root/reference/model.py:47-50,58,61 used at :69,73) and the 5x5 s2 p2 stage-entry
// convs conv2..conv4 (/root/reference/model.py:98,102,106 used at :192,197,202), with the
// BatchNorm affine (:59,62,99,103,107), the residual add (:79) and the clipped ReLU (:36-39)
// folded into the epilogue.
//
// Data layout: activations are NHWC, 16-bit (fp16 or bf16).  GEMM view: M = output pixels,
// N = output channels, K = taps x input channels.  One K-step = one filter tap x 64 input channels:
//   A tile  (128 pixels x 64 ch)   one TMA box {64, wt, 1, hb, nb} of the (tap-shifted) input; zero padding
//                                  comes from TMA out-of-bounds fill
//   B tile  (N_TILE cout x 64 ch)  one TMA box of the [tap][cout][cin] weight tensor
// both land in 128B-swizzled shared memory and feed tcgen05.mma (M=128, N=N_TILE, K=16) x 4.
// Accumulators live in TMEM (double-buffered), the epilogue reads them with tcgen05.ld, applies
// scale/bias (+residual) (+clip), converts to 16 bit and TMA-stores NHWC.
#pragma once
#include "dsk_ptx.cuh"

namespace dsk {

constexpr int kMaxTaps = 25;
constexpr int kTileM = 128;
constexpr int kKStep = 64;               // 16-bit elements per K-step = 128 bytes = one swizzle row
constexpr int kATileBytes = kTileM * 128;  // 16 KB

enum ConvFlags : int {
  CONV_RESIDUAL = 1,  // add residual tile (same shape as output) before the clip
  CONV_CLIP = 2,      // clamp to [0, clip_hi]
};

struct ConvParams {
  // tile geometry
  int tiles_w, tiles_h, tiles_n, tiles_c;  // output tile grid: width, height, batch, cout
  int wt, hb, nb;                          // pixels per tile along w, h, batch (wt*hb*nb == 128)
  int taps, cin_chunks;
  int cout;
  int flags;
  float clip_hi;
  const float* scale;  // [cout] or nullptr (=1)
  const float* bias;   // [cout] or nullptr (=0)
  // output placement in the 5-D output view (c, w, ph, h, n): channel base and parity row.  Plain NHWC
  // outputs use (0, 0); the stride-2 data-gradient writes parity class (ph, pw) with out_c_base = pw*C.
  int out_c_base, out_ph;
  // per-tap source offsets in the 5-D input view (c, w2, ph, h2, n) and weight slice index
  int16_t tap_c[kMaxTaps];
  int8_t tap_w[kMaxTaps];
  int8_t tap_dw[kMaxTaps];
  int8_t tap_ph[kMaxTaps];
  int8_t tap_dh[kMaxTaps];
};

constexpr int kConvThreads = 384;  // 4 control warps (TMA, MMA, TMEM alloc, residual prefetch) + 8 epilogue warps

template <int N_TILE>
struct ConvSmem {
  static constexpr int kStages = (N_TILE == 64) ? 6 : (N_TILE == 128 ? 4 : 3);
  static constexpr int kBTileBytes = N_TILE * 128;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStagingBytes = 2 * kATileBytes;  // two 128-row x 128-byte output chunks
  static constexpr int kResBytes = 2 * kATileBytes;      // residual tiles, prefetched by their own warp
  static constexpr int kScaleBiasBytes = 2 * 512 * 4;
  static constexpr int kBarBytes = 256;
  static constexpr int kAccStages = (N_TILE == 256) ? 2 : 4;  // TMEM accumulators (<= 512 columns)
  static constexpr int kTotal =
      kStages * kStageBytes + kStagingBytes + kResBytes + kScaleBiasBytes + kBarBytes + 1024;
};

// OUT_F32: the epilogue stores fp32 (no residual / clip): used for the pre-BatchNorm conv output of the
// train-mode forward, which must not be rounded to 16 bit before the batch statistics are applied.
template <int N_TILE, bool BF16, bool OUT_F32 = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                 const ConvParams p) {
  using S = ConvSmem<N_TILE>;
  constexpr int kStages = S::kStages;
  constexpr int kAcc = S::kAccStages;
  constexpr int kTmemCols = kAcc * N_TILE;
  constexpr int kChunks = N_TILE / 64;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kATileBytes;
  uint8_t* smem_stg = smem + kStages * S::kStageBytes;
  uint8_t* smem_res = smem_stg + S::kStagingBytes;
  float* smem_scale = reinterpret_cast<float*>(smem_res + S::kResBytes);
  float* smem_bias = smem_scale + 512;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_scale) + S::kScaleBiasBytes);
  uint64_t* full_bar = bars;                  // [kStages]
  uint64_t* empty_bar = bars + kStages;       // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;   // [kAcc]
  uint64_t* tmem_empty = tmem_full + kAcc;    // [kAcc]
  uint64_t* res_full = tmem_empty + kAcc;     // [2]
  uint64_t* res_empty = res_full + 2;         // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  const int ksteps = p.taps * p.cin_chunks;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  const int num_tiles = tiles_m * p.tiles_c;

  // ---- one-time setup -------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (p.flags & CONV_RESIDUAL) tma_prefetch_desc(&tmRes);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);  // one arrive per epilogue warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&res_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < p.cout && i < 512; i += blockDim.x) {
    smem_scale[i] = p.scale ? p.scale[i] : 1.0f;
    smem_bias[i] = p.bias ? p.bias[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above touched only parameters; activations of the previous kernel are read/written below

  // tile -> coordinates. Tile order: cout tile fastest so CTAs running concurrently share the A tile in L2.
  auto decode = [&](int tile, int& c0, int& w0, int& h0, int& n0) {
    int ct = tile % p.tiles_c;
    int mt = tile / p.tiles_c;
    int wt_i = mt % p.tiles_w;
    int r = mt / p.tiles_w;
    int ht_i = r % p.tiles_h;
    int nt_i = r / p.tiles_h;
    c0 = ct * N_TILE;
    w0 = wt_i * p.wt;
    h0 = ht_i * p.hb;
    n0 = nt_i * p.nb;
  };

  if (warp == 0) {
    // ===================== TMA producer (warp-converged loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int c0, w0, h0, n0;
      decode(tile, c0, w0, h0, n0);
      for (int ks = 0; ks < ksteps; ++ks) {
        const int tap = ks / p.cin_chunks;
        const int ch = ks - tap * p.cin_chunks;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_5d(smem_a + stage * kATileBytes, &tmA, &full_bar[stage], p.tap_c[tap] + ch * kKStep,
                      w0 + p.tap_dw[tap], p.tap_ph[tap], h0 + p.tap_dh[tap], n0);
          tma_load_3d(smem_b + stage * S::kBTileBytes, &tmB, &full_bar[stage], ch * kKStep, c0, p.tap_w[tap]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-converged loop, one elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc_f16(kTileM, N_TILE, BF16);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * N_TILE;
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * kATileBytes));
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * S::kBTileBytes));
#pragma unroll
          for (int k = 0; k < kKStep / 16; ++k) {
            // advancing 16 elements (32 B) along K inside the swizzle atom = +2 in the (addr>>4) field
            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (ks > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (ks == ksteps - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == kAcc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp == 3) {
    // ===================== residual prefetcher: one 128 x 64 tile per output chunk, two buffers =====================
    if (!OUT_F32 && (p.flags & CONV_RESIDUAL)) {
      int rb = 0;
      uint32_t rph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int c0, w0, h0, n0;
        decode(tile, c0, w0, h0, n0);
        for (int j = 0; j < kChunks; ++j) {
          mbar_wait(&res_empty[rb], rph ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&res_full[rb], kATileBytes);
            tma_load_5d(smem_res + rb * kATileBytes, &tmRes, &res_full[rb], p.out_c_base + c0 + j * 64, w0, p.out_ph, h0, n0);
          }
          __syncwarp();
          if (++rb == 2) {
            rb = 0;
            rph ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps; thread = one output pixel x 32 of the 64 channels of a chunk =========
    const int ew = (warp - 4) & 3;          // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    const int half = (warp - 4) >> 2;       // which 32 accumulator columns of each 64-column group
    const int row = ew * 32 + lane;         // row of the 128-row tile
    const int etid = threadIdx.x - 128;
    const bool has_res = (p.flags & CONV_RESIDUAL) != 0;
    const bool do_clip = (p.flags & CONV_CLIP) != 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int rb = 0;
    uint32_t rph = 0;
    int buf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int c0, w0, h0, n0;
      decode(tile, c0, w0, h0, n0);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < kChunks; ++j) {
        uint32_t v[32];
        const float* sc = smem_scale + c0 + j * 64 + half * 32;
        const float* bi = smem_bias + c0 + j * 64 + half * 32;
        if constexpr (OUT_F32) {
          // fp32 output: the two column halves fill one 128-row x 32-float staging buffer each (both buffers per group)
          if (etid == 0) tma_store_wait_read<0>();
          named_bar_sync(1, 256);
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * N_TILE + j * 64 + half * 32, v);
          tmem_ld_wait();
          uint8_t* my_row = smem_stg + half * kATileBytes + row * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint4 o;
            o.x = __float_as_uint(fmaf(__uint_as_float(v[q * 4 + 0]), sc[q * 4 + 0], bi[q * 4 + 0]));
            o.y = __float_as_uint(fmaf(__uint_as_float(v[q * 4 + 1]), sc[q * 4 + 1], bi[q * 4 + 1]));
            o.z = __float_as_uint(fmaf(__uint_as_float(v[q * 4 + 2]), sc[q * 4 + 2], bi[q * 4 + 2]));
            o.w = __float_as_uint(fmaf(__uint_as_float(v[q * 4 + 3]), sc[q * 4 + 3], bi[q * 4 + 3]));
            *reinterpret_cast<uint4*>(my_row + ((q ^ (row & 7)) << 4)) = o;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 256);
          if (etid == 0) {
            tma_store_5d(&tmOut, smem_stg, p.out_c_base + c0 + j * 64, w0, p.out_ph, h0, n0);
            tma_store_5d(&tmOut, smem_stg + kATileBytes, p.out_c_base + c0 + j * 64 + 32, w0, p.out_ph, h0, n0);
            tma_store_commit();
          }
        } else {
          uint8_t* stg = smem_stg + buf * kATileBytes;
          // staging buffer `buf` was last used two chunks ago: its TMA store must have finished reading smem
          if (etid == 0) tma_store_wait_read<1>();
          named_bar_sync(1, 256);
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * N_TILE + j * 64 + half * 32, v);
          tmem_ld_wait();
          if (has_res) mbar_wait(&res_full[rb], rph);
          const uint8_t* res_row = smem_res + rb * kATileBytes + row * 128;
          uint8_t* my_row = stg + row * 128;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {  // 4 x 16-byte chunks = this thread's 32 channels
            float f[8];
            const float4 s0 = *reinterpret_cast<const float4*>(sc + qq * 8);
            const float4 s1 = *reinterpret_cast<const float4*>(sc + qq * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(bi + qq * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bi + qq * 8 + 4);
            const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float biv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaf(__uint_as_float(v[qq * 8 + e]), scv[e], biv[e]);
            const int chunk16 = ((half * 4 + qq) ^ (row & 7)) << 4;
            if (has_res) {
              const uint4 r = *reinterpret_cast<const uint4*>(res_row + chunk16);
              float2 t;
              t = unpack2<BF16>(r.x); f[0] += t.x; f[1] += t.y;
              t = unpack2<BF16>(r.y); f[2] += t.x; f[3] += t.y;
              t = unpack2<BF16>(r.z); f[4] += t.x; f[5] += t.y;
              t = unpack2<BF16>(r.w); f[6] += t.x; f[7] += t.y;
            }
            if (do_clip) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fminf(fmaxf(f[e], 0.0f), p.clip_hi);
            }
            uint4 o;
            o.x = pack2<BF16>(f[0], f[1]);
            o.y = pack2<BF16>(f[2], f[3]);
            o.z = pack2<BF16>(f[4], f[5]);
            o.w = pack2<BF16>(f[6], f[7]);
            *reinterpret_cast<uint4*>(my_row + chunk16) = o;
          }
          fence_proxy_async_smem();
          if (has_res) {  // residual buffer consumed: hand it back to the prefetcher
            __syncwarp();
            if (lane == 0) mbar_arrive(&res_empty[rb]);
            if (++rb == 2) {
              rb = 0;
              rph ^= 1;
            }
          }
          named_bar_sync(1, 256);
          if (etid == 0) {
            tma_store_5d(&tmOut, stg, p.out_c_base + c0 + j * 64, w0, p.out_ph, h0, n0);
            tma_store_commit();
          }
          buf ^= 1;
        }
      }
      // accumulator fully read: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == kAcc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (etid == 0) tma_store_wait_all<0>();
  }

  // ---- teardown ---------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace dsk
