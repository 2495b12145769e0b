// Stage-entry conv1 of the eval forward — 5x5 stride-2 pad-2, 1 -> 64 channels, + folded bn1 + clipped ReLU
// (/root/reference/model.py:94-97 used at :187-189) — on the tcgen05 tensor cores at fp32-level accuracy.
//
// The SIMT form of this layer is FP32-issue bound (1600 FMAs per output pixel, 21.5 us per batch-64 forward on
// B200, as long as a 12-GFLOP tensor-core conv).  Here a CTA builds the im2col operand of 128 output pixels in
// shared memory (one thread = one pixel = one 128-byte K-major SWIZZLE_128B row) and lets six UMMAs do the math:
//   x = x_hi + x_lo, w = w_hi + w_lo (each half a 16-bit float);  x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
//   A row  = [ x_hi(taps 0..24), 0 x 7 | x_lo(taps 0..24), 0 x 7 ]                      (K = 64)
//   B1 row = [ w_hi,             0 x 7 | w_hi,             0 x 7 ]  -> 4 UMMAs (K = 64): (x_hi + x_lo) * w_hi
//   B2 row = [ w_lo,             0 x 7 |        unused            ]  -> 2 UMMAs (K = 32):  x_hi * w_lo
// The dropped x_lo*w_lo term is 2^-22 relative (fp16 halves; 2^-16 with bf16 halves), far below the 16-bit
// rounding of the activation this kernel stores.  Accumulation is fp32 in TMEM.
// One tile per CTA, 128 threads, ~36 KB shared memory and 64 TMEM columns: several CTAs share an SM, so one CTA's
// operand build overlaps another's MMA / epilogue without any intra-CTA pipeline.
#pragma once
#include "dsk_ptx.cuh"

namespace dsk {

constexpr int kConv1ImgHalfs = 2 * 64 * 64;  // B1 | B2, each 64 rows x 64 halfs, pre-swizzled

// w [64][25] fp32 -> the pre-swizzled shared-memory image of B1 | B2 (byte offset of element (n, k) inside a
// matrix: n*128 + ((k/8) ^ (n&7))*16 + (k%8)*2 — the SWIZZLE_128B K-major layout TMA would have produced).
template <bool BF16>
__global__ void pack_conv1_umma_kernel(const float* __restrict__ w, uint16_t* __restrict__ img) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kConv1ImgHalfs; i += gridDim.x * blockDim.x) {
    const int mat = i / (64 * 64), n = (i / 64) % 64, k = i % 64;
    const int tap = k & 31;
    uint16_t v = 0;
    if (tap < 25 && (mat == 0 || k < 32)) {
      const float wf = w[n * 25 + tap];
      const uint16_t hi = to16<BF16>(wf);
      v = mat == 0 ? hi : to16<BF16>(wf - from16<BF16>(hi));
    }
    img[mat * 64 * 64 + n * 64 + (((k >> 3) ^ (n & 7)) << 3) + (k & 7)] = v;
  }
}

// x (B, T, 64) fp32; out: zero-padded NHWC 16-bit activation (rows n*(T/2+1)+h+1, 33 pixels per row, 64 channels).
// grid = B * (T/2) / 4 tiles of 4 output rows x 32 pixels; T/2 must be a multiple of 4.
template <bool BF16>
__global__ void __launch_bounds__(128)
conv1_umma_kernel(const float* __restrict__ x, const uint4* __restrict__ wimg, const float* __restrict__ scale,
                  const float* __restrict__ bias, uint16_t* __restrict__ out, int T, float clip_hi) {
  constexpr int WIN = 64, WOUT = 32, ROWS = 4, PATCH_ROWS = 2 * ROWS + 3, PATCH_W = WIN + 4;
  __shared__ __align__(1024) uint8_t sA[128 * 128];
  __shared__ __align__(1024) uint8_t sB[2 * 64 * 128];
  __shared__ float patch[PATCH_ROWS][PATCH_W];
  __shared__ float s_scale[64], s_bias[64];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_ptr;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int hout = T / 2, tiles_h = hout / ROWS;
  const int n = blockIdx.x / tiles_h, h0 = (blockIdx.x % tiles_h) * ROWS;
  pdl_launch_dependents();
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 64);
    tmem_relinquish();
  }
  if (tid == 32) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  // parameters are safe to read before the dependency wait; the input batch may come from the preceding kernel of
  // the stream and the output buffer is still read by the previous forward
#pragma unroll
  for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(sB)[tid + 128 * i] = wimg[tid + 128 * i];
  if (tid < 64) {
    s_scale[tid] = scale[tid];
    s_bias[tid] = bias[tid];
  }
  pdl_wait();
  {
    const float* xin = x + static_cast<long>(n) * T * WIN;
    constexpr int NEL = PATCH_ROWS * PATCH_W, NIT = (NEL + 127) / 128;
    float t[NIT];
#pragma unroll
    for (int j = 0; j < NIT; ++j) {  // all loads first (independent), then the stores
      const int i = tid + 128 * j;
      const int pr = i / PATCH_W, pc = i - pr * PATCH_W;
      const int ih = 2 * h0 - 2 + pr, iw = pc - 2;
      t[j] = (i < NEL && ih >= 0 && ih < T && iw >= 0 && iw < WIN) ? xin[ih * WIN + iw] : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < NIT; ++j) {
      const int i = tid + 128 * j;
      if (i < NEL) patch[i / PATCH_W][i % PATCH_W] = t[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;

  // ---- operand build: thread = output pixel (r, ow) of the tile = row tid of A
  {
    const int r = tid >> 5, ow = tid & 31;
    uint32_t hi[16], lo[16];  // 32 halfs each, taps 25..31 are zero
#pragma unroll
    for (int q = 0; q < 16; ++q) hi[q] = lo[q] = 0u;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int tap = i * 5 + j;
        const float v = patch[2 * r + i][2 * ow + j];
        const uint16_t h16 = to16<BF16>(v);
        const uint16_t l16 = to16<BF16>(v - from16<BF16>(h16));
        hi[tap >> 1] |= static_cast<uint32_t>(h16) << ((tap & 1) * 16);
        lo[tap >> 1] |= static_cast<uint32_t>(l16) << ((tap & 1) * 16);
      }
    }
    uint8_t* row = sA + tid * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      *reinterpret_cast<uint4*>(row + ((c ^ (tid & 7)) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
      *reinterpret_cast<uint4*>(row + (((c + 4) ^ (tid & 7)) << 4)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
    }
  }
  fence_proxy_async_smem();  // generic-proxy writes of A (and B) -> visible to the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (elect_one_sync()) {
      constexpr uint32_t idesc = umma_idesc_f16(128, 64, BF16);
      const uint64_t da = umma_desc_sw128(smem_u32(sA));
      const uint64_t db1 = umma_desc_sw128(smem_u32(sB));
      const uint64_t db2 = umma_desc_sw128(smem_u32(sB + 64 * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base, da + 2 * k, db1 + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 2; ++k) umma_f16(tmem_base, da + 2 * k, db2 + 2 * k, idesc, 1u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();

  // ---- epilogue: TMEM lane = pixel; folded BN, clip, 16-bit pack into the (now free) A tile, then cooperative
  // stores: 8 lanes write one 128-byte pixel row, so a warp store covers four full lines (a thread storing its own
  // row would touch 32 lines per instruction)
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + half * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = half * 32 + g * 8 + e;
        f[e] = fminf(fmaxf(fmaf(__uint_as_float(v[g * 8 + e]), s_scale[c], s_bias[c]), 0.0f), clip_hi);
      }
      *reinterpret_cast<uint4*>(sA + tid * 128 + (((half * 4 + g) ^ (tid & 7)) << 4)) =
          make_uint4(pack2<BF16>(f[0], f[1]), pack2<BF16>(f[2], f[3]), pack2<BF16>(f[4], f[5]), pack2<BF16>(f[6], f[7]));
    }
  }
  __syncthreads();
  {
    const int chunk = tid & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * 16 + (tid >> 3);
      const uint4 val = *reinterpret_cast<const uint4*>(sA + row * 128 + ((chunk ^ (row & 7)) << 4));
      const long pix = (static_cast<long>(n) * (hout + 1) + h0 + (row >> 5) + 1) * (WOUT + 1) + 1 + (row & 31);
      reinterpret_cast<uint4*>(out + pix * 64)[chunk] = val;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace dsk
