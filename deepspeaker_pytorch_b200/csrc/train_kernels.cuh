// Layout converters and (below) the training-mode kernels: batch-statistics BatchNorm, backward passes.
#pragma once
#include "dsk_ptx.cuh"

namespace dsk {

// fp32 NCHW -> 16-bit NHWC (boundary / test helper).
template <bool BF16>
__global__ void nchw_to_nhwc16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int B, int C, int HW) {
  const long total = static_cast<long>(B) * C * HW;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    const long r = i / C;
    const int hw = r % HW;
    const int b = r / HW;
    out[i] = to16<BF16>(in[(static_cast<long>(b) * C + c) * HW + hw]);
  }
}

// 16-bit NHWC -> fp32 NCHW.
template <bool BF16>
__global__ void nhwc16_to_nchw_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int B, int C, int HW) {
  const long total = static_cast<long>(B) * C * HW;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = i % C;
    const long r = i / C;
    const int hw = r % HW;
    const int b = r / HW;
    out[(static_cast<long>(b) * C + c) * HW + hw] = from16<BF16>(in[i]);
  }
}

}  // namespace dsk
