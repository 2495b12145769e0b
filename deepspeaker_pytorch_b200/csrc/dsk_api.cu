// libdsk.so — host side: engine handle, TMA descriptor construction, launch plans and the C ABI
// declared in include/dsk.h.  No torch types; raw device pointers + cudaStream_t only.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/dsk.h"
#include "conv_umma.cuh"
#include "loss_kernels.cuh"
#include "simt_kernels.cuh"
#include "train_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return fail(DSK_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define KERNEL_CHECK()                                                                              \
  do {                                                                                              \
    cudaError_t e_ = cudaGetLastError();                                                            \
    if (e_ != cudaSuccess)                                                                          \
      return fail(DSK_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 16-bit tensor map, 128B swizzle, zero OOB fill. dims/strides innermost first; strides[i] is the byte
// stride of dim i+1.
int make_tmap(CUtensorMap* out, bool bf16, const void* ptr, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(DSK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i < rank - 1) gs[i] = strides_bytes[i];
  }
  CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                  const_cast<void*>(ptr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string d;
    for (int i = 0; i < rank; ++i) d += std::to_string(dims[i]) + "/" + std::to_string(box[i]) + " ";
    return fail(DSK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims/box %s)", (int)r, rank,
                d.c_str());
  }
  return DSK_OK;
}

struct ConvLaunch {
  CUtensorMap tmA, tmB, tmOut, tmRes;
  dsk::ConvParams p;
  int n_tile = 0;
  int grid = 0;
};

struct LayerCfg {
  int cin, cout, ksize, stride;
};

// conv index i = 3*stage + {0: 5x5 s2 entry conv, 1,2: 3x3 block convs}
LayerCfg layer_cfg(int i) {
  static const int ch[4] = {64, 128, 256, 512};
  const int st = i / 3, k = i % 3;
  if (k == 0) return {st == 0 ? 1 : ch[st - 1], ch[st], 5, 2};
  return {ch[st], ch[st], 3, 1};
}

}  // namespace

struct dsk_handle_s {
  int device = 0;
  bool bf16 = false;
  int num_sms = 148;
  bool weights_loaded = false;
  int emb = 512;
  // packed parameters
  void* wpk[DSK_NUM_CONV] = {};       // 16-bit [tap][cout][cin]   (conv1: nullptr)
  void* wpk_dgrad[DSK_NUM_CONV] = {}; // 16-bit rotated/transposed for stride-1 dgrad
  float* conv1_w = nullptr;           // fp32 [64][25]
  float* scale[DSK_NUM_CONV] = {};    // folded eval BN
  float* bias[DSK_NUM_CONV] = {};
  float* fc_wq = nullptr;             // fp32 [E][w*512+c]
  const float* fc_b = nullptr;        // borrowed (valid until next load_weights)
  dsk_weights w = {};                 // borrowed parameter pointers (train mode reads gamma/beta, updates running stats)
  // workspace
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // plans keyed by (B, T)
  struct Plan {
    int B = 0, T = 0;
    std::vector<void*> act;  // 12 activation buffers (16-bit NHWC), index = conv index
    float* pooled = nullptr;
    float* fc_out = nullptr;
    std::vector<ConvLaunch> conv;  // index = conv index (0 unused)
  };
  std::map<std::pair<int, int>, Plan> plans;
  // optional per-launch timing (dsk_set_profiling): events recorded around every kernel of a forward
  bool profiling = false;
  std::vector<cudaEvent_t> events;
  int n_marks = 0;
};

namespace {

// Choose the pixel box (wt, hb, nb) with wt*hb*nb == 128 that wastes the fewest rows.
void choose_tile(int B, int Hout, int Wout, int& wt, int& hb, int& nb) {
  wt = Wout < 128 ? Wout : 128;
  const int rows = 128 / wt;  // h*n rows per tile
  long best = -1;
  hb = 1;
  nb = rows;
  for (int h = rows; h >= 1; h >>= 1) {
    const int n = rows / h;
    const long padded = static_cast<long>((Hout + h - 1) / h) * h * ((B + n - 1) / n) * n;
    if (best < 0 || padded < best) {
      best = padded;
      hb = h;
      nb = n;
    }
  }
}

template <int N_TILE, bool BF16>
int launch_conv_t(const ConvLaunch& L, cudaStream_t s) {
  auto kern = dsk::conv_umma_kernel<N_TILE, BF16>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dsk::ConvSmem<N_TILE>::kTotal));
    attr_set = true;
  }
  kern<<<L.grid, 256, dsk::ConvSmem<N_TILE>::kTotal, s>>>(L.tmA, L.tmB, L.tmOut, L.tmRes, L.p);
  KERNEL_CHECK();
  return DSK_OK;
}

int launch_conv(const dsk_handle_s* h, const ConvLaunch& L, cudaStream_t s) {
  if (h->bf16) {
    switch (L.n_tile) {
      case 64: return launch_conv_t<64, true>(L, s);
      case 128: return launch_conv_t<128, true>(L, s);
      case 256: return launch_conv_t<256, true>(L, s);
    }
  } else {
    switch (L.n_tile) {
      case 64: return launch_conv_t<64, false>(L, s);
      case 128: return launch_conv_t<128, false>(L, s);
      case 256: return launch_conv_t<256, false>(L, s);
    }
  }
  return fail(DSK_ERR_INVALID, "unsupported N tile %d", L.n_tile);
}

// Build descriptors + parameters for one fused conv layer on NHWC 16-bit tensors.
int build_conv(const dsk_handle_s* h, ConvLaunch* L, const void* in, const void* wpk, const float* scale,
               const float* bias, const void* res, void* out, int B, int Hin, int Win, int cin, int cout, int ksize,
               int stride, int flags, float clip_hi) {
  if (!((ksize == 3 && stride == 1) || (ksize == 5 && stride == 2)))
    return fail(DSK_ERR_INVALID, "conv: only 3x3 s1 p1 and 5x5 s2 p2 are supported (got k=%d s=%d)", ksize, stride);
  if (cin % 64 || cout % 64 || cin < 64 || cout < 64 || cout > 512)
    return fail(DSK_ERR_INVALID, "conv: cin/cout must be multiples of 64 and cout <= 512 (got %d, %d)", cin, cout);
  if (stride == 2 && ((Hin & 1) || (Win & 1)))
    return fail(DSK_ERR_INVALID, "conv: stride-2 input must have even H and W (got %d x %d)", Hin, Win);
  const int Hout = Hin / stride, Wout = Win / stride;
  if (Wout > 128 && Wout % 128) return fail(DSK_ERR_INVALID, "conv: unsupported output width %d", Wout);
  if (128 % (Wout < 128 ? Wout : 128)) return fail(DSK_ERR_INVALID, "conv: output width %d must divide 128", Wout);
  const bool bf = h->bf16;
  dsk::ConvParams& p = L->p;
  memset(&p, 0, sizeof(p));
  choose_tile(B, Hout, Wout, p.wt, p.hb, p.nb);
  p.tiles_w = (Wout + p.wt - 1) / p.wt;
  p.tiles_h = (Hout + p.hb - 1) / p.hb;
  p.tiles_n = (B + p.nb - 1) / p.nb;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  // N tile: the largest of {256,128,64} dividing cout that still gives every SM a tile; else the smallest.
  int n_tile = 64;
  for (int cand : {256, 128, 64}) {
    if (cout % cand) continue;
    n_tile = cand;
    if (static_cast<long>(tiles_m) * (cout / cand) >= h->num_sms) break;
  }
  L->n_tile = n_tile;
  p.tiles_c = cout / n_tile;
  p.taps = ksize * ksize;
  p.cin_chunks = cin / 64;
  p.cout = cout;
  p.flags = flags;
  p.clip_hi = clip_hi;
  p.scale = scale;
  p.bias = bias;
  for (int r = 0; r < ksize; ++r)
    for (int s = 0; s < ksize; ++s) {
      const int t = r * ksize + s;
      if (stride == 1) {
        p.tap_c[t] = 0;
        p.tap_dw[t] = static_cast<int8_t>(s - 1);
        p.tap_ph[t] = 0;
        p.tap_dh[t] = static_cast<int8_t>(r - 1);
      } else {
        // input col = 2*w - 2 + s  ->  (w2 = w + floor((s-2)/2), parity = s & 1); same for rows
        p.tap_c[t] = static_cast<int16_t>((s & 1) * cin);
        p.tap_dw[t] = static_cast<int8_t>((s - 2) >> 1);  // arithmetic shift = floor
        p.tap_ph[t] = static_cast<int8_t>(r & 1);
        p.tap_dh[t] = static_cast<int8_t>((r - 2) >> 1);
      }
    }
  const int num_tiles = tiles_m * p.tiles_c;
  L->grid = num_tiles < h->num_sms ? num_tiles : h->num_sms;

  // A: 5-D view (c, w2, ph, h2, n) of the NHWC input
  {
    uint64_t dims[5], str[4];
    if (stride == 1) {
      dims[0] = cin; dims[1] = Win; dims[2] = 1; dims[3] = Hin; dims[4] = B;
      str[0] = 2ull * cin; str[1] = 2ull * Win * cin; str[2] = 2ull * Win * cin; str[3] = 2ull * Hin * Win * cin;
    } else {
      dims[0] = 2ull * cin; dims[1] = Win / 2; dims[2] = 2; dims[3] = Hin / 2; dims[4] = B;
      str[0] = 4ull * cin; str[1] = 2ull * Win * cin; str[2] = 4ull * Win * cin; str[3] = 2ull * Hin * Win * cin;
    }
    uint32_t box[5] = {64, (uint32_t)p.wt, 1, (uint32_t)p.hb, (uint32_t)p.nb};
    int rc = make_tmap(&L->tmA, bf, in, 5, dims, str, box);
    if (rc) return rc;
  }
  {  // B: (cin, cout, taps)
    uint64_t dims[3] = {(uint64_t)cin, (uint64_t)cout, (uint64_t)p.taps};
    uint64_t str[2] = {2ull * cin, 2ull * cin * cout};
    uint32_t box[3] = {64, (uint32_t)n_tile, 1};
    int rc = make_tmap(&L->tmB, bf, wpk, 3, dims, str, box);
    if (rc) return rc;
  }
  {  // out / residual: (cout, Wout, Hout, B)
    uint64_t dims[4] = {(uint64_t)cout, (uint64_t)Wout, (uint64_t)Hout, (uint64_t)B};
    uint64_t str[3] = {2ull * cout, 2ull * Wout * cout, 2ull * Hout * Wout * cout};
    uint32_t box[4] = {64, (uint32_t)p.wt, (uint32_t)p.hb, (uint32_t)p.nb};
    int rc = make_tmap(&L->tmOut, bf, out, 4, dims, str, box);
    if (rc) return rc;
    rc = make_tmap(&L->tmRes, bf, (flags & dsk::CONV_RESIDUAL) ? res : out, 4, dims, str, box);
    if (rc) return rc;
  }
  return DSK_OK;
}

int check_handle(dsk_handle h) {
  if (!h) return fail(DSK_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  return DSK_OK;
}

template <typename T>
int dev_alloc(T** p, size_t n) {
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  return DSK_OK;
}

// per-utterance element counts of the 12 activation tensors at time length T
void act_shape(int i, int T, int& H, int& W, int& C) {
  const int st = i / 3;
  H = T >> (st + 1);
  W = 64 >> (st + 1);
  C = 64 << st;
}

int get_plan(dsk_handle h, int B, int T, dsk_handle_s::Plan** out) {
  auto key = std::make_pair(B, T);
  auto it = h->plans.find(key);
  if (it != h->plans.end()) {
    *out = &it->second;
    return DSK_OK;
  }
  // A new shape: (re)allocate the workspace for it alone and rebuild all plans lazily.
  size_t bytes = 0;
  size_t off[DSK_NUM_CONV];
  for (int i = 0; i < DSK_NUM_CONV; ++i) {
    int H, W, C;
    act_shape(i, T, H, W, C);
    off[i] = bytes;
    bytes += ((static_cast<size_t>(B) * H * W * C * 2 + 1023) / 1024) * 1024;
  }
  const size_t off_pooled = bytes;
  bytes += static_cast<size_t>(B) * 2048 * 4;
  const size_t off_fc = bytes;
  bytes += static_cast<size_t>(B) * h->emb * 4;
  if (bytes > h->ws_bytes) {
    // drop cached plans: their descriptors point into the old workspace
    h->plans.clear();
    if (h->ws) CUDA_TRY(cudaFree(h->ws));
    h->ws = nullptr;
    h->ws_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->ws, bytes));
    h->ws_bytes = bytes;
  } else {
    // one workspace is shared by all shapes; descriptors of other shapes stay valid (same base pointer)
  }
  dsk_handle_s::Plan pl;
  pl.B = B;
  pl.T = T;
  pl.act.resize(DSK_NUM_CONV);
  uint8_t* base = static_cast<uint8_t*>(h->ws);
  for (int i = 0; i < DSK_NUM_CONV; ++i) pl.act[i] = base + off[i];
  pl.pooled = reinterpret_cast<float*>(base + off_pooled);
  pl.fc_out = reinterpret_cast<float*>(base + off_fc);
  pl.conv.resize(DSK_NUM_CONV);
  for (int i = 1; i < DSK_NUM_CONV; ++i) {
    const LayerCfg c = layer_cfg(i);
    int Hi, Wi, Ci;
    act_shape(i - 1, T, Hi, Wi, Ci);  // input of conv i is activation i-1
    const int k = i % 3;
    const void* res = (k == 2) ? pl.act[i - 2] : nullptr;  // block output adds the block input
    const int flags = dsk::CONV_CLIP | (k == 2 ? dsk::CONV_RESIDUAL : 0);
    int rc = build_conv(h, &pl.conv[i], pl.act[i - 1], h->wpk[i], h->scale[i], h->bias[i], res, pl.act[i], B, Hi, Wi,
                        c.cin, c.cout, c.ksize, c.stride, flags, 20.0f);
    if (rc) return rc;
  }
  auto ins = h->plans.emplace(key, std::move(pl));
  *out = &ins.first->second;
  return DSK_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* dsk_last_error(void) { return g_err.c_str(); }
int32_t dsk_version(void) { return 100; }

int32_t dsk_create(dsk_handle* out, int32_t device, int32_t operand) {
  if (!out) return fail(DSK_ERR_INVALID, "dsk_create: out is null");
  if (operand != DSK_F16 && operand != DSK_BF16) return fail(DSK_ERR_INVALID, "dsk_create: bad operand type %d", operand);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(DSK_ERR_ARCH, "dsk_create: device %d is sm_%d%d; this library contains sm_100a code only", device,
                prop.major, prop.minor);
  dsk_handle h = new dsk_handle_s();
  h->device = device;
  h->bf16 = operand == DSK_BF16;
  h->num_sms = prop.multiProcessorCount;
  *out = h;
  return DSK_OK;
}

int32_t dsk_destroy(dsk_handle h) {
  if (!h) return DSK_OK;
  cudaSetDevice(h->device);
  for (int i = 0; i < DSK_NUM_CONV; ++i) {
    cudaFree(h->wpk[i]);
    cudaFree(h->wpk_dgrad[i]);
    cudaFree(h->scale[i]);
    cudaFree(h->bias[i]);
  }
  cudaFree(h->conv1_w);
  cudaFree(h->fc_wq);
  cudaFree(h->ws);
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  delete h;
  return DSK_OK;
}

int32_t dsk_load_weights(dsk_handle h, const dsk_weights* w, void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!w) return fail(DSK_ERR_INVALID, "dsk_load_weights: null weights");
  if (w->embedding_size <= 0 || w->embedding_size % 64)
    return fail(DSK_ERR_INVALID, "dsk_load_weights: embedding_size must be a positive multiple of 64");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->weights_loaded && h->emb != w->embedding_size) return fail(DSK_ERR_INVALID, "embedding_size changed");
  h->emb = w->embedding_size;
  for (int i = 0; i < DSK_NUM_CONV; ++i) {
    const LayerCfg c = layer_cfg(i);
    const int taps = c.ksize * c.ksize;
    const long n = static_cast<long>(c.cout) * c.cin * taps;
    if (!w->conv_w[i] || !w->bn_gamma[i] || !w->bn_beta[i] || !w->bn_running_mean[i] || !w->bn_running_var[i])
      return fail(DSK_ERR_INVALID, "dsk_load_weights: null parameter pointer for conv/bn %d", i);
    if (!h->scale[i]) {
      rc = dev_alloc(&h->scale[i], c.cout);
      if (rc) return rc;
      rc = dev_alloc(&h->bias[i], c.cout);
      if (rc) return rc;
    }
    dsk::bn_fold_kernel<<<(c.cout + 127) / 128, 128, 0, s>>>(w->bn_gamma[i], w->bn_beta[i], w->bn_running_mean[i],
                                                             w->bn_running_var[i], 1e-5f, h->scale[i], h->bias[i],
                                                             c.cout);
    KERNEL_CHECK();
    if (i == 0) {
      if (!h->conv1_w) {
        rc = dev_alloc(&h->conv1_w, 64 * 25);
        if (rc) return rc;
      }
      CUDA_TRY(cudaMemcpyAsync(h->conv1_w, w->conv_w[0], 64 * 25 * sizeof(float), cudaMemcpyDeviceToDevice, s));
      continue;
    }
    if (!h->wpk[i]) {
      CUDA_TRY(cudaMalloc(&h->wpk[i], n * 2));
      if (c.stride == 1) CUDA_TRY(cudaMalloc(&h->wpk_dgrad[i], n * 2));
    }
    const int blocks = static_cast<int>((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
    if (h->bf16) {
      dsk::pack_conv_weight_kernel<true><<<blocks, 256, 0, s>>>(w->conv_w[i], (uint16_t*)h->wpk[i], c.cout, c.cin, taps);
      if (c.stride == 1)
        dsk::pack_conv_weight_dgrad_kernel<true><<<blocks, 256, 0, s>>>(w->conv_w[i], (uint16_t*)h->wpk_dgrad[i], c.cout, c.cin, taps);
    } else {
      dsk::pack_conv_weight_kernel<false><<<blocks, 256, 0, s>>>(w->conv_w[i], (uint16_t*)h->wpk[i], c.cout, c.cin, taps);
      if (c.stride == 1)
        dsk::pack_conv_weight_dgrad_kernel<false><<<blocks, 256, 0, s>>>(w->conv_w[i], (uint16_t*)h->wpk_dgrad[i], c.cout, c.cin, taps);
    }
    KERNEL_CHECK();
  }
  if (!w->fc_w || !w->fc_b) return fail(DSK_ERR_INVALID, "dsk_load_weights: null fc pointer");
  if (!h->fc_wq) {
    rc = dev_alloc(&h->fc_wq, static_cast<size_t>(h->emb) * 2048);
    if (rc) return rc;
  }
  dsk::pack_fc_weight_kernel<<<1024, 256, 0, s>>>(w->fc_w, h->fc_wq, h->emb, 512, 4);
  KERNEL_CHECK();
  h->fc_b = w->fc_b;
  h->w = *w;
  h->weights_loaded = true;
  return DSK_OK;
}

int32_t dsk_rescnn_forward(dsk_handle h, const float* x, int32_t B, int32_t T, float* emb, int32_t mode,
                           void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!h->weights_loaded) return fail(DSK_ERR_STATE, "dsk_rescnn_forward: call dsk_load_weights first");
  if (!x || !emb || B <= 0) return fail(DSK_ERR_INVALID, "dsk_rescnn_forward: bad arguments");
  if (T < 16 || T % 16) return fail(DSK_ERR_INVALID, "dsk_rescnn_forward: T must be a positive multiple of 16 (got %d)", T);
  if (mode != DSK_EVAL) return fail(DSK_ERR_INVALID, "dsk_rescnn_forward: use dsk_rescnn_forward_train for batch-statistics BN");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dsk_handle_s::Plan* pl;
  rc = get_plan(h, B, T, &pl);
  if (rc) return rc;
  h->n_marks = 0;
  auto mark = [&]() {
    if (!h->profiling) return;
    if (h->n_marks >= static_cast<int>(h->events.size())) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      h->events.push_back(e);
    }
    cudaEventRecord(h->events[h->n_marks++], s);
  };
  mark();
  // conv1 (+bn1 +clip)
  {
    const int hout = T / 2;
    const int blocks = B * ((hout + 3) / 4);
    if (h->bf16)
      dsk::conv1_kernel<true><<<blocks, 256, 0, s>>>(x, h->conv1_w, h->scale[0], h->bias[0], (uint16_t*)pl->act[0], T, 1, 20.0f);
    else
      dsk::conv1_kernel<false><<<blocks, 256, 0, s>>>(x, h->conv1_w, h->scale[0], h->bias[0], (uint16_t*)pl->act[0], T, 1, 20.0f);
    KERNEL_CHECK();
    mark();
  }
  for (int i = 1; i < DSK_NUM_CONV; ++i) {
    rc = launch_conv(h, pl->conv[i], s);
    if (rc) return rc;
    mark();
  }
  // tail
  {
    const int H4 = T / 16, WC = 4 * 512;
    if (h->bf16)
      dsk::pool_time_kernel<true><<<B, 256, 0, s>>>((const uint16_t*)pl->act[11], pl->pooled, H4, WC);
    else
      dsk::pool_time_kernel<false><<<B, 256, 0, s>>>((const uint16_t*)pl->act[11], pl->pooled, H4, WC);
    KERNEL_CHECK();
    mark();
    static bool fc_attr = false;
    const int fc_smem = 8 * 2048 * 4;
    if (!fc_attr) {
      CUDA_TRY(cudaFuncSetAttribute(dsk::fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fc_smem));
      fc_attr = true;
    }
    dim3 g((B + 7) / 8, h->emb / 64);
    dsk::fc_kernel<<<g, 256, fc_smem, s>>>(pl->pooled, h->fc_wq, h->fc_b, pl->fc_out, B, 2048, h->emb);
    KERNEL_CHECK();
    mark();
    dsk::l2norm_kernel<<<B, 128, 0, s>>>(pl->fc_out, emb, nullptr, h->emb, 10.0f);
    KERNEL_CHECK();
    mark();
  }
  return DSK_OK;
}

int32_t dsk_set_profiling(dsk_handle h, int32_t enable) {
  if (!h) return fail(DSK_ERR_INVALID, "null handle");
  h->profiling = enable != 0;
  h->n_marks = 0;
  return DSK_OK;
}

int32_t dsk_get_launch_times(dsk_handle h, float* ms_out, int32_t cap, int32_t* n_out) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!ms_out || !n_out) return fail(DSK_ERR_INVALID, "dsk_get_launch_times: null output");
  const int n = h->n_marks > 0 ? h->n_marks - 1 : 0;
  if (n > cap) return fail(DSK_ERR_INVALID, "dsk_get_launch_times: need room for %d values", n);
  if (n > 0) CUDA_TRY(cudaEventSynchronize(h->events[h->n_marks - 1]));
  for (int i = 0; i < n; ++i) CUDA_TRY(cudaEventElapsedTime(&ms_out[i], h->events[i], h->events[i + 1]));
  *n_out = n;
  return DSK_OK;
}

int32_t dsk_conv2d_nhwc(dsk_handle h, const void* in, const void* w_packed, const float* scale, const float* bias,
                        const void* res, void* out, int32_t B, int32_t Hin, int32_t Win, int32_t cin, int32_t cout,
                        int32_t ksize, int32_t stride, int32_t flags, float clip_hi, void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  if (!in || !w_packed || !out) return fail(DSK_ERR_INVALID, "dsk_conv2d_nhwc: null pointer");
  if ((flags & dsk::CONV_RESIDUAL) && !res) return fail(DSK_ERR_INVALID, "dsk_conv2d_nhwc: residual flag without res");
  ConvLaunch L;
  rc = build_conv(h, &L, in, w_packed, scale, bias, res, out, B, Hin, Win, cin, cout, ksize, stride, flags, clip_hi);
  if (rc) return rc;
  return launch_conv(h, L, static_cast<cudaStream_t>(stream));
}

int32_t dsk_pack_conv_weight(dsk_handle h, const float* w_oihw, void* w_packed, int32_t cout, int32_t cin,
                             int32_t ksize, void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  const long n = static_cast<long>(cout) * cin * ksize * ksize;
  const int blocks = static_cast<int>((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->bf16)
    dsk::pack_conv_weight_kernel<true><<<blocks, 256, 0, s>>>(w_oihw, (uint16_t*)w_packed, cout, cin, ksize * ksize);
  else
    dsk::pack_conv_weight_kernel<false><<<blocks, 256, 0, s>>>(w_oihw, (uint16_t*)w_packed, cout, cin, ksize * ksize);
  KERNEL_CHECK();
  return DSK_OK;
}

int32_t dsk_nchw_f32_to_nhwc16(dsk_handle h, const float* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W,
                               void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  const long n = static_cast<long>(B) * C * H * W;
  const int blocks = static_cast<int>((n + 255) / 256 < 65535 ? (n + 255) / 256 : 65535);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->bf16)
    dsk::nchw_to_nhwc16_kernel<true><<<blocks, 256, 0, s>>>(in, (uint16_t*)out, B, C, H * W);
  else
    dsk::nchw_to_nhwc16_kernel<false><<<blocks, 256, 0, s>>>(in, (uint16_t*)out, B, C, H * W);
  KERNEL_CHECK();
  return DSK_OK;
}

int32_t dsk_nhwc16_to_nchw_f32(dsk_handle h, const void* in, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                               void* stream) {
  int rc = check_handle(h);
  if (rc) return rc;
  const long n = static_cast<long>(B) * C * H * W;
  const int blocks = static_cast<int>((n + 255) / 256 < 65535 ? (n + 255) / 256 : 65535);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->bf16)
    dsk::nhwc16_to_nchw_kernel<true><<<blocks, 256, 0, s>>>((const uint16_t*)in, out, B, C, H * W);
  else
    dsk::nhwc16_to_nchw_kernel<false><<<blocks, 256, 0, s>>>((const uint16_t*)in, out, B, C, H * W);
  KERNEL_CHECK();
  return DSK_OK;
}

// ---- distances / loss / selection -----------------------------------------------------------------
static inline float pd_eps(int D) { return static_cast<float>(1e-4 / static_cast<double>(D)); }

int32_t dsk_pairwise_distance(const float* x1, const float* x2, int32_t B, int32_t D, float* out, void* stream) {
  if (!x1 || !x2 || !out || B <= 0 || D <= 0) return fail(DSK_ERR_INVALID, "dsk_pairwise_distance: bad arguments");
  dsk::pairwise_distance_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x1, x2, B, D, pd_eps(D), out);
  KERNEL_CHECK();
  return DSK_OK;
}

int32_t dsk_pairwise_distance_bwd(const float* x1, const float* x2, const float* dist, const float* grad_out,
                                  int32_t B, int32_t D, float* grad_x1, float* grad_x2, void* stream) {
  if (!x1 || !x2 || !dist || !grad_out || B <= 0 || D <= 0)
    return fail(DSK_ERR_INVALID, "dsk_pairwise_distance_bwd: bad arguments");
  const long n = static_cast<long>(B) * D;
  dsk::pairwise_distance_bwd_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x1, x2, dist, grad_out, B, D, grad_x1, grad_x2);
  KERNEL_CHECK();
  return DSK_OK;
}

int32_t dsk_triplet_loss(const float* a, const float* p, const float* n, int32_t B, int32_t D, float margin,
                         float* loss, float* d_p, float* d_n, void* stream) {
  if (!a || !p || !n || !loss || !d_p || !d_n || B <= 0 || D <= 0)
    return fail(DSK_ERR_INVALID, "dsk_triplet_loss: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dsk::triplet_dist_kernel<<<(B + 7) / 8, 256, 0, s>>>(a, p, n, B, D, pd_eps(D), d_p, d_n);
  KERNEL_CHECK();
  dsk::hinge_mean_kernel<<<1, 1024, 0, s>>>(d_p, d_n, B, margin, loss);
  KERNEL_CHECK();
  return DSK_OK;
}

int32_t dsk_triplet_loss_bwd(const float* a, const float* p, const float* n, const float* d_p, const float* d_n,
                             const float* grad_loss, int32_t B, int32_t D, float margin, float* ga, float* gp,
                             float* gn, void* stream) {
  if (!a || !p || !n || !d_p || !d_n || !grad_loss || !ga || !gp || !gn || B <= 0 || D <= 0)
    return fail(DSK_ERR_INVALID, "dsk_triplet_loss_bwd: bad arguments");
  const long cnt = static_cast<long>(B) * D;
  dsk::triplet_loss_bwd_kernel<<<(cnt + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, p, n, d_p, d_n, grad_loss, B, D, margin, ga, gp, gn);
  KERNEL_CHECK();
  return DSK_OK;
}

int32_t dsk_margin_select(const float* d_p, const float* d_n, int32_t B, float margin, int64_t* idx,
                          int32_t* count, void* stream) {
  if (!d_p || !d_n || !idx || !count || B <= 0) return fail(DSK_ERR_INVALID, "dsk_margin_select: bad arguments");
  dsk::margin_select_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(d_p, d_n, B, margin, idx, count);
  KERNEL_CHECK();
  return DSK_OK;
}

int32_t dsk_gather_rows(const float* src, const int64_t* idx, const int32_t* count, int32_t max_rows,
                        int64_t row_elems, float* out, void* stream) {
  if (!src || !idx || !count || !out || max_rows <= 0 || row_elems <= 0)
    return fail(DSK_ERR_INVALID, "dsk_gather_rows: bad arguments");
  dsk::gather_rows_kernel<<<max_rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, idx, count, row_elems, out);
  KERNEL_CHECK();
  return DSK_OK;
}

int32_t dsk_allpairs_topk(const float* E, const int64_t* labels, int32_t N, int32_t D, int32_t k, int64_t* idx,
                          float* val, void* stream) {
  if (!E || !labels || !idx || !val || N <= 0 || D <= 0 || k <= 0 || k > N)
    return fail(DSK_ERR_INVALID, "dsk_allpairs_topk: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* S = nullptr;
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&S), static_cast<size_t>(N) * N * sizeof(float), s));
  dim3 g((N + 63) / 64, (N + 63) / 64);
  dsk::allpairs_sqdist_kernel<<<g, 256, 0, s>>>(E, N, D, S);
  KERNEL_CHECK();
  dsk::topk_rows_kernel<<<(N + 7) / 8, 256, 0, s>>>(S, labels, N, pd_eps(D), k, idx, val);
  KERNEL_CHECK();
  CUDA_TRY(cudaFreeAsync(S, s));
  return DSK_OK;
}

}  // extern "C"
