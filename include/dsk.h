/*
 * dsk.h — C ABI of the B200-native Deep Speaker hot path (libdsk.so).
 *
 * The reference (qqueing/DeepSpeaker-pytorch) is 100 % Python and reaches the GPU only through
 * torch.nn library calls; it has no FFI of its own.  Each entry point below therefore cites the
 * reference *Python* call site whose arithmetic it replaces (file:line in /root/reference).
 * A maintainer binds these with ctypes; the host-side mirror of the
 * reference's classes lives in deepspeaker_pytorch_b200/model.py.
 *
 * Conventions
 *  - every pointer is a raw CUDA device pointer owned by the caller (PyTorch); the library
 *    borrows it for the duration of the call and never frees it;
 *  - every function is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *    synchronises the device and never reads results on the host;
 *  - return value: 0 on success, negative dsk_status on error; dsk_last_error() gives the
 *    message of the last failure on the calling thread;
 *  - no exceptions cross this boundary, there is no CPU fallback.
 */
#ifndef DSK_H_
#define DSK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  DSK_OK = 0,
  DSK_ERR_INVALID = -1, /* bad argument / unsupported shape */
  DSK_ERR_CUDA = -2,    /* a CUDA runtime / driver call failed */
  DSK_ERR_STATE = -3,   /* call sequence error (e.g. forward before load_weights) */
  DSK_ERR_ARCH = -4     /* device is not sm_100 */
} dsk_status;

/* 16-bit tensor-core operand format (fp32 accumulation either way) */
typedef enum { DSK_F16 = 0, DSK_BF16 = 1 } dsk_operand_t;

/* BatchNorm behaviour of a forward: running statistics (eval) or batch statistics (train) */
typedef enum { DSK_EVAL = 0, DSK_TRAIN = 1 } dsk_mode_t;

#define DSK_NUM_CONV 12 /* conv1, layer1.0.conv1, layer1.0.conv2, conv2, layer2.0.conv1, ... layer4.0.conv2 */

/* Parameters of DeepSpeakerModel.model (fp32, PyTorch layouts), /root/reference/model.py:91-112,162-164.
 * conv_w[i]: OIHW.  i = 3*stage + {0: convK (5x5 s2), 1: layerK.0.conv1, 2: layerK.0.conv2 (3x3)}.
 * bn_*[i]  : the BatchNorm2d that follows conv i (bnK, layerK.0.bn1, layerK.0.bn2).
 * fc_w     : (embedding_size, 2048) with column index c*4 + w (model.py:164,208-209). */
typedef struct {
  const float* conv_w[DSK_NUM_CONV];
  const float* bn_gamma[DSK_NUM_CONV];
  const float* bn_beta[DSK_NUM_CONV];
  float* bn_running_mean[DSK_NUM_CONV]; /* read in eval; updated in place in train (momentum 0.1) */
  float* bn_running_var[DSK_NUM_CONV];
  const float* fc_w;
  const float* fc_b;
  int32_t embedding_size; /* 512 */
} dsk_weights;

/* Gradients, same layouts as dsk_weights (fp32, written not accumulated). */
typedef struct {
  float* conv_w[DSK_NUM_CONV];
  float* bn_gamma[DSK_NUM_CONV];
  float* bn_beta[DSK_NUM_CONV];
  float* fc_w;
  float* fc_b;
} dsk_grads;

typedef struct dsk_handle_s* dsk_handle;
typedef struct dsk_train_ctx_s* dsk_train_ctx; /* what one train-mode forward saved for its backward */
typedef struct dsk_pipeline_s* dsk_pipeline;   /* serving pipeline: copy streams + compute lanes around one weight image */

const char* dsk_last_error(void);
int32_t dsk_version(void);

/* Per-module engine state (repacked weights, folded BN, workspace, TMA descriptors).
 * Created at DeepSpeakerModel.cuda()/first call, released in __del__ (model.py:153-167). */
int32_t dsk_create(dsk_handle* out, int32_t device, int32_t operand /* dsk_operand_t */);
int32_t dsk_destroy(dsk_handle h);

/* Repack conv weights to [tap][cout][cin] 16-bit, fold eval BatchNorm to scale/bias, reorder fc.
 * Must be called after every parameter update (the Python shim tracks parameter versions). */
int32_t dsk_load_weights(dsk_handle h, const dsk_weights* w, void* stream);
/* The same for a handle that is about to TRAIN (every optimizer step changes every parameter, so this runs once per
 * step, before the forwards): rebuilds only what dsk_rescnn_forward_train / dsk_rescnn_backward read - the forward and
 * data-gradient operand images of the eleven tensor-core convs (one kernel launch), conv1's filter and the reordered fc
 * weight - and skips the eval-only work (BatchNorm folding, conv1's split image, the plane-major 5x5 images).
 * dsk_rescnn_forward (eval) then fails with DSK_ERR_STATE until dsk_load_weights is called again. */
int32_t dsk_load_weights_train(dsk_handle h, const dsk_weights* w, void* stream);
/* Serving with several forwards in flight (one handle + activation workspace per compute stream): `h` borrows the
 * packed weights / folded BN of `src` instead of holding its own copy, so all lanes read one 21 MB weight image (it
 * has to stay L2-resident: the convs re-read it per tile).  `h` follows later dsk_load_weights(src) calls at its next
 * forward; it is inference-only, must not be given weights of its own, and `src` must outlive it. */
int32_t dsk_share_weights(dsk_handle h, dsk_handle src);

/* DeepSpeakerModel.forward (/root/reference/model.py:185-218), BN in eval mode
 * (train_triplet.py:332,347): x (B,1,T,64) fp32 contiguous -> emb (B,E) fp32 with ||emb||=10.
 * T must be a multiple of 16.
 * Asynchronous on `stream`, with one exception that synchronises the DEVICE: the first forward after
 * dsk_load_weights (the folded BN affine is copied to the host and baked into the conv kernels' parameter block).
 * The first call of a new (B, T) shape rebuilds the plan and re-zeroes the padded activation workspace with a
 * memset ORDERED ON `stream` (after the forwards this handle still has in flight there); if the workspace has to
 * grow, cudaFree synchronises the device.  One handle must only be driven from one stream at a time (use one handle
 * per compute lane, dsk_share_weights).  From the second call of a shape on, the 15 kernels are replayed as one CUDA graph whose first / last
 * nodes are re-pointed at x / emb; inside a caller's own stream capture the plain launches are recorded instead (warm the
 * shape up before capturing). */
int32_t dsk_rescnn_forward(dsk_handle h, const float* x, int32_t B, int32_t T, float* emb, int32_t mode,
                           void* stream);

/* DeepSpeakerModel.forward with the module in train mode (train_triplet.py:203,215): BatchNorm uses the batch
 * statistics of THIS call (the reference forwards a, p and n separately, so statistics are per call) and updates
 * running_mean / running_var in place (momentum 0.1, unbiased variance).  Saves activations in *ctx for
 * dsk_rescnn_backward; x must stay alive until then.  Contexts are pooled inside the handle. */
int32_t dsk_rescnn_forward_train(dsk_handle h, const float* x, int32_t B, int32_t T, float* emb, dsk_train_ctx* ctx,
                                 void* stream);
/* Several train-mode forwards of one step in flight at once (DeepSpeakerModel.forward_triplet runs the anchor / positive /
 * negative forwards of train_triplet.py:215 on three streams so that the HBM-bound BatchNorm passes of one overlap the
 * tensor-core convs of another): with `on` != 0 a train forward only RECORDS its batch statistics in the context and
 * leaves running_mean / running_var alone; dsk_train_ctx_commit_stats then applies that forward's momentum update
 * (all 12 layers, one launch).  Committing the contexts in the order of the reference's sequential calls (a, p, n) on
 * one stream gives bit-identical running statistics.  A context must be committed before its backward consumes it. */
int32_t dsk_set_defer_running_stats(dsk_handle h, int32_t on);
int32_t dsk_train_ctx_commit_stats(dsk_handle h, dsk_train_ctx ctx, void* stream);
/* Backward of that forward (what loss.backward() triggers, train_triplet.py:223): grad_emb (B,E) fp32 ->
 * gradients of every conv / BN / fc parameter, written (not accumulated) into `grads`.  Consumes the context. */
int32_t dsk_rescnn_backward(dsk_handle h, dsk_train_ctx ctx, const float* grad_emb, const dsk_grads* grads,
                            void* stream);
/* Debug / test read-back of what a train-mode forward saved: which = 0 the pre-BatchNorm conv output (fp32),
 * 1 the post-activation tensor (16-bit), of conv layer `layer` (0..11), converted to fp32 NCHW. */
int32_t dsk_train_ctx_read(dsk_handle h, dsk_train_ctx ctx, int32_t which, int32_t layer, float* out_nchw, void* stream);
/* Return an unused context to the pool (forward without backward, e.g. under no_grad). */
int32_t dsk_train_ctx_release(dsk_handle h, dsk_train_ctx ctx);
/* fp16 operands: inside the backward the 16-bit gradient tensors are multiplied by a power of two S and every parameter
 * gradient is divided by it again.  scale = 0 (default): S is chosen PER BACKWARD ON THE DEVICE from the largest incoming
 * gradient, S = 2^floor(log2(512 / max|dL/d(fc output)|)) - no host synchronisation, follows the loss as it shrinks during
 * training; scale > 0: that fixed value.  bf16 operands: 1 unless set. */
int32_t dsk_set_loss_scale(dsk_handle h, float scale);

/* Device timing of the next dsk_rescnn_forward calls (which then launch kernel by kernel, not as a graph).
 * enable = 1: CUDA events are recorded on `stream` around every kernel of the forward (order: conv1, the 11
 * tensor-core convs in network order, pool, fc, l2norm); enable = 2: only at the section boundaries
 * conv1 | 11 tensor-core convs | tail (3 values), so the conv chain runs back to back as in production.  dsk_get_launch_times waits for the last profiled forward (the only call in this
 * library that blocks the host) and returns its per-launch milliseconds. Used by bench.py's roofline. */
int32_t dsk_set_profiling(dsk_handle h, int32_t enable);
int32_t dsk_get_launch_times(dsk_handle h, float* ms_out, int32_t cap, int32_t* n_out);

/* One fused conv layer on NHWC 16-bit tensors: out = clip(conv(in, w)*scale + bias (+res)).
 * Building block of dsk_rescnn_forward, exported for unit tests against F.conv2d.
 * ksize/stride in {(3,1),(5,2)}; cin, cout multiples of 64; flags: 1 = add residual, 2 = clip to [0,clip_hi].
 * w_packed comes from dsk_pack_conv_weight. */
int32_t dsk_conv2d_nhwc(dsk_handle h, const void* in, const void* w_packed, const float* scale, const float* bias,
                        const void* res, void* out, int32_t B, int32_t Hin, int32_t Win, int32_t cin, int32_t cout,
                        int32_t ksize, int32_t stride, int32_t flags, float clip_hi, void* stream);
/* Backward building blocks of dsk_rescnn_backward, exported for unit tests against torch autograd.
 * dgrad: g_in (B,Hin,Win,cin) = d/d(input) of conv(input, w) given G (B,Hout,Wout,cout) (+ res, stride 1 only).
 * wgrad: dw (cout,cin,k,k) fp32 = mult * d/d(w) given G and the conv input X (B,Hin,Win,cin). */
int32_t dsk_conv2d_dgrad_nhwc(dsk_handle h, const void* G, const float* w_oihw, const void* res, void* gin, int32_t B,
                              int32_t Hin, int32_t Win, int32_t cin, int32_t cout, int32_t ksize, int32_t stride,
                              void* stream);
int32_t dsk_conv2d_wgrad_nhwc(dsk_handle h, const void* G, const void* X, float* dw_oihw, int32_t B, int32_t Hin,
                              int32_t Win, int32_t cin, int32_t cout, int32_t ksize, int32_t stride, float mult,
                              void* stream);
/* y = clip(BatchNorm_train(raw) (+res), 0, 20) on an NHWC tensor viewed as [M][C] (raw fp32, y/res 16-bit); writes the
 * batch mean / rstd and updates the running statistics (model.py:59,62 in train mode + :79-80). */
int32_t dsk_bn_act_train_forward(dsk_handle h, const float* raw, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, const void* res, void* y, float* mean,
                                 float* rstd, int64_t M, int32_t C, void* stream);
/* its backward: gy -> G (w.r.t. raw), gres (w.r.t. res, may be NULL), dgamma, dbeta (x inv_scale). */
int32_t dsk_bn_act_train_backward(dsk_handle h, const void* gy, const void* y, const float* raw, const float* gamma,
                                  const float* mean, const float* rstd, void* G, void* gres, float* dgamma,
                                  float* dbeta, int64_t M, int32_t C, float inv_scale, void* stream);
/* The eval forward's convs on the zero-padded NHWC layout with halo reuse (csrc/conv3x3_halo.cuh), exported for
 * unit tests.  Standard padded tensor: [dsk_padded_positions(N,H,W)][C] 16-bit, pixel (n,h,w) at position
 * (n*(H+1)+h+1)*(W+1) + w+1, every other position zero.  Parity-planar tensor of an (N,H,W,C) image (H, W even):
 * four planes p = (h&1)*2 + (w&1), plane p = standard padded tensor of geometry (N, H/2, W/2) holding pixel
 * (n, h>>1, w>>1), planes dsk_padded_positions(N,H/2,W/2) positions apart.  W <= 34.
 * dsk_conv3x3_padded: 3x3 s1 p1, C -> C, standard in; out standard (out_planar = 0) or parity-planar (1).
 * dsk_conv5x5s2_planar: 5x5 s2 p2, parity-planar input of the (N, 2*Hout, 2*Wout, cin) image -> standard padded out. */
int32_t dsk_conv3x3_padded(dsk_handle h, const void* in, const void* w_packed, const float* scale, const float* bias,
                           const void* res, void* out, int32_t N, int32_t H, int32_t W, int32_t C, int32_t flags,
                           float clip_hi, int32_t out_planar, void* stream);
int32_t dsk_conv5x5s2_planar(dsk_handle h, const void* in_planar, const float* w_oihw, const float* scale,
                             const float* bias, void* out, int32_t N, int32_t Hout, int32_t Wout, int32_t cin,
                             int32_t cout, int32_t flags, float clip_hi, void* stream);
int64_t dsk_padded_positions(int32_t N, int32_t H, int32_t W);
/* Debug: device buffer of 3*512 int64 that dsk_conv3x3_padded fills with clock64 stamps of CTA 0
 * (producer / MMA / epilogue roles); NULL switches tracing off. */
int32_t dsk_debug_set_trace(dsk_handle h, void* device_buffer);
int32_t dsk_pack_conv_weight(dsk_handle h, const float* w_oihw, void* w_packed, int32_t cout, int32_t cin,
                             int32_t ksize, void* stream);
/* fp32 NCHW <-> 16-bit NHWC converters (test helpers; also used at the boundary for C>1 inputs) */
int32_t dsk_nchw_f32_to_nhwc16(dsk_handle h, const float* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W,
                               void* stream);
int32_t dsk_nhwc16_to_nchw_f32(dsk_handle h, const void* in, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                               void* stream);

/* PairwiseDistance(p=2).forward (/root/reference/model.py:8-18): out[i] = sqrt(sum_j (x1-x2)^2 + 1e-4/D). */
int32_t dsk_pairwise_distance(const float* x1, const float* x2, int32_t B, int32_t D, float* out, void* stream);
/* d(out)/d(x1) and d(out)/d(x2) given grad_out (B,), dist (B,) from the forward. Either grad pointer may be NULL. */
int32_t dsk_pairwise_distance_bwd(const float* x1, const float* x2, const float* dist, const float* grad_out,
                                  int32_t B, int32_t D, float* grad_x1, float* grad_x2, void* stream);

/* TripletMarginLoss(margin).forward (/root/reference/model.py:19-33):
 * loss = mean(clamp(margin + d_p - d_n, 0)). Writes loss (1,), d_p (B,), d_n (B,). */
int32_t dsk_triplet_loss(const float* a, const float* p, const float* n, int32_t B, int32_t D, float margin,
                         float* loss, float* d_p, float* d_n, void* stream);
/* Gradients of the loss w.r.t. a, p, n scaled by grad_loss (device scalar). */
int32_t dsk_triplet_loss_bwd(const float* a, const float* p, const float* n, const float* d_p, const float* d_n,
                             const float* grad_loss, int32_t B, int32_t D, float margin, float* ga, float* gp,
                             float* gn, void* stream);

/* "Choose the hard negatives" (/root/reference/train_triplet.py:251-262):
 * idx = ascending indices i with d_n[i] - d_p[i] < margin  (== np.where(mask == 1)); count on device. */
int32_t dsk_margin_select(const float* d_p, const float* d_n, int32_t B, float margin, int64_t* idx,
                          int32_t* count, void* stream);
/* Row gather out[j] = src[idx[j]] for j < *count (train_triplet.py:265-274), row = row_elems fp32. */
int32_t dsk_gather_rows(const float* src, const int64_t* idx, const int32_t* count, int32_t max_rows,
                        int64_t row_elems, float* out, void* stream);

/* All-pairs distance + per-row k smallest over different-label columns (BASELINE config 4; no reference
 * implementation exists — defined from PairwiseDistance, model.py:13-18):
 * D[i][j] = sqrt(sum_d (E[i]-E[j])^2 + 1e-4/Dim), candidates j with labels[j] != labels[i];
 * ties broken by lower j. Writes idx (N,k) int64 and val (N,k) fp32, ascending distance. */
/* Same result (bit-identical indices and values), computed with a tcgen05 fp16 Gram GEMM + candidate selection + exact
 * fp32 refinement of the k+8 best candidates per row (exact row scan on the device when the safety margin is not met).
 * Falls back to dsk_allpairs_topk when D % 64 != 0 or k > 8. */
int32_t dsk_allpairs_topk_tc(dsk_handle h, const float* E, const int64_t* labels, int32_t N, int32_t D, int32_t k,
                             int64_t* idx, float* val, void* stream);
int32_t dsk_allpairs_topk(const float* E, const int64_t* labels, int32_t N, int32_t D, int32_t k, int64_t* idx,
                          float* val, void* stream);

/* nn.Linear of DeepSpeakerModel.forward_classifier (/root/reference/model.py:167,220-223): y (M,N) = x (M,K) w(N,K)^T + b.
 * fp32 on the CUDA cores, fixed summation order (deterministic).  b may be NULL. */
int32_t dsk_linear_forward(const float* x, const float* w, const float* b, int32_t M, int32_t N, int32_t K, float* y,
                           void* stream);
/* its backward: gx (M,K) = gy w, gw (N,K) = gy^T x, gb (N) = column sums of gy; any output pointer may be NULL. */
int32_t dsk_linear_backward(const float* x, const float* w, const float* gy, int32_t M, int32_t N, int32_t K, float* gx,
                            float* gw, float* gb, void* stream);
/* nn.CrossEntropyLoss() over (M,C) logits and int64 labels (/root/reference/train_triplet.py:281-285):
 * loss (1,) = mean_i (logsumexp_j logits[i] - logits[i][label_i]); also writes lse (M,) and row_loss (M,) (workspace
 * the backward reads).  A label outside [0,C) yields NaN. */
int32_t dsk_cross_entropy(const float* logits, const int64_t* labels, int32_t M, int32_t C, float* loss, float* lse,
                          float* row_loss, void* stream);
/* dlogits (M,C) = (softmax(logits) - onehot(labels)) * grad_loss / M; grad_loss is a device scalar. */
int32_t dsk_cross_entropy_bwd(const float* logits, const int64_t* labels, const float* lse, const float* grad_loss,
                              int32_t M, int32_t C, float* dlogits, void* stream);

/* torch.optim.Adagrad step (/root/reference/train_triplet.py:369-383, called at :224,291) on ONE flat bucket of n fp32
 * elements (parameters, gradients and the running sum of squares laid out identically), fused with the gradient scale
 * that follows the data-parallel allreduce:  g = grad * grad_mult (/ *grad_denom if non-NULL, a device scalar);
 * g += weight_decay * p;  sum = fma(g, g, sum);  p += (g * -clr) / (sqrt(sum) + eps),  clr = lr / (1 + (step-1) lr_decay).
 * `step` counts from 1.  Operation order = torch's foreach Adagrad, so the result is bit-identical to it.
 * Buffers must be 16-byte aligned. */
int32_t dsk_adagrad_step(float* param, const float* grad, float* state_sum, int64_t n, double lr, double lr_decay,
                         double weight_decay, double eps, int64_t step, float grad_mult, const float* grad_denom,
                         void* stream);

/* Serving pipeline for the reference's test() loop (/root/reference/train_triplet.py:337-350: batch to the GPU, model(x),
 * result back - serialised on one stream there).  `primary` owns the weights (dsk_load_weights); the pipeline adds
 * `lanes` handles that borrow them (dsk_share_weights; `primary` stays free for its caller), one compute stream per lane and two copy streams, and keeps depth*lanes device
 * slots.  dsk_pipeline_submit queues, without blocking the host: H2D of the PINNED input batch (B,1,T,64) fp32 -> eval forward
 * on the next lane -> D2H of the (B,E) embeddings into the PINNED output; *ticket identifies the batch.  Both host buffers
 * are accessed asynchronously: leave x_host / emb_host alone until dsk_pipeline_wait(ticket) (blocks the host until that
 * batch's output is complete) or dsk_pipeline_sync.
 * dsk_pipeline_submit_device runs device-resident batches through the same lanes (inputs ordered after `after_stream`);
 * dsk_pipeline_join makes `stream` wait for everything submitted so far; dsk_pipeline_lane_stream returns the cudaStream_t of
 * compute lane 0..lanes-1 (-1: the H2D copy stream, -2: the D2H copy stream) for event timing.  One thread drives a pipeline. */
int32_t dsk_pipeline_create(dsk_pipeline* out, dsk_handle primary, int32_t lanes, int32_t depth);
int32_t dsk_pipeline_destroy(dsk_pipeline p);
int32_t dsk_pipeline_submit(dsk_pipeline p, const float* x_host, int32_t B, int32_t T, float* emb_host, int64_t* ticket);
int32_t dsk_pipeline_submit_device(dsk_pipeline p, const float* x_dev, int32_t B, int32_t T, float* emb_dev, void* after_stream,
                                   int64_t* ticket);
int32_t dsk_pipeline_join(dsk_pipeline p, void* stream);
int32_t dsk_pipeline_wait(dsk_pipeline p, int64_t ticket);
int32_t dsk_pipeline_sync(dsk_pipeline p);
int32_t dsk_pipeline_lane_stream(dsk_pipeline p, int32_t lane, void** stream_out);

/* Log mel-filterbank front-end of the reference (/root/reference/audio_processing.py:9-36 mk_MFB with constants.py:
 * python_speech_features.fbank(audio, samplerate, nfilt=64, winlen=0.025) -> 20*log10(max(., 1e-5)) (log_scale) -> minus the
 * per-bin mean over the utterance (subtract_mean, normalize_frames with Scale=False).  audio: n_samples fp32 mono on the
 * device; feat: (dsk_fbank_num_frames(n_samples, sample_rate), 64) fp32 row-major, the (T, 64) layout the network's input
 * is cropped from.  python_speech_features is not vendored in the reference (parity against it unpinned): its published
 * algorithm is restated (pre-emphasis 0.97, 25 ms / 10 ms rectangular frames, NFFT 512, |rfft|^2 / NFFT, triangular mel
 * filters, zeros -> eps). */
int64_t dsk_fbank_num_frames(int64_t n_samples, int32_t sample_rate);
int32_t dsk_fbank(const float* audio, int64_t n_samples, int32_t sample_rate, int32_t log_scale, int32_t subtract_mean,
                  float* feat, void* stream);

/* Threshold sweep of the verification metric (/root/reference/eval_metrics.py:16-37 calculate_roc, :53-88 calculate_val /
 * calculate_val_far; called from train_triplet.py:361): for every threshold t (double, as numpy's arange yields them)
 * tp[t] = #{i : same[i] && (double)dist[i] < t}, fp[t] = #{i : !same[i] && (double)dist[i] < t} — numpy's
 * np.less(dist, t) with its float32 -> float64 promotion, so the counts are exactly the reference's.  All other sweep
 * quantities (tn, fn, tpr, fpr, accuracy, val, far) are integer arithmetic on tp, fp, n_same, n_diff. */
int32_t dsk_threshold_counts(const float* dist, const uint8_t* same, int32_t P, const double* thresholds, int32_t nT,
                             int32_t* tp, int32_t* fp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSK_H_ */
