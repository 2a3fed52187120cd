/*
 * kws_b200.h — C ABI of libkws_b200.so, the B200-native (sm_100a) replacement for the hot path of
 * harvard-edge/multilingual_kws:   PCM -> micro-frontend log-mel -> EfficientNet-B0 embedding ->
 * 3-way few-shot head (+ Adam step), and the sliding-window streaming use of the same path.
 *
 * The reference exposes no FFI of its own (it is pure Python on TensorFlow); each entry point below
 * names the reference call site whose third-party native implementation it replaces.  All pointers
 * prefixed d_ are DEVICE pointers owned by the caller; `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 on success and a negative code on failure (message: kws_last_error()).
 * There is no CPU fallback: without a CUDA device every compute entry fails with KWS_ERR_CUDA.
 * Handles are immutable after creation except kws_head_t; distinct handles may be used from
 * distinct threads concurrently.
 */
#ifndef KWS_B200_H_
#define KWS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KWS_OK 0
#define KWS_ERR_ARG (-1)
#define KWS_ERR_CUDA (-2)
#define KWS_ERR_UNSUPPORTED (-3)

typedef struct kws_frontend kws_frontend_t;
typedef struct kws_embed kws_embed_t;
typedef struct kws_head kws_head_t;

const char* kws_last_error(void);
int kws_abi_version(void);

/* Page-locked host staging memory for callers that hold their clips in HOST buffers (the reference feeds numpy
 * arrays: multilingual_kws/embedding/input_data.py:66-83 file2spec, batch_streaming_analysis.py:100-118).
 * write_combined = 1 -> cudaHostAllocWriteCombined: upload-only buffers (the host writes PCM into them, the device
 * reads them over PCIe at line rate); 0 -> ordinary pinned memory (download buffers the host reads). */
int kws_host_alloc(void** out, size_t bytes, int write_combined);
int kws_host_free(void* p);

/* ---------------------------------------------------------------------------------------------
 * Frontend — replaces frontend_op.audio_microfrontend(...) as called at
 * multilingual_kws/embedding/input_data.py:25-33 (and its per-window use at
 * multilingual_kws/embedding/batch_streaming_analysis.py:108-115).  Arguments are the op's
 * attributes; the reference passes (16000, 30, 20, 40) and the defaults
 * (125, 7500, 10, .025, .06, .05, 1, .95, 80, 21, 1, 6).
 * ------------------------------------------------------------------------------------------- */
int kws_frontend_create(kws_frontend_t** out, int sample_rate, int window_ms, int step_ms, int num_channels,
                        float lower_hz, float upper_hz, int smoothing_bits, float even_smoothing,
                        float odd_smoothing, float min_signal_remaining, int enable_pcan, float pcan_strength,
                        float pcan_offset, int gain_bits, int enable_log, int scale_shift);
void kws_frontend_destroy(kws_frontend_t* fe);
/* frames the op emits for an n_samples-long vector: (n - window)/step + 1, or 0 */
int kws_frontend_num_frames(const kws_frontend_t* fe, int n_samples);
/* Host copy of the fixed-point tables (for cross-checking against an independent implementation). */
int kws_frontend_tables(const kws_frontend_t* fe, void* out, size_t out_bytes, size_t* needed_bytes);

/* Batch of independent clips.  d_pcm int16 [batch, n_samples]; every clip starts from a fresh zero
 * noise-estimate state (the op allocates a new FrontendState per call).  Writes
 * d_out_f32[b, frame, channel] = (float)uint16 * out_scale  (the reference multiplies by 10/256,
 * input_data.py:34) and/or the raw uint16 values; either output pointer may be NULL. */
int kws_frontend_forward(kws_frontend_t* fe, const int16_t* d_pcm, int batch, int n_samples, float out_scale,
                         float* d_out_f32, uint16_t* d_out_u16, void* stream);

/* Sliding windows over one long signal (batch_streaming_analysis.py:66-115): window w covers samples
 * [w*hop, w*hop + clip_samples), w < n_windows, each from a fresh zero state.  Per-frame work
 * (window/FFT/mel/sqrt) is computed once per step-aligned frame and shared by all windows, which is
 * bit-identical to recomputing it per window; requires hop_samples % window_step == 0.
 * d_scratch: kws_frontend_stream_scratch_bytes() bytes of device memory. */
int64_t kws_frontend_stream_num_windows(const kws_frontend_t* fe, int64_t total_samples, int clip_samples,
                                        int hop_samples);
size_t kws_frontend_stream_scratch_bytes(const kws_frontend_t* fe, int64_t total_samples);
int kws_frontend_stream(kws_frontend_t* fe, const int16_t* d_pcm, int64_t total_samples, int clip_samples,
                        int hop_samples, int64_t first_window, int64_t n_windows, float out_scale,
                        float* d_out_f32, void* d_scratch, int scratch_ready, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Embedding tower — replaces `tf.keras.models.load_model(base)` + `Model(inputs,
 * base.get_layer("dense_2").output)` + `.predict(specs)` (multilingual_kws/embedding/
 * transfer_learning.py:36-43, distance_filtering.py:12-27,55,69; architecture defined at
 * multilingual_kws/train_multilingual_embedding.py:66-83 = Keras EfficientNetB0 + GAP + 3 Dense).
 * `blob` is the "KWSW0001" container of Keras-named fp32 tensors (multilingual_kws_b200/model.py
 * writes it).  BatchNorm is folded at create time (inference mode, as every reference forward outside
 * pre-training).  Activations and tensor-core operands are 16-bit with fp32 accumulation: act_dtype 0 = IEEE
 * fp16 (11-bit significand; default, what the parity tests run) or 1 = bf16 (wider range, 8-bit significand);
 * output is fp32 [batch, out_dim].
 * ------------------------------------------------------------------------------------------- */
int kws_embed_create(kws_embed_t** out, const void* blob, size_t bytes, int act_dtype /* 0 fp16, 1 bf16 */);
void kws_embed_destroy(kws_embed_t* m);
int kws_embed_info(const kws_embed_t* m, int* in_h, int* in_w, int* out_dim, int* n_ops, double* flops_per_clip);
int kws_embed_op_name(const kws_embed_t* m, int op, char* buf, size_t buf_bytes, int64_t* out_elems_per_clip);
/* op kind: 0 stem conv, 1 tcgen05 GEMM, 2 depthwise+SE; flops / algorithmic activation bytes per clip */
int kws_embed_op_info(const kws_embed_t* m, int op, int* kind, double* flops_per_clip, double* bytes_per_clip,
                      int* gemm_n, int* gemm_k, int* rows_per_clip);
/* clips per pass through the layer list (keeps one chunk's activations L2-resident) */
int kws_embed_set_chunk(kws_embed_t* m, int chunk);
/* clips per pass for the late (small-activation) layers; large values fill the SMs */
int kws_embed_set_chunk_late(kws_embed_t* m, int chunk);
/* 1 (default): capture the launch list of a forward pass into a CUDA graph per (buffers, batch, schedule) the second
   time that key is met, and replay it from then on (one-off buffers run plain launches); 0: plain launches always */
int kws_embed_set_graph(kws_embed_t* m, int enable);
/* The frozen part of the network only: ops 0 .. tap_op, whose output (16-bit NHWC) is copied to d_tap; no embedding is
   written.  The trainable tail of phase 2 (transfer_learning.py:97-112) starts from that tensor. */
int kws_embed_forward_until(kws_embed_t* m, const float* d_feats, int batch, void* d_workspace, size_t ws_bytes, int tap_op,
                            void* d_tap, void* stream);
/* Schedule of the network's tail (blocks 4a..7a + top conv, maps of <= 7x5 pixels): 0 (default) = layer by layer (expand
   GEMM, depthwise + pool, SE GEMMs, gating, project GEMM: six launches per block), 1 = one fused tcgen05 launch per
   MBConv block, 2 = runs of consecutive blocks (and the top conv + average pool) per launch, 3 = runs of the blocks
   with <= 672 expanded channels only.  Same results up to 16-bit rounding of the intermediates the fused kernel keeps
   in fp32.  The fused schedules cut launches (81 -> 18) and DRAM traffic, not time. */
int kws_embed_set_fuse(kws_embed_t* m, int mode);
/* kernel launches one forward pass of `batch` clips issues */
int kws_embed_launches(const kws_embed_t* m, int batch);
size_t kws_embed_workspace_bytes(const kws_embed_t* m, int batch);
/* d_feats fp32 [batch, 49, 40] (the frontend's output) -> d_emb fp32 [batch, out_dim] */
int kws_embed_forward(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                      size_t ws_bytes, void* stream);
/* Throughput schedule for callers that keep several forward passes in flight on different streams: the grids of the
 * network's head (stem ... block3b) are sized for sm_head SMs, those of its launch/latency-bound tail for sm_tail SMs
 * (0 = all SMs), so the tail's persistent CTAs leave room for a neighbouring pass.  Same results as kws_embed_forward. */
int kws_embed_forward_budget(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                             size_t ws_bytes, int sm_head, int sm_tail, void* stream);
/* same, additionally copying the output of op `tap_op` (16-bit NHWC; fp32 for the last op) to d_tap */
int kws_embed_forward_tap(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                          size_t ws_bytes, int tap_op, void* d_tap, void* stream);

/* profiling: CUDA events around every op on `stream`; host_op_ms[n_ops] = device ms per op; synchronises */
int kws_embed_forward_timed(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                            size_t ws_bytes, float* host_op_ms, void* stream);

/* The pointwise-conv / dense operator on its own (tcgen05 GEMM + fused epilogue):
 * out[M,N] = act(A[M,K] x W[N,K]^T + bias[N]) (+ residual[M,N]); A, W, residual 16-bit K-major (dtype 0 fp16,
 * 1 bf16); out same type
 * or fp32; act 0 none, 1 swish, 2 relu, 3 selu; gap4 averages aligned groups of 4 rows (out [M/4,N]);
 * block_n 0 = automatic.  N and K must be multiples of 8. */
int kws_gemm_h16(const void* d_a, const void* d_w, int M, int N, int K, const float* d_bias, int act,
                 const void* d_residual, void* d_out, int out_f32, int gap4, int block_n, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Few-shot head — replaces what Keras runs for the trainable part of
 *   Sequential[frozen embedding, Dense(18, tanh), Dense(3, softmax)] compiled with Adam(lr) and
 *   SparseCategoricalCrossentropy  (multilingual_kws/embedding/transfer_learning.py:47-59) during
 *   xfer.fit(...) / .predict(...) (:86-93, :196-198).
 * Parameters are host fp32 arrays in Keras layouts: w1 [in_dim, hidden], b1 [hidden], w2 [hidden, classes],
 * b2 [classes].  Flat gradient buffer layout (kws_head_flat_size() floats):
 *   [dW1 | db1 | dW2 | db2 | loss_sum | n_correct | n_samples], gradients SUMMED over the local batch so one
 *   all-reduce(sum) across ranks gives the global batch; kws_head_apply_adam divides by n_samples.
 * ------------------------------------------------------------------------------------------- */
int kws_head_create(kws_head_t** out, int in_dim, int hidden, int classes, const float* w1, const float* b1,
                    const float* w2, const float* b2, float beta1, float beta2, float eps);
void kws_head_destroy(kws_head_t* h);
size_t kws_head_flat_size(const kws_head_t* h);
int kws_head_num_params(const kws_head_t* h);
long long kws_head_step_count(const kws_head_t* h);
int kws_head_forward(kws_head_t* h, const float* d_emb, int batch, float* d_probs, void* stream);
int kws_head_grad(kws_head_t* h, const float* d_emb, const int32_t* d_labels, int batch, float* d_flat, void* stream);
int kws_head_apply_adam(kws_head_t* h, const float* d_flat, float lr, void* stream);
int kws_head_get_params(const kws_head_t* h, float* host_out);
int kws_head_reset_optimizer(kws_head_t* h);
/* d(sum of losses) / d(embedding) [B, in_dim] fp32 of the batch of the last kws_head_grad call: what flows into the
   embedding when its top layers train too (transfer_learning.py:97-112, backprop_into_embedding=True). */
int kws_head_input_grad(kws_head_t* h, int B, float* d_demb, void* stream);
/* kws_head_apply_adam with the step size read from device memory (kws_train_lr_step); graph-capturable; does NOT advance
   kws_head_step_count (a capture executes nothing, a replay is not seen by the host) */
int kws_head_apply_adam_dev(kws_head_t* h, const float* d_flat, const float* d_lr_t, void* stream);
/* reports `steps` executed kws_head_apply_adam_dev updates (keeps kws_head_step_count in step) */
int kws_head_advance_step_count(kws_head_t* h, long long steps);

/* ---------------------------------------------------------------------------------------------
 * Fine-tune backward through the top of the embedding — replaces what Keras / TensorFlow autodiff executes in the
 * second xfer.fit of transfer_learn (multilingual_kws/embedding/transfer_learning.py:97-112: "unfreeze the top 20
 * layers while leaving BatchNorm layers frozen" + Adam(embedding_lr)) for those 20 layers: block7a, top_conv, the
 * dense tower.  All contractions (forward, data gradient, weight gradient) are kws_gemm_h16 calls; these are the
 * kernels between them.  16-bit tensors are IEEE half; activation gradients carry a static loss scale.
 * --------------------------------------------------------------------------------------------- */
/* out[c * ld_out + r] = in[r][c]; ld_out >= rows (the weight-gradient GEMMs contract over the batch, padded to 8) */
int kws_train_transpose_h16(const void* d_in, int rows, int cols, void* d_out, int ld_out, void* stream);
int kws_train_swish_fwd(const void* d_z, size_t n, void* d_a, void* stream);
int kws_train_gap_swish_fwd(const void* d_t, int B, int P, int C, void* d_h0, void* stream);    /* mean_p swish(t) */
int kws_train_gap_swish_bwd(const void* d_dh0, const void* d_t, int B, int P, int C, void* d_dt, void* stream);
/* depthwise conv of a <= 7x7 map (TF SAME for stride 1, correct_pad + VALID for stride 2), NHWC 16-bit:
   z = conv(e) * folded BN + shift, d = swish(z), pooled[b][c] = mean_p d */
int kws_train_dw_fwd(const void* d_e, const float* d_w, const float* d_shift, int B, int C, int H, int W, int K, int S,
                     int pad_top, int pad_left, void* d_z, void* d_d, void* d_pooled, void* stream);
size_t kws_train_dw_bwd_scratch_floats(int B, int C, int K);
/* dz = (dd + dp / P_out) * swish'(z); de = conv^T(dz) (16-bit); dw [K*K][C] fp32 = sum over the batch (fixed order) */
int kws_train_dw_bwd(const void* d_dd, const void* d_dp, const void* d_z, const void* d_e, const float* d_w, int B, int C,
                     int H, int W, int K, int S, int pad_top, int pad_left, void* d_de, float* d_dw, float* d_scratch,
                     void* stream);
int kws_train_gate_fwd(const void* d_d, const void* d_g, int B, int P, int C, void* d_out, void* stream);   /* d * g[b][c] */
/* dd = ddg * g; dgpre[b][c] = (sum_p ddg * d) * g (1 - g) */
int kws_train_gate_bwd(const void* d_ddg, const void* d_d, const void* d_g, int B, int P, int C, void* d_dd,
                       void* d_dgpre, void* stream);
/* dz = scale * dy * act'(ref): kind 0 relu (ref = its fp16 output), 1 selu (dy and ref = fp32: the embedding and the
   head's input gradient), 2 swish (ref = fp16 pre-activation) */
int kws_train_act_bwd(int kind, const void* d_dy, const void* d_ref, size_t n, float scale, void* d_dz, void* stream);
int kws_train_colsum(const void* d_x, int rows, int cols, float* d_out, void* stream);          /* fp32 bias gradient */
/* Keras Adam (eps outside the sqrt, bias-corrected step) on an fp32 master [n / cols][cols]: the gradient is a SUM over
   the global batch with respect to the BN-folded weight, g = grad * row_scale[row] / (*d_count * loss_scale);
   d_out16 / d_out32 (optional) receive the folded 16-bit / fp32 copy the forward kernels read. */
int kws_train_adam(float* d_param, float* d_m, float* d_v, const float* d_grad, size_t n, int cols, const float* d_row_scale,
                   const float* d_count, float loss_scale, float lr, long long step, const float* d_lr_t, float beta1,
                   float beta2, float eps, void* d_out16, float* d_out32, void* stream);
/* *d_step += 1; *d_lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t): with d_lr_t passed to kws_train_adam /
   kws_head_apply_adam_dev (lr, step ignored) a whole optimisation step can be captured in a CUDA graph and replayed */
int kws_train_lr_step(long long* d_step, float lr, float beta1, float beta2, float* d_lr_t, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Streaming post-processor — replaces the per-window Python loop
 *   for ix, offset in enumerate(range(0, audio_data_end, clip_stride_samples)):
 *       recognize_commands.process_latest_result(inferences[ix], current_time_ms, recognize_element)
 * of multilingual_kws/embedding/batch_streaming_analysis.py:131-163, i.e.
 * SingleTargetRecognizeCommands.process_latest_result (single_target_recognize_commands.py:94-207), for every
 * detection threshold of a sweep at once.  d_probs float32 [n_windows, n_labels] (the softmax rows, left on the
 * device by the head), d_times_ms int64 [n_windows] ascending (the reference raises ValueError on out-of-order
 * input; the caller checks).  Outputs: d_scores double [n_windows] = recognize_element.score at every step
 * (bit-identical: same float64 operation order), d_valid uint8 [n_windows] = 0 where the reference bails out
 * ("too few results"), d_found_idx int32 [n_thresholds, max_found] = window indices of the new non-silence
 * commands, d_found_count int32 [n_thresholds] (may exceed max_found: rerun with a larger buffer).
 * ------------------------------------------------------------------------------------------- */
int kws_stream_detect(const float* d_probs, int n_windows, int n_labels, int target_id, const int64_t* d_times_ms,
                      double average_window_duration_ms, double suppression_ms, int minimum_count,
                      const double* d_thresholds, int n_thresholds, double* d_scores, uint8_t* d_valid,
                      int32_t* d_found_idx, int32_t* d_found_count, int max_found, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Waveform augmentation of the fine-tune input pipeline — replaces, per batch, the tf.data map
 *   ds.map(self.augment) -> ds.map(self.get_spectrogram_and_label_id) -> ds.map(self.map_spec_aug)
 * of multilingual_kws/embedding/input_data.py:447-471, i.e. random_timeshift (:243-267),
 * random_background_sample (:227-241), add_background (:141-157), the branches of augment (:275-304), the int16
 * cast at :23 and spec_augment (:306-364).  Random decisions are drawn by the caller (one kws_aug_item per output
 * clip); samples stay on the device: d_fg int16 [n_fg, fg_stride] holds the source clips (decode_wav output x 32768,
 * exact), d_bg int16 [n_bg, bg_stride] the zero-padded background recordings; both 16-byte aligned with strides that
 * are multiples of 8.  d_pcm_out int16 [batch, n_samples] feeds kws_frontend_forward; d_audio_out (optional, may be
 * NULL) receives the float32 waveform before the int16 cast.
 *   mode 0: out[i] = fg[i - shift] (zero outside the clip)                       plain / time-shifted / "unknown" clip
 *   mode 1: out[i] = bg[bg_offset + i] * volume                                  "silence" sample
 *   mode 2: out[i] = clip(bg[bg_offset + i] * (rms_fg / rms_bg) * volume + fg[i - shift], -1, 1)   background mix
 * kws_spec_mask zeroes, in place, up to two frequency bands and two time bands per clip of d_feats
 * [batch, frames, channels]; d_bands int32 [batch, 8] = (f_start, f_size) x 2, (t_start, t_size) x 2, size 0 = unused.
 * ------------------------------------------------------------------------------------------- */
typedef struct kws_aug_item {
  int32_t mode, fg_index, shift, bg_index, bg_offset;
  float volume;
  int32_t reserved[2];
} kws_aug_item;
int kws_augment_pcm(const int16_t* d_fg, int n_fg, int fg_stride, const int16_t* d_bg, int n_bg, int bg_stride,
                    const void* d_plan, int batch, int n_samples, int16_t* d_pcm_out, float* d_audio_out, void* stream);
int kws_spec_mask(float* d_feats, int batch, int frames, int channels, const int32_t* d_bands, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KWS_B200_H_ */
