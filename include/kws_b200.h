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

#ifdef __cplusplus
}
#endif
#endif /* KWS_B200_H_ */
