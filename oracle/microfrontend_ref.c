/*
 * oracle/microfrontend_ref.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, integer-exact) of the algorithm behind the reference call
 *     multilingual_kws/embedding/input_data.py:19-35   (to_micro_spectrogram)
 * i.e. TensorFlow 2.7's `audio_microfrontend` op (pinned by docker/Dockerfile:1,
 * `FROM tensorflow/tensorflow:2.7.0-gpu`) with the arguments the reference passes:
 * sample_rate 16000, window 30 ms, step 20 ms, 40 channels, every other attribute at the
 * op's default (PCAN on, log on, noise reduction on, out_scale 1, float32 out).
 *
 * The arithmetic itself lives in an un-vendored third-party dependency that is ABSENT from
 * /root/reference: tensorflow/lite/experimental/microfrontend/lib/{window,fft,filterbank,
 * noise_reduction,pcan_gain_control,log_scale,frontend}*.c and third_party kissfft built with
 * FIXED_POINT=16.  This file restates that published algorithm from the
 * public description of those sources; nothing is copied.
 *
 * PINS: the reference ships no tests / golden vectors for this path and TensorFlow cannot be
 * executed in the build container.  The oracle reproduces exactly the known-answer vectors of the
 * upstream op's own unit tests (audio_microfrontend_op_test.py testSimple / testSimpleFloatScaled and
 * the lib/ unit-test chain (`<stage>_test.cc`): window, filterbank sqrt, noise reduction, PCAN, log scale) — restated from
 * the published test files, tests/golden/tf_microfrontend_kat.json, tests/test_oracle_tf_kat.py — at
 * the op's 1 kHz / 25 ms / 2-channel test configuration; the reference's 16 kHz / 30 ms / 40-channel
 * configuration runs the same code and is covered by closed-form table checks, FFT-vs-numpy
 * consistency and analytic invariants (tests/test_oracle_frontend.py), and by an independent floating-point
 * statement of the same signal chain (oracle/frontend_float_model.py, tests/test_oracle_float_model.py: the
 * final features agree to < 5 units of ~400 on 16 000 entries, without bias, wherever this integer pipeline's
 * own resolution permits the comparison).  A TensorFlow-generated vector at the production configuration is
 * still missing.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call into this file.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off: plain SSE2 float/double semantics so the
 * init tables round the way an x86-64 TF build rounds them).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KWS_MAX_CHANNELS 64
#define KWS_FFT_MAX 2048

/* ---- fixed-point constants ---- */
enum {
  kFrontendWindowBits = 12,
  kFilterbankBits = 12,
  kNoiseReductionBits = 14,
  kPcanSnrBits = 12,
  kPcanOutputBits = 6,
  kWideDynamicFunctionBits = 32,
  kWideDynamicFunctionLUTSize = 4 * 32 - 3,
  kLogScaleLog2 = 16,
  kLogSegmentsLog2 = 7,
  kLogCoeff = 45426,
  kFracBits = 15, /* kissfft FIXED_POINT=16 */
  kSampMax = 32767
};

typedef struct { int16_t r, i; } cpx16;

typedef struct kws_ref_frontend {
  /* config */
  int sample_rate, window_size, window_step, num_channels;
  int fft_size;             /* smallest power of two >= window_size */
  int smoothing_bits, enable_pcan, enable_log, scale_shift, gain_bits;
  /* window */
  int16_t window_coef[KWS_FFT_MAX];
  /* kissfft tables for the complex FFT of size fft_size/2 */
  int ncfft;
  int factors[64];
  cpx16 twiddles[KWS_FFT_MAX / 2];
  cpx16 super_twiddles[KWS_FFT_MAX / 4];
  /* filterbank, un-padded: per FFT bin its band and (weight, unweight) */
  int start_index, end_index;
  int16_t bin_band[KWS_FFT_MAX / 2 + 1];
  int16_t bin_weight[KWS_FFT_MAX / 2 + 1];
  int16_t bin_unweight[KWS_FFT_MAX / 2 + 1];
  int band_start[KWS_MAX_CHANNELS + 2]; /* first bin of band i; band_start[nch+1] = end */
  /* noise reduction */
  uint16_t even_smoothing, odd_smoothing, min_signal_remaining;
  /* pcan */
  int16_t gain_lut[kWideDynamicFunctionLUTSize];
  int snr_shift;
  /* log */
  uint16_t log_lut[130];
  int correction_bits;
} kws_ref_frontend;

/* ------------------------------------------------------------------ bits */
static int msb32(uint32_t x) { return x ? 32 - __builtin_clz(x) : 0; }
static int msb64(uint64_t x) { return x ? 64 - __builtin_clzll(x) : 0; }

/* bit-by-bit integer sqrt with final round-to-nearest (App. A.4) */
static uint16_t sqrt32_ref(uint32_t num) {
  if (num == 0) return 0;
  uint32_t res = 0;
  int max_bit_number = 32 - msb32(num);
  max_bit_number |= 1;
  uint32_t bit = 1u << (31 - max_bit_number);
  int iterations = (31 - max_bit_number) / 2 + 1;
  while (iterations--) {
    if (num >= res + bit) {
      num -= res + bit;
      res = (res >> 1) + bit;
    } else {
      res >>= 1;
    }
    bit >>= 2;
  }
  if (num > res && res != 0xFFFF) ++res;
  return (uint16_t)res;
}

static uint32_t sqrt64_ref(uint64_t num) {
  if ((num >> 32) == 0) return sqrt32_ref((uint32_t)num);
  uint64_t res = 0;
  int max_bit_number = 64 - msb64(num);
  max_bit_number |= 1;
  uint64_t bit = 1ull << (63 - max_bit_number);
  int iterations = (63 - max_bit_number) / 2 + 1;
  while (iterations--) {
    if (num >= res + bit) {
      num -= res + bit;
      res = (res >> 1) + bit;
    } else {
      res >>= 1;
    }
    bit >>= 2;
  }
  if (num > res && res != 0xFFFFFFFFull) ++res;
  return (uint32_t)res;
}

uint32_t kws_ref_sqrt64(uint64_t x) { return sqrt64_ref(x); }

/* ------------------------------------------------------------------ kissfft (FIXED_POINT=16) */
static inline int16_t sround(int32_t x) { return (int16_t)((x + (1 << (kFracBits - 1))) >> kFracBits); }
static inline int16_t fixdiv(int16_t x, int div) { return sround((int32_t)x * (kSampMax / div)); }
static inline cpx16 cmul(cpx16 a, cpx16 b) {
  cpx16 m;
  m.r = sround((int32_t)a.r * b.r - (int32_t)a.i * b.i);
  m.i = sround((int32_t)a.r * b.i + (int32_t)a.i * b.r);
  return m;
}
static inline cpx16 cadd(cpx16 a, cpx16 b) { cpx16 c = {(int16_t)(a.r + b.r), (int16_t)(a.i + b.i)}; return c; }
static inline cpx16 csub(cpx16 a, cpx16 b) { cpx16 c = {(int16_t)(a.r - b.r), (int16_t)(a.i - b.i)}; return c; }

static void kf_factor(int n, int* facbuf) {
  int p = 4;
  double floor_sqrt = floor(sqrt((double)n));
  do {
    while (n % p) {
      switch (p) {
        case 4: p = 2; break;
        case 2: p = 3; break;
        default: p += 2; break;
      }
      if (p > floor_sqrt) p = n;
    }
    n /= p;
    *facbuf++ = p;
    *facbuf++ = n;
  } while (n > 1);
}

static void kf_bfly2(cpx16* Fout, size_t fstride, const kws_ref_frontend* st, int m) {
  cpx16* Fout2 = Fout + m;
  const cpx16* tw1 = st->twiddles;
  do {
    Fout->r = fixdiv(Fout->r, 2); Fout->i = fixdiv(Fout->i, 2);
    Fout2->r = fixdiv(Fout2->r, 2); Fout2->i = fixdiv(Fout2->i, 2);
    cpx16 t = cmul(*Fout2, *tw1);
    tw1 += fstride;
    *Fout2 = csub(*Fout, t);
    *Fout = cadd(*Fout, t);
    ++Fout2; ++Fout;
  } while (--m);
}

static void kf_bfly4(cpx16* Fout, size_t fstride, const kws_ref_frontend* st, size_t m) {
  const cpx16 *tw1, *tw2, *tw3;
  cpx16 s[6];
  size_t k = m;
  const size_t m2 = 2 * m, m3 = 3 * m;
  tw3 = tw2 = tw1 = st->twiddles;
  do {
    Fout[0].r = fixdiv(Fout[0].r, 4);   Fout[0].i = fixdiv(Fout[0].i, 4);
    Fout[m].r = fixdiv(Fout[m].r, 4);   Fout[m].i = fixdiv(Fout[m].i, 4);
    Fout[m2].r = fixdiv(Fout[m2].r, 4); Fout[m2].i = fixdiv(Fout[m2].i, 4);
    Fout[m3].r = fixdiv(Fout[m3].r, 4); Fout[m3].i = fixdiv(Fout[m3].i, 4);

    s[0] = cmul(Fout[m], *tw1);
    s[1] = cmul(Fout[m2], *tw2);
    s[2] = cmul(Fout[m3], *tw3);

    s[5] = csub(Fout[0], s[1]);
    Fout[0] = cadd(Fout[0], s[1]);
    s[3] = cadd(s[0], s[2]);
    s[4] = csub(s[0], s[2]);
    Fout[m2] = csub(Fout[0], s[3]);
    tw1 += fstride; tw2 += fstride * 2; tw3 += fstride * 3;
    Fout[0] = cadd(Fout[0], s[3]);

    Fout[m].r = (int16_t)(s[5].r + s[4].i);
    Fout[m].i = (int16_t)(s[5].i - s[4].r);
    Fout[m3].r = (int16_t)(s[5].r - s[4].i);
    Fout[m3].i = (int16_t)(s[5].i + s[4].r);
    ++Fout;
  } while (--k);
}

static void kf_work(cpx16* Fout, const cpx16* f, size_t fstride, const int* factors,
                    const kws_ref_frontend* st) {
  cpx16* Fout_beg = Fout;
  const int p = *factors++;
  const int m = *factors++;
  const cpx16* Fout_end = Fout + p * m;
  if (m == 1) {
    do { *Fout = *f; f += fstride; } while (++Fout != Fout_end);
  } else {
    do { kf_work(Fout, f, fstride * p, factors, st); f += fstride; } while ((Fout += m) != Fout_end);
  }
  Fout = Fout_beg;
  switch (p) {
    case 2: kf_bfly2(Fout, fstride, st, m); break;
    case 4: kf_bfly4(Fout, fstride, st, m); break;
    default: abort(); /* power-of-two sizes only */
  }
}

static void make_cexp(cpx16* x, double phase) {
  x->r = (int16_t)floor(.5 + kSampMax * cos(phase));
  x->i = (int16_t)floor(.5 + kSampMax * sin(phase));
}

/* real FFT of `fft_size` int16 samples → fft_size/2+1 complex bins (kiss_fftr semantics) */
static void fftr(const kws_ref_frontend* st, const int16_t* timedata, cpx16* freq) {
  const int ncfft = st->ncfft;
  cpx16 tmp[KWS_FFT_MAX / 2];
  kf_work(tmp, (const cpx16*)timedata, 1, st->factors, st);

  cpx16 tdc;
  tdc.r = fixdiv(tmp[0].r, 2);
  tdc.i = fixdiv(tmp[0].i, 2);
  freq[0].r = (int16_t)(tdc.r + tdc.i);
  freq[ncfft].r = (int16_t)(tdc.r - tdc.i);
  freq[ncfft].i = freq[0].i = 0;
  for (int k = 1; k <= ncfft / 2; ++k) {
    cpx16 fpk = tmp[k], fpnk, f1k, f2k, tw;
    fpnk.r = tmp[ncfft - k].r;
    fpnk.i = (int16_t)(-tmp[ncfft - k].i);
    fpk.r = fixdiv(fpk.r, 2);   fpk.i = fixdiv(fpk.i, 2);
    fpnk.r = fixdiv(fpnk.r, 2); fpnk.i = fixdiv(fpnk.i, 2);
    f1k = cadd(fpk, fpnk);
    f2k = csub(fpk, fpnk);
    tw = cmul(f2k, st->super_twiddles[k - 1]);
    freq[k].r = (int16_t)(((int32_t)f1k.r + tw.r) >> 1);
    freq[k].i = (int16_t)(((int32_t)f1k.i + tw.i) >> 1);
    freq[ncfft - k].r = (int16_t)(((int32_t)f1k.r - tw.r) >> 1);
    freq[ncfft - k].i = (int16_t)(((int32_t)tw.i - f1k.i) >> 1);
  }
}

/* ------------------------------------------------------------------ init */
static float freq_to_mel(float freq) { return 1127.0 * log1p(freq / 700.0); }

static int16_t pcan_gain_lookup(float strength, float offset, int gain_bits, int32_t input_bits, uint32_t x) {
  const float x_as_float = ((float)x) / ((uint32_t)1 << input_bits);
  const float gain_as_float = ((uint32_t)1 << gain_bits) * powf(x_as_float + offset, -strength);
  if (gain_as_float > INT16_MAX) return INT16_MAX;
  return (int16_t)(gain_as_float + 0.5f);
}

kws_ref_frontend* kws_ref_frontend_create(int sample_rate, int window_ms, int step_ms, int num_channels,
                                          float lower_hz, float upper_hz, int smoothing_bits,
                                          float even_smoothing, float odd_smoothing,
                                          float min_signal_remaining, int enable_pcan, float pcan_strength,
                                          float pcan_offset, int gain_bits, int enable_log, int scale_shift) {
  if (num_channels < 1 || num_channels > KWS_MAX_CHANNELS) return NULL;
  kws_ref_frontend* st = (kws_ref_frontend*)calloc(1, sizeof(*st));
  if (!st) return NULL;
  st->sample_rate = sample_rate;
  st->window_size = window_ms * sample_rate / 1000;
  st->window_step = step_ms * sample_rate / 1000;
  st->num_channels = num_channels;
  st->smoothing_bits = smoothing_bits;
  st->enable_pcan = enable_pcan;
  st->enable_log = enable_log;
  st->scale_shift = scale_shift;
  st->gain_bits = gain_bits;
  int fft_size = 1;
  while (fft_size < st->window_size) fft_size <<= 1;
  if (fft_size > KWS_FFT_MAX || fft_size < 8) { free(st); return NULL; }
  st->fft_size = fft_size;

  /* A.2 window: float arg, double cos, float store, double +0.5, floor */
  {
    const float arg = M_PI * 2.0 / ((float)st->window_size);
    for (int i = 0; i < st->window_size; ++i) {
      float float_value = 0.5 - (0.5 * cos(arg * (i + 0.5)));
      st->window_coef[i] = (int16_t)floor(float_value * (1 << kFrontendWindowBits) + 0.5);
    }
  }
  /* A.3 kissfft tables */
  {
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    st->ncfft = fft_size / 2;
    for (int i = 0; i < st->ncfft; ++i) make_cexp(&st->twiddles[i], -2 * pi * i / st->ncfft);
    kf_factor(st->ncfft, st->factors);
    for (int i = 0; i < st->ncfft / 2; ++i)
      make_cexp(&st->super_twiddles[i], -3.14159265358979323846264338327 * ((double)(i + 1) / st->ncfft + .5));
  }
  /* A.4 filterbank (the C source's aligned/padded weight layout only adds zero weights; the
     un-padded per-bin form below is arithmetically identical) */
  {
    const int nch1 = num_channels + 1;
    const int spectrum_size = fft_size / 2 + 1;
    float center[KWS_MAX_CHANNELS + 1];
    const float mel_low = freq_to_mel(lower_hz);
    const float mel_hi = freq_to_mel(upper_hz);
    const float mel_span = mel_hi - mel_low;
    const float mel_spacing = mel_span / ((float)nch1);
    for (int i = 0; i < nch1; ++i) center[i] = mel_low + (mel_spacing * (i + 1));
    const float hz_per_sbin = 0.5 * sample_rate / ((float)spectrum_size - 1);
    st->start_index = 1.5 + lower_hz / hz_per_sbin;
    st->end_index = 0;
    int chan_start = st->start_index;
    for (int chan = 0; chan < nch1; ++chan) {
      int freq_index = chan_start;
      while (freq_index < spectrum_size + 8 && freq_to_mel((freq_index)*hz_per_sbin) <= center[chan]) ++freq_index;
      st->band_start[chan] = chan_start;
      const float denom_val = (chan == 0) ? mel_low : center[chan - 1];
      for (int f = chan_start; f < freq_index; ++f) {
        if (f >= spectrum_size) { free(st); return NULL; } /* reference errors out too */
        const float weight = (center[chan] - freq_to_mel(f * hz_per_sbin)) / (center[chan] - denom_val);
        st->bin_band[f] = (int16_t)chan;
        st->bin_weight[f] = (int16_t)floor(weight * (1 << kFilterbankBits) + 0.5);
        st->bin_unweight[f] = (int16_t)floor((1.0 - weight) * (1 << kFilterbankBits) + 0.5);
      }
      if (freq_index > chan_start && freq_index > st->end_index) st->end_index = freq_index;
      chan_start = freq_index;
    }
    st->band_start[nch1] = chan_start;
    if (st->end_index >= spectrum_size) { free(st); return NULL; }
  }
  /* A.5 noise reduction */
  st->even_smoothing = (uint16_t)(even_smoothing * (1 << kNoiseReductionBits));
  st->odd_smoothing = (uint16_t)(odd_smoothing * (1 << kNoiseReductionBits));
  st->min_signal_remaining = (uint16_t)(min_signal_remaining * (1 << kNoiseReductionBits));
  /* A.6 PCAN LUT */
  st->correction_bits = msb32((uint32_t)fft_size) - 1 - (kFilterbankBits / 2);
  {
    const int32_t input_bits = smoothing_bits - st->correction_bits;
    st->snr_shift = gain_bits - st->correction_bits - kPcanSnrBits;
    int16_t* lut = st->gain_lut;
    lut[0] = pcan_gain_lookup(pcan_strength, pcan_offset, gain_bits, input_bits, 0);
    lut[1] = pcan_gain_lookup(pcan_strength, pcan_offset, gain_bits, input_bits, 1);
    for (int interval = 2; interval <= kWideDynamicFunctionBits; ++interval) {
      const uint32_t x0 = (uint32_t)1 << (interval - 1);
      const uint32_t x1 = x0 + (x0 >> 1);
      const uint32_t x2 = (interval == kWideDynamicFunctionBits) ? x0 + (x0 - 1) : 2 * x0;
      const int16_t y0 = pcan_gain_lookup(pcan_strength, pcan_offset, gain_bits, input_bits, x0);
      const int16_t y1 = pcan_gain_lookup(pcan_strength, pcan_offset, gain_bits, input_bits, x1);
      const int16_t y2 = pcan_gain_lookup(pcan_strength, pcan_offset, gain_bits, input_bits, x2);
      const int32_t diff1 = (int32_t)y1 - y0;
      const int32_t diff2 = (int32_t)y2 - y0;
      const int32_t a1 = 4 * diff1 - diff2;
      const int32_t a2 = diff2 - a1;
      lut[4 * interval - 6] = y0;
      lut[4 * interval - 6 + 1] = (int16_t)a1;
      lut[4 * interval - 6 + 2] = (int16_t)a2;
    }
  }
  /* A.7 log LUT: round(65536*(log2(1+i/128) - i/128)), i = 0..128, plus a trailing 0 */
  for (int i = 0; i <= 128; ++i) {
    double v = 65536.0 * (log2(1.0 + i / 128.0) - i / 128.0);
    st->log_lut[i] = (uint16_t)floor(v + 0.5);
  }
  st->log_lut[129] = 0;
  return st;
}

void kws_ref_frontend_destroy(kws_ref_frontend* st) { free(st); }

int kws_ref_frontend_num_frames(const kws_ref_frontend* st, int n_samples) {
  if (n_samples < st->window_size) return 0;
  return (n_samples - st->window_size) / st->window_step + 1;
}

/* ------------------------------------------------------------------ per-frame stages */
static int16_t wide_dynamic_function(uint32_t x, const int16_t* lut) {
  if (x <= 2) return lut[x];
  const int16_t interval = (int16_t)msb32(x);
  lut += 4 * interval - 6;
  const int16_t frac = (int16_t)(((interval < 11) ? (x << (11 - interval)) : (x >> (interval - 11))) & 0x3FF);
  int32_t result = ((int32_t)lut[2] * frac) >> 5;
  result += (int32_t)((uint32_t)lut[1] << 5);
  result *= frac;
  result = (result + (1 << 14)) >> 15;
  result += lut[0];
  return (int16_t)result;
}

static uint32_t pcan_shrink(uint32_t x) {
  if (x < (2u << kPcanSnrBits)) return (x * x) >> (2 + 2 * kPcanSnrBits - kPcanOutputBits);
  return (x >> (kPcanSnrBits - kPcanOutputBits)) - (1u << kPcanOutputBits);
}

static uint32_t log2_fraction_part(const uint16_t* log_lut, uint32_t x, uint32_t log2x) {
  int32_t frac = (int32_t)(x - (1ll << log2x));
  if (log2x < kLogScaleLog2) frac <<= kLogScaleLog2 - log2x;
  else frac >>= log2x - kLogScaleLog2;
  const uint32_t base_seg = (uint32_t)frac >> (kLogScaleLog2 - kLogSegmentsLog2);
  const uint32_t seg_unit = (((uint32_t)1) << kLogScaleLog2) >> kLogSegmentsLog2;
  const int32_t c0 = log_lut[base_seg];
  const int32_t c1 = log_lut[base_seg + 1];
  const int32_t seg_base = (int32_t)(seg_unit * base_seg);
  const int32_t rel_pos = ((c1 - c0) * (frac - seg_base)) >> kLogScaleLog2;
  return (uint32_t)(frac + c0 + rel_pos);
}

static uint32_t log_scale(const uint16_t* log_lut, uint32_t x, uint32_t scale_shift) {
  const uint32_t integer = (uint32_t)msb32(x) - 1;
  const uint32_t fraction = log2_fraction_part(log_lut, x, integer);
  const uint32_t log2 = (integer << kLogScaleLog2) + fraction;
  const uint32_t round = (1u << kLogScaleLog2) / 2;
  const uint32_t loge = (uint32_t)((((uint64_t)kLogCoeff) * log2 + round) >> kLogScaleLog2);
  return ((loge << scale_shift) + round) >> kLogScaleLog2;
}

/* Frame-local part: window → input_shift → FFT → energy → filterbank → sqrt >> shift.
 * Depends only on the 480 samples of the frame (the property the streaming frame-reuse path uses). */
void kws_ref_frontend_frame_magnitudes(const kws_ref_frontend* st, const int16_t* frame, uint32_t* scaled /*[nch]*/) {
  int16_t win[KWS_FFT_MAX];
  int16_t max_abs = 0;
  for (int i = 0; i < st->window_size; ++i) {
    int16_t v = (int16_t)((((int32_t)frame[i]) * st->window_coef[i]) >> kFrontendWindowBits);
    win[i] = v;
    if (v < 0) v = (int16_t)(-v);
    if (v > max_abs) max_abs = v;
  }
  const int input_shift = 15 - msb32((uint32_t)max_abs);
  int16_t fft_in[KWS_FFT_MAX];
  int i = 0;
  for (; i < st->window_size; ++i) fft_in[i] = (int16_t)(uint16_t)(((uint32_t)(uint16_t)win[i]) << input_shift);
  for (; i < st->fft_size; ++i) fft_in[i] = 0;
  cpx16 freq[KWS_FFT_MAX / 2 + 1];
  fftr(st, fft_in, freq);

  uint64_t work[KWS_MAX_CHANNELS + 2];
  uint64_t weight_acc = 0, unweight_acc = 0;
  int f = st->start_index;
  for (int chan = 0; chan <= st->num_channels; ++chan) {
    const int fend = st->band_start[chan + 1];
    for (; f < fend; ++f) {
      const int32_t re = freq[f].r, im = freq[f].i;
      const uint32_t mag_squared = (uint32_t)(re * re) + (uint32_t)(im * im);
      const int32_t energy = (int32_t)mag_squared; /* the C source re-reads it through an int32_t* */
      weight_acc += st->bin_weight[f] * ((uint64_t)energy);
      unweight_acc += st->bin_unweight[f] * ((uint64_t)energy);
    }
    work[chan] = weight_acc;
    weight_acc = unweight_acc;
    unweight_acc = 0;
  }
  for (int c = 0; c < st->num_channels; ++c) scaled[c] = sqrt64_ref(work[c + 1]) >> input_shift;
}

/* Sequential part for one frame given the running noise estimate. */
void kws_ref_frontend_frame_finish(const kws_ref_frontend* st, uint32_t* signal /*[nch] in/out scratch*/,
                                   uint32_t* estimate /*[nch] state*/, uint16_t* out /*[nch]*/) {
  for (int i = 0; i < st->num_channels; ++i) {
    const uint32_t smoothing = ((i & 1) == 0) ? st->even_smoothing : st->odd_smoothing;
    const uint32_t one_minus_smoothing = (1u << kNoiseReductionBits) - smoothing;
    const uint32_t signal_scaled_up = signal[i] << st->smoothing_bits;
    uint32_t est = (uint32_t)((((uint64_t)signal_scaled_up * smoothing) +
                               ((uint64_t)estimate[i] * one_minus_smoothing)) >> kNoiseReductionBits);
    estimate[i] = est;
    if (est > signal_scaled_up) est = signal_scaled_up;
    const uint32_t floor_ = (uint32_t)(((uint64_t)signal[i] * st->min_signal_remaining) >> kNoiseReductionBits);
    const uint32_t subtracted = (signal_scaled_up - est) >> st->smoothing_bits;
    signal[i] = subtracted > floor_ ? subtracted : floor_;
  }
  if (st->enable_pcan) {
    for (int i = 0; i < st->num_channels; ++i) {
      const uint32_t gain = (uint32_t)wide_dynamic_function(estimate[i], st->gain_lut);
      const uint32_t snr = (uint32_t)(((uint64_t)signal[i] * gain) >> st->snr_shift);
      signal[i] = pcan_shrink(snr);
    }
  }
  for (int i = 0; i < st->num_channels; ++i) {
    uint32_t value = signal[i];
    if (st->enable_log) {
      if (st->correction_bits < 0) value >>= -st->correction_bits;
      else value <<= st->correction_bits;
      value = (value > 1) ? log_scale(st->log_lut, value, (uint32_t)st->scale_shift) : 0;
    }
    out[i] = (value < 0xFFFF) ? (uint16_t)value : 0xFFFF;
  }
}

/* One clip: fresh zero state, all frames (the op allocates a new FrontendState per call). */
int kws_ref_frontend_clip(const kws_ref_frontend* st, const int16_t* pcm, int n_samples, uint16_t* out /*[frames,nch]*/) {
  const int frames = kws_ref_frontend_num_frames(st, n_samples);
  uint32_t estimate[KWS_MAX_CHANNELS];
  uint32_t signal[KWS_MAX_CHANNELS];
  memset(estimate, 0, sizeof(estimate));
  for (int t = 0; t < frames; ++t) {
    kws_ref_frontend_frame_magnitudes(st, pcm + (size_t)t * st->window_step, signal);
    kws_ref_frontend_frame_finish(st, signal, estimate, out + (size_t)t * st->num_channels);
  }
  return frames;
}

/* Batch of clips → float features = uint16 * (10/256)  (input_data.py:34). */
int kws_ref_frontend_batch(const kws_ref_frontend* st, const int16_t* pcm, int batch, int n_samples,
                           float* out_f32 /*[B,frames,nch] or NULL*/, uint16_t* out_u16 /*or NULL*/) {
  const int frames = kws_ref_frontend_num_frames(st, n_samples);
  const int nch = st->num_channels;
  uint16_t* tmp = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)(frames > 0 ? frames : 1) * nch);
  if (!tmp) return -1;
  for (int b = 0; b < batch; ++b) {
    kws_ref_frontend_clip(st, pcm + (size_t)b * n_samples, n_samples, tmp);
    for (int i = 0; i < frames * nch; ++i) {
      if (out_u16) out_u16[(size_t)b * frames * nch + i] = tmp[i];
      if (out_f32) out_f32[(size_t)b * frames * nch + i] = (float)tmp[i] * (10.0f / 256.0f);
    }
  }
  free(tmp);
  return frames;
}

/* ------------------------------------------------------------------ table access for tests and for
 * cross-checking the tables the CUDA library builds independently. */
int kws_ref_frontend_tables(const kws_ref_frontend* st, int16_t* window /*[window_size]*/,
                            int16_t* twiddles /*[ncfft*2]*/, int16_t* super_twiddles /*[ncfft/2*2]*/,
                            int16_t* bin_band, int16_t* bin_weight, int16_t* bin_unweight /*[fft/2+1]*/,
                            int16_t* gain_lut /*[125]*/, uint16_t* log_lut /*[130]*/, int32_t* scalars /*[16]*/) {
  if (window) memcpy(window, st->window_coef, sizeof(int16_t) * st->window_size);
  if (twiddles) memcpy(twiddles, st->twiddles, sizeof(cpx16) * st->ncfft);
  if (super_twiddles) memcpy(super_twiddles, st->super_twiddles, sizeof(cpx16) * (st->ncfft / 2));
  if (bin_band) memcpy(bin_band, st->bin_band, sizeof(int16_t) * (st->fft_size / 2 + 1));
  if (bin_weight) memcpy(bin_weight, st->bin_weight, sizeof(int16_t) * (st->fft_size / 2 + 1));
  if (bin_unweight) memcpy(bin_unweight, st->bin_unweight, sizeof(int16_t) * (st->fft_size / 2 + 1));
  if (gain_lut) memcpy(gain_lut, st->gain_lut, sizeof(st->gain_lut));
  if (log_lut) memcpy(log_lut, st->log_lut, sizeof(st->log_lut));
  if (scalars) {
    scalars[0] = st->window_size; scalars[1] = st->window_step; scalars[2] = st->fft_size;
    scalars[3] = st->start_index; scalars[4] = st->end_index; scalars[5] = st->even_smoothing;
    scalars[6] = st->odd_smoothing; scalars[7] = st->min_signal_remaining; scalars[8] = st->snr_shift;
    scalars[9] = st->correction_bits; scalars[10] = st->num_channels; scalars[11] = st->ncfft;
  }
  return 0;
}

/* Raw fixed-point real FFT of one already-scaled frame (for the FFT-vs-numpy consistency test). */
void kws_ref_fftr(const kws_ref_frontend* st, const int16_t* timedata /*[fft_size]*/, int16_t* freq /*[(fft/2+1)*2]*/) {
  fftr(st, timedata, (cpx16*)freq);
}
