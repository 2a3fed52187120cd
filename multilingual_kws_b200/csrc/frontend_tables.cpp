// frontend_tables.cpp — host-side construction of the fixed-point frontend tables.
// Product code: built into libkws_b200.so and uploaded to the GPU by kws_frontend_create.
// (The CPU oracle builds its own tables independently; tests cross-check the two.)
//
// Formulas: SURVEY.md Appendix A.2-A.7 (TF 2.7 microfrontend `*_util.c` initialisers), evaluated
// with the same float/double mix an x86-64 build of the op uses; compile with -ffp-contract=off.
#include "frontend_tables.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace kws {

static float freq_to_mel(float freq) { return 1127.0 * log1p(freq / 700.0); }

static uint32_t quant_cexp(double phase) {
  const int16_t r = (int16_t)floor(.5 + 32767 * cos(phase));
  const int16_t i = (int16_t)floor(.5 + 32767 * sin(phase));
  return pack16(r, i);
}

static int16_t pcan_gain(float strength, float offset, int gain_bits, int32_t input_bits, uint32_t x) {
  const float x_as_float = ((float)x) / ((uint32_t)1 << input_bits);
  const float gain_as_float = ((uint32_t)1 << gain_bits) * powf(x_as_float + offset, -strength);
  if (gain_as_float > 32767) return 32767;
  return (int16_t)(gain_as_float + 0.5f);
}

const char* build_frontend_tables(const FrontendConfig& cfg, FrontendTables* T) {
  memset(T, 0, sizeof(*T));
  const int window_size = cfg.window_ms * cfg.sample_rate / 1000;
  const int window_step = cfg.step_ms * cfg.sample_rate / 1000;
  if (cfg.num_channels < 1 || cfg.num_channels > kMaxChannels) return "num_channels must be in [1, 64]";
  if (window_size <= kFftSize / 2 || window_size > kFftSize)
    return "CUDA frontend is specialised for fft_size 512: need 256 < window_size_samples <= 512";
  if (window_step < 2 || (window_step & 1)) return "window_step_samples must be even (32-bit frame addressing)";
  if (cfg.smoothing_bits < 0 || cfg.smoothing_bits > 31) return "smoothing_bits out of range";
  T->window_size = window_size;
  T->window_step = window_step;
  T->num_channels = cfg.num_channels;
  T->smoothing_bits = cfg.smoothing_bits;
  T->enable_pcan = cfg.enable_pcan;
  T->enable_log = cfg.enable_log;
  T->scale_shift = cfg.scale_shift;

  // A.2 Hann window, 12-bit
  {
    int16_t coef[kFftSize] = {0};
    const float arg = M_PI * 2.0 / ((float)window_size);
    for (int i = 0; i < window_size; ++i) {
      float float_value = 0.5 - (0.5 * cos(arg * (i + 0.5)));
      coef[i] = (int16_t)floor(float_value * (1 << 12) + 0.5);
    }
    for (int i = 0; i < kFftSize / 2; ++i) T->window_pairs[i] = pack16(coef[2 * i], coef[2 * i + 1]);
  }
  // A.3 kissfft twiddles (complex FFT of 256) and real-FFT super twiddles
  {
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    for (int i = 0; i < kNcfft; ++i) T->twiddles[i] = quant_cexp(-2 * pi * i / kNcfft);
    for (int i = 0; i < kNcfft / 2; ++i)
      T->super_twiddles[i] = quant_cexp(-3.14159265358979323846264338327 * ((double)(i + 1) / kNcfft + .5));
  }
  // A.4 mel filterbank
  std::vector<int> widths(cfg.num_channels + 1, 0);
  {
    const int nch1 = cfg.num_channels + 1;
    const int spectrum_size = kFftSize / 2 + 1;
    float center[kMaxChannels + 1];
    const float mel_low = freq_to_mel(cfg.lower_hz);
    const float mel_hi = freq_to_mel(cfg.upper_hz);
    if (!(cfg.lower_hz >= 0.0f) || !(cfg.upper_hz > cfg.lower_hz)) return "bad band limits";
    const float mel_span = mel_hi - mel_low;
    const float mel_spacing = mel_span / ((float)nch1);
    for (int i = 0; i < nch1; ++i) center[i] = mel_low + (mel_spacing * (i + 1));
    const float hz_per_sbin = 0.5 * cfg.sample_rate / ((float)spectrum_size - 1);
    T->start_index = 1.5 + cfg.lower_hz / hz_per_sbin;
    T->end_index = 0;
    int chan_start = T->start_index;
    for (int chan = 0; chan < nch1; ++chan) {
      int freq_index = chan_start;
      while (freq_index < spectrum_size + 8 && freq_to_mel((freq_index)*hz_per_sbin) <= center[chan]) ++freq_index;
      if (freq_index > spectrum_size - 1) return "filterbank upper limit reaches the Nyquist bin";
      T->band_start[chan] = (int16_t)chan_start;
      widths[chan] = freq_index - chan_start;
      const float denom_val = (chan == 0) ? mel_low : center[chan - 1];
      for (int f = chan_start; f < freq_index; ++f) {
        const float weight = (center[chan] - freq_to_mel(f * hz_per_sbin)) / (center[chan] - denom_val);
        T->bin_weight[f] = (int16_t)floor(weight * (1 << 12) + 0.5);
        T->bin_unweight[f] = (int16_t)floor((1.0 - weight) * (1 << 12) + 0.5);
      }
      if (freq_index > chan_start && freq_index > T->end_index) T->end_index = freq_index;
      chan_start = freq_index;
    }
    T->band_start[nch1] = (int16_t)chan_start;
  }
  // band -> lane schedule (longest-processing-time first) for the 16 lanes of a frame's half-warp
  {
    const int nb = cfg.num_channels + 1;
    std::vector<int> order(nb);
    for (int i = 0; i < nb; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return widths[a] > widths[b]; });
    int load[kHalfWarp] = {0}, count[kHalfWarp] = {0};
    memset(T->lane_bands, -1, sizeof(T->lane_bands));
    for (int band : order) {
      int best = -1;
      for (int l = 0; l < kHalfWarp; ++l)
        if (count[l] < kMaxLaneBands && (best < 0 || load[l] < load[best])) best = l;
      if (best < 0) return "band schedule overflow";
      T->lane_bands[best][count[best]++] = (int8_t)band;
      load[best] += widths[band] + 2;   // +2: per-band loop / store overhead
    }
  }
  // A.5 noise reduction constants
  T->even_smoothing = (uint16_t)(cfg.even_smoothing * (1 << 14));
  T->odd_smoothing = (uint16_t)(cfg.odd_smoothing * (1 << 14));
  T->min_signal_remaining = (uint16_t)(cfg.min_signal_remaining * (1 << 14));
  // A.6 PCAN gain LUT
  T->correction_bits = 10 - 1 - (12 / 2);   // MSB32(512) - 1 - kFilterbankBits/2
  {
    const int32_t input_bits = cfg.smoothing_bits - T->correction_bits;
    T->snr_shift = cfg.gain_bits - T->correction_bits - 12;
    if (input_bits < 0 || input_bits > 31 || T->snr_shift < 0 || T->snr_shift > 63 || cfg.gain_bits < 0 || cfg.gain_bits > 31)
      return "pcan bit widths out of range";
    int16_t* lut = T->gain_lut;
    lut[0] = pcan_gain(cfg.pcan_strength, cfg.pcan_offset, cfg.gain_bits, input_bits, 0);
    lut[1] = pcan_gain(cfg.pcan_strength, cfg.pcan_offset, cfg.gain_bits, input_bits, 1);
    for (int interval = 2; interval <= 32; ++interval) {
      const uint32_t x0 = (uint32_t)1 << (interval - 1);
      const uint32_t x1 = x0 + (x0 >> 1);
      const uint32_t x2 = (interval == 32) ? x0 + (x0 - 1) : 2 * x0;
      const int16_t y0 = pcan_gain(cfg.pcan_strength, cfg.pcan_offset, cfg.gain_bits, input_bits, x0);
      const int16_t y1 = pcan_gain(cfg.pcan_strength, cfg.pcan_offset, cfg.gain_bits, input_bits, x1);
      const int16_t y2 = pcan_gain(cfg.pcan_strength, cfg.pcan_offset, cfg.gain_bits, input_bits, x2);
      const int32_t diff1 = (int32_t)y1 - y0;
      const int32_t diff2 = (int32_t)y2 - y0;
      const int32_t a1 = 4 * diff1 - diff2;
      const int32_t a2 = diff2 - a1;
      lut[4 * interval - 6] = y0;
      lut[4 * interval - 5] = (int16_t)a1;
      lut[4 * interval - 4] = (int16_t)a2;
    }
  }
  // A.7 log2 fraction LUT
  for (int i = 0; i <= 128; ++i) T->log_lut[i] = (uint16_t)floor(65536.0 * (log2(1.0 + i / 128.0) - i / 128.0) + 0.5);
  T->log_lut[129] = 0;
  return nullptr;
}

}  // namespace kws
