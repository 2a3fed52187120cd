// tests/emu/frontend_emulator.cpp — TEST-ONLY host replay of the CUDA frontend's half-warp
// choreography (phases P0..P5 of csrc/frontend_core.cuh executed lane by lane on the CPU).
// Lets the no-GPU test suite check the kernel's indexing / padding / twiddle selection / band
// schedule bit-for-bit against the oracle.  Never linked into libkws_b200.so.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../multilingual_kws_b200/csrc/frontend_core.cuh"
#include "../../multilingual_kws_b200/csrc/frontend_tables.h"

using namespace kws;

static void frame_mags(const FrontendTables& T, const uint32_t* frame_words, uint32_t* mags) {
  LaneRegs R[kHalfWarp];
  uint32_t fftbuf[kFftBufWords];
  uint32_t energy[kEnergyWords];
  uint64_t WU[2 * (kMaxChannels + 1)];
  memset(fftbuf, 0xAB, sizeof(fftbuf));
  memset(energy, 0xCD, sizeof(energy));
  memset(WU, 0xEF, sizeof(WU));
  int32_t mx = 0;
  for (int l = 0; l < kHalfWarp; ++l) { fe_p0_window(l, frame_words, T, R[l]); mx = R[l].local_max > mx ? R[l].local_max : mx; }
  const int shift = 15 - msb32((uint32_t)mx);
  for (int l = 0; l < kHalfWarp; ++l) fe_p1_fft_pass1(l, shift, T.twiddles, R[l], fftbuf);
  for (int l = 0; l < kHalfWarp; ++l) { uint32_t tw2[15]; fe_load_tw2(l, T.twiddles, tw2); fe_p2_fft_pass2(l, tw2, fftbuf); }
  for (int l = 0; l < kHalfWarp; ++l) fe_p3_real_energy(l, fftbuf, T.super_twiddles, energy);
  for (int l = 0; l < kHalfWarp; ++l) fe_p4_band_sums(l, energy, T, WU);
  for (int c = 0; c < T.num_channels; ++c) mags[c] = fe_p5_channel(c, shift, WU);
}

extern "C" int emu_frontend_batch(int sample_rate, int window_ms, int step_ms, int num_channels, float lower_hz,
                                  float upper_hz, int smoothing_bits, float even_s, float odd_s, float min_sig,
                                  int enable_pcan, float pcan_strength, float pcan_offset, int gain_bits,
                                  int enable_log, int scale_shift, const int16_t* pcm, int B, int n_samples,
                                  uint16_t* out, uint32_t* mags_out /*or null*/) {
  FrontendConfig cfg;
  cfg.sample_rate = sample_rate; cfg.window_ms = window_ms; cfg.step_ms = step_ms; cfg.num_channels = num_channels;
  cfg.lower_hz = lower_hz; cfg.upper_hz = upper_hz; cfg.smoothing_bits = smoothing_bits; cfg.even_smoothing = even_s;
  cfg.odd_smoothing = odd_s; cfg.min_signal_remaining = min_sig; cfg.enable_pcan = enable_pcan;
  cfg.pcan_strength = pcan_strength; cfg.pcan_offset = pcan_offset; cfg.gain_bits = gain_bits;
  cfg.enable_log = enable_log; cfg.scale_shift = scale_shift;
  static FrontendTables T;
  if (build_frontend_tables(cfg, &T)) return -1;
  const int frames = n_samples < T.window_size ? 0 : (n_samples - T.window_size) / T.window_step + 1;
  const int C = T.num_channels;
  std::vector<uint32_t> mags((size_t)(frames > 0 ? frames : 1) * C), est(C);
  for (int b = 0; b < B; ++b) {
    const int16_t* clip = pcm + (size_t)b * n_samples;
    for (int t = 0; t < frames; ++t) frame_mags(T, (const uint32_t*)(clip + (size_t)t * T.window_step), &mags[(size_t)t * C]);
    std::fill(est.begin(), est.end(), 0u);
    for (int t = 0; t < frames; ++t)
      for (int c = 0; c < C; ++c) {
        const uint32_t s = mags[(size_t)t * C + c];
        est[c] = fe_noise_estimate(s, est[c], c, T);
        out[((size_t)b * frames + t) * C + c] = (uint16_t)fe_pointwise(s, est[c], T);
        if (mags_out) mags_out[((size_t)b * frames + t) * C + c] = s;
      }
  }
  return frames;
}

extern "C" int emu_frontend_tables(FrontendTables* out) {
  FrontendConfig cfg;
  return build_frontend_tables(cfg, out) ? -1 : (int)sizeof(FrontendTables);
}

extern "C" uint32_t emu_isqrt64_round(uint64_t x) { return isqrt64_round(x); }
