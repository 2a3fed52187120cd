// frontend_tables.h — host-side table builder for the fixed-point frontend (product code).
#pragma once
#include "frontend_core.cuh"

namespace kws {

// Attributes of TF's audio_microfrontend op; the reference
// (multilingual_kws/embedding/input_data.py:25-33) overrides only the first four.
struct FrontendConfig {
  int sample_rate = 16000, window_ms = 30, step_ms = 20, num_channels = 40;
  float lower_hz = 125.0f, upper_hz = 7500.0f;
  int smoothing_bits = 10;
  float even_smoothing = 0.025f, odd_smoothing = 0.06f, min_signal_remaining = 0.05f;
  int enable_pcan = 1;
  float pcan_strength = 0.95f, pcan_offset = 80.0f;
  int gain_bits = 21, enable_log = 1, scale_shift = 6;
};

// Fills *T; returns nullptr on success or a static error string.
const char* build_frontend_tables(const FrontendConfig& cfg, FrontendTables* T);

}  // namespace kws
