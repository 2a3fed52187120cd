// frontend.cu — sm_100a kernels + C ABI for the fixed-point micro-frontend.
//
// Replaces the TF `audio_microfrontend` CPU op at reference
// multilingual_kws/embedding/input_data.py:25-33 (clip batches) and its per-window use at
// multilingual_kws/embedding/batch_streaming_analysis.py:108-115 (streaming).
//
// Kernels
//   frontend_clip_kernel     one CTA per clip.  The clip's PCM (31 680 B for 1 s) is staged into
//                            shared memory by ONE TMA bulk copy (cp.async.bulk + mbarrier) while the
//                            CTA copies the 5 KB table block; 8 half-warps then each own a frame per
//                            round (P0..P5 of frontend_core.cuh, radix-16 register FFT passes,
//                            half-warp shuffles for max|x|); the 49-step noise-estimate recurrence
//                            runs on 40 threads out of smem, and the pointwise tail (spectral
//                            subtraction, PCAN, log) is spread over all threads and written out
//                            coalesced.  HBM traffic = PCM in + features out (39 840 B / clip).
//   frontend_frame_mags_kernel / frontend_window_tail_kernel
//                            the same arithmetic split at the frame/window boundary for long signals:
//                            magnitudes once per 20 ms frame, then the per-window recurrence + tail.
#include <stdint.h>
#include <string.h>

#include <new>

#include "common.h"
#include "frontend_core.cuh"
#include "frontend_tables.h"
#include "ptx.cuh"

namespace kws {

__constant__ uint32_t c_twiddles[kNcfft];   // fft-512 twiddles do not depend on the op's attributes

struct __align__(16) FrameScratch {
  uint32_t fftbuf[kFftBufWords];   // later aliased by the band sums W/U (2*(C+1) uint64 <= 1040 B)
  uint32_t energy[kEnergyWords];
};
static_assert(sizeof(uint64_t) * 2 * (kMaxChannels + 1) <= sizeof(uint32_t) * kFftBufWords, "WU alias");
static_assert(sizeof(FrontendTables) % 16 == 0, "tables are copied as uint4");

constexpr int kClipThreads = 128;
constexpr int kHalfWarpsPerCta = kClipThreads / kHalfWarp;

using ptx::mbar_expect_tx;
using ptx::mbar_wait;
using ptx::tma_bulk_g2s;

// ---------------------------------------------------------------- one frame on one half-warp
__device__ __forceinline__ void halfwarp_frame_mags(const uint32_t* frame_words, const FrontendTables& T,
                                                    const uint32_t* tw2, FrameScratch& S, int lane,
                                                    uint32_t* mags_out, bool store) {
  LaneRegs R;
  fe_p0_window(lane, frame_words, T, R);
  int mx = R.local_max;
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));   // stays inside the half-warp
  const int shift = 15 - msb32((uint32_t)mx);
  fe_p1_fft_pass1(lane, shift, c_twiddles, R, S.fftbuf);
  __syncwarp();
  fe_p2_fft_pass2(lane, tw2, S.fftbuf);
  __syncwarp();
  fe_p3_real_energy(lane, S.fftbuf, T.super_twiddles, S.energy);
  __syncwarp();
  uint64_t* WU = reinterpret_cast<uint64_t*>(S.fftbuf);
  fe_p4_band_sums(lane, S.energy, T, WU);
  __syncwarp();
  for (int c = lane; c < T.num_channels; c += kHalfWarp) {
    const uint32_t v = fe_p5_channel(c, shift, WU);
    if (store) mags_out[c] = v;
  }
  __syncwarp();
}

__device__ __forceinline__ void load_tables(FrontendTables* dst_smem, const FrontendTables* src, int tid, int nthreads) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst_smem);
  for (int i = tid; i < (int)(sizeof(FrontendTables) / 16); i += nthreads) d[i] = __ldg(s + i);
}

// ---------------------------------------------------------------- fused clip kernel
__global__ void __launch_bounds__(kClipThreads, 5)   // <= 96 registers: five CTAs per SM (shared memory allows five)
frontend_clip_kernel(const int16_t* __restrict__ pcm, int n_samples, int n_frames, int num_channels,
                     const FrontendTables* __restrict__ tables, float out_scale, float* __restrict__ out_f32,
                     uint16_t* __restrict__ out_u16, int stage_bytes, int region_bytes, int use_tma, int slot_bytes,
                     int T_step, int window) {
  extern __shared__ __align__(128) uint8_t smem[];
  FrontendTables& T = *reinterpret_cast<FrontendTables*>(smem);
  uint8_t* p = smem + sizeof(FrontendTables);
  int16_t* s_pcm = reinterpret_cast<int16_t*>(p);            p += region_bytes;
  FrameScratch* scratch = reinterpret_cast<FrameScratch*>(p); p += sizeof(FrameScratch) * kHalfWarpsPerCta;
  uint32_t* s_mags = reinterpret_cast<uint32_t*>(p);          p += ((n_frames * num_channels * 4 + 15) & ~15);
  uint64_t* bar = reinterpret_cast<uint64_t*>(p);      // two barriers (ring slots)

  const int tid = threadIdx.x;
  const int clip = blockIdx.x;
  const int16_t* clip_pcm = pcm + (size_t)clip * n_samples;

  // PCM staging.  TMA path: a two-slot ring — round r (8 frames = 7 steps + one window of samples, 5 440 B for the
  // reference's 30 ms / 20 ms framing) is bulk-copied into slot r & 1 two rounds ahead of its use, so the CTA holds
  // 11 KB of PCM instead of the whole 32 KB clip and five CTAs (20 warps) share an SM instead of three.
  const int step = T_step;
  const int rounds = (n_frames + kHalfWarpsPerCta - 1) / kHalfWarpsPerCta;
  const int slot_samples = slot_bytes / 2;
  auto issue_round = [&](int r) {                      // one thread
    const int first = r * kHalfWarpsPerCta, last = min(first + kHalfWarpsPerCta, n_frames) - 1;
    const uint32_t bytes = (uint32_t)((((last - first) * step + window) * 2 + 15) & ~15);
    mbar_expect_tx(bar + (r & 1), bytes);
    tma_bulk_g2s(s_pcm + (r & 1) * slot_samples, clip_pcm + (size_t)first * step, bytes, bar + (r & 1));
  };
  if (use_tma) {
    if (tid == 0) {
      ptx::mbar_init(bar, 1);
      ptx::mbar_init(bar + 1, 1);
      ptx::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      issue_round(0);
      if (rounds > 1) issue_round(1);
    }
  } else {
    const int n = stage_bytes / 2;
    for (int i = tid; i < n; i += kClipThreads) s_pcm[i] = i < n_samples ? clip_pcm[i] : (int16_t)0;
  }
  load_tables(&T, tables, tid, kClipThreads);          // overlaps the TMA transfer
  const int lane = tid & (kHalfWarp - 1);
  const int hw = tid >> 4;
  uint32_t tw2[15];
  fe_load_tw2(lane, tables->twiddles, tw2);
  __syncthreads();

  // Phase A: frames -> magnitudes (half-warp per frame)
  for (int r = 0; r < rounds; ++r) {
    const int base = r * kHalfWarpsPerCta;
    const int f = base + hw;
    const bool valid = f < n_frames;
    const int ff = valid ? f : base;
    const int16_t* frame = use_tma ? s_pcm + (r & 1) * slot_samples + (ff - base) * step : s_pcm + ff * step;
    if (use_tma) mbar_wait(bar + (r & 1), (uint32_t)(r >> 1) & 1u);
    halfwarp_frame_mags(reinterpret_cast<const uint32_t*>(frame), T, tw2, scratch[hw], lane,
                        s_mags + ff * num_channels, valid);
    if (use_tma && r + 2 < rounds) {
      __syncthreads();                                 // every half-warp has left slot r & 1
      if (tid == 0) issue_round(r + 2);
    }
  }
  __syncthreads();

  // Phase B: the sequential noise-estimate recurrence, one thread per channel (PCM region is dead)
  uint32_t* s_est = reinterpret_cast<uint32_t*>(s_pcm);
  if (tid < num_channels) {
    uint32_t est = 0;
    for (int t = 0; t < n_frames; ++t) {
      est = fe_noise_estimate(s_mags[t * num_channels + tid], est, tid, T);
      s_est[t * num_channels + tid] = est;
    }
  }
  __syncthreads();

  // Phase C: pointwise tail, coalesced stores
  const int total = n_frames * num_channels;
  const size_t obase = (size_t)clip * total;
  for (int e = tid; e < total; e += kClipThreads) {
    const uint32_t v = fe_pointwise(s_mags[e], s_est[e], T);
    if (out_f32) out_f32[obase + e] = (float)v * out_scale;
    if (out_u16) out_u16[obase + e] = (uint16_t)v;
  }
}

// ---------------------------------------------------------------- split path: per-frame magnitudes
// Frame g lives in row g / frames_per_row at sample offset (g % frames_per_row) * step.
__global__ void __launch_bounds__(kClipThreads)
frontend_frame_mags_kernel(const int16_t* __restrict__ pcm, long long n_frames_total, int frames_per_row,
                           long long row_stride_samples, int num_channels,
                           const FrontendTables* __restrict__ tables, uint32_t* __restrict__ mags) {
  extern __shared__ __align__(128) uint8_t smem[];
  FrontendTables& T = *reinterpret_cast<FrontendTables*>(smem);
  FrameScratch* scratch = reinterpret_cast<FrameScratch*>(smem + sizeof(FrontendTables));
  const int tid = threadIdx.x, lane = tid & (kHalfWarp - 1), hw = tid >> 4;
  load_tables(&T, tables, tid, kClipThreads);
  uint32_t tw2[15];
  fe_load_tw2(lane, tables->twiddles, tw2);
  __syncthreads();
  const int step = T.window_step;
  for (long long base = (long long)blockIdx.x * kHalfWarpsPerCta; base < n_frames_total;
       base += (long long)gridDim.x * kHalfWarpsPerCta) {
    const long long g = base + hw;
    const bool valid = g < n_frames_total;
    const long long gg = valid ? g : 0;
    const long long row = gg / frames_per_row;
    const int t = (int)(gg - row * frames_per_row);
    const int16_t* frame = pcm + row * row_stride_samples + (long long)t * step;
    halfwarp_frame_mags(reinterpret_cast<const uint32_t*>(frame), T, tw2, scratch[hw], lane,
                        mags + gg * num_channels, valid);
  }
}

// ---------------------------------------------------------------- split path: per-window recurrence + tail
constexpr int kTailThreads = 256;
constexpr int kTailGroup = 64;                       // threads per window
constexpr int kTailWindows = kTailThreads / kTailGroup;

// Window w (global index first_window + w) of row r starts at frame r*frames_per_row + wi*hop_frames.
template <bool kSeq>
__global__ void __launch_bounds__(kTailThreads)
frontend_window_tail_kernel(const uint32_t* __restrict__ mags, long long first_window, long long n_windows,
                            long long windows_per_row, int frames_per_row, int hop_frames, int win_frames,
                            int num_channels, const FrontendTables* __restrict__ tables, float out_scale,
                            float* __restrict__ out_f32, uint16_t* __restrict__ out_u16) {
  extern __shared__ __align__(128) uint8_t smem[];
  FrontendTables& T = *reinterpret_cast<FrontendTables*>(smem);
  uint32_t* s_est_all = reinterpret_cast<uint32_t*>(smem + sizeof(FrontendTables));
  const int tid = threadIdx.x, grp = tid / kTailGroup, gt = tid % kTailGroup;
  const int total = win_frames * num_channels;
  uint32_t* s_est = s_est_all + (size_t)grp * total;
  load_tables(&T, tables, tid, kTailThreads);
  __syncthreads();
  for (long long base = (long long)blockIdx.x * kTailWindows; base < n_windows;
       base += (long long)gridDim.x * kTailWindows) {
    const long long w = base + grp;
    const bool valid = w < n_windows;
    const long long gw = first_window + (valid ? w : 0);
    const long long row = gw / windows_per_row;
    const long long wi = gw - row * windows_per_row;
    const uint32_t* m = mags + (row * frames_per_row + wi * hop_frames) * num_channels;
    if (kSeq) {   // windows too long to stage the estimates in smem: recurrence + tail per channel thread
      if (valid && gt < num_channels) {
        uint32_t est = 0;
        const size_t obase = (size_t)w * total;
        for (int t = 0; t < win_frames; ++t) {
          const uint32_t sig = __ldg(m + t * num_channels + gt);
          est = fe_noise_estimate(sig, est, gt, T);
          const uint32_t v = fe_pointwise(sig, est, T);
          if (out_f32) out_f32[obase + t * num_channels + gt] = (float)v * out_scale;
          if (out_u16) out_u16[obase + t * num_channels + gt] = (uint16_t)v;
        }
      }
      continue;
    }
    if (valid && gt < num_channels) {
      uint32_t est = 0;
      for (int t = 0; t < win_frames; ++t) {
        est = fe_noise_estimate(__ldg(m + t * num_channels + gt), est, gt, T);
        s_est[t * num_channels + gt] = est;
      }
    }
    __syncthreads();
    if (valid) {
      const size_t obase = (size_t)w * total;
      for (int e = gt; e < total; e += kTailGroup) {
        const uint32_t v = fe_pointwise(__ldg(m + e), s_est[e], T);
        if (out_f32) out_f32[obase + e] = (float)v * out_scale;
        if (out_u16) out_u16[obase + e] = (uint16_t)v;
      }
    }
    __syncthreads();
  }
}

}  // namespace kws

// ==================================================================== C ABI
using namespace kws;

struct kws_frontend {
  FrontendTables host;
  FrontendTables* dev = nullptr;
  int device = -1;
  int max_smem_optin = 0;
  int sm_count = 0;
};

extern "C" int kws_frontend_create(kws_frontend_t** out, int sample_rate, int window_ms, int step_ms,
                                   int num_channels, float lower_hz, float upper_hz, int smoothing_bits,
                                   float even_smoothing, float odd_smoothing, float min_signal_remaining,
                                   int enable_pcan, float pcan_strength, float pcan_offset, int gain_bits,
                                   int enable_log, int scale_shift) {
  KWS_REQUIRE(out != nullptr, "kws_frontend_create: out is NULL");
  *out = nullptr;
  FrontendConfig cfg;
  cfg.sample_rate = sample_rate; cfg.window_ms = window_ms; cfg.step_ms = step_ms; cfg.num_channels = num_channels;
  cfg.lower_hz = lower_hz; cfg.upper_hz = upper_hz; cfg.smoothing_bits = smoothing_bits;
  cfg.even_smoothing = even_smoothing; cfg.odd_smoothing = odd_smoothing; cfg.min_signal_remaining = min_signal_remaining;
  cfg.enable_pcan = enable_pcan; cfg.pcan_strength = pcan_strength; cfg.pcan_offset = pcan_offset;
  cfg.gain_bits = gain_bits; cfg.enable_log = enable_log; cfg.scale_shift = scale_shift;
  kws_frontend* fe = new (std::nothrow) kws_frontend();
  KWS_REQUIRE(fe != nullptr, "out of host memory");
  if (const char* err = build_frontend_tables(cfg, &fe->host)) {
    set_error("kws_frontend_create: %s", err);
    delete fe;
    return KWS_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaGetDevice(&fe->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&fe->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, fe->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&fe->sm_count, cudaDevAttrMultiProcessorCount, fe->device);
  if (e == cudaSuccess) e = cudaMalloc(&fe->dev, sizeof(FrontendTables));
  if (e == cudaSuccess) e = cudaMemcpy(fe->dev, &fe->host, sizeof(FrontendTables), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_twiddles, fe->host.twiddles, sizeof(uint32_t) * kNcfft);
  if (e != cudaSuccess) {
    set_error("kws_frontend_create: CUDA device required (%s); there is no CPU fallback", cudaGetErrorString(e));
    if (fe->dev) cudaFree(fe->dev);
    delete fe;
    return KWS_ERR_CUDA;
  }
  *out = fe;
  return KWS_OK;
}

extern "C" void kws_frontend_destroy(kws_frontend_t* fe) {
  if (!fe) return;
  if (fe->dev) cudaFree(fe->dev);
  delete fe;
}

static int num_frames_of(const FrontendTables& T, long long n_samples) {
  if (n_samples < T.window_size) return 0;
  return (int)((n_samples - T.window_size) / T.window_step + 1);
}

extern "C" int kws_frontend_num_frames(const kws_frontend_t* fe, int n_samples) {
  if (!fe) return KWS_ERR_ARG;
  return num_frames_of(fe->host, n_samples);
}

extern "C" int kws_frontend_tables(const kws_frontend_t* fe, void* out, size_t out_bytes, size_t* needed) {
  KWS_REQUIRE(fe != nullptr, "kws_frontend_tables: NULL handle");
  if (needed) *needed = sizeof(FrontendTables);
  if (out) {
    KWS_REQUIRE(out_bytes >= sizeof(FrontendTables), "kws_frontend_tables: buffer too small");
    memcpy(out, &fe->host, sizeof(FrontendTables));
  }
  return KWS_OK;
}

static int launch_split(kws_frontend* fe, const int16_t* d_pcm, long long rows, long long row_stride, int frames_per_row,
                        bool run_mags, uint32_t* d_mags, long long first_window, long long n_windows,
                        long long windows_per_row, int hop_frames, int win_frames, float out_scale, float* d_out_f32,
                        uint16_t* d_out_u16, cudaStream_t st) {
  const FrontendTables& T = fe->host;
  const int C = T.num_channels;
  if (run_mags) {
    const long long total = rows * frames_per_row;
    const size_t smem = sizeof(FrontendTables) + sizeof(FrameScratch) * kHalfWarpsPerCta;
    KWS_CUDA_CHECK(cudaFuncSetAttribute(frontend_frame_mags_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (total + kHalfWarpsPerCta - 1) / kHalfWarpsPerCta;
    const long long cap = (long long)fe->sm_count * 16;
    if (grid > cap) grid = cap;
    if (grid > 0) {
      frontend_frame_mags_kernel<<<(unsigned)grid, kClipThreads, smem, st>>>(d_pcm, total, frames_per_row, row_stride, C,
                                                                           fe->dev, d_mags);
      KWS_CUDA_CHECK(cudaGetLastError());
    }
  }
  if (n_windows > 0) {
    size_t smem = sizeof(FrontendTables) + (size_t)kTailWindows * win_frames * C * 4;
    const bool seq = smem > (size_t)fe->max_smem_optin;
    if (seq) smem = sizeof(FrontendTables);
    auto kern = seq ? frontend_window_tail_kernel<true> : frontend_window_tail_kernel<false>;
    KWS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (n_windows + kTailWindows - 1) / kTailWindows;
    const long long cap = (long long)fe->sm_count * 8;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, kTailThreads, smem, st>>>(d_mags, first_window, n_windows, windows_per_row, frames_per_row,
                                                    hop_frames, win_frames, C, fe->dev, out_scale, d_out_f32, d_out_u16);
    KWS_CUDA_CHECK(cudaGetLastError());
  }
  return KWS_OK;
}

extern "C" int kws_frontend_forward(kws_frontend_t* fe, const int16_t* d_pcm, int batch, int n_samples, float out_scale,
                                    float* d_out_f32, uint16_t* d_out_u16, void* stream) {
  KWS_REQUIRE(fe != nullptr, "kws_frontend_forward: NULL handle");
  KWS_REQUIRE(batch >= 0 && n_samples >= 0, "kws_frontend_forward: negative size");
  KWS_REQUIRE(d_out_f32 != nullptr || d_out_u16 != nullptr, "kws_frontend_forward: no output buffer");
  const FrontendTables& T = fe->host;
  const int n_frames = num_frames_of(T, n_samples);
  if (batch == 0 || n_frames == 0) return KWS_OK;   // the op returns an empty [0, C] tensor
  KWS_REQUIRE(d_pcm != nullptr, "kws_frontend_forward: d_pcm is NULL");
  KWS_REQUIRE(((uintptr_t)d_pcm & 3) == 0, "kws_frontend_forward: d_pcm must be 4-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int C = T.num_channels;
  const int needed = (n_frames - 1) * T.window_step + T.window_size;       // samples actually consumed
  const bool tma_ok = ((size_t)n_samples * 2 % 16 == 0) && (((uintptr_t)d_pcm & 15) == 0);
  const int stage_bytes = (int)round_up((size_t)needed * 2, 16);
  // TMA path: two ring slots of one 8-frame round each; fallback (unaligned PCM): the whole clip.  The region is
  // reused for the noise estimates after the magnitudes are done, so it is at least n_frames * C words.
  const int slot_bytes = (int)round_up((size_t)((kHalfWarpsPerCta - 1) * T.window_step + T.window_size) * 2, 16);
  const int pcm_bytes = tma_ok ? 2 * slot_bytes : stage_bytes;
  const int region_bytes = (int)round_up(pcm_bytes > n_frames * C * 4 ? pcm_bytes : n_frames * C * 4, 16);
  const size_t smem = sizeof(FrontendTables) + region_bytes + sizeof(FrameScratch) * kHalfWarpsPerCta +
                      round_up((size_t)n_frames * C * 4, 16) + 32;
  if (smem <= (size_t)fe->max_smem_optin) {
    KWS_CUDA_CHECK(cudaFuncSetAttribute(frontend_clip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    frontend_clip_kernel<<<batch, kClipThreads, smem, st>>>(d_pcm, n_samples, n_frames, C, fe->dev, out_scale, d_out_f32,
                                                          d_out_u16, stage_bytes, region_bytes, tma_ok ? 1 : 0, slot_bytes,
                                                          T.window_step, T.window_size);
    KWS_CUDA_CHECK(cudaGetLastError());
    return KWS_OK;
  }
  // Long clips: split path with a temporary magnitude buffer (stream-ordered allocation).
  KWS_REQUIRE(n_samples % 2 == 0, "kws_frontend_forward: odd n_samples with batch > 1 is not supported");
  uint32_t* d_mags = nullptr;
  KWS_CUDA_CHECK(cudaMallocAsync(&d_mags, (size_t)batch * n_frames * C * 4, st));
  int rc = launch_split(fe, d_pcm, batch, n_samples, n_frames, true, d_mags, 0, batch, 1, 0, n_frames, out_scale,
                        d_out_f32, d_out_u16, st);
  cudaFreeAsync(d_mags, st);
  return rc;
}

extern "C" int64_t kws_frontend_stream_num_windows(const kws_frontend_t* fe, int64_t total_samples, int clip_samples,
                                                   int hop_samples) {
  if (!fe || hop_samples <= 0 || clip_samples <= 0) return KWS_ERR_ARG;
  // batch_streaming_analysis.py:108-110: range(0, total - clip, hop)
  const int64_t end = total_samples - clip_samples;
  if (end <= 0) return 0;
  return (end + hop_samples - 1) / hop_samples;
}

extern "C" size_t kws_frontend_stream_scratch_bytes(const kws_frontend_t* fe, int64_t total_samples) {
  if (!fe || total_samples < 0) return 0;
  const long long frames = total_samples < fe->host.window_size
                               ? 0
                               : (total_samples - fe->host.window_size) / fe->host.window_step + 1;
  return round_up((size_t)frames * fe->host.num_channels * sizeof(uint32_t), 256) + 256;
}

extern "C" int kws_frontend_stream(kws_frontend_t* fe, const int16_t* d_pcm, int64_t total_samples, int clip_samples,
                                   int hop_samples, int64_t first_window, int64_t n_windows, float out_scale,
                                   float* d_out_f32, void* d_scratch, int scratch_ready, void* stream) {
  KWS_REQUIRE(fe != nullptr, "kws_frontend_stream: NULL handle");
  const FrontendTables& T = fe->host;
  KWS_REQUIRE(hop_samples > 0 && hop_samples % T.window_step == 0,
              "kws_frontend_stream: hop_samples (%d) must be a positive multiple of the frame step (%d)", hop_samples,
              T.window_step);
  KWS_REQUIRE(clip_samples >= T.window_size, "kws_frontend_stream: clip shorter than one analysis window");
  const int64_t all_windows = kws_frontend_stream_num_windows(fe, total_samples, clip_samples, hop_samples);
  KWS_REQUIRE(first_window >= 0 && n_windows >= 0 && first_window + n_windows <= all_windows,
              "kws_frontend_stream: window range [%lld, %lld) outside [0, %lld)", (long long)first_window,
              (long long)(first_window + n_windows), (long long)all_windows);
  if (n_windows == 0 && scratch_ready) return KWS_OK;
  KWS_REQUIRE(d_pcm != nullptr && d_scratch != nullptr, "kws_frontend_stream: NULL device buffer");
  KWS_REQUIRE(n_windows == 0 || d_out_f32 != nullptr, "kws_frontend_stream: NULL output");
  KWS_REQUIRE(((uintptr_t)d_pcm & 3) == 0, "kws_frontend_stream: d_pcm must be 4-byte aligned");
  const long long frames_total = num_frames_of(T, total_samples) < 0 ? 0 : (total_samples - T.window_size) / T.window_step + 1;
  const int win_frames = num_frames_of(T, clip_samples);
  return launch_split(fe, d_pcm, 1, 0, (int)frames_total, !scratch_ready, (uint32_t*)d_scratch, first_window, n_windows,
                      all_windows > 0 ? all_windows : 1, hop_samples / T.window_step, win_frames, out_scale, d_out_f32,
                      nullptr, (cudaStream_t)stream);
}
