// augment.cu — waveform augmentation of the fine-tune input pipeline on the device.
//
// Replaces, for a whole batch at once, the per-element tf.data map of the reference
// (multilingual_kws/embedding/input_data.py): AudioDataset.random_timeshift (:243-267), random_background_sample
// (:227-241), add_background (:141-157), the three branches of AudioDataset.augment (:275-304), the int16 cast of
// to_micro_spectrogram (:23), and spec_augment (:306-364).  The random DECISIONS (shift, branch, background file /
// offset / volume, mask bands) are drawn on the host exactly as the host mirror draws them and arrive as a 32-byte
// plan item per clip; the SAMPLES never leave the GPU: source clips and background recordings live in int16 banks in
// HBM, the output is the [batch, n_samples] int16 PCM tensor the frontend kernel consumes.
//
// augment_pcm_kernel: one CTA per output clip.  The (unaligned) source windows are staged into shared memory by two
// TMA bulk copies of their 16-byte-aligned hulls; pass 1 = the two mean squares (float32 squares as tf.square gives,
// accumulated in float64 in a fixed order -> deterministic), pass 2 = mix / clip / truncate-to-int16 with every float32
// operation issued as an explicit round-to-nearest intrinsic (no FMA contraction), so the result equals the host
// mirror's numpy float32 arithmetic bit for bit.  HBM-bound: 2 B/sample in per source + 2 B/sample out.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "ptx.cuh"

namespace kws {
namespace {

struct AugItem {       // mirrors kws_aug_item in include/kws_b200.h
  int32_t mode, fg_index, shift, bg_index, bg_offset;
  float volume;
  int32_t reserved[2];
};
static_assert(sizeof(AugItem) == 32, "plan item layout");

constexpr int kAugThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5;
  __syncthreads();                                   // s_red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kAugThreads / 32; ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(kAugThreads)
augment_pcm_kernel(const int16_t* __restrict__ fg, int n_fg, int fg_stride, const int16_t* __restrict__ bg, int n_bg,
                   int bg_stride, const AugItem* __restrict__ plan, int n_samples, int16_t* __restrict__ pcm_out,
                   float* __restrict__ audio_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int N = n_samples;
  const int cap = N + 16;                            // staged elements per source (aligned hull of an N-long window)
  int16_t* s_fg = reinterpret_cast<int16_t*>(smem);
  int16_t* s_bg = s_fg + cap;
  double* s_red = reinterpret_cast<double*>(smem + (((size_t)cap * 4 + 15) & ~(size_t)15));
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + kAugThreads / 32);

  const int tid = threadIdx.x;
  const AugItem it = plan[blockIdx.x];
  const bool use_fg = (it.mode == 0 || it.mode == 2) && it.fg_index >= 0 && it.fg_index < n_fg;
  const bool use_bg = (it.mode == 1 || it.mode == 2) && it.bg_index >= 0 && it.bg_index < n_bg;
  // source index of output sample i:  foreground i - shift (zero outside [0, N)),  background bg_offset + i
  const int f_lo = max(0, -it.shift), f_hi = min(N, N - it.shift);                 // foreground source range
  const int b_lo = max(0, it.bg_offset), b_hi = min(bg_stride, it.bg_offset + N);  // background source range
  const int f_a0 = f_lo & ~7, b_a0 = b_lo & ~7;
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
    uint32_t bytes = 0;
    const uint32_t f_bytes = (use_fg && f_hi > f_lo) ? (uint32_t)(min((f_hi + 7) & ~7, fg_stride) - f_a0) * 2u : 0u;
    const uint32_t b_bytes = (use_bg && b_hi > b_lo) ? (uint32_t)(min((b_hi + 7) & ~7, bg_stride) - b_a0) * 2u : 0u;
    bytes = f_bytes + b_bytes;
    if (bytes) {
      ptx::mbar_expect_tx(bar, bytes);
      if (f_bytes) ptx::tma_bulk_g2s(s_fg, fg + (size_t)it.fg_index * fg_stride + f_a0, f_bytes, bar);
      if (b_bytes) ptx::tma_bulk_g2s(s_bg, bg + (size_t)it.bg_index * bg_stride + b_a0, b_bytes, bar);
    } else {
      ptx::mbar_arrive(bar);
    }
  }
  __syncthreads();
  ptx::mbar_wait(bar, 0);

  const float kInv = 1.0f / 32768.0f;                // int16 -> float exactly as decode_wav does
  auto fg_at = [&](int i) -> float {
    const int s = i - it.shift;
    return (use_fg && s >= f_lo && s < f_hi) ? (float)s_fg[s - f_a0] * kInv : 0.0f;
  };
  auto bg_at = [&](int i) -> float {
    const int s = it.bg_offset + i;
    return (use_bg && s >= b_lo && s < b_hi) ? (float)s_bg[s - b_a0] * kInv : 0.0f;
  };

  float snr = 0.0f;
  if (it.mode == 2) {                                // uniform per CTA: no divergence at the barriers inside block_sum
    double sf = 0.0, sb = 0.0;
    for (int i = tid; i < N; i += kAugThreads) {
      const float f = fg_at(i), b = bg_at(i);
      sf += (double)__fmul_rn(f, f);
      sb += (double)__fmul_rn(b, b);
    }
    sf = block_sum(sf, s_red);
    sb = block_sum(sb, s_red);
    const float rms_f = __fsqrt_rn((float)(sf / (double)N));
    const float rms_b = __fsqrt_rn((float)(sb / (double)N));
    snr = rms_b > 0.0f ? __fdiv_rn(rms_f, rms_b) : 0.0f;
  }
  int16_t* out = pcm_out + (size_t)blockIdx.x * N;
  float* aout = audio_out ? audio_out + (size_t)blockIdx.x * N : nullptr;
  for (int i = 2 * tid; i < N; i += 2 * kAugThreads) {
    float x[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = i + u;
      float v = 0.0f;
      if (j < N) {
        if (it.mode == 0) {
          v = fg_at(j);
        } else if (it.mode == 1) {
          v = __fmul_rn(bg_at(j), it.volume);
        } else {
          v = __fadd_rn(__fmul_rn(__fmul_rn(bg_at(j), snr), it.volume), fg_at(j));
          v = fminf(fmaxf(v, -1.0f), 1.0f);
        }
      }
      x[u] = v;
    }
    // tf.cast(audio * 32768, int16): truncate toward zero, wrap to 16 bits (+1.0 -> -32768)
    const uint32_t lo = (uint32_t)(uint16_t)(int16_t)__float2int_rz(__fmul_rn(x[0], 32768.0f));
    const uint32_t hi = (uint32_t)(uint16_t)(int16_t)__float2int_rz(__fmul_rn(x[1], 32768.0f));
    if (i + 1 < N) {
      *reinterpret_cast<uint32_t*>(out + i) = lo | (hi << 16);
      if (aout) *reinterpret_cast<float2*>(aout + i) = make_float2(x[0], x[1]);
    } else {
      out[i] = (int16_t)lo;
      if (aout) aout[i] = x[0];
    }
  }
}

// spec_augment: zero up to two frequency bands and two time bands per clip, in place.
__global__ void __launch_bounds__(256)
spec_mask_kernel(float* __restrict__ feats, int frames, int channels, const int32_t* __restrict__ bands) {
  const int32_t* b = bands + (size_t)blockIdx.x * 8;
  const int f0 = b[0], f0n = b[1], f1 = b[2], f1n = b[3], t0 = b[4], t0n = b[5], t1 = b[6], t1n = b[7];
  if ((f0n | f1n | t0n | t1n) == 0) return;
  float* x = feats + (size_t)blockIdx.x * frames * channels;
  for (int i = threadIdx.x; i < frames * channels; i += blockDim.x) {
    const int t = i / channels, f = i - t * channels;
    const bool hit = (f >= f0 && f < f0 + f0n) || (f >= f1 && f < f1 + f1n) || (t >= t0 && t < t0 + t0n) ||
                     (t >= t1 && t < t1 + t1n);
    if (hit) x[i] = 0.0f;
  }
}

}  // namespace
}  // namespace kws

extern "C" int kws_augment_pcm(const int16_t* d_fg, int n_fg, int fg_stride, const int16_t* d_bg, int n_bg, int bg_stride,
                               const void* d_plan, int batch, int n_samples, int16_t* d_pcm_out, float* d_audio_out,
                               void* stream) {
  using namespace kws;
  KWS_REQUIRE(batch >= 0 && n_samples > 0 && n_samples % 8 == 0, "kws_augment_pcm: n_samples (%d) must be a positive multiple of 8", n_samples);
  if (batch == 0) return KWS_OK;
  KWS_REQUIRE(d_plan && d_pcm_out, "kws_augment_pcm: NULL plan / output");
  KWS_REQUIRE(n_fg == 0 || (d_fg && fg_stride >= n_samples && fg_stride % 8 == 0 && ((uintptr_t)d_fg & 15) == 0),
              "kws_augment_pcm: foreground bank must be 16-byte aligned with a row stride >= n_samples, multiple of 8");
  KWS_REQUIRE(n_bg == 0 || (d_bg && bg_stride > 0 && bg_stride % 8 == 0 && ((uintptr_t)d_bg & 15) == 0),
              "kws_augment_pcm: background bank must be 16-byte aligned with a row stride that is a multiple of 8");
  const size_t cap = (size_t)n_samples + 16;
  const size_t smem = ((cap * 4 + 15) & ~(size_t)15) + sizeof(double) * (kAugThreads / 32) + 16;
  KWS_REQUIRE(smem <= 227 * 1024, "kws_augment_pcm: clip of %d samples does not fit shared memory", n_samples);
  KWS_CUDA_CHECK(cudaFuncSetAttribute(augment_pcm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  augment_pcm_kernel<<<batch, kAugThreads, smem, (cudaStream_t)stream>>>(
      d_fg, n_fg, fg_stride, d_bg, n_bg, bg_stride, static_cast<const AugItem*>(d_plan), n_samples, d_pcm_out, d_audio_out);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

extern "C" int kws_spec_mask(float* d_feats, int batch, int frames, int channels, const int32_t* d_bands, void* stream) {
  using namespace kws;
  KWS_REQUIRE(batch >= 0 && frames > 0 && channels > 0, "kws_spec_mask: bad shape");
  if (batch == 0) return KWS_OK;
  KWS_REQUIRE(d_feats && d_bands, "kws_spec_mask: NULL buffer");
  spec_mask_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(d_feats, frames, channels, d_bands);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}
