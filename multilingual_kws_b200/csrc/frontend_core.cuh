// frontend_core.cuh — integer-exact micro-frontend arithmetic shared by the sm_100a kernels
// (frontend.cu) and by a host-side lane emulator (tests/test_frontend_lane_emulation.py via
// kws_debug_frontend_emulate) that replays the exact half-warp choreography on the CPU.
//
// Replaces, for the hot path, TF 2.7's `audio_microfrontend` op as called at
// reference multilingual_kws/embedding/input_data.py:25-33.
//
// Work decomposition: one HALF-WARP (16 lanes) per 30 ms frame.
//   P0  window multiply (int16*int16>>12), per-lane max|.|          -> half-warp max -> input_shift
//   P1  radix-16 register FFT pass = kissfft stages (4,1)+(4,4)       -> smem (padded 17/16)
//   P2  radix-16 register FFT pass = kissfft stages (4,16)+(4,64)     -> smem
//   P3  kiss_fftr real post-pass + |X|^2                              -> smem energy
//   P4  mel band partial sums (64-bit IMAD.WIDE), balanced over lanes -> smem W/U
//   P5  work[c+1] = W[c+1]+U[c]; rounded isqrt64 >> input_shift        -> magnitudes
// Every intermediate is wrapped to int16 exactly where kissfft (FIXED_POINT=16) stores to
// kiss_fft_scalar, so results are bit-identical to the CPU oracle for ANY int16 input.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KWS_HD __host__ __device__ __forceinline__
#else
#define KWS_HD inline
#endif

namespace kws {

constexpr int kFftSize = 512;        // the CUDA path is specialised for 256 < window_size <= 512
constexpr int kNcfft = 256;
constexpr int kHalfWarp = 16;
constexpr int kMaxChannels = 64;
constexpr int kFftBufWords = 272;    // 256 complex words, padded idx + idx/16
constexpr int kEnergyWords = 260;
constexpr int kMaxLaneBands = 8;

struct FrontendTables {
  int32_t window_size, window_step, num_channels, start_index, end_index;
  int32_t smoothing_bits, even_smoothing, odd_smoothing, min_signal_remaining;
  int32_t snr_shift, correction_bits, enable_pcan, enable_log, scale_shift;
  int32_t pad_[2];
  uint32_t window_pairs[256];        // word i = coef[2i] | coef[2i+1] << 16 (zero past window_size)
  uint32_t twiddles[256];            // r | i << 16, phase -2*pi*i/256
  uint32_t super_twiddles[128];      // phase -pi*((i+1)/256 + .5)
  int16_t bin_weight[260];
  int16_t bin_unweight[260];
  int16_t band_start[kMaxChannels + 4];   // band i covers bins [band_start[i], band_start[i+1])
  int8_t lane_bands[kHalfWarp][kMaxLaneBands];  // balanced band -> lane schedule, -1 terminated
  int16_t gain_lut[128];
  uint16_t log_lut[132];
};

// ---------------------------------------------------------------- fixed-point primitives
KWS_HD int32_t wrap16(int32_t x) { return (int32_t)(int16_t)x; }
KWS_HD int32_t sround(int32_t x) { return (int32_t)(int16_t)((x + 16384) >> 15); }
KWS_HD int32_t lo16(uint32_t w) { return (int32_t)(int16_t)(w & 0xFFFFu); }
KWS_HD int32_t hi16(uint32_t w) { return (int32_t)(int16_t)(w >> 16); }
KWS_HD uint32_t pack16(int32_t lo, int32_t hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

struct cpx { int32_t r, i; };   // int16-range values held in int32 registers

KWS_HD cpx unpack(uint32_t w) { cpx c; c.r = lo16(w); c.i = hi16(w); return c; }
KWS_HD uint32_t packc(cpx c) { return pack16(c.r, c.i); }
KWS_HD cpx cmul(cpx a, cpx b) {
  cpx m;
  m.r = sround(a.r * b.r - a.i * b.i);
  m.i = sround(a.r * b.i + a.i * b.r);
  return m;
}
KWS_HD cpx cadd(cpx a, cpx b) { cpx c; c.r = wrap16(a.r + b.r); c.i = wrap16(a.i + b.i); return c; }
KWS_HD cpx csub(cpx a, cpx b) { cpx c; c.r = wrap16(a.r - b.r); c.i = wrap16(a.i - b.i); return c; }
KWS_HD cpx fixdiv(cpx a, int32_t mult) { cpx c; c.r = sround(a.r * mult); c.i = sround(a.i * mult); return c; }

// Range analysis that lets most int16 re-wraps be dropped (a wrap is the identity when the value fits):
//   C_FIXDIV(x, 4):  |sround(x*8191)| <= 8191 for every int16 x;   C_FIXDIV(x, 2): <= 16383.
//   C_MUL(f, tw) with |f.r|,|f.i| <= 8191 and |tw| <= 32768:  |result| <= 8191*sqrt(2)+1 = 11585.
//   s5 = f0 - s1, f0 + s1: <= 19776;  s3 = s0 + s2, s4 = s0 - s2: <= 23170 — all inside int16.
//   Only the four final sums (up to 42946) can leave int16, and those keep their wrap.
KWS_HD int32_t sround_nw(int32_t x) { return (x + 16384) >> 15; }
KWS_HD cpx cmul_nw(cpx a, cpx b) {
  cpx m;
  m.r = sround_nw(a.r * b.r - a.i * b.i);
  m.i = sround_nw(a.r * b.i + a.i * b.r);
  return m;
}
KWS_HD cpx fixdiv_nw(cpx a, int32_t mult) { cpx c; c.r = sround_nw(a.r * mult); c.i = sround_nw(a.i * mult); return c; }

// kissfft kf_bfly4 body for one k (forward transform), with the C_FIXDIV(.,4) prologue.
KWS_HD void bfly4(cpx& f0, cpx& f1, cpx& f2, cpx& f3, cpx tw1, cpx tw2, cpx tw3) {
  f0 = fixdiv_nw(f0, 8191); f1 = fixdiv_nw(f1, 8191); f2 = fixdiv_nw(f2, 8191); f3 = fixdiv_nw(f3, 8191);
  const cpx s0 = cmul_nw(f1, tw1), s1 = cmul_nw(f2, tw2), s2 = cmul_nw(f3, tw3);
  const int32_t s5r = f0.r - s1.r, s5i = f0.i - s1.i;
  const int32_t t0r = f0.r + s1.r, t0i = f0.i + s1.i;
  const int32_t s3r = s0.r + s2.r, s3i = s0.i + s2.i;
  const int32_t s4r = s0.r - s2.r, s4i = s0.i - s2.i;
  f2.r = wrap16(t0r - s3r); f2.i = wrap16(t0i - s3i);
  f0.r = wrap16(t0r + s3r); f0.i = wrap16(t0i + s3i);
  f1.r = wrap16(s5r + s4i); f1.i = wrap16(s5i - s4r);
  f3.r = wrap16(s5r - s4i); f3.i = wrap16(s5i + s4r);
}
// Same butterfly when all three twiddles are tw[0] = (32767, 0).  After C_FIXDIV(.,4) every
// component satisfies |x| <= 8191, for which sround(x*32767) == x, so the C_MULs are identities.
KWS_HD void bfly4_unit(cpx& f0, cpx& f1, cpx& f2, cpx& f3) {
  f0 = fixdiv_nw(f0, 8191); f1 = fixdiv_nw(f1, 8191); f2 = fixdiv_nw(f2, 8191); f3 = fixdiv_nw(f3, 8191);
  const int32_t s5r = f0.r - f2.r, s5i = f0.i - f2.i;
  const int32_t t0r = f0.r + f2.r, t0i = f0.i + f2.i;
  const int32_t s3r = f1.r + f3.r, s3i = f1.i + f3.i;
  const int32_t s4r = f1.r - f3.r, s4i = f1.i - f3.i;
  f2.r = t0r - s3r; f2.i = t0i - s3i;          // four terms of magnitude <= 8191: no wrap possible
  f0.r = t0r + s3r; f0.i = t0i + s3i;
  f1.r = s5r + s4i; f1.i = s5i - s4r;
  f3.r = s5r - s4i; f3.i = s5i + s4r;
}

KWS_HD int msb32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return 32 - __clz((int)x);
#else
  return x ? 32 - __builtin_clz(x) : 0;
#endif
}

KWS_HD int fft_pad(int idx) { return idx + (idx >> 4); }

// Rounded integer sqrt with the reference's saturation quirks (Sqrt32 caps at 0xFFFF, Sqrt64 at
// 0xFFFFFFFF; Sqrt64 takes the 32-bit path when the high word is clear).  Mathematically:
// r = floor(sqrt(x)); if (x - r*r > r && r != cap) ++r.
KWS_HD uint32_t isqrt64_round(uint64_t x) {
  if (x == 0) return 0;
#if defined(__CUDA_ARCH__)
  uint64_t r = (uint64_t)__dsqrt_rn((double)x);
#else
  uint64_t r = (uint64_t)__builtin_sqrt((double)x);
#endif
  if (r > 0xFFFFFFFFull) r = 0xFFFFFFFFull;
  while (r * r > x) --r;                                     // at most a step or two
  while (r < 0xFFFFFFFFull && (r + 1) * (r + 1) <= x) ++r;
  const uint64_t rem = x - r * r;
  const uint64_t cap = (x >> 32) ? 0xFFFFFFFFull : 0xFFFFull;
  if (rem > r && r != cap) ++r;
  return (uint32_t)r;
}

// ---------------------------------------------------------------- per-lane register state
struct LaneRegs {
  uint32_t w[16];      // packed complex int16 working set (P0..P2)
  int32_t local_max;   // P0: max |windowed sample| seen by this lane
};

// P0: lane j loads words j+16n (n = 0..15) of the frame, windows them.
// `frame_words` points at the frame's first sample viewed as 32-bit words (two int16 samples each).
KWS_HD void fe_p0_window(int lane, const uint32_t* frame_words, const FrontendTables& T, LaneRegs& R) {
  const int nwords = (T.window_size + 1) >> 1;
  int32_t mx = 0;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    const int wi = lane + 16 * n;
    uint32_t out = 0;
    if (wi < nwords) {
      const uint32_t s = frame_words[wi];
      const uint32_t c = T.window_pairs[wi];
      int32_t a = (lo16(s) * lo16(c)) >> 12;           // |s| <= 32768, 0 <= coef <= 4096: always inside int16
      int32_t b = (hi16(s) * hi16(c)) >> 12;           // coef is 0 past window_size
      out = pack16(a, b);
      int32_t aa = a < 0 ? wrap16(-a) : a;               // int16 negate: -(-32768) stays negative
      int32_t bb = b < 0 ? wrap16(-b) : b;
      mx = aa > mx ? aa : mx;
      mx = bb > mx ? bb : mx;
    }
    R.w[n] = out;
  }
  R.local_max = mx;
}

// P1: scale by input_shift, then kissfft stages (4,1) and (4,4) on the lane's 16 points.
// Lane j = a + 4b holds x[n] = f[j + 16n]; leaf order gives y[4c+d] = x[c+4d].
// tw_const: the full 256-entry twiddle table (uniform indices here -> constant-bank operands).
KWS_HD void fe_p1_fft_pass1(int lane, int input_shift, const uint32_t* tw_const, LaneRegs& R, uint32_t* fftbuf) {
  cpx y[16];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const uint32_t s = R.w[c + 4 * d];
      cpx v;
      v.r = (int32_t)(int16_t)(uint16_t)(((uint32_t)(s & 0xFFFFu)) << input_shift);
      v.i = (int32_t)(int16_t)(uint16_t)(((uint32_t)(s >> 16)) << input_shift);
      y[4 * c + d] = v;
    }
  // stage (p=4, m=1): four butterflies over y[4c..4c+3], all twiddles tw[0]
#pragma unroll
  for (int c = 0; c < 4; ++c) bfly4_unit(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
  // stage (p=4, m=4), fstride 16: butterfly k over y[k], y[k+4], y[k+8], y[k+12]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k == 0) bfly4_unit(y[0], y[4], y[8], y[12]);
    else bfly4(y[k], y[k + 4], y[k + 8], y[k + 12], unpack(tw_const[16 * k]), unpack(tw_const[32 * k]),
               unpack(tw_const[48 * k]));
  }
  // lane (a,b) owns Fout[64a + 16b + t]
  const int a = lane & 3, b = lane >> 2;
  const int base = 64 * a + 16 * b;
#pragma unroll
  for (int t = 0; t < 16; ++t) fftbuf[fft_pad(base + t)] = packc(y[t]);
}

// P2: lane k' gathers z[q] = Fout[16q + k'] (q = 4a+b), runs stages (4,16) and (4,64).
// tw2[0..2] = tw[4k'], tw[8k'], tw[12k'];  tw2[3+3b+{0,1,2}] = tw[k], tw[2k], tw[3k] with k = 16b+k'.
KWS_HD void fe_p2_fft_pass2(int lane, const uint32_t* tw2, uint32_t* fftbuf) {
  cpx z[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) z[q] = unpack(fftbuf[fft_pad(16 * q + lane)]);
  const cpx t1 = unpack(tw2[0]), t2 = unpack(tw2[1]), t3 = unpack(tw2[2]);
#pragma unroll
  for (int a = 0; a < 4; ++a) bfly4(z[4 * a], z[4 * a + 1], z[4 * a + 2], z[4 * a + 3], t1, t2, t3);
#pragma unroll
  for (int b = 0; b < 4; ++b)
    bfly4(z[b], z[4 + b], z[8 + b], z[12 + b], unpack(tw2[3 + 3 * b]), unpack(tw2[4 + 3 * b]), unpack(tw2[5 + 3 * b]));
#pragma unroll
  for (int q = 0; q < 16; ++q) fftbuf[fft_pad(16 * q + lane)] = packc(z[q]);
}

// Load the lane's 15 pass-2 twiddles (called once per kernel).
KWS_HD void fe_load_tw2(int lane, const uint32_t* tw, uint32_t* tw2) {
  tw2[0] = tw[(4 * lane) & 255]; tw2[1] = tw[(8 * lane) & 255]; tw2[2] = tw[(12 * lane) & 255];
  for (int b = 0; b < 4; ++b) {
    const int k = 16 * b + lane;
    tw2[3 + 3 * b] = tw[k]; tw2[4 + 3 * b] = tw[2 * k]; tw2[5 + 3 * b] = tw[3 * k];   // 3k <= 237
  }
}

// P3: kiss_fftr post-pass for k = 1 + lane + 16 i (i = 0..7) and |X|^2 of bins k and 256-k.
KWS_HD void fe_p3_real_energy(int lane, const uint32_t* fftbuf, const uint32_t* super_tw, uint32_t* energy) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = 1 + lane + 16 * i;
    cpx fpk = unpack(fftbuf[fft_pad(k)]);
    cpx t = unpack(fftbuf[fft_pad((256 - k) & 255)]);
    cpx fpnk; fpnk.r = t.r; fpnk.i = wrap16(-t.i);
    fpk = fixdiv_nw(fpk, 16383);                 // |.| <= 16383: sums / differences below stay inside int16
    fpnk = fixdiv_nw(fpnk, 16383);
    cpx f1k, f2k;
    f1k.r = fpk.r + fpnk.r; f1k.i = fpk.i + fpnk.i;
    f2k.r = fpk.r - fpnk.r; f2k.i = fpk.i - fpnk.i;
    const cpx tw = cmul(f2k, unpack(super_tw[k - 1]));      // up to 46338: C_MUL's int16 store wraps
    const int32_t ar = (f1k.r + tw.r) >> 1, ai = (f1k.i + tw.i) >> 1;   // |sum| <= 65534 -> halves fit int16
    const int32_t br = (f1k.r - tw.r) >> 1, bi = (tw.i - f1k.i) >> 1;
    const uint32_t ea = (uint32_t)(ar * ar) + (uint32_t)(ai * ai);
    const uint32_t eb = (uint32_t)(br * br) + (uint32_t)(bi * bi);
    if (k != 128) energy[k] = ea;     // for k == ncfft/2 the second store wins in kiss_fftr
    energy[256 - k] = eb;
  }
}

// P4: weighted / unweighted sums of the bands scheduled on this lane.
KWS_HD void fe_p4_band_sums(int lane, const uint32_t* energy, const FrontendTables& T, uint64_t* WU /*[2*(nch+1)]*/) {
  for (int s = 0; s < kMaxLaneBands; ++s) {
    const int band = T.lane_bands[lane][s];
    if (band < 0) break;
    uint64_t w = 0, u = 0;
    const int f1 = T.band_start[band + 1];
    for (int f = T.band_start[band]; f < f1; ++f) {
      const uint64_t e = (uint64_t)(int64_t)(int32_t)energy[f];   // the C source re-reads energy as int32
      w += (uint64_t)(int64_t)T.bin_weight[f] * e;
      u += (uint64_t)(int64_t)T.bin_unweight[f] * e;
    }
    WU[2 * band] = w;
    WU[2 * band + 1] = u;
  }
}

// P5: channel c (output) = work[c+1] = W[c+1] + U[c]; rounded sqrt; >> input_shift.
KWS_HD uint32_t fe_p5_channel(int c, int input_shift, const uint64_t* WU) {
  const uint64_t work = WU[2 * (c + 1)] + WU[2 * c + 1];
  return isqrt64_round(work) >> input_shift;
}

// ---------------------------------------------------------------- sequential + pointwise tail
// Noise-estimate recurrence (the only truly sequential part): returns the new estimate.
KWS_HD uint32_t fe_noise_estimate(uint32_t signal, uint32_t prev_est, int c, const FrontendTables& T) {
  const uint32_t smoothing = (c & 1) ? (uint32_t)T.odd_smoothing : (uint32_t)T.even_smoothing;
  const uint32_t one_minus = (1u << 14) - smoothing;
  const uint32_t s_up = signal << T.smoothing_bits;
  return (uint32_t)((((uint64_t)s_up * smoothing) + ((uint64_t)prev_est * one_minus)) >> 14);
}

KWS_HD int32_t fe_wide_dynamic(uint32_t x, const int16_t* lut) {
  if (x <= 2) return lut[x];
  const int interval = msb32(x);
  const int16_t* p = lut + 4 * interval - 6;
  const int32_t frac = (int32_t)(int16_t)(((interval < 11) ? (x << (11 - interval)) : (x >> (interval - 11))) & 0x3FF);
  int32_t result = ((int32_t)p[2] * frac) >> 5;
  result += (int32_t)((uint32_t)(int32_t)p[1] << 5);
  result *= frac;
  result = (result + (1 << 14)) >> 15;
  result += p[0];
  return (int32_t)(int16_t)result;
}

KWS_HD uint32_t fe_log_scale(uint32_t x, const uint16_t* log_lut, int scale_shift) {
  const uint32_t integer = (uint32_t)msb32(x) - 1;
  int32_t frac = (int32_t)(x - (uint32_t)(1ull << integer));
  if (integer < 16) frac <<= 16 - integer; else frac >>= integer - 16;
  const uint32_t base_seg = (uint32_t)frac >> 9;
  const int32_t c0 = log_lut[base_seg], c1 = log_lut[base_seg + 1];
  const int32_t seg_base = (int32_t)(512u * base_seg);
  const int32_t rel_pos = ((c1 - c0) * (frac - seg_base)) >> 16;
  const uint32_t fraction = (uint32_t)(frac + c0 + rel_pos);
  const uint32_t log2 = (integer << 16) + fraction;
  const uint32_t loge = (uint32_t)(((uint64_t)45426u * log2 + 32768u) >> 16);
  return ((loge << scale_shift) + 32768u) >> 16;
}

// Pointwise tail for one (frame, channel): spectral subtraction, PCAN, log, saturate.
// `est` is the estimate AFTER this frame's update.
KWS_HD uint32_t fe_pointwise(uint32_t signal, uint32_t est, const FrontendTables& T) {
  const uint32_t s_up = signal << T.smoothing_bits;
  uint32_t e = est > s_up ? s_up : est;
  const uint32_t floor_ = (uint32_t)(((uint64_t)signal * (uint32_t)T.min_signal_remaining) >> 14);
  const uint32_t sub = (s_up - e) >> T.smoothing_bits;
  uint32_t v = sub > floor_ ? sub : floor_;
  if (T.enable_pcan) {
    const uint32_t gain = (uint32_t)fe_wide_dynamic(est, T.gain_lut);
    const uint32_t snr = (uint32_t)(((uint64_t)v * gain) >> T.snr_shift);
    v = (snr < 8192u) ? (snr * snr) >> 20 : (snr >> 6) - 64u;
  }
  if (T.enable_log) {
    if (T.correction_bits < 0) v >>= -T.correction_bits; else v <<= T.correction_bits;
    v = (v > 1) ? fe_log_scale(v, T.log_lut, T.scale_shift) : 0;
  }
  return v < 0xFFFFu ? v : 0xFFFFu;
}

}  // namespace kws
