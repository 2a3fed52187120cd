// conv_kernels.cu — the non-GEMM kernels of the EfficientNet-B0 embedding tower (sm_100a):
//   stem_conv_kernel   Rescaling(1/255) + ZeroPadding2D(correct_pad) + Conv3x3 s2 (1->32) + BN + swish,
//                      fp32 log-mel features in, NHWC 16-bit (fp16 default, bf16 optional) out.
//   dwse_kernel        one MBConv middle section per launch: depthwise kxk conv (TF SAME / correct_pad+VALID)
//                      + folded BN + swish + squeeze-excite (global pool -> FC+swish -> FC+sigmoid -> scale),
//                      whole clips staged in shared memory by TMA bulk copies; the pooled vector never
//                      leaves the SM, the gated activation is written once as the next GEMM's A operand.
// Replaces the cuDNN depthwise / Eigen kernels behind Keras' EfficientNetB0 block()
// (model defined at reference multilingual_kws/train_multilingual_embedding.py:66-83).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "conv_kernels.h"
#include "ptx.cuh"

namespace kws {

namespace {

__device__ __forceinline__ float swish(float x) { return ptx::swish_f(x); }
__device__ __forceinline__ float sigmoidf(float x) { return ptx::sigmoid_f(x); }

// ---------------------------------------------------------------- stem
constexpr int kStemThreads = 256;
constexpr int kStemC = 32;

__global__ void __launch_bounds__(kStemThreads)
stem_conv_kernel(const float* __restrict__ feats, int batch, StemParams P, uint16_t* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_in = reinterpret_cast<float*>(smem);                     // [(H+2)*(W+2)]: the clip inside a ring of zeros
  const int HP = P.H + 2, WP = P.W + 2;
  float* s_w = s_in + ((HP * WP + 3) & ~3);                          // [9][32]
  float* s_b = s_w + 9 * kStemC;                                     // [32]
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * kStemC; i += kStemThreads) s_w[i] = P.w[i];
  if (tid < kStemC) s_b[tid] = P.bias[tid];
  for (int i = tid; i < HP * WP; i += kStemThreads) s_in[i] = 0.0f;  // ZeroPadding2D pads the NORMALISED input with 0
  const int npix = P.Ho * P.Wo;
  __syncthreads();
  // item = (pixel, group of 8 output channels); 256 % 4 == 0, so a thread's channel group never changes and its
  // 72 weights + 8 biases live in registers for the whole kernel
  const int g = tid & 3;
  float wr[9][8], br[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[t][j] = s_w[t * kStemC + g * 8 + j];
#pragma unroll
  for (int j = 0; j < 8; ++j) br[j] = s_b[g * 8 + j];
  // a thread's pixels are p0, p0 + 64, ...: (ho, wo) advance by (64 / Wo, 64 % Wo) with a carry, no division in the loop
  const int p0 = tid >> 2, dq = (kStemThreads / 4) / P.Wo, dr = (kStemThreads / 4) % P.Wo;
  const int ho0 = p0 / P.Wo, wo0 = p0 - ho0 * P.Wo;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  for (int clip = blockIdx.x; clip < batch; clip += gridDim.x) {
    __syncthreads();
    const float* src = feats + (size_t)clip * P.H * P.W;
    for (int i = tid; i < P.H * P.W; i += kStemThreads) {
      const int r = i / P.W, c = i - r * P.W;
      s_in[(r + 1) * WP + c + 1] = fmaf(__ldg(src + i), P.in_scale, P.in_shift);
    }
    __syncthreads();
    int ho = ho0, wo = wo0;
    for (int p = p0; p < npix; p += kStemThreads / 4) {
      // taps (kh, kw) read input (2 ho + kh - pad_top, 2 wo + kw - pad_left); +1 for the ring: always inside the buffer
      const float* xin = s_in + (ho * 2 - P.pad_top + 1) * WP + (wo * 2 - P.pad_left + 1);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = br[j];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float x = xin[kh * WP + kw];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(x, wr[kh * 3 + kw][j], acc[j]);
        }
      uint4 pk;
      pk.x = ptx::pack_h2(swish(acc[0]), swish(acc[1]), P.bf16);
      pk.y = ptx::pack_h2(swish(acc[2]), swish(acc[3]), P.bf16);
      pk.z = ptx::pack_h2(swish(acc[4]), swish(acc[5]), P.bf16);
      pk.w = ptx::pack_h2(swish(acc[6]), swish(acc[7]), P.bf16);
      *reinterpret_cast<uint4*>(out + ((size_t)clip * npix + p) * kStemC + g * 8) = pk;
      ho += dq; wo += dr;
      if (wo >= P.Wo) { wo -= P.Wo; ++ho; }
    }
  }
}

// ---------------------------------------------------------------- depthwise + SE
constexpr int kDwThreads = 256;

struct DwSmem {
  uint32_t in_bytes, out_off, pooled_off, part_off, s_off, red_off, bar_off, w_off, w_floats, total;
  uint32_t in2_off;   // second input buffer (0: single-buffered)
  uint32_t zero_off;  // one all-zero input row (the padded rows of the early-layer specialisations read it)
  int PL;   // pixel lanes per channel pair
};
constexpr uint32_t kDwSmemCap = 227 * 1024;
__host__ __device__ inline DwSmem dw_smem(const DwseParams& P, int G) {
  DwSmem L;
  const int C2 = P.C >> 1;
  L.PL = C2 >= kDwThreads ? 1 : kDwThreads / C2;
  L.in_bytes = (uint32_t)G * P.H * P.W * P.C * 2;
  L.out_off = (L.in_bytes + 127) & ~127u;
  L.pooled_off = L.out_off + (((uint32_t)G * P.Ho * P.Wo * P.C * 2 + 127) & ~127u);
  L.part_off = L.pooled_off + (uint32_t)G * P.C * 4;                       // [PL][G][C] partial channel sums
  L.s_off = L.part_off + (L.PL > 1 ? (uint32_t)L.PL * G * P.C * 4 : 0u);
  L.red_off = (L.s_off + (uint32_t)G * P.se * 4 + 15) & ~15u;                            // [se][8 warps][4 clips] FC1 partials
  L.bar_off = (L.red_off + (uint32_t)P.se * (kDwThreads / 32) * 4 * 4 + 15) & ~15u;
  // Narrow layers keep ALL their weights (depthwise, both SE matrices, biases) resident in smem for the lifetime of
  // the persistent CTA, so no phase waits on a global-memory round trip; wide layers read them through L1/L2.
  L.w_floats = (uint32_t)(P.K * P.K * P.C + P.C + (P.se_external ? 0 : 2 * P.se * P.C + P.se + P.C));
  L.w_off = (L.bar_off + 16 + 15) & ~15u;
  if (L.w_floats * 4 > 40 * 1024) L.w_floats = 0;
  L.total = L.w_off + L.w_floats * 4;
  L.zero_off = 0;
  if ((uint32_t)P.W * (P.C >> 1) * 4 <= 4096) {
    L.zero_off = (L.total + 15) & ~15u;
    L.total = L.zero_off + 4096;
  }
  // A layer whose group does not leave room for a second CTA on the SM (block2a: 96 KB of input per clip) would
  // expose the whole bulk-load latency of every group; if two input buffers fit, the next group's load is issued
  // while the current one is convolved.
  L.in2_off = 0;
  const uint32_t in2 = (L.total + 127) & ~127u;
  if (L.total > 100 * 1024 && in2 + L.in_bytes <= kDwSmemCap) {
    L.in2_off = in2;
    L.total = in2 + L.in_bytes;
  }
  return L;
}

// Tiny late-stage maps (<= 7x5 in, <= 4x3 out): the whole (clip, channel-pair) image lives in registers and every
// bound is a compile-time constant, so only the taps that actually overlap the image are executed, with no branches.
template <int K, int S, int H, int W, int PT, int PLFT>
struct SmallGeom {
  static constexpr int HO = S == 1 ? H : (H + PT + K / 2 - K) / 2 + 1;
  static constexpr int WO = S == 1 ? W : (W + PLFT + K / 2 - K) / 2 + 1;
};
template <int K, int S, int H, int W, int PT, int PLFT>
__device__ __forceinline__ void dw_small_item(const uint32_t* __restrict__ in_g, int C2, const float2 (&wreg)[K * K],
                                              float2 bias, int bf16, uint32_t* __restrict__ out_g, float& sum0,
                                              float& sum1) {
  using G = SmallGeom<K, S, H, W, PT, PLFT>;
  float2 x[H * W];
#pragma unroll
  for (int q = 0; q < H * W; ++q) x[q] = ptx::unpack_h2(in_g[(size_t)q * C2], bf16);
#pragma unroll
  for (int ho = 0; ho < G::HO; ++ho)
#pragma unroll
    for (int wo = 0; wo < G::WO; ++wo) {
      float a0 = bias.x, a1 = bias.y;
#pragma unroll
      for (int kh = 0; kh < K; ++kh)
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const int r = ho * S + kh - PT, c = wo * S + kw - PLFT;     // compile-time after unrolling
          if (r >= 0 && r < H && c >= 0 && c < W) {
            a0 = fmaf(x[r * W + c].x, wreg[kh * K + kw].x, a0);
            a1 = fmaf(x[r * W + c].y, wreg[kh * K + kw].y, a1);
          }
        }
      a0 = swish(a0);
      a1 = swish(a1);
      sum0 += a0;
      sum1 += a1;
      out_g[(size_t)(ho * G::WO + wo) * C2] = ptx::pack_h2(a0, a1, bf16);
    }
}

// Early-stage maps (25x20 ... 7x5): one output ROW per call with every column bound, smem offset and window slot a
// compile-time constant (geometry, stride, padding and channel count are template parameters), so the row costs
// K*S shared loads + K*K*2 FMAs per output pixel and no branches; rows in the padding read an all-zero smem row.
template <int K, int S, int W, int WO, int PLFT, int C2>
__device__ __forceinline__ void dw_strip_row(const uint32_t* const (&rowp)[K],
                                             const float2 (&wreg)[K * K], float2 bias, int bf16,
                                             uint32_t* __restrict__ orow, float& sum0, float& sum1) {
  float2 win[K][K];
#pragma unroll
  for (int wo = 0; wo < WO; ++wo) {
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      if (kw >= K - S || wo == 0) {                      // compile-time: the columns that enter the window here
        const int c = wo * S + kw - PLFT;
        const int slot = (wo * S + kw) % K;
#pragma unroll
        for (int kh = 0; kh < K; ++kh) {
          float2 v = make_float2(0.0f, 0.0f);
          if (c >= 0 && c < W) v = ptx::unpack_h2(rowp[kh][c * C2], bf16);   // compile-time bound; padded rows
          win[kh][slot] = v;                                                  // point at the all-zero row
        }
      }
    }
    float a0 = bias.x, a1 = bias.y;
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int kw = 0; kw < K; ++kw) {
        const int slot = (wo * S + kw) % K;
        a0 = fmaf(win[kh][slot].x, wreg[kh * K + kw].x, a0);
        a1 = fmaf(win[kh][slot].y, wreg[kh * K + kw].y, a1);
      }
    a0 = swish(a0);
    a1 = swish(a1);
    sum0 += a0;
    sum1 += a1;
    orow[wo * C2] = ptx::pack_h2(a0, a1, bf16);
  }
}

// GEOM: 0 = generic (runtime geometry, row strips); 1..6 = the tiny-map geometries of EfficientNet-B0 at 49x40 input
// (whole image in registers); 7..11 = its early-stage geometries (compile-time row strips).
// Squeeze-excite of the narrow layers (C <= 256: every block up to 3b), all clips of a group, written to need only
// a handful of registers (it shares the kernel with the convolution's weight / window registers).
// FC1: eight lanes per (clip, squeeze unit) pair, each summing every eighth channel, then a 3-step butterfly: the
// order of the additions depends only on the clip, never on how clips are grouped.
// FC2: thread = channel; gate[g][c] = sigmoid(b2[c] + s[g] . w2[:][c]) overwrites the pooled sums.
__device__ __forceinline__ void se_narrow(float* s_pool, float* s_se, const float* w_se1, const float* b_se1,
                                          const float* w_se2, const float* b_se2, int C, int se, int gn, float inv_npix) {
  const int tid = threadIdx.x, sub = tid & 7;
  for (int pair = tid >> 3; pair < ((gn * se + 3) & ~3); pair += kDwThreads / 8) {   // whole warps stay in the loop
    const bool ok = pair < gn * se;
    const int g = ok ? pair / se : 0, j = ok ? pair - g * se : 0;
    const float* pp = s_pool + g * C;
    const float* wp = w_se1 + j * C;
    float acc = 0.0f;
    for (int c = sub; c < C; c += 8) acc = fmaf(pp[c], wp[c], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (ok && sub == 0) s_se[pair] = swish(acc * inv_npix + b_se1[j]);
  }
  __syncthreads();
  if (tid < C) {
    const float b2 = b_se2[tid];
    for (int g = 0; g < gn; ++g) {
      float a = b2;
      for (int j = 0; j < se; ++j) a = fmaf(w_se2[j * C + tid], s_se[g * se + j], a);
      s_pool[g * C + tid] = sigmoidf(a);
    }
  }
  __syncthreads();
}

// block1a (GEOM 7: 64 KB of tiles per clip, 3x3 taps) fits three CTAs per SM in shared memory; its register cap follows
template <int K, int S, int GEOM>
__global__ void __launch_bounds__(kDwThreads, GEOM == 7 ? 3 : 2)     // <= 128 (80) registers: two (three) CTAs per SM
dwse_kernel(const uint16_t* __restrict__ x, int batch, int G, DwseParams P, uint16_t* __restrict__ y) {
  extern __shared__ __align__(128) uint8_t smem[];
  const DwSmem L = dw_smem(P, G);
  const uint32_t* s_in = reinterpret_cast<const uint32_t*>(smem);              // bf16x2 words, [G][H][W][C/2]
  const bool two_in = L.in2_off != 0;
  uint32_t* s_out = reinterpret_cast<uint32_t*>(smem + L.out_off);            // bf16x2, [G][Ho*Wo][C/2]
  float* s_pool = reinterpret_cast<float*>(smem + L.pooled_off);              // [G][C] sums, later gates
  float* s_part = reinterpret_cast<float*>(smem + L.part_off);                // [PL][G][C] (deterministic pool reduce)
  float* s_se = reinterpret_cast<float*>(smem + L.s_off);                     // [G][se]
  float* s_red = reinterpret_cast<float*>(smem + L.red_off);                  // [se][warps][4]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  const uint32_t* s_zero = reinterpret_cast<const uint32_t*>(smem + L.zero_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = P.C, C2 = C >> 1, npix = P.Ho * P.Wo;
  const int clip_words = P.H * P.W * C2;
  const float inv_npix = 1.0f / (float)npix;

  // weight pointers: smem-resident copies for narrow layers, global (L1/L2) otherwise
  const float *w_dw = P.w_dw, *b_dw = P.b_dw, *w_se1 = P.w_se1, *b_se1 = P.b_se1, *w_se2 = P.w_se2, *b_se2 = P.b_se2;
  if (L.w_floats) {
    float* sw = reinterpret_cast<float*>(smem + L.w_off);
    const int n_dw = K * K * P.C;
    for (int i = tid; i < n_dw; i += kDwThreads) sw[i] = __ldg(P.w_dw + i);
    for (int i = tid; i < P.C; i += kDwThreads) sw[n_dw + i] = __ldg(P.b_dw + i);
    w_dw = sw; b_dw = sw + n_dw;
    if (!P.se_external) {
      float* q = sw + n_dw + P.C;
      const int n_se = P.se * P.C;
      for (int i = tid; i < n_se; i += kDwThreads) { q[i] = __ldg(P.w_se1 + i); q[n_se + P.se + i] = __ldg(P.w_se2 + i); }
      for (int i = tid; i < P.se; i += kDwThreads) q[n_se + i] = __ldg(P.b_se1 + i);
      for (int i = tid; i < P.C; i += kDwThreads) q[2 * n_se + P.se + i] = __ldg(P.b_se2 + i);
      w_se1 = q; b_se1 = q + n_se; w_se2 = q + n_se + P.se; b_se2 = q + 2 * n_se + P.se;
    }
  }
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(bar + 1, 1);
    ptx::fence_barrier_init();
  }
  if (L.zero_off)
    for (int i = tid; i < 1024; i += kDwThreads) reinterpret_cast<uint32_t*>(smem + L.zero_off)[i] = 0u;
  __syncthreads();
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();              // weights / barrier set-up above overlapped the predecessor's tail

  // thread -> (channel pair, pixel lane)
  const int PL = L.PL;
  const int cp0 = C2 >= kDwThreads ? tid : tid % C2;
  const int pl = C2 >= kDwThreads ? 0 : tid / C2;
  const bool active = pl < PL;
  const int n_groups = (batch + G - 1) / G;
  auto issue_load = [&](int grp_, int buf) {                 // one thread: bulk copy of a whole group into buffer `buf`
    const int g0_ = grp_ * G;
    const uint32_t bytes = (uint32_t)min(G, batch - g0_) * clip_words * 4;
    ptx::mbar_expect_tx(bar + buf, bytes);
    ptx::tma_bulk_g2s(smem + (buf ? L.in2_off : 0u), x + (size_t)g0_ * clip_words * 2, bytes, bar + buf);
  };
  if (two_in && tid == 0 && (int)blockIdx.x < n_groups) issue_load(blockIdx.x, 0);
  int it = 0;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
    const int g0 = grp * G;
    const int gn = min(G, batch - g0);
    const int cur = two_in ? (it & 1) : 0;
    if (!two_in && tid == 0) issue_load(grp, 0);
    ptx::mbar_wait(bar + cur, two_in ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u));
    __syncthreads();
    // the other buffer was last read by the previous group's convolution, which every thread has left
    if (two_in && tid == 0 && grp + (int)gridDim.x < n_groups) issue_load(grp + gridDim.x, cur ^ 1);
    s_in = reinterpret_cast<const uint32_t*>(smem + (cur ? L.in2_off : 0u));

    // ---- depthwise conv + BN + swish -> s_out (16-bit), channel sums -> s_pool
    if constexpr (GEOM >= 1 && GEOM <= 6) {
      // balanced flat loop over (clip, channel pair) items: one thread computes every output pixel of its item
      constexpr int H = GEOM == 1 ? 7 : (GEOM <= 4 ? 4 : 2);
      constexpr int W = GEOM == 1 ? 5 : (GEOM <= 4 ? 3 : 2);
      constexpr int PT = (GEOM == 3 || GEOM == 5) ? 2 : 1;
      constexpr int PLFT = (GEOM == 3 || GEOM == 4 || GEOM == 5) ? 2 : 1;
      for (int item = tid; item < C2 * gn; item += kDwThreads) {
        const int g = item / C2, cp = item - g * C2;
        float2 wreg[K * K];
#pragma unroll
        for (int kk = 0; kk < K * K; ++kk) wreg[kk] = *(reinterpret_cast<const float2*>(w_dw + (size_t)kk * C) + cp);
        const float2 bias = *(reinterpret_cast<const float2*>(b_dw) + cp);
        float sum0 = 0.0f, sum1 = 0.0f;
        dw_small_item<K, S, H, W, PT, PLFT>(s_in + (size_t)g * clip_words + cp, C2, wreg, bias, P.bf16,
                                            s_out + (size_t)g * npix * C2 + cp, sum0, sum1);
        s_pool[g * C + 2 * cp] = sum0;
        s_pool[g * C + 2 * cp + 1] = sum1;
      }
    } else if (active) {
      for (int cp = cp0; cp < C2; cp += kDwThreads) {
        float2 wreg[K * K];
#pragma unroll
        for (int kk = 0; kk < K * K; ++kk) wreg[kk] = *(reinterpret_cast<const float2*>(w_dw + (size_t)kk * C) + cp);
        const float2 bias = *(reinterpret_cast<const float2*>(b_dw) + cp);
        for (int g = 0; g < gn; ++g) {
          const uint32_t* in_g = s_in + (size_t)g * clip_words + cp;
          float sum0 = 0.0f, sum1 = 0.0f;
          // Row strips with a K x K register window: the pixel lane `pl` owns output rows ho = pl, pl + PL, ...;
          // along a row the window slides by S columns, so each output costs K*S shared-memory loads instead of K*K.
          // Window columns live in slot (wo*S + kw) % K; the wo loop is unrolled by K so slots are compile-time.
          for (int ho = pl; ho < P.Ho; ho += PL) {
            const uint32_t* rowp[K];
            bool row_ok[K];
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
              const int r = ho * S + kh - P.pad_top;
              row_ok[kh] = (r >= 0) && (r < P.H);
              rowp[kh] = in_g + (size_t)(row_ok[kh] ? r : 0) * P.W * C2;
              if (GEOM >= 7 && !row_ok[kh]) rowp[kh] = s_zero + cp;
            }
            uint32_t* orow = s_out + ((size_t)g * npix + (size_t)ho * P.Wo) * C2 + cp;
            if constexpr (GEOM >= 7) {
              // block1a 25x20 k3 s1 C32 | block2a 25x20 k3 s2 C96 | block2b 13x10 k3 s1 C144 | block3a 13x10 k5 s2 C144 |
              // block3b 7x5 k5 s1 C240
              constexpr int GW = GEOM <= 8 ? 20 : (GEOM <= 10 ? 10 : 5);
              constexpr int GWO = GEOM == 7 ? 20 : (GEOM <= 9 ? 10 : 5);
              constexpr int GPL = (GEOM == 8) ? 0 : ((GEOM == 11) ? 2 : 1);
              constexpr int GC2 = GEOM == 7 ? 16 : (GEOM == 8 ? 48 : (GEOM <= 10 ? 72 : 120));
              dw_strip_row<K, S, GW, GWO, GPL, GC2>(rowp, wreg, bias, P.bf16, orow, sum0, sum1);
            } else {
            float2 win[K][K];
            for (int wo_base = 0; wo_base < P.Wo; wo_base += K) {
#pragma unroll
              for (int j = 0; j < K; ++j) {
                const int wo = wo_base + j;
                if (wo < P.Wo) {
#pragma unroll
                  for (int kw = 0; kw < K; ++kw) {
                    if (kw >= K - S || wo == 0) {                 // new columns (all K of them for the first output)
                      const int c = wo * S + kw - P.pad_left;
                      const bool col_ok = (c >= 0) && (c < P.W);
                      constexpr int dummy = 0; (void)dummy;
                      const int slot = (j * S + kw) % K;
#pragma unroll
                      for (int kh = 0; kh < K; ++kh) {
                        if (!row_ok[kh]) continue;
                        float2 v = make_float2(0.0f, 0.0f);
                        if (col_ok) v = ptx::unpack_h2(rowp[kh][c * C2], P.bf16);
                        win[kh][slot] = v;
                      }
                    }
                  }
                  float a0 = bias.x, a1 = bias.y;
#pragma unroll
                  for (int kh = 0; kh < K; ++kh) {
                    if (!row_ok[kh]) continue;                    // padded rows contribute nothing (tiny late maps: most rows)
#pragma unroll
                    for (int kw = 0; kw < K; ++kw) {
                      const int slot = (j * S + kw) % K;
                      a0 = fmaf(win[kh][slot].x, wreg[kh * K + kw].x, a0);
                      a1 = fmaf(win[kh][slot].y, wreg[kh * K + kw].y, a1);
                    }
                  }
                  a0 = swish(a0);
                  a1 = swish(a1);
                  sum0 += a0;
                  sum1 += a1;
                  orow[(size_t)wo * C2] = ptx::pack_h2(a0, a1, P.bf16);
                }
              }
            }
            }   // generic (run-time geometry) row
          }
          if (PL == 1) {
            s_pool[g * C + 2 * cp] = sum0;
            s_pool[g * C + 2 * cp + 1] = sum1;
          } else {
            s_part[(pl * gn + g) * C + 2 * cp] = sum0;
            s_part[(pl * gn + g) * C + 2 * cp + 1] = sum1;
          }
        }
      }
    }
    __syncthreads();
    if ((GEOM == 0 || GEOM >= 7) && PL > 1) {   // fixed-order sum over pixel lanes: results do not depend on scheduling
      for (int i = tid; i < gn * C; i += kDwThreads) {
        float a = 0.0f;
        for (int q = 0; q < PL; ++q) a += s_part[q * gn * C + i];
        s_pool[i] = a;
      }
      __syncthreads();
    }

    if (P.se_external) {
      // un-gated activation and channel means go to global; the SE GEMMs and the gating pass follow as separate launches
      const int vecs = gn * npix * (C >> 3);
      const uint4* src = reinterpret_cast<const uint4*>(s_out);
      uint4* dst = reinterpret_cast<uint4*>(y) + (size_t)g0 * npix * (C >> 3);
      for (int i = tid; i < vecs; i += kDwThreads) dst[i] = src[i];
      uint32_t* pdst = reinterpret_cast<uint32_t*>(P.pooled_out) + (size_t)g0 * C2;
      for (int i = tid; i < gn * C2; i += kDwThreads)
        pdst[i] = ptx::pack_h2(s_pool[2 * i] * inv_npix, s_pool[2 * i + 1] * inv_npix, P.bf16);
      __syncthreads();
      continue;
    }

    const bool narrow = C <= kDwThreads;
    if (narrow) {
      se_narrow(s_pool, s_se, w_se1, b_se1, w_se2, b_se2, C, P.se, gn, inv_npix);
    } else {
    // ---- SE reduce: s[g][j] = swish(b1[j] + mean_g . w1[j][:]).  Every thread owns channels c = tid + 256 i and
    // streams its slice of every weight row (coalesced, independent loads -> deep memory-level parallelism);
    // partial dot products are reduced by warp shuffles, then across the 8 warps through smem.
    {
      constexpr int kMaxCi = 5;                                   // ceil(1152 / 256)
      const int nci = (C + kDwThreads - 1) / kDwThreads;
      for (int gb = 0; gb < gn; gb += 4) {
        float pv[4][kMaxCi];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < kMaxCi; ++i) {
            const int c = tid + i * kDwThreads;
            pv[u][i] = (i < nci && c < C && gb + u < gn) ? s_pool[(gb + u) * C + c] : 0.0f;
          }
#pragma unroll 4
        for (int j = 0; j < P.se; ++j) {
          const float* wrow = w_se1 + (size_t)j * C;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int i = 0; i < kMaxCi; ++i) {
            const int c = tid + i * kDwThreads;
            if (i < nci && c < C) {
              const float wv = wrow[c];
              a0 = fmaf(wv, pv[0][i], a0); a1 = fmaf(wv, pv[1][i], a1);
              a2 = fmaf(wv, pv[2][i], a2); a3 = fmaf(wv, pv[3][i], a3);
            }
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
          }
          if (lane == 0) *reinterpret_cast<float4*>(&s_red[(j * (kDwThreads / 32) + warp) * 4]) = make_float4(a0, a1, a2, a3);
        }
        __syncthreads();
        for (int t = tid; t < P.se * 4; t += kDwThreads) {
          const int j = t >> 2, u = t & 3;
          if (gb + u < gn) {
            float a = 0.f;
#pragma unroll
            for (int wv = 0; wv < kDwThreads / 32; ++wv) a += s_red[(j * (kDwThreads / 32) + wv) * 4 + u];
            s_se[(gb + u) * P.se + j] = swish(a * inv_npix + b_se1[j]);
          }
        }
        __syncthreads();
      }
    }

    // ---- SE expand: gate[g][c] = sigmoid(b2[c] + s[g] . w2[:][c])  (overwrites the pooled sums)
    for (int c = tid; c < C; c += kDwThreads) {
      const float b2 = b_se2[c];
      for (int gb = 0; gb < gn; gb += 8) {
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = b2;
#pragma unroll 16
        for (int j = 0; j < P.se; ++j) {
          const float wv = w_se2[(size_t)j * C + c];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (gb + u < gn) acc[u] = fmaf(wv, s_se[(gb + u) * P.se + j], acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (gb + u < gn) s_pool[(gb + u) * C + c] = sigmoidf(acc[u]);
      }
    }
    __syncthreads();

    }   // wide in-kernel SE

    // ---- scale and store: 16-byte vectors (8 channels).  A thread keeps one channel vector (its 8 gates stay in
    // registers for the clip) and walks the pixels; consecutive threads write consecutive 16-byte pieces.
    {
      const int C8 = C >> 3;                                  // uint4 vectors per pixel
      const int PLg = kDwThreads / C8;                        // pixel lanes (C <= 1152 -> C8 <= 144 -> PLg >= 1)
      const int c8 = tid % C8, pl8 = tid / C8;
      if (pl8 < PLg) {
        for (int g = 0; g < gn; ++g) {
          const uint4* src = reinterpret_cast<const uint4*>(s_out) + (size_t)g * npix * C8 + c8;
          uint4* dst = reinterpret_cast<uint4*>(y) + ((size_t)(g0 + g) * npix * C8) + c8;
          const float4 g0v = *reinterpret_cast<const float4*>(s_pool + g * C + 8 * c8);
          const float4 g1v = *reinterpret_cast<const float4*>(s_pool + g * C + 8 * c8 + 4);
          for (int px = pl8; px < npix; px += PLg) {
            const uint4 v = src[(size_t)px * C8];
            float2 x;
            uint4 o;
            x = ptx::unpack_h2(v.x, P.bf16); o.x = ptx::pack_h2(x.x * g0v.x, x.y * g0v.y, P.bf16);
            x = ptx::unpack_h2(v.y, P.bf16); o.y = ptx::pack_h2(x.x * g0v.z, x.y * g0v.w, P.bf16);
            x = ptx::unpack_h2(v.z, P.bf16); o.z = ptx::pack_h2(x.x * g1v.x, x.y * g1v.y, P.bf16);
            x = ptx::unpack_h2(v.w, P.bf16); o.w = ptx::pack_h2(x.x * g1v.z, x.y * g1v.w, P.bf16);
            dst[(size_t)px * C8] = o;
          }
        }
      }
    }
    __syncthreads();
  }
}

// gating pass of the external-SE path: one CTA per clip (grid-stride), 16-byte vectors
__global__ void __launch_bounds__(256)
se_scale_kernel(uint16_t* __restrict__ y, const uint16_t* __restrict__ gates, int batch, int npix, int C, int bf16) {
  const int C8 = C >> 3, tid = threadIdx.x;
  const int vec_per_clip = npix * C8;
  const int step = 256 % C8;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  for (int clip = blockIdx.x; clip < batch; clip += gridDim.x) {
    uint4* dst = reinterpret_cast<uint4*>(y) + (size_t)clip * vec_per_clip;
    const uint4* g = reinterpret_cast<const uint4*>(gates) + (size_t)clip * C8;
    int c8 = tid % C8;
    for (int i = tid; i < vec_per_clip; i += 256) {
      const uint4 v = dst[i];
      const uint4 gv = __ldg(g + c8);
      float2 a, b;
      uint4 o;
      a = ptx::unpack_h2(v.x, bf16); b = ptx::unpack_h2(gv.x, bf16); o.x = ptx::pack_h2(a.x * b.x, a.y * b.y, bf16);
      a = ptx::unpack_h2(v.y, bf16); b = ptx::unpack_h2(gv.y, bf16); o.y = ptx::pack_h2(a.x * b.x, a.y * b.y, bf16);
      a = ptx::unpack_h2(v.z, bf16); b = ptx::unpack_h2(gv.z, bf16); o.z = ptx::pack_h2(a.x * b.x, a.y * b.y, bf16);
      a = ptx::unpack_h2(v.w, bf16); b = ptx::unpack_h2(gv.w, bf16); o.w = ptx::pack_h2(a.x * b.x, a.y * b.y, bf16);
      dst[i] = o;
      c8 += step;
      if (c8 >= C8) c8 -= C8;
    }
  }
}

}  // namespace

int launch_se_scale(void* d_y, const void* d_gates, int batch, int npix, int C, int bf16, int sm_count, cudaStream_t st) {
  if (batch == 0) return KWS_OK;
  const int grid = batch < sm_count * 8 ? batch : sm_count * 8;
  KWS_CUDA_CHECK(launch_pdl(se_scale_kernel, dim3(grid), dim3(256), 0, st, static_cast<uint16_t*>(d_y),
                            static_cast<const uint16_t*>(d_gates), batch, npix, C, bf16));
  return KWS_OK;
}

int launch_stem(const float* d_feats, int batch, const StemParams& P, void* d_out, int sm_count,
                cudaStream_t st) {
  if (batch == 0) return KWS_OK;
  const size_t smem = (size_t)((((P.H + 2) * (P.W + 2) + 3) & ~3) + 9 * kStemC + kStemC) * 4;
  const int grid = batch < sm_count * 4 ? batch : sm_count * 4;
  KWS_CUDA_CHECK(launch_pdl(stem_conv_kernel, dim3(grid), dim3(kStemThreads), smem, st, d_feats, batch, P,
                            static_cast<uint16_t*>(d_out)));
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

int dwse_pick_group(const DwseParams& P, int max_smem, int batch, int sm_count) {
  const int cands[5] = {16, 8, 4, 2, 1};
  int fit = 0;
  for (int i = 0; i < 5 && !fit; ++i)
    if ((int)dw_smem(P, cands[i]).total <= 100 * 1024) fit = cands[i];
  for (int i = 0; i < 5 && !fit; ++i)
    if ((int)dw_smem(P, cands[i]).total <= max_smem) fit = cands[i];
  if (!fit) return 0;
  while (fit > 1 && (batch + fit - 1) / fit < sm_count) fit >>= 1;       // at least one group per SM
  return fit;
}

int launch_dwse(const void* d_x, int batch, const DwseParams& P, void* d_y, int G, int sm_count,
                cudaStream_t st) {
  if (batch == 0) return KWS_OK;
  KWS_REQUIRE(G >= 1, "dwse: layer does not fit shared memory");
  KWS_REQUIRE(P.C % 8 == 0, "dwse: channels must be a multiple of 8");
  KWS_REQUIRE((P.K == 3 || P.K == 5) && (P.S == 1 || P.S == 2), "dwse: unsupported kernel %d / stride %d", P.K, P.S);
  const size_t smem = dw_smem(P, G).total;
  void (*kern)(const uint16_t*, int, int, DwseParams, uint16_t*) =
      P.K == 3 ? (P.S == 1 ? dwse_kernel<3, 1, 0> : dwse_kernel<3, 2, 0>) : (P.S == 1 ? dwse_kernel<5, 1, 0> : dwse_kernel<5, 2, 0>);
  // tiny-map specialisations (geometry must match exactly, else the generic kernel handles it)
  auto is = [&](int k, int s_, int h, int w, int pt, int pl) {
    return P.K == k && P.S == s_ && P.H == h && P.W == w && P.pad_top == pt && P.pad_left == pl;
  };
  if (is(3, 2, 7, 5, 1, 1)) kern = dwse_kernel<3, 2, 1>;
  else if (is(3, 1, 4, 3, 1, 1)) kern = dwse_kernel<3, 1, 2>;
  else if (is(5, 1, 4, 3, 2, 2)) kern = dwse_kernel<5, 1, 3>;
  else if (is(5, 2, 4, 3, 1, 2)) kern = dwse_kernel<5, 2, 4>;
  else if (is(5, 1, 2, 2, 2, 2)) kern = dwse_kernel<5, 1, 5>;
  else if (is(3, 1, 2, 2, 1, 1)) kern = dwse_kernel<3, 1, 6>;
  else if (is(3, 1, 25, 20, 1, 1) && P.C == 32) kern = dwse_kernel<3, 1, 7>;
  else if (is(3, 2, 25, 20, 1, 0) && P.C == 96) kern = dwse_kernel<3, 2, 8>;
  else if (is(3, 1, 13, 10, 1, 1) && P.C == 144) kern = dwse_kernel<3, 1, 9>;
  else if (is(5, 2, 13, 10, 2, 1) && P.C == 144) kern = dwse_kernel<5, 2, 10>;
  else if (is(5, 1, 7, 5, 2, 2) && P.C == 240) kern = dwse_kernel<5, 1, 11>;
  KWS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_groups = (batch + G - 1) / G;
  const int per_sm = (kern == dwse_kernel<3, 1, 7> && smem <= 74 * 1024) ? 3 : (smem <= 100 * 1024 ? 2 : 1);
  KWS_REQUIRE(smem <= (size_t)kDwSmemCap, "dwse: %zu bytes of shared memory exceed the per-CTA limit", smem);
  const int grid = n_groups < sm_count * per_sm ? n_groups : sm_count * per_sm;
  KWS_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kDwThreads), smem, st, static_cast<const uint16_t*>(d_x), batch, G, P,
                            static_cast<uint16_t*>(d_y)));
  return KWS_OK;
}

}  // namespace kws
