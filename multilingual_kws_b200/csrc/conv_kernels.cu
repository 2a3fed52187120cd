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

__device__ __forceinline__ float swish(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoidf(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// ---------------------------------------------------------------- stem
constexpr int kStemThreads = 256;
constexpr int kStemC = 32;

__global__ void __launch_bounds__(kStemThreads)
stem_conv_kernel(const float* __restrict__ feats, int batch, StemParams P, uint16_t* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* s_in = reinterpret_cast<float*>(smem);                     // [H*W]
  float* s_w = s_in + ((P.H * P.W + 3) & ~3);                        // [9][32]
  float* s_b = s_w + 9 * kStemC;                                     // [32]
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * kStemC; i += kStemThreads) s_w[i] = P.w[i];
  if (tid < kStemC) s_b[tid] = P.bias[tid];
  const int npix = P.Ho * P.Wo;
  for (int clip = blockIdx.x; clip < batch; clip += gridDim.x) {
    __syncthreads();
    const float* src = feats + (size_t)clip * P.H * P.W;
    for (int i = tid; i < P.H * P.W; i += kStemThreads) s_in[i] = fmaf(__ldg(src + i), P.in_scale, P.in_shift);
    __syncthreads();
    // item = (pixel, group of 8 output channels)
    for (int item = tid; item < npix * 4; item += kStemThreads) {
      const int p = item >> 2, g = item & 3;
      const int ho = p / P.Wo, wo = p - ho * P.Wo;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = s_b[g * 8 + j];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int r = ho * 2 + kh - P.pad_top;
        if (r < 0 || r >= P.H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int c = wo * 2 + kw - P.pad_left;
          if (c < 0 || c >= P.W) continue;
          const float x = s_in[r * P.W + c];
          const float* wv = s_w + (kh * 3 + kw) * kStemC + g * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(x, wv[j], acc[j]);
        }
      }
      uint4 pk;
      pk.x = ptx::pack_h2(swish(acc[0]), swish(acc[1]), P.bf16);
      pk.y = ptx::pack_h2(swish(acc[2]), swish(acc[3]), P.bf16);
      pk.z = ptx::pack_h2(swish(acc[4]), swish(acc[5]), P.bf16);
      pk.w = ptx::pack_h2(swish(acc[6]), swish(acc[7]), P.bf16);
      *reinterpret_cast<uint4*>(out + ((size_t)clip * npix + p) * kStemC + g * 8) = pk;
    }
  }
}

// ---------------------------------------------------------------- depthwise + SE
// Two kernels per MBConv middle section (the inputs were just written by the expand GEMM and are L2-resident):
//   dwpool_kernel   depthwise conv + BN + swish straight from / to global memory (no staging, small smem, so
//                   several CTAs per SM hide the load latency) and the per-clip channel sums (global average pool);
//   se_scale_kernel per group of clips: SE reduce FC + swish, SE expand FC + sigmoid, then the gated activation is
//                   written in place as the next GEMM's A operand.
constexpr int kDwThreads = 256;

// thread -> (channel pair, pixel lane): PL lanes share one channel pair when C/2 < 256
__host__ __device__ inline int dw_pixel_lanes(int C) { return (C >> 1) >= kDwThreads ? 1 : kDwThreads / (C >> 1); }

template <int K, int S>
__global__ void __launch_bounds__(kDwThreads, 2)
dwpool_kernel(const uint16_t* __restrict__ x, int batch, DwseParams P, uint16_t* __restrict__ y,
              float* __restrict__ pooled) {
  extern __shared__ __align__(128) uint8_t smem_dw[];
  float* s_part = reinterpret_cast<float*>(smem_dw);         // [2][PL][C] partial channel sums (PL > 1 only)
  const int tid = threadIdx.x;
  const int C = P.C, C2 = C >> 1, npix = P.Ho * P.Wo;
  const size_t clip_words = (size_t)P.H * P.W * C2;
  const int PL = dw_pixel_lanes(C);
  const int cp0 = C2 >= kDwThreads ? tid : tid % C2;
  const int pl = C2 >= kDwThreads ? 0 : tid / C2;
  const bool active = pl < PL;
  const uint32_t* x32 = reinterpret_cast<const uint32_t*>(x);
  uint32_t* y32 = reinterpret_cast<uint32_t*>(y);

  for (int cp = cp0; cp < C2; cp += kDwThreads) {             // one pass when C2 <= 256
    float2 wreg[K * K];
#pragma unroll
    for (int kk = 0; kk < K * K; ++kk) wreg[kk] = __ldg(reinterpret_cast<const float2*>(P.w_dw + (size_t)kk * C) + cp);
    const float2 bias = __ldg(reinterpret_cast<const float2*>(P.b_dw) + cp);
    int buf = 0;
    for (int clip = blockIdx.x; clip < batch; clip += gridDim.x) {
      float sum0 = 0.0f, sum1 = 0.0f;
      if (active) {
        const uint32_t* in_g = x32 + (size_t)clip * clip_words + cp;
        // Row strips with a K x K register window: the pixel lane `pl` owns output rows ho = pl, pl + PL, ...; along a
        // row the window slides by S columns, so each output costs K*S loads instead of K*K.  Window columns live in
        // slot (wo*S + kw) % K; the wo loop is unrolled by K so slots are compile-time.
        for (int ho = pl; ho < P.Ho; ho += PL) {
          const uint32_t* rowp[K];
          bool row_ok[K];
#pragma unroll
          for (int kh = 0; kh < K; ++kh) {
            const int r = ho * S + kh - P.pad_top;
            row_ok[kh] = (r >= 0) && (r < P.H);
            rowp[kh] = in_g + (size_t)(row_ok[kh] ? r : 0) * P.W * C2;
          }
          float2 win[K][K];
          uint32_t* orow = y32 + ((size_t)clip * npix + (size_t)ho * P.Wo) * C2 + cp;
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
                    const int slot = (j * S + kw) % K;
#pragma unroll
                    for (int kh = 0; kh < K; ++kh) {
                      if (!row_ok[kh]) continue;
                      float2 v = make_float2(0.0f, 0.0f);
                      if (col_ok) v = ptx::unpack_h2(__ldg(rowp[kh] + (size_t)c * C2), P.bf16);
                      win[kh][slot] = v;
                    }
                  }
                }
                float a0 = bias.x, a1 = bias.y;
#pragma unroll
                for (int kh = 0; kh < K; ++kh) {
                  if (!row_ok[kh]) continue;                    // padded rows contribute nothing
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
        }
      }
      if (PL == 1) {
        if (active) *reinterpret_cast<float2*>(pooled + (size_t)clip * C + 2 * cp) = make_float2(sum0, sum1);
      } else {
        // fixed-order reduce over the pixel lanes (deterministic); double-buffered so one barrier per clip suffices
        float* part = s_part + (size_t)buf * PL * C;
        if (active) *reinterpret_cast<float2*>(part + (size_t)pl * C + 2 * cp) = make_float2(sum0, sum1);
        __syncthreads();
        if (tid < C) {
          float a = 0.0f;
          for (int q = 0; q < PL; ++q) a += part[q * C + tid];
          pooled[(size_t)clip * C + tid] = a;
        }
        buf ^= 1;
      }
    }
  }
}

struct SeSmem {
  uint32_t s_off, red_off, total;
};
__host__ __device__ inline SeSmem se_smem(const DwseParams& P, int G) {
  SeSmem L;
  L.s_off = (uint32_t)G * P.C * 4;                                              // after s_pool [G][C]
  L.red_off = (L.s_off + (uint32_t)G * P.se * 4 + 15) & ~15u;                   // [se][8 warps][4 clips] FC1 partials
  L.total = L.red_off + (uint32_t)P.se * (kDwThreads / 32) * 4 * 4;
  return L;
}

__global__ void __launch_bounds__(kDwThreads)
se_scale_kernel(const float* __restrict__ pooled, int batch, int G, DwseParams P, uint16_t* __restrict__ y) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SeSmem L = se_smem(P, G);
  float* s_pool = reinterpret_cast<float*>(smem);                               // [G][C] sums, later gates
  float* s_se = reinterpret_cast<float*>(smem + L.s_off);                      // [G][se]
  float* s_red = reinterpret_cast<float*>(smem + L.red_off);                   // [se][warps][4]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = P.C, npix = P.Ho * P.Wo;
  const float inv_npix = 1.0f / (float)npix;
  const int n_groups = (batch + G - 1) / G;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int g0 = grp * G;
    const int gn = min(G, batch - g0);
    __syncthreads();
    for (int i = tid; i < gn * C; i += kDwThreads) s_pool[i] = pooled[(size_t)g0 * C + i];
    __syncthreads();

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
          const float* wrow = P.w_se1 + (size_t)j * C;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int i = 0; i < kMaxCi; ++i) {
            const int c = tid + i * kDwThreads;
            if (i < nci && c < C) {
              const float wv = __ldg(wrow + c);
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
            s_se[(gb + u) * P.se + j] = swish(a * inv_npix + __ldg(P.b_se1 + j));
          }
        }
        __syncthreads();
      }
    }

    // ---- SE expand: gate[g][c] = sigmoid(b2[c] + s[g] . w2[:][c])  (overwrites the pooled sums)
    for (int c = tid; c < C; c += kDwThreads) {
      const float b2 = __ldg(P.b_se2 + c);
      for (int gb = 0; gb < gn; gb += 8) {
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = b2;
#pragma unroll 16
        for (int j = 0; j < P.se; ++j) {
          const float wv = __ldg(P.w_se2 + (size_t)j * C + c);
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

    // ---- scale in place: 16-byte vectors (8 channels), coalesced; no div/mod in the loop
    {
      const int C8 = C >> 3;                                  // uint4 vectors per pixel
      const int vec_per_clip = npix * C8;
      const int step = kDwThreads % C8;
      for (int g = 0; g < gn; ++g) {
        uint4* dst = reinterpret_cast<uint4*>(y) + ((size_t)(g0 + g) * vec_per_clip);
        const float* gate = s_pool + g * C;
        int c8 = tid % C8;
        for (int i = tid; i < vec_per_clip; i += kDwThreads) {
          const uint4 v = dst[i];
          const float4 g0v = *reinterpret_cast<const float4*>(gate + 8 * c8);
          const float4 g1v = *reinterpret_cast<const float4*>(gate + 8 * c8 + 4);
          float2 xv;
          uint4 o;
          xv = ptx::unpack_h2(v.x, P.bf16); o.x = ptx::pack_h2(xv.x * g0v.x, xv.y * g0v.y, P.bf16);
          xv = ptx::unpack_h2(v.y, P.bf16); o.y = ptx::pack_h2(xv.x * g0v.z, xv.y * g0v.w, P.bf16);
          xv = ptx::unpack_h2(v.z, P.bf16); o.z = ptx::pack_h2(xv.x * g1v.x, xv.y * g1v.y, P.bf16);
          xv = ptx::unpack_h2(v.w, P.bf16); o.w = ptx::pack_h2(xv.x * g1v.z, xv.y * g1v.w, P.bf16);
          dst[i] = o;
          c8 += step;
          if (c8 >= C8) c8 -= C8;
        }
      }
    }
  }
}

}  // namespace

int launch_stem(const float* d_feats, int batch, const StemParams& P, void* d_out, int sm_count,
                cudaStream_t st) {
  if (batch == 0) return KWS_OK;
  const size_t smem = (size_t)(((P.H * P.W + 3) & ~3) + 9 * kStemC + kStemC) * 4;
  const int grid = batch < sm_count * 4 ? batch : sm_count * 4;
  stem_conv_kernel<<<grid, kStemThreads, smem, st>>>(d_feats, batch, P, static_cast<uint16_t*>(d_out));
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

int launch_dwse(const void* d_x, int batch, const DwseParams& P, void* d_y, float* d_pooled, int sm_count,
                cudaStream_t st) {
  if (batch == 0) return KWS_OK;
  KWS_REQUIRE(P.C % 8 == 0, "dwse: channels must be a multiple of 8");
  KWS_REQUIRE((P.K == 3 || P.K == 5) && (P.S == 1 || P.S == 2), "dwse: unsupported kernel %d / stride %d", P.K, P.S);
  KWS_REQUIRE(d_pooled != nullptr, "dwse: pooled scratch is NULL");
  // 1. depthwise + BN + swish + channel sums
  {
    void (*kern)(const uint16_t*, int, DwseParams, uint16_t*, float*) =
        P.K == 3 ? (P.S == 1 ? dwpool_kernel<3, 1> : dwpool_kernel<3, 2>)
                 : (P.S == 1 ? dwpool_kernel<5, 1> : dwpool_kernel<5, 2>);
    const int PL = dw_pixel_lanes(P.C);
    const size_t smem = PL > 1 ? (size_t)2 * PL * P.C * 4 : 0;
    const int grid = batch < sm_count * 4 ? batch : sm_count * 4;
    kern<<<grid, kDwThreads, smem, st>>>(static_cast<const uint16_t*>(d_x), batch, P, static_cast<uint16_t*>(d_y), d_pooled);
    KWS_CUDA_CHECK(cudaGetLastError());
  }
  // 2. squeeze-excite gates + in-place scaling; clips per CTA chosen so every SM gets a group
  {
    int G = 8;
    while (G > 1 && (batch + G - 1) / G < 2 * sm_count) G >>= 1;
    const size_t smem = se_smem(P, G).total;
    const int n_groups = (batch + G - 1) / G;
    const int grid = n_groups < sm_count * 4 ? n_groups : sm_count * 4;
    se_scale_kernel<<<grid, kDwThreads, smem, st>>>(d_pooled, batch, G, P, static_cast<uint16_t*>(d_y));
    KWS_CUDA_CHECK(cudaGetLastError());
  }
  return KWS_OK;
}

}  // namespace kws
