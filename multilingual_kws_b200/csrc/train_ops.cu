// train_ops.cu — the non-GEMM kernels of the fine-tune backward pass through the top of the embedding (C ABI kws_train_*).
//
// Phase 2 of the reference's transfer_learn (multilingual_kws/embedding/transfer_learning.py:97-112: "unfreeze the top 20
// layers while leaving BatchNorm layers frozen", then a second fit with Adam(embedding_lr)) trains, next to the 3-way head,
// the last 20 layers of the embedding: block7a (expand / depthwise / squeeze-excite / project), top_conv and the dense
// tower.  Every contraction of that slice — forward, data gradient and weight gradient — runs on the tcgen05 GEMM
// (kws_gemm_h16); this file holds what sits between the GEMMs: activations and their derivatives, the 2x2 depthwise
// conv and its two gradients, the squeeze-excite gate, layout transposes for the weight-gradient GEMMs (which contract
// over the batch), column sums for bias gradients, and Keras-Adam with the frozen BatchNorm scale folded in.
// Activations and activation gradients are fp16 (gradients carry a static loss scale), sums and parameters fp32.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.h"

using namespace kws;

namespace {

constexpr int kT = 256;

__device__ __forceinline__ float h2f(uint16_t v) { return __half2float(__ushort_as_half(v)); }
__device__ __forceinline__ uint16_t f2h(float v) {
  // saturating: a gradient that overflows fp16 must not become inf
  v = fminf(fmaxf(v, -65504.0f), 65504.0f);
  return __half_as_ushort(__float2half_rn(v));
}
__device__ __forceinline__ float sigmoid_x(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float swish_x(float x) { return x * sigmoid_x(x); }
__device__ __forceinline__ float dswish_x(float x) {
  const float s = sigmoid_x(x);
  return s * (1.0f + x * (1.0f - s));
}

// out[c][r] = in[r][c]
__global__ void __launch_bounds__(256) transpose_h16_kernel(const uint16_t* __restrict__ in, int rows, int cols, int ld_out,
                                                            uint16_t* __restrict__ out) {
  __shared__ uint16_t tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : (uint16_t)0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) out[(size_t)c * ld_out + r] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(kT) swish_fwd_kernel(const uint16_t* __restrict__ z, uint16_t* __restrict__ a, size_t n) {
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) a[i] = f2h(swish_x(h2f(z[i])));
}

// h0[b][c] = mean_p swish(t[b][p][c])
__global__ void __launch_bounds__(kT) gap_swish_fwd_kernel(const uint16_t* __restrict__ t, int B, int P, int C,
                                                           uint16_t* __restrict__ h0) {
  const size_t n = (size_t)B * C;
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) {
    const size_t b = i / C, c = i - b * C;
    float s = 0.0f;
    for (int p = 0; p < P; ++p) s += swish_x(h2f(t[(b * P + p) * C + c]));
    h0[i] = f2h(s / (float)P);
  }
}
// dt[b][p][c] = scale * dh0[b][c] / P * swish'(t[b][p][c])
__global__ void __launch_bounds__(kT) gap_swish_bwd_kernel(const uint16_t* __restrict__ dh0, const uint16_t* __restrict__ t,
                                                           int B, int P, int C, uint16_t* __restrict__ dt) {
  const size_t n = (size_t)B * P * C;
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) {
    const size_t c = i % C, b = i / ((size_t)P * C);
    dt[i] = f2h(h2f(dh0[b * C + c]) / (float)P * dswish_x(h2f(t[i])));
  }
}

struct DwGeom { int H, W, K, S, pad_top, pad_left, Ho, Wo; };

// depthwise conv of a tiny map, thread = (clip, channel): z = conv(e) + shift, d = swish(z), pooled = mean_p d
__global__ void __launch_bounds__(kT) dw_fwd_kernel(const uint16_t* __restrict__ e, const float* __restrict__ w,
                                                    const float* __restrict__ shift, int B, int C, DwGeom G,
                                                    uint16_t* __restrict__ z, uint16_t* __restrict__ d,
                                                    uint16_t* __restrict__ pooled) {
  const size_t n = (size_t)B * C;
  const int Pi = G.H * G.W, Po = G.Ho * G.Wo;
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) {
    const size_t b = i / C, c = i - b * C;
    const uint16_t* eb = e + b * Pi * C + c;
    const float sh = shift[c];
    float sum = 0.0f;
    for (int ho = 0; ho < G.Ho; ++ho)
      for (int wo = 0; wo < G.Wo; ++wo) {
        float acc = sh;
        for (int kh = 0; kh < G.K; ++kh) {
          const int r = ho * G.S + kh - G.pad_top;
          if (r < 0 || r >= G.H) continue;
          for (int kw = 0; kw < G.K; ++kw) {
            const int cl = wo * G.S + kw - G.pad_left;
            if (cl < 0 || cl >= G.W) continue;
            acc = fmaf(h2f(eb[(size_t)(r * G.W + cl) * C]), w[(size_t)(kh * G.K + kw) * C + c], acc);
          }
        }
        const size_t o = (b * Po + ho * G.Wo + wo) * C + c;
        const float a = swish_x(acc);
        z[o] = f2h(acc);
        d[o] = f2h(a);
        sum += a;
      }
    pooled[i] = f2h(sum / (float)Po);
  }
}

// dg[b][p][c] = d[b][p][c] * g[b][c]
__global__ void __launch_bounds__(kT) gate_fwd_kernel(const uint16_t* __restrict__ d, const uint16_t* __restrict__ g, int B,
                                                      int P, int C, uint16_t* __restrict__ out) {
  const size_t n = (size_t)B * P * C;
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) {
    const size_t c = i % C, b = i / ((size_t)P * C);
    out[i] = f2h(h2f(d[i]) * h2f(g[b * C + c]));
  }
}
// dd[b][p][c] = ddg * g;   dgpre[b][c] = (sum_p ddg * d) * g (1 - g)
__global__ void __launch_bounds__(kT) gate_bwd_kernel(const uint16_t* __restrict__ ddg, const uint16_t* __restrict__ d,
                                                      const uint16_t* __restrict__ g, int B, int P, int C,
                                                      uint16_t* __restrict__ dd, uint16_t* __restrict__ dgpre) {
  const size_t n = (size_t)B * C;
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) {
    const size_t b = i / C, c = i - b * C;
    const float gv = h2f(g[i]);
    float s = 0.0f;
    for (int p = 0; p < P; ++p) {
      const size_t o = (b * P + p) * C + c;
      const float v = h2f(ddg[o]);
      s = fmaf(v, h2f(d[o]), s);
      dd[o] = f2h(v * gv);
    }
    dgpre[i] = f2h(s * gv * (1.0f - gv));
  }
}

// backward of dw_fwd for one batch slab per blockIdx.y: dz = (dd + dp / Po) * swish'(z); de = conv^T(dz);
// per-slab partial sums of dw (fp32), reduced in slab order by dw_bwd_reduce_kernel (deterministic)
__global__ void __launch_bounds__(kT) dw_bwd_kernel(const uint16_t* __restrict__ dd, const uint16_t* __restrict__ dp,
                                                    const uint16_t* __restrict__ z, const uint16_t* __restrict__ e,
                                                    const float* __restrict__ w, int B, int C, DwGeom G, int slab,
                                                    uint16_t* __restrict__ de, float* __restrict__ partial) {
  const int c = blockIdx.x * kT + threadIdx.x;
  if (c >= C) return;
  const int Pi = G.H * G.W, Po = G.Ho * G.Wo, KK = G.K * G.K;
  float dwacc[25];
  for (int t = 0; t < KK; ++t) dwacc[t] = 0.0f;
  const int b0 = blockIdx.y * slab, b1 = min(B, b0 + slab);
  for (int b = b0; b < b1; ++b) {
    const float dpv = h2f(dp[(size_t)b * C + c]) / (float)Po;
    float dein[49];
    for (int q = 0; q < Pi; ++q) dein[q] = 0.0f;
    for (int ho = 0; ho < G.Ho; ++ho)
      for (int wo = 0; wo < G.Wo; ++wo) {
        const size_t o = ((size_t)b * Po + ho * G.Wo + wo) * C + c;
        const float dz = (h2f(dd[o]) + dpv) * dswish_x(h2f(z[o]));
        for (int kh = 0; kh < G.K; ++kh) {
          const int r = ho * G.S + kh - G.pad_top;
          if (r < 0 || r >= G.H) continue;
          for (int kw = 0; kw < G.K; ++kw) {
            const int cl = wo * G.S + kw - G.pad_left;
            if (cl < 0 || cl >= G.W) continue;
            const int t = kh * G.K + kw, q = r * G.W + cl;
            dwacc[t] = fmaf(dz, h2f(e[((size_t)b * Pi + q) * C + c]), dwacc[t]);
            dein[q] = fmaf(dz, w[(size_t)t * C + c], dein[q]);
          }
        }
      }
    for (int q = 0; q < Pi; ++q) de[((size_t)b * Pi + q) * C + c] = f2h(dein[q]);
  }
  for (int t = 0; t < KK; ++t) partial[((size_t)blockIdx.y * KK + t) * C + c] = dwacc[t];
}
__global__ void __launch_bounds__(kT) dw_bwd_reduce_kernel(const float* __restrict__ partial, int slabs, int n,
                                                           float* __restrict__ out) {
  const int i = blockIdx.x * kT + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int k = 0; k < slabs; ++k) s += partial[(size_t)k * n + i];
  out[i] = s;
}

// dz = dy * act'(ref).  kind 0: relu from its output (fp16), 1: selu from its output (fp32), 2: swish from its
// pre-activation (fp16).  dy16 (kind 0, 2) or dy32 (kind 1: the head hands over an fp32 gradient, scaled here).
__global__ void __launch_bounds__(kT) act_bwd_kernel(int kind, const void* __restrict__ dy, const void* __restrict__ ref,
                                                     size_t n, float scale, uint16_t* __restrict__ dz) {
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) {
    float g, d;
    if (kind == 1) {
      g = static_cast<const float*>(dy)[i] * scale;
      const float o = static_cast<const float*>(ref)[i];
      d = o > 0.0f ? 1.0507009873554805f : o + 1.7580993408473766f;     // lambda, or lambda alpha e^z = out + lambda alpha
    } else {
      g = h2f(static_cast<const uint16_t*>(dy)[i]) * scale;
      const float r = h2f(static_cast<const uint16_t*>(ref)[i]);
      d = kind == 0 ? (r > 0.0f ? 1.0f : 0.0f) : dswish_x(r);
    }
    dz[i] = f2h(g * d);
  }
}

// out[c] = sum_r x[r][c]  (fixed order; one thread per column)
__global__ void __launch_bounds__(kT) colsum_kernel(const uint16_t* __restrict__ x, int rows, int cols, float* __restrict__ out) {
  const int c = blockIdx.x * kT + threadIdx.x;
  if (c >= cols) return;
  float s = 0.0f;
  for (int r = 0; r < rows; ++r) s += h2f(x[(size_t)r * cols + c]);
  out[c] = s;
}

// Keras Adam on an fp32 master tensor [rows][cols]; the gradient arrives as a SUM over the global batch of the loss-scaled
// per-sample gradients with respect to the FOLDED weight (w * row_scale): g = grad * row_scale / (count * loss_scale).
// Writes the 16-bit forward copy (w * row_scale) the GEMMs read.
__global__ void __launch_bounds__(kT) adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                  const float* __restrict__ grad, size_t n, int cols,
                                                  const float* __restrict__ row_scale, const float* __restrict__ count,
                                                  float inv_loss_scale, float lr_host, const float* __restrict__ lr_dev,
                                                  float b1, float b2, float eps,
                                                  uint16_t* __restrict__ out16, float* __restrict__ out32) {
  const float inv = inv_loss_scale / fmaxf(*count, 1.0f);
  const float lr_t = lr_dev ? *lr_dev : lr_host;
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) {
    const float rs = row_scale ? row_scale[i / cols] : 1.0f;
    const float g = grad[i] * rs * inv;
    const float mi = b1 * m[i] + (1.0f - b1) * g;
    const float vi = b2 * v[i] + (1.0f - b2) * g * g;
    m[i] = mi; v[i] = vi;
    const float w = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    p[i] = w;
    if (out16) out16[i] = __half_as_ushort(__float2half_rn(w * rs));
    if (out32) out32[i] = w * rs;
  }
}

// step counter and bias-corrected step size on the device, so that a captured CUDA graph of a whole optimisation step can
// be replayed without the host: t <- t + 1, lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)
__global__ void lr_step_kernel(long long* __restrict__ t, float lr, float b1, float b2, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const long long s = ++(*t);
    out[0] = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)s)) / (1.0 - pow((double)b1, (double)s)));
  }
}

inline int grid_for(size_t n) { const size_t g = (n + kT - 1) / kT; return (int)(g < 4096 ? (g ? g : 1) : 4096); }

}  // namespace

#define ST(s) ((cudaStream_t)(s))
#define LAUNCH_OK() KWS_CUDA_CHECK(cudaGetLastError()); return KWS_OK

extern "C" int kws_train_transpose_h16(const void* d_in, int rows, int cols, void* d_out, int ld_out, void* stream) {
  KWS_REQUIRE(d_in && d_out && rows > 0 && cols > 0 && ld_out >= rows, "kws_train_transpose_h16: bad argument");
  transpose_h16_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, ST(stream)>>>(
      static_cast<const uint16_t*>(d_in), rows, cols, ld_out, static_cast<uint16_t*>(d_out));
  LAUNCH_OK();
}
extern "C" int kws_train_swish_fwd(const void* d_z, size_t n, void* d_a, void* stream) {
  KWS_REQUIRE(d_z && d_a, "kws_train_swish_fwd: bad argument");
  if (n == 0) return KWS_OK;
  swish_fwd_kernel<<<grid_for(n), kT, 0, ST(stream)>>>(static_cast<const uint16_t*>(d_z), static_cast<uint16_t*>(d_a), n);
  LAUNCH_OK();
}
extern "C" int kws_train_gap_swish_fwd(const void* d_t, int B, int P, int C, void* d_h0, void* stream) {
  KWS_REQUIRE(d_t && d_h0 && B > 0 && P > 0 && C > 0, "kws_train_gap_swish_fwd: bad argument");
  gap_swish_fwd_kernel<<<grid_for((size_t)B * C), kT, 0, ST(stream)>>>(static_cast<const uint16_t*>(d_t), B, P, C,
                                                                       static_cast<uint16_t*>(d_h0));
  LAUNCH_OK();
}
extern "C" int kws_train_gap_swish_bwd(const void* d_dh0, const void* d_t, int B, int P, int C, void* d_dt, void* stream) {
  KWS_REQUIRE(d_dh0 && d_t && d_dt && B > 0 && P > 0 && C > 0, "kws_train_gap_swish_bwd: bad argument");
  gap_swish_bwd_kernel<<<grid_for((size_t)B * P * C), kT, 0, ST(stream)>>>(
      static_cast<const uint16_t*>(d_dh0), static_cast<const uint16_t*>(d_t), B, P, C, static_cast<uint16_t*>(d_dt));
  LAUNCH_OK();
}
static int make_geom(int H, int W, int K, int S, int pad_top, int pad_left, DwGeom* g) {
  KWS_REQUIRE(H >= 1 && W >= 1 && H * W <= 49 && (K == 3 || K == 5) && (S == 1 || S == 2), "depthwise geometry not supported");
  g->H = H; g->W = W; g->K = K; g->S = S; g->pad_top = pad_top; g->pad_left = pad_left;
  if (S == 1) { g->Ho = H; g->Wo = W; }
  else { g->Ho = (H + pad_top + K / 2 - K) / 2 + 1; g->Wo = (W + pad_left + K / 2 - K) / 2 + 1; }
  return KWS_OK;
}
extern "C" int kws_train_dw_fwd(const void* d_e, const float* d_w, const float* d_shift, int B, int C, int H, int W, int K,
                                int S, int pad_top, int pad_left, void* d_z, void* d_d, void* d_pooled, void* stream) {
  KWS_REQUIRE(d_e && d_w && d_shift && d_z && d_d && d_pooled && B > 0 && C > 0, "kws_train_dw_fwd: bad argument");
  DwGeom G;
  const int rc = make_geom(H, W, K, S, pad_top, pad_left, &G);
  if (rc != KWS_OK) return rc;
  dw_fwd_kernel<<<grid_for((size_t)B * C), kT, 0, ST(stream)>>>(static_cast<const uint16_t*>(d_e), d_w, d_shift, B, C, G,
                                                                static_cast<uint16_t*>(d_z), static_cast<uint16_t*>(d_d),
                                                                static_cast<uint16_t*>(d_pooled));
  LAUNCH_OK();
}
extern "C" size_t kws_train_dw_bwd_scratch_floats(int B, int C, int K) { return (size_t)((B + 15) / 16) * K * K * C; }
extern "C" int kws_train_dw_bwd(const void* d_dd, const void* d_dp, const void* d_z, const void* d_e, const float* d_w, int B,
                                int C, int H, int W, int K, int S, int pad_top, int pad_left, void* d_de, float* d_dw,
                                float* d_scratch, void* stream) {
  KWS_REQUIRE(d_dd && d_dp && d_z && d_e && d_w && d_de && d_dw && d_scratch && B > 0 && C > 0, "kws_train_dw_bwd: bad argument");
  DwGeom G;
  const int rc = make_geom(H, W, K, S, pad_top, pad_left, &G);
  if (rc != KWS_OK) return rc;
  const int slab = 16, slabs = (B + slab - 1) / slab;
  dw_bwd_kernel<<<dim3((C + kT - 1) / kT, slabs), kT, 0, ST(stream)>>>(
      static_cast<const uint16_t*>(d_dd), static_cast<const uint16_t*>(d_dp), static_cast<const uint16_t*>(d_z),
      static_cast<const uint16_t*>(d_e), d_w, B, C, G, slab, static_cast<uint16_t*>(d_de), d_scratch);
  KWS_CUDA_CHECK(cudaGetLastError());
  const int n = K * K * C;
  dw_bwd_reduce_kernel<<<(n + kT - 1) / kT, kT, 0, ST(stream)>>>(d_scratch, slabs, n, d_dw);
  LAUNCH_OK();
}
extern "C" int kws_train_gate_fwd(const void* d_d, const void* d_g, int B, int P, int C, void* d_out, void* stream) {
  KWS_REQUIRE(d_d && d_g && d_out && B > 0 && P > 0 && C > 0, "kws_train_gate_fwd: bad argument");
  gate_fwd_kernel<<<grid_for((size_t)B * P * C), kT, 0, ST(stream)>>>(static_cast<const uint16_t*>(d_d),
                                                                      static_cast<const uint16_t*>(d_g), B, P, C,
                                                                      static_cast<uint16_t*>(d_out));
  LAUNCH_OK();
}
extern "C" int kws_train_gate_bwd(const void* d_ddg, const void* d_d, const void* d_g, int B, int P, int C, void* d_dd,
                                  void* d_dgpre, void* stream) {
  KWS_REQUIRE(d_ddg && d_d && d_g && d_dd && d_dgpre && B > 0 && P > 0 && C > 0, "kws_train_gate_bwd: bad argument");
  gate_bwd_kernel<<<grid_for((size_t)B * C), kT, 0, ST(stream)>>>(
      static_cast<const uint16_t*>(d_ddg), static_cast<const uint16_t*>(d_d), static_cast<const uint16_t*>(d_g), B, P, C,
      static_cast<uint16_t*>(d_dd), static_cast<uint16_t*>(d_dgpre));
  LAUNCH_OK();
}
extern "C" int kws_train_act_bwd(int kind, const void* d_dy, const void* d_ref, size_t n, float scale, void* d_dz, void* stream) {
  KWS_REQUIRE(kind >= 0 && kind <= 2 && d_dy && d_ref && d_dz, "kws_train_act_bwd: bad argument");
  if (n == 0) return KWS_OK;
  act_bwd_kernel<<<grid_for(n), kT, 0, ST(stream)>>>(kind, d_dy, d_ref, n, scale, static_cast<uint16_t*>(d_dz));
  LAUNCH_OK();
}
extern "C" int kws_train_colsum(const void* d_x, int rows, int cols, float* d_out, void* stream) {
  KWS_REQUIRE(d_x && d_out && rows > 0 && cols > 0, "kws_train_colsum: bad argument");
  colsum_kernel<<<(cols + kT - 1) / kT, kT, 0, ST(stream)>>>(static_cast<const uint16_t*>(d_x), rows, cols, d_out);
  LAUNCH_OK();
}
extern "C" int kws_train_lr_step(long long* d_step, float lr, float beta1, float beta2, float* d_lr_t, void* stream) {
  KWS_REQUIRE(d_step && d_lr_t, "kws_train_lr_step: bad argument");
  lr_step_kernel<<<1, 32, 0, ST(stream)>>>(d_step, lr, beta1, beta2, d_lr_t);
  LAUNCH_OK();
}
extern "C" int kws_train_adam(float* d_param, float* d_m, float* d_v, const float* d_grad, size_t n, int cols,
                              const float* d_row_scale, const float* d_count, float loss_scale, float lr, long long step,
                              const float* d_lr_t, float beta1, float beta2, float eps, void* d_out16, float* d_out32,
                              void* stream) {
  KWS_REQUIRE(d_param && d_m && d_v && d_grad && d_count && cols > 0 && (d_lr_t || step >= 1) && loss_scale > 0.0f,
              "kws_train_adam: bad argument");
  if (n == 0) return KWS_OK;
  const double lr_t = d_lr_t ? 0.0 : (double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step));
  adam_kernel<<<grid_for(n), kT, 0, ST(stream)>>>(d_param, d_m, d_v, d_grad, n, cols, d_row_scale, d_count, 1.0f / loss_scale,
                                                  (float)lr_t, d_lr_t, beta1, beta2, eps, static_cast<uint16_t*>(d_out16), d_out32);
  LAUNCH_OK();
}
