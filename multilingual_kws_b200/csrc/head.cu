// head.cu — the few-shot head: Dense(in->hidden, tanh) -> Dense(hidden->classes, softmax), sparse
// cross-entropy, backward through the two Dense layers, and the Keras-Adam update (C ABI kws_head_*).
//
// Replaces, for reference multilingual_kws/embedding/transfer_learning.py:47-59,86-93, what Keras
// executes per `fit` step on the trainable part of the model (the embedding is frozen there, :43).
//
//   kws_head_grad        two launches.  (1) one warp per sample: z1 = e . W1 (W1 staged in smem), then tanh, logits,
//                        softmax, CE on the logits, dz2, dz1 inside the warp; (2) the gradient sums over the batch,
//                        16 input features per CTA, written in a fixed order (deterministic, no atomics) into a flat
//                        buffer  [dW1 | db1 | dW2 | db2 | loss_sum | correct | count].  ~12 us at batch 64 ... 512 (the
//                        first version — 32 samples per CTA, partials reduced by the last CTA — took 140 us).
//                        Gradients are SUMS over samples so that one NCCL all-reduce(sum) of the flat buffer
//                        over NVLink yields the global-batch gradient (+ the loss / accuracy scalars).
//   kws_head_apply_adam  mean = flat / count, Adam (beta1 .9, beta2 .999, eps 1e-7 outside the sqrt,
//                        bias-corrected step size), identical on every rank.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.h"

using namespace kws;

namespace {

constexpr int kHeadThreads = 256;
constexpr int kSamplesPerWarp = 2;
constexpr int kSamplesPerCta = (kHeadThreads / 32) * kSamplesPerWarp;   // 16
constexpr int kMaxHidden = 32;
constexpr int kMaxClasses = 8;
constexpr int kDwK = 16;                  // input features per CTA of the dW1 kernel
constexpr int kDwLanes = kHeadThreads / kDwK;   // sample lanes per feature

struct HeadDims {
  int in_dim, hidden, classes;
  int off_b1, off_w2, off_b2, n_params;   // flat offsets
};

__device__ __forceinline__ float warp_sum(float a) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

// ---- per-sample part: one warp per sample (two samples per warp, 16 per CTA; W1 staged in shared memory with a padded
// pitch).  z1 = e . W1 with the lanes striding over the 1024 inputs, then the whole tail (tanh, logits, softmax, CE on
// the logits, dz2, dz1) inside the warp: lane j holds hidden unit j, class sums are warp reductions.
// TRAIN: writes h [B,H], dz1 [B,H], dz2 [B,C], stat [B,2] (loss, correct) for the gradient kernel; else probs [B,C].
template <bool TRAIN, int HT>      // HT: compile-time hidden size (18 = the reference's head), 0 = run-time
__global__ void __launch_bounds__(kHeadThreads)
head_sample_kernel(const float* __restrict__ emb, const int32_t* __restrict__ labels, int B, HeadDims D,
                   const float* __restrict__ params, float* __restrict__ probs, float* __restrict__ hbuf,
                   float* __restrict__ dz1, float* __restrict__ dz2, float* __restrict__ stat) {
  extern __shared__ float s_w1[];                        // [in_dim][hidden + 1]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = HT > 0 ? HT : D.hidden, HP = H + 1, C = D.classes;
  constexpr int JN = HT > 0 ? HT : kMaxHidden;           // unrolled hidden units (no predicated-off slots when HT > 0)
#pragma unroll 8
  for (int k = warp; k < D.in_dim; k += kHeadThreads / 32)        // one W1 row (H contiguous floats) per warp and step
    if (lane < H) s_w1[k * HP + lane] = __ldg(params + (size_t)k * H + lane);
  __syncthreads();
  for (int u = 0; u < kSamplesPerWarp; ++u) {
    const int s = blockIdx.x * kSamplesPerCta + warp * kSamplesPerWarp + u;
    if (s >= B) break;                                   // warp-uniform
    float acc[JN];
#pragma unroll
    for (int j = 0; j < JN; ++j) acc[j] = 0.f;
    // eight independent global loads in flight per lane (a plain k loop would wait ~700 cycles for each of its 32 loads)
    for (int k0 = lane; k0 < D.in_dim; k0 += 256) {
      float e[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) e[q] = (k0 + 32 * q < D.in_dim) ? __ldg(emb + (size_t)s * D.in_dim + k0 + 32 * q) : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (k0 + 32 * q < D.in_dim) {
          const float* wr = s_w1 + (k0 + 32 * q) * HP;
#pragma unroll
          for (int j = 0; j < JN; ++j)
            if (j < H) acc[j] = fmaf(e[q], wr[j], acc[j]);
        }
      }
    }
    float z1 = 0.f;                                      // lane j keeps hidden unit j
#pragma unroll
    for (int j = 0; j < JN; ++j)
      if (j < H) {
        const float a = warp_sum(acc[j]);
        if (lane == j) z1 = a;
      }
    const bool unit = lane < H;
    const float hj = unit ? tanhf(z1 + __ldg(params + D.off_b1 + lane)) : 0.f;
    float z[kMaxClasses], mx = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) {
        z[c] = warp_sum(unit ? hj * __ldg(params + D.off_w2 + lane * C + c) : 0.f) + __ldg(params + D.off_b2 + c);
        if (z[c] > mx) { mx = z[c]; arg = c; }
      }
    float p[kMaxClasses], sum = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) { p[c] = expf(z[c] - mx); sum += p[c]; }
    if (!TRAIN) {
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < kMaxClasses; ++c)
          if (c < C) probs[(size_t)s * C + c] = p[c] / sum;
      }
      continue;
    }
    const int y = labels[s];
    float dh = 0.f, zy = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) {
        const float d = p[c] / sum - (c == y ? 1.f : 0.f);          // d(sum of losses) / dz2
        if (c == y) zy = z[c];
        if (unit) dh = fmaf(d, __ldg(params + D.off_w2 + lane * C + c), dh);
        if (lane == 0) dz2[(size_t)s * C + c] = d;
      }
    if (unit) {
      hbuf[(size_t)s * H + lane] = hj;
      dz1[(size_t)s * H + lane] = dh * (1.f - hj * hj);
    }
    if (lane == 0) {
      stat[(size_t)s * 2] = logf(sum) + mx - zy;                    // cross-entropy on the logits
      stat[(size_t)s * 2 + 1] = arg == y ? 1.f : 0.f;
    }
  }
}

// ---- gradient sums, written straight into the flat buffer [dW1 | db1 | dW2 | db2 | loss_sum | correct | count] in a
// fixed order (deterministic).  CTAs 0 .. in_dim/16-1: dW1[k][j] = sum_s e[s][k] dz1[s][j] for 16 inputs each (16
// sample lanes per input, combined through shared memory in lane order); the last CTA: db1, dW2, db2 and the scalars.
__global__ void __launch_bounds__(kHeadThreads)
head_grad_sums_kernel(const float* __restrict__ emb, int B, HeadDims D, const float* __restrict__ hbuf,
                      const float* __restrict__ dz1, const float* __restrict__ dz2, const float* __restrict__ stat,
                      float* __restrict__ flat) {
  __shared__ float red[kDwLanes * kDwK * (kMaxHidden + 1)];
  const int tid = threadIdx.x;
  const int H = D.hidden, HP = D.hidden + 1, C = D.classes;
  if (blockIdx.x + 1 < gridDim.x) {
    const int kk = tid % kDwK, sl = tid / kDwK;
    const int k = blockIdx.x * kDwK + kk;
    float acc[kMaxHidden];
#pragma unroll
    for (int j = 0; j < kMaxHidden; ++j) acc[j] = 0.f;
    // samples in tiles of 128: the tile's dz1 rows (9 KB) are staged in shared memory once per CTA instead of being
    // fetched from L2 by every thread (all 64 CTAs read the same rows at the same time)
    float* s_dz = red;                                   // [128][H] (the reduction below reuses the buffer afterwards)
    for (int t0 = 0; t0 < B; t0 += 128) {
      const int nt = min(128, B - t0);
      __syncthreads();
      for (int i = tid; i < nt * H; i += kHeadThreads) s_dz[i] = __ldg(dz1 + (size_t)t0 * H + i);
      __syncthreads();
      if (k < D.in_dim) {
        float e[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) e[q] = (sl + kDwLanes * q < nt) ? __ldg(emb + (size_t)(t0 + sl + kDwLanes * q) * D.in_dim + k) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (sl + kDwLanes * q < nt) {
            const float* d = s_dz + (sl + kDwLanes * q) * H;
#pragma unroll
            for (int j = 0; j < kMaxHidden; ++j)
              if (j < H) acc[j] = fmaf(e[q], d[j], acc[j]);
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMaxHidden; ++j)
      if (j < H) red[(sl * kDwK + kk) * HP + j] = acc[j];
    __syncthreads();
    for (int o = tid; o < kDwK * H; o += kHeadThreads) {
      const int kq = o / H, j = o - kq * H;
      float a = 0.f;
      for (int l = 0; l < kDwLanes; ++l) a += red[(l * kDwK + kq) * HP + j];
      if (blockIdx.x * kDwK + kq < D.in_dim) flat[(size_t)(blockIdx.x * kDwK + kq) * H + j] = a;
    }
    return;
  }
  // db1 [H], dW2 [H*C], db2 [C], loss_sum, correct: warp w sums the samples s = w (mod 8), lane l the outputs l, l + 32,
  // l + 64 (independent loads in flight); the eight partial sums per output are combined in warp order
  const int n_out = H + H * C + C + 2;
  const int warp = tid >> 5, lane = tid & 31;
  // every output is sum_s p[s * sp] * q[s * sq] (q = a constant 1 for the plain sums): the index arithmetic is done once
  // per lane, the sample loop is loads + FMAs only and is unrolled so that the loads of 8 samples are in flight
  const float one = 1.0f;
  const float *p[3], *q[3];
  int sp[3], sq[3];
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int o = lane + 32 * u;
    p[u] = &one; q[u] = &one; sp[u] = 0; sq[u] = 0;
    if (o < H) { p[u] = dz1 + o; sp[u] = H; }
    else if (o < H + H * C) { const int j = (o - H) / C, c = (o - H) - j * C; p[u] = hbuf + j; sp[u] = H; q[u] = dz2 + c; sq[u] = C; }
    else if (o < H + H * C + C) { p[u] = dz2 + (o - H - H * C); sp[u] = C; }
    else if (o < n_out) { p[u] = stat + (o - H - H * C - C); sp[u] = 2; }
  }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 8
  for (int s = warp; s < B; s += kHeadThreads / 32) {
    a0 = fmaf(p[0][(size_t)s * sp[0]], q[0][(size_t)s * sq[0]], a0);
    a1 = fmaf(p[1][(size_t)s * sp[1]], q[1][(size_t)s * sq[1]], a1);
    a2 = fmaf(p[2][(size_t)s * sp[2]], q[2][(size_t)s * sq[2]], a2);
  }
  if (lane < n_out) red[warp * n_out + lane] = a0;
  if (lane + 32 < n_out) red[warp * n_out + lane + 32] = a1;
  if (lane + 64 < n_out) red[warp * n_out + lane + 64] = a2;
  __syncthreads();
  if (tid < n_out) {
    float t = 0.f;
    for (int w = 0; w < kHeadThreads / 32; ++w) t += red[w * n_out + tid];
    int dst;
    if (tid < H) dst = D.off_b1 + tid;
    else if (tid < H + H * C) dst = D.off_w2 + (tid - H);
    else if (tid < H + H * C + C) dst = D.off_b2 + (tid - H - H * C);
    else dst = D.n_params + (tid - H - H * C - C);
    flat[dst] = t;
  }
  if (tid == 0) flat[D.n_params + 2] = (float)B;
}

__global__ void head_adam_kernel(float* __restrict__ params, float* __restrict__ m, float* __restrict__ v,
                                 const float* __restrict__ flat, int n_params, float lr_host,
                                 const float* __restrict__ lr_dev, float beta1, float beta2, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_params) return;
  const float lr_t = lr_dev ? *lr_dev : lr_host;
  const float count = flat[n_params + 2];
  const float g = flat[i] / fmaxf(count, 1.0f);
  const float mi = beta1 * m[i] + (1.0f - beta1) * g;
  const float vi = beta2 * v[i] + (1.0f - beta2) * g * g;
  m[i] = mi;
  v[i] = vi;
  params[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace

struct kws_head {
  HeadDims D;
  float *params = nullptr, *m = nullptr, *v = nullptr;
  float* scratch = nullptr;            // per-sample h, dz1, dz2, (loss, correct) of the last kws_head_grad batch
  int scratch_rows = 0;
  float beta1, beta2, eps;
  long long t = 0;
};

extern "C" int kws_head_create(kws_head_t** out, int in_dim, int hidden, int classes, const float* w1, const float* b1,
                               const float* w2, const float* b2, float beta1, float beta2, float eps) {
  KWS_REQUIRE(out && w1 && b1 && w2 && b2, "kws_head_create: NULL argument");
  *out = nullptr;
  KWS_REQUIRE(in_dim >= 1 && hidden >= 1 && hidden <= kMaxHidden && classes >= 2 && classes <= kMaxClasses,
              "kws_head_create: unsupported sizes (hidden <= %d, classes <= %d)", kMaxHidden, kMaxClasses);
  kws_head* h = new (std::nothrow) kws_head();
  KWS_REQUIRE(h != nullptr, "out of host memory");
  HeadDims& D = h->D;
  D.in_dim = in_dim; D.hidden = hidden; D.classes = classes;
  D.off_b1 = in_dim * hidden; D.off_w2 = D.off_b1 + hidden; D.off_b2 = D.off_w2 + hidden * classes;
  D.n_params = D.off_b2 + classes;
  h->beta1 = beta1; h->beta2 = beta2; h->eps = eps;
  std::vector<float> flat(D.n_params);
  memcpy(flat.data(), w1, sizeof(float) * in_dim * hidden);
  memcpy(flat.data() + D.off_b1, b1, sizeof(float) * hidden);
  memcpy(flat.data() + D.off_w2, w2, sizeof(float) * hidden * classes);
  memcpy(flat.data() + D.off_b2, b2, sizeof(float) * classes);
  cudaError_t e = cudaMalloc(&h->params, sizeof(float) * D.n_params);
  if (e == cudaSuccess) e = cudaMalloc(&h->m, sizeof(float) * D.n_params);
  if (e == cudaSuccess) e = cudaMalloc(&h->v, sizeof(float) * D.n_params);
  if (e == cudaSuccess) e = cudaMemcpy(h->params, flat.data(), sizeof(float) * D.n_params, cudaMemcpyHostToDevice);
  const int sample_smem = (int)(sizeof(float) * (size_t)in_dim * (hidden + 1));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(head_sample_kernel<false, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, sample_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(head_sample_kernel<true, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, sample_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(head_sample_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sample_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(head_sample_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sample_smem);
  if (e == cudaSuccess) e = cudaMemset(h->m, 0, sizeof(float) * D.n_params);
  if (e == cudaSuccess) e = cudaMemset(h->v, 0, sizeof(float) * D.n_params);
  if (e != cudaSuccess) {
    set_error("kws_head_create: CUDA device required (%s); there is no CPU fallback", cudaGetErrorString(e));
    cudaFree(h->params); cudaFree(h->m); cudaFree(h->v);
    delete h;
    return KWS_ERR_CUDA;
  }
  *out = h;
  return KWS_OK;
}

extern "C" void kws_head_destroy(kws_head_t* h) {
  if (!h) return;
  cudaFree(h->params); cudaFree(h->m); cudaFree(h->v); cudaFree(h->scratch);
  delete h;
}

extern "C" size_t kws_head_flat_size(const kws_head_t* h) { return h ? (size_t)h->D.n_params + 3 : 0; }
extern "C" int kws_head_num_params(const kws_head_t* h) { return h ? h->D.n_params : KWS_ERR_ARG; }
extern "C" long long kws_head_step_count(const kws_head_t* h) { return h ? h->t : KWS_ERR_ARG; }

static int head_sample_smem(const HeadDims& D) { return (int)(sizeof(float) * (size_t)D.in_dim * (D.hidden + 1)); }

extern "C" int kws_head_forward(kws_head_t* h, const float* d_emb, int B, float* d_probs, void* stream) {
  KWS_REQUIRE(h && B >= 0, "kws_head_forward: bad argument");
  if (B == 0) return KWS_OK;
  KWS_REQUIRE(d_emb && d_probs, "kws_head_forward: NULL device buffer");
  const int smem = head_sample_smem(h->D);
  const int grid = (B + kSamplesPerCta - 1) / kSamplesPerCta;
  if (h->D.hidden == 18)
    head_sample_kernel<false, 18><<<grid, kHeadThreads, smem, (cudaStream_t)stream>>>(
        d_emb, nullptr, B, h->D, h->params, d_probs, nullptr, nullptr, nullptr, nullptr);
  else
    head_sample_kernel<false, 0><<<grid, kHeadThreads, smem, (cudaStream_t)stream>>>(
        d_emb, nullptr, B, h->D, h->params, d_probs, nullptr, nullptr, nullptr, nullptr);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

extern "C" int kws_head_grad(kws_head_t* h, const float* d_emb, const int32_t* d_labels, int B, float* d_flat,
                             void* stream) {
  KWS_REQUIRE(h && B >= 1 && d_emb && d_labels && d_flat, "kws_head_grad: bad argument");
  const HeadDims& D = h->D;
  const size_t row = (size_t)2 * D.hidden + D.classes + 2;      // h, dz1, dz2, (loss, correct) per sample
  if (B > h->scratch_rows) {
    cudaFree(h->scratch);
    h->scratch = nullptr;
    h->scratch_rows = 0;
    const int rows = B < 1024 ? 1024 : B;
    KWS_CUDA_CHECK(cudaMalloc(&h->scratch, sizeof(float) * row * rows));
    h->scratch_rows = rows;
  }
  float* hbuf = h->scratch;
  float* dz1 = hbuf + (size_t)h->scratch_rows * D.hidden;
  float* dz2 = dz1 + (size_t)h->scratch_rows * D.hidden;
  float* stat = dz2 + (size_t)h->scratch_rows * D.classes;
  KWS_REQUIRE(D.hidden + D.hidden * D.classes + D.classes + 2 <= 96, "kws_head_grad: head too wide for the sums kernel");
  const int smem = head_sample_smem(D);
  const int grid = (B + kSamplesPerCta - 1) / kSamplesPerCta;
  if (D.hidden == 18)
    head_sample_kernel<true, 18><<<grid, kHeadThreads, smem, (cudaStream_t)stream>>>(
        d_emb, d_labels, B, D, h->params, nullptr, hbuf, dz1, dz2, stat);
  else
    head_sample_kernel<true, 0><<<grid, kHeadThreads, smem, (cudaStream_t)stream>>>(
        d_emb, d_labels, B, D, h->params, nullptr, hbuf, dz1, dz2, stat);
  KWS_CUDA_CHECK(cudaGetLastError());
  head_grad_sums_kernel<<<(D.in_dim + kDwK - 1) / kDwK + 1, kHeadThreads, 0, (cudaStream_t)stream>>>(
      d_emb, B, D, hbuf, dz1, dz2, stat, d_flat);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

// d(sum of losses) / d(embedding) of the batch of the last kws_head_grad call: demb[s][k] = sum_j dz1[s][j] W1[k][j].
// Needed when the layers below the head train too (transfer_learn phase 2, reference transfer_learning.py:97-112).
__global__ void __launch_bounds__(256) head_input_grad_kernel(const float* __restrict__ dz1, const float* __restrict__ params,
                                                              int B, int in_dim, int H, float* __restrict__ demb) {
  const size_t n = (size_t)B * in_dim;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t s = i / in_dim, k = i - s * in_dim;
    const float* w = params + k * H;
    const float* d = dz1 + s * H;
    float a = 0.f;
    for (int j = 0; j < H; ++j) a = fmaf(d[j], __ldg(w + j), a);
    demb[i] = a;
  }
}

extern "C" int kws_head_input_grad(kws_head_t* h, int B, float* d_demb, void* stream) {
  KWS_REQUIRE(h && d_demb && B >= 1 && B <= h->scratch_rows, "kws_head_input_grad: call kws_head_grad on the same batch first");
  const float* dz1 = h->scratch + (size_t)h->scratch_rows * h->D.hidden;
  const size_t n = (size_t)B * h->D.in_dim;
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  head_input_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dz1, h->params, B, h->D.in_dim, h->D.hidden, d_demb);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

extern "C" int kws_head_apply_adam(kws_head_t* h, const float* d_flat, float lr, void* stream) {
  KWS_REQUIRE(h && d_flat, "kws_head_apply_adam: bad argument");
  h->t += 1;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)h->beta2, (double)h->t)) / (1.0 - pow((double)h->beta1, (double)h->t));
  const int n = h->D.n_params;
  head_adam_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->params, h->m, h->v, d_flat, n, (float)lr_t, nullptr,
                                                                      h->beta1, h->beta2, h->eps);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

// Same update with the bias-corrected step size read from device memory (kws_train_lr_step): the call can be captured
// in a CUDA graph and replayed.  It does not count the step on the host (a capture executes nothing, a replay is not seen):
// the caller reports executed steps with kws_head_advance_step_count.
extern "C" int kws_head_apply_adam_dev(kws_head_t* h, const float* d_flat, const float* d_lr_t, void* stream) {
  KWS_REQUIRE(h && d_flat && d_lr_t, "kws_head_apply_adam_dev: bad argument");
  const int n = h->D.n_params;
  head_adam_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->params, h->m, h->v, d_flat, n, 0.0f, d_lr_t,
                                                                      h->beta1, h->beta2, h->eps);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}
extern "C" int kws_head_advance_step_count(kws_head_t* h, long long steps) {
  KWS_REQUIRE(h && steps >= 0, "kws_head_advance_step_count: bad argument");
  h->t += steps;
  return KWS_OK;
}

extern "C" int kws_head_get_params(const kws_head_t* h, float* host_out) {
  KWS_REQUIRE(h && host_out, "kws_head_get_params: bad argument");
  KWS_CUDA_CHECK(cudaMemcpy(host_out, h->params, sizeof(float) * h->D.n_params, cudaMemcpyDeviceToHost));
  return KWS_OK;
}

extern "C" int kws_head_reset_optimizer(kws_head_t* h) {
  KWS_REQUIRE(h != nullptr, "kws_head_reset_optimizer: NULL handle");
  KWS_CUDA_CHECK(cudaMemset(h->m, 0, sizeof(float) * h->D.n_params));
  KWS_CUDA_CHECK(cudaMemset(h->v, 0, sizeof(float) * h->D.n_params));
  h->t = 0;
  return KWS_OK;
}
