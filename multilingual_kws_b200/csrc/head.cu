// head.cu — the few-shot head: Dense(in->hidden, tanh) -> Dense(hidden->classes, softmax), sparse
// cross-entropy, backward through the two Dense layers, and the Keras-Adam update (C ABI kws_head_*).
//
// Replaces, for reference multilingual_kws/embedding/transfer_learning.py:47-59,86-93, what Keras
// executes per `fit` step on the trainable part of the model (the embedding is frozen there, :43).
//
//   kws_head_grad        one launch: forward + loss + backward for the local batch.  Each CTA owns a tile of
//                        32 samples (warp-cooperative 1024-long dot products, W1 staged in smem), writes its
//                        partial gradient, and the LAST CTA to finish sums the partials in a fixed order
//                        (deterministic) into a flat buffer  [dW1 | db1 | dW2 | db2 | loss_sum | correct | count].
//                        Gradients are SUMS over samples so that one NCCL all-reduce(sum) of the flat buffer
//                        over NVLink yields the global-batch gradient (+ the loss / accuracy scalars).
//   kws_head_apply_adam  mean = flat / count, Adam (beta1 .9, beta2 .999, eps 1e-7 outside the sqrt,
//                        bias-corrected step size), identical on every rank.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.h"

using namespace kws;

namespace {

constexpr int kHeadThreads = 256;
constexpr int kTile = 32;                 // samples per CTA
constexpr int kMaxHidden = 32;
constexpr int kMaxClasses = 8;

struct HeadDims {
  int in_dim, hidden, classes;
  int off_b1, off_w2, off_b2, n_params;   // flat offsets
};

// ---- forward only: probs[B, classes]
__global__ void __launch_bounds__(kHeadThreads)
head_forward_kernel(const float* __restrict__ emb, int B, HeadDims D, const float* __restrict__ params,
                    float* __restrict__ probs) {
  extern __shared__ float sm[];
  float* s_z1 = sm;                                  // [kTile][hidden]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * kTile;
  const float* w1 = params;
  // each warp: 4 samples x hidden units
  for (int si = warp * 4; si < warp * 4 + 4; ++si) {
    const int s = s0 + si;
    for (int j = 0; j < D.hidden; ++j) {
      float acc = 0.f;
      if (s < B)
        for (int k = lane; k < D.in_dim; k += 32) acc = fmaf(__ldg(emb + (size_t)s * D.in_dim + k), __ldg(w1 + (size_t)k * D.hidden + j), acc);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) s_z1[si * D.hidden + j] = acc;
    }
  }
  __syncthreads();
  if (tid < kTile && s0 + tid < B) {
    float h[kMaxHidden], z[kMaxClasses];
    for (int j = 0; j < D.hidden; ++j) h[j] = tanhf(s_z1[tid * D.hidden + j] + params[D.off_b1 + j]);
    float mx = -INFINITY;
    for (int c = 0; c < D.classes; ++c) {
      float a = params[D.off_b2 + c];
      for (int j = 0; j < D.hidden; ++j) a = fmaf(h[j], params[D.off_w2 + j * D.classes + c], a);
      z[c] = a;
      mx = fmaxf(mx, a);
    }
    float sum = 0.f;
    for (int c = 0; c < D.classes; ++c) { z[c] = expf(z[c] - mx); sum += z[c]; }
    for (int c = 0; c < D.classes; ++c) probs[(size_t)(s0 + tid) * D.classes + c] = z[c] / sum;
  }
}

// ---- forward + loss + backward; partial gradients per CTA, last CTA reduces
__global__ void __launch_bounds__(kHeadThreads)
head_grad_kernel(const float* __restrict__ emb, const int32_t* __restrict__ labels, int B, HeadDims D,
                 const float* __restrict__ params, float* __restrict__ partials, unsigned int* __restrict__ counter,
                 float* __restrict__ flat) {
  extern __shared__ float sm[];
  float* s_w1 = sm;                                      // [in_dim][hidden+1] (padded: conflict-free)
  float* s_z1 = s_w1 + (size_t)D.in_dim * (D.hidden + 1);  // [kTile][hidden]  z1, later dz1
  float* s_h = s_z1 + kTile * D.hidden;                  // [kTile][hidden]
  float* s_dz2 = s_h + kTile * D.hidden;                 // [kTile][classes]
  float* s_stat = s_dz2 + kTile * D.classes;             // [kTile][2] loss, correct
  __shared__ bool is_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * kTile;
  const int H = D.hidden, HP = D.hidden + 1;
  const int flat_n = D.n_params + 3;
  float* mine = partials + (size_t)blockIdx.x * flat_n;

  for (int i = tid; i < D.in_dim * H; i += kHeadThreads) s_w1[(i / H) * HP + (i % H)] = __ldg(params + i);
  __syncthreads();

  // z1: each warp owns 4 samples; lanes stride over k, accumulate all hidden units, then warp-reduce
  for (int si = warp * 4; si < warp * 4 + 4; ++si) {
    const int s = s0 + si;
    float acc[kMaxHidden];
#pragma unroll
    for (int j = 0; j < kMaxHidden; ++j) acc[j] = 0.f;
    if (s < B) {
      for (int k = lane; k < D.in_dim; k += 32) {
        const float e = __ldg(emb + (size_t)s * D.in_dim + k);
        const float* wr = s_w1 + k * HP;
#pragma unroll
        for (int j = 0; j < kMaxHidden; ++j)
          if (j < H) acc[j] = fmaf(e, wr[j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxHidden; ++j) {
      if (j < H) {
        float a = acc[j];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) s_z1[si * H + j] = a;
      }
    }
  }
  __syncthreads();

  // per-sample tail: tanh, logits, softmax, CE on logits, dz2, dz1
  if (tid < kTile) {
    const int s = s0 + tid;
    float loss = 0.f, correct = 0.f;
    if (s < B) {
      float h[kMaxHidden], z[kMaxClasses], p[kMaxClasses];
      for (int j = 0; j < H; ++j) { h[j] = tanhf(s_z1[tid * H + j] + params[D.off_b1 + j]); s_h[tid * H + j] = h[j]; }
      float mx = -INFINITY;
      int arg = 0;
      for (int c = 0; c < D.classes; ++c) {
        float a = params[D.off_b2 + c];
        for (int j = 0; j < H; ++j) a = fmaf(h[j], params[D.off_w2 + j * D.classes + c], a);
        z[c] = a;
        if (a > mx) { mx = a; arg = c; }
      }
      float sum = 0.f;
      for (int c = 0; c < D.classes; ++c) { p[c] = expf(z[c] - mx); sum += p[c]; }
      const int y = labels[s];
      loss = logf(sum) + mx - z[y];
      correct = arg == y ? 1.f : 0.f;
      for (int c = 0; c < D.classes; ++c) {
        const float d = p[c] / sum - (c == y ? 1.f : 0.f);      // d(sum of losses)/dz2
        s_dz2[tid * D.classes + c] = d;
      }
      for (int j = 0; j < H; ++j) {
        float dh = 0.f;
        for (int c = 0; c < D.classes; ++c) dh = fmaf(s_dz2[tid * D.classes + c], params[D.off_w2 + j * D.classes + c], dh);
        s_z1[tid * H + j] = dh * (1.f - h[j] * h[j]);             // dz1
      }
    } else {
      for (int j = 0; j < H; ++j) { s_h[tid * H + j] = 0.f; s_z1[tid * H + j] = 0.f; }
      for (int c = 0; c < D.classes; ++c) s_dz2[tid * D.classes + c] = 0.f;
    }
    s_stat[tid * 2] = loss;
    s_stat[tid * 2 + 1] = correct;
  }
  __syncthreads();

  // partial dW1[k][j] = sum_s e[s][k] * dz1[s][j]
  const int ns = min(kTile, B - s0);
  for (int k = tid; k < D.in_dim; k += kHeadThreads) {
    float acc[kMaxHidden];
#pragma unroll
    for (int j = 0; j < kMaxHidden; ++j) acc[j] = 0.f;
    for (int si = 0; si < ns; ++si) {
      const float e = __ldg(emb + (size_t)(s0 + si) * D.in_dim + k);
#pragma unroll
      for (int j = 0; j < kMaxHidden; ++j)
        if (j < H) acc[j] = fmaf(e, s_z1[si * H + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < kMaxHidden; ++j)
      if (j < H) mine[(size_t)k * H + j] = acc[j];
  }
  // db1, dW2, db2, stats
  for (int i = tid; i < H + H * D.classes + D.classes + 3; i += kHeadThreads) {
    float a = 0.f;
    int dst;
    if (i < H) {
      for (int si = 0; si < kTile; ++si) a += s_z1[si * H + i];
      dst = D.off_b1 + i;
    } else if (i < H + H * D.classes) {
      const int j = (i - H) / D.classes, c = (i - H) % D.classes;
      for (int si = 0; si < kTile; ++si) a = fmaf(s_h[si * H + j], s_dz2[si * D.classes + c], a);
      dst = D.off_w2 + j * D.classes + c;
    } else if (i < H + H * D.classes + D.classes) {
      const int c = i - H - H * D.classes;
      for (int si = 0; si < kTile; ++si) a += s_dz2[si * D.classes + c];
      dst = D.off_b2 + c;
    } else {
      const int w = i - (H + H * D.classes + D.classes);
      if (w == 2) a = (float)ns;
      else for (int si = 0; si < kTile; ++si) a += s_stat[si * 2 + w];
      dst = D.n_params + w;
    }
    mine[dst] = a;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int i = tid; i < flat_n; i += kHeadThreads) {
      float a = 0.f;
      for (unsigned int b = 0; b < gridDim.x; ++b) a += __ldcg(partials + (size_t)b * flat_n + i);
      flat[i] = a;
    }
    if (tid == 0) *counter = 0;
  }
}

__global__ void head_adam_kernel(float* __restrict__ params, float* __restrict__ m, float* __restrict__ v,
                                 const float* __restrict__ flat, int n_params, float lr_t, float beta1, float beta2,
                                 float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_params) return;
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
  float *params = nullptr, *m = nullptr, *v = nullptr, *partials = nullptr;
  unsigned int* counter = nullptr;
  int partial_ctas = 0;
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
  if (e == cudaSuccess) e = cudaMalloc(&h->counter, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemcpy(h->params, flat.data(), sizeof(float) * D.n_params, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->m, 0, sizeof(float) * D.n_params);
  if (e == cudaSuccess) e = cudaMemset(h->v, 0, sizeof(float) * D.n_params);
  if (e == cudaSuccess) e = cudaMemset(h->counter, 0, sizeof(unsigned int));
  if (e != cudaSuccess) {
    set_error("kws_head_create: CUDA device required (%s); there is no CPU fallback", cudaGetErrorString(e));
    cudaFree(h->params); cudaFree(h->m); cudaFree(h->v); cudaFree(h->counter);
    delete h;
    return KWS_ERR_CUDA;
  }
  *out = h;
  return KWS_OK;
}

extern "C" void kws_head_destroy(kws_head_t* h) {
  if (!h) return;
  cudaFree(h->params); cudaFree(h->m); cudaFree(h->v); cudaFree(h->counter); cudaFree(h->partials);
  delete h;
}

extern "C" size_t kws_head_flat_size(const kws_head_t* h) { return h ? (size_t)h->D.n_params + 3 : 0; }
extern "C" int kws_head_num_params(const kws_head_t* h) { return h ? h->D.n_params : KWS_ERR_ARG; }
extern "C" long long kws_head_step_count(const kws_head_t* h) { return h ? h->t : KWS_ERR_ARG; }

extern "C" int kws_head_forward(kws_head_t* h, const float* d_emb, int B, float* d_probs, void* stream) {
  KWS_REQUIRE(h && B >= 0, "kws_head_forward: bad argument");
  if (B == 0) return KWS_OK;
  KWS_REQUIRE(d_emb && d_probs, "kws_head_forward: NULL device buffer");
  const int grid = (B + kTile - 1) / kTile;
  head_forward_kernel<<<grid, kHeadThreads, sizeof(float) * kTile * h->D.hidden, (cudaStream_t)stream>>>(
      d_emb, B, h->D, h->params, d_probs);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

extern "C" int kws_head_grad(kws_head_t* h, const float* d_emb, const int32_t* d_labels, int B, float* d_flat,
                             void* stream) {
  KWS_REQUIRE(h && B >= 1 && d_emb && d_labels && d_flat, "kws_head_grad: bad argument");
  const HeadDims& D = h->D;
  const int grid = (B + kTile - 1) / kTile;
  const size_t flat_n = (size_t)D.n_params + 3;
  if (grid > h->partial_ctas) {
    cudaFree(h->partials);
    h->partials = nullptr;
    KWS_CUDA_CHECK(cudaMalloc(&h->partials, sizeof(float) * flat_n * grid));
    h->partial_ctas = grid;
  }
  const size_t smem = sizeof(float) * ((size_t)D.in_dim * (D.hidden + 1) + 2 * kTile * D.hidden + kTile * D.classes + kTile * 2);
  KWS_CUDA_CHECK(cudaFuncSetAttribute(head_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_grad_kernel<<<grid, kHeadThreads, smem, (cudaStream_t)stream>>>(d_emb, d_labels, B, D, h->params, h->partials,
                                                                     h->counter, d_flat);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}

extern "C" int kws_head_apply_adam(kws_head_t* h, const float* d_flat, float lr, void* stream) {
  KWS_REQUIRE(h && d_flat, "kws_head_apply_adam: bad argument");
  h->t += 1;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)h->beta2, (double)h->t)) / (1.0 - pow((double)h->beta1, (double)h->t));
  const int n = h->D.n_params;
  head_adam_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->params, h->m, h->v, d_flat, n, (float)lr_t,
                                                                      h->beta1, h->beta2, h->eps);
  KWS_CUDA_CHECK(cudaGetLastError());
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
