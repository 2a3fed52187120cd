// conv_kernels.h — launchers for the stem and depthwise+SE kernels (see conv_kernels.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace kws {

struct StemParams {
  int H, W, Ho, Wo, pad_top, pad_left;
  int bf16;                   // output storage: 1 bf16, 0 fp16
  float in_scale, in_shift;   // Rescaling(1/255) + Normalization: x' = x*in_scale + in_shift
  const float* w;      // [9][32]  conv kernel * BN scale   (device)
  const float* bias;   // [32]     folded BN shift
};

struct DwseParams {
  int H, W, C, Ho, Wo, K, S, pad_top, pad_left, se;
  int bf16;            // activation storage: 1 bf16, 0 fp16
  const float* w_dw;   // [K*K][C]  depthwise kernel * BN scale
  const float* b_dw;   // [C]       folded BN shift
  const float* w_se1;  // [se][C]   se_reduce kernel (transposed)
  const float* b_se1;  // [se]
  const float* w_se2;  // [se][C]   se_expand kernel
  const float* b_se2;  // [C]
  // Wide layers: the SE fully-connected pair runs as two batched tensor-core GEMMs over the whole chunk instead of per
  // clip group inside this kernel (the SE weights, up to 442 KB, are then streamed once per chunk, not once per group).
  // The kernel then stops after the pool: y receives the UN-gated activation, pooled_out [batch, C] the channel means.
  int se_external;
  uint16_t* pooled_out;
};

int launch_stem(const float* d_feats, int batch, const StemParams& P, void* d_out, int sm_count, cudaStream_t st);
// clips per CTA iteration: fits shared memory (<= 100 KB when possible) and leaves >= 2 groups per SM
int dwse_pick_group(const DwseParams& P, int max_smem, int batch, int sm_count);
int launch_dwse(const void* d_x, int batch, const DwseParams& P, void* d_y, int G, int sm_count,
                cudaStream_t st);

// y[clip, p, c] *= gates[clip, c]   (16-bit, in place; the gating pass of the external-SE path)
int launch_se_scale(void* d_y, const void* d_gates, int batch, int npix, int C, int bf16, int sm_count, cudaStream_t st);

}  // namespace kws
