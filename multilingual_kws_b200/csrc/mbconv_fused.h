// mbconv_fused.h — one launch per MBConv block (or run of blocks) of the network's tail: see mbconv_fused.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace kws {

// Device-resident description of one MBConv block (built once at kws_embed_create; the four tensor maps describe the
// block's constant 16-bit weight matrices, so a forward pass encodes only the activation map).
constexpr int kFusedBoxRows = 32;   // rows per TMA box of the 128-row weight tiles (4 boxes per tile, split over a cluster)

struct alignas(64) FusedBlockDev {
  CUtensorMap tm_exp;    // expand   [cexp][cin]     (BN scale folded), box kFusedBoxRows x 64
  CUtensorMap tm_se1;    // se_reduce [se_pad][cexp]                     box se_pad x 64
  CUtensorMap tm_se2;    // se_expand [cexp][se_pad]                     box kFusedBoxRows x 64
  CUtensorMap tm_proj;   // project  [cout][cexp]    (BN scale folded), box kFusedBoxRows x 64
  const float* b_exp;    // [cexp]  expand BN shift
  const float* w_dw;     // [k*k][cexp] depthwise kernel * BN scale
  const float* b_dw;     // [cexp]  depthwise BN shift
  const float* b_se1;    // [se_pad] (zero padded)
  const float* b_se2;    // [cexp]
  const float* b_proj;   // [cout]  project BN shift
  int cin, cexp, cout, se, se_pad;
  int geom;              // depthwise geometry (kernel, stride, map, padding): 1..6, see mbconv_fused.cu
  int pin, pout;         // pixels per clip before / after the depthwise conv
  int residual;          // 1: out += block input (stride 1, cin == cout)
  int pool_out;          // 1: "head" pseudo-block (top conv): expand + swish + mean over the pixels only, no dw/SE/project
  int reserved[6];
};

// Host mirror of the fields the launcher needs (shared-memory / TMEM sizing)
struct FusedBlockInfo {
  int cin, cexp, cout, se_pad, geom, pin, pout, residual, pool_out;
};

// geometry id for (k, stride, H, W, pad_top, pad_left), 0 if the fused kernel has no specialisation for it
int fused_geom_id(int k, int s, int h, int w, int pad_top, int pad_left);

// Largest clip group (<= 16) for which the blocks fit shared memory and TMEM; 0 = not fusable.
int fused_max_group(const FusedBlockInfo* blocks, int nblocks, int max_smem);

// x: [batch * pin0, cin0] 16-bit activations; out: [batch * pout_last, cout_last] 16-bit (pool_out: [batch, cexp]).
// d_blocks / h_blocks: nblocks consecutive blocks executed back to back inside one launch (the block output stays in
// shared memory between them).
int launch_mbconv_fused(const void* d_x, int batch, const FusedBlockDev* d_blocks, const FusedBlockInfo* h_blocks,
                        int nblocks, void* d_out, int bf16, int sm_count, int max_smem, cudaStream_t st);

}  // namespace kws
