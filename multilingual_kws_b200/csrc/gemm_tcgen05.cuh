// gemm_tcgen05.cuh — persistent, warp-specialised 16-bit (fp16 / bf16) GEMM on Blackwell tensor cores (tcgen05 + TMEM + TMA)
// with a fused pointwise epilogue.  Used for every dense contraction of the embedding tower: the 1x1
// (pointwise) convolutions of the MBConv stack, the top conv (+ global average pool) and the dense
// tower.  Replaces the cuDNN/cuBLAS calls Keras makes for
// reference multilingual_kws/train_multilingual_embedding.py:66-83 (model definition).
//
//   D[M,N] = epilogue( A[M,K] (fp16|bf16, K-major) x W[N,K]^T (same type, K-major) )      fp32 accumulate in TMEM
//   epilogue: + bias[N] (folded BatchNorm), activation, + residual[M,N], optional 4-row mean (GAP of the
//   2x2 top activation).  An epilogue thread owns one accumulator row and 16 consecutive columns at a time = 32
//   contiguous output bytes, written (and, for the skip connection, read) with one 256-bit access: full sectors,
//   no staging tile and no synchronisation between the epilogue warps.
//
// Roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane)
// + TMEM allocation, warps 2..9 = epilogue (TMEM lane quarter = warp_id % 4, two warps per quarter split the
// 16-column chunks; the bias is read through L1 as warp-uniform 16-byte loads).  Pipelines: smem ring
// full/empty mbarriers (TMA <-> MMA), 2 TMEM accumulator stages full/empty (MMA <-> epilogue), so the
// epilogue of tile i overlaps the loads + MMAs of tile i+1.  Tiles: 128 x BLOCK_N x 64, SWIZZLE_128B.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace kws {

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;           // widest k-block: 64 x 16-bit = 128 B = one SWIZZLE_128B atom row
constexpr int kGemmThreads = 320;         // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kGemmMaxStages = 12;

enum GemmAct : int { kActNone = 0, kActSwish = 1, kActRelu = 2, kActSelu = 3, kActSigmoid = 4 };

struct GemmEpilogue {
  const float* bias;                 // [N] or nullptr
  const void* residual;              // 16-bit [M, ldr] or nullptr
  void* out;                         // bf16 or fp32, row pitch ldo elements
  int ldo, ldr;
  int act;                           // GemmAct
  int out_f32;                       // 1: store float
  int gap4;                          // 1: average each aligned group of 4 rows -> row m/4
  int bf16;                          // storage/operand type of A, W, residual, 16-bit out: 1 bf16, 0 fp16
};

struct GemmShape {
  int M, N, K;
  int block_n;                       // multiple of 16, <= 256
  int stages;
  int block_k;                       // 16 / 32 / 64 elements per k-block = SWIZZLE_32B / 64B / 128B tiles: small-K layers
                                     // get narrow tiles, so many more of them fit in flight (memory-level parallelism)
  int m_tiles, n_tiles;
  int acc_stages;                    // TMEM accumulator stages per CTA: 2, or 1 for wide tiles (block_n > 128) that
                                     // should still leave room for a second CTA on the SM (256 of the 512 columns)
};

// Host: build the tensor maps, pick the tile shape (block_n = 0 -> cost model) and launch on `stream`.
// a: [M,K], w: [N,K] 16-bit K-major device pointers.
int gemm_h16(const void* a, const void* w, int M, int N, int K, int block_n, const GemmEpilogue& ep, int sm_count,
             cudaStream_t stream);

// Host: 2-D 16-bit tensor map [rows, cols] (cols contiguous), box = {64, box_rows}, 128B swizzle, zero OOB fill.
int make_tmap_h16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, int bf16,
                  uint32_t box_cols = 64);

}  // namespace kws
