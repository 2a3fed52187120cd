// gemm_tcgen05.cuh — persistent, warp-specialised 16-bit (fp16 / bf16) GEMM on Blackwell tensor cores (tcgen05 + TMEM + TMA)
// with a fused pointwise epilogue.  Used for every dense contraction of the embedding tower: the 1x1
// (pointwise) convolutions of the MBConv stack, the top conv (+ global average pool) and the dense
// tower.  Replaces the cuDNN/cuBLAS calls Keras makes for
// reference multilingual_kws/train_multilingual_embedding.py:66-83 (model definition).
//
//   D[M,N] = epilogue( A[M,K] (fp16|bf16, K-major) x W[N,K]^T (same type, K-major) )      fp32 accumulate in TMEM
//   epilogue: + bias[N] (folded BatchNorm), activation, + residual[M,N], optional 4-row mean (GAP of the
//   2x2 top activation), store bf16 or fp32.
//
// Roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected lane)
// + TMEM allocation, warps 2..9 = epilogue (TMEM lane quarter = warp_id % 4, two warps per quarter split the
// 16-column chunks; the tile's bias slice is staged in smem once per tile).  Pipelines: smem ring
// full/empty mbarriers (TMA <-> MMA), 2 TMEM accumulator stages full/empty (MMA <-> epilogue), so the
// epilogue of tile i overlaps the loads + MMAs of tile i+1.  Tiles: 128 x BLOCK_N x 64, SWIZZLE_128B.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace kws {

constexpr int kGemmBlockM = 128;
constexpr int kGemmBlockK = 64;           // 64 bf16 = 128 B = one SWIZZLE_128B atom row
constexpr int kGemmThreads = 320;         // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kGemmMaxStages = 8;

enum GemmAct : int { kActNone = 0, kActSwish = 1, kActRelu = 2, kActSelu = 3 };

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
  int m_tiles, n_tiles;
};

// Host: launch on `stream`.  tmap_a: [M,K] box {64,128}; tmap_b: [N,K] box {64,block_n}; both SWIZZLE_128B.
int launch_gemm_tcgen05(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmShape& shape,
                        const GemmEpilogue& ep, int sm_count, cudaStream_t stream);

// Host: 2-D bf16 K-major tensor map, box = {64, box_rows}, 128B swizzle, zero OOB fill.
int make_tmap_h16_kmajor(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, int bf16);

size_t gemm_smem_bytes(int block_n, int stages);

}  // namespace kws
