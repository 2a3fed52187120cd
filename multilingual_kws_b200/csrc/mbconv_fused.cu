// mbconv_fused.cu — a whole MBConv block of the network's tail (blocks 4a ... 7a, maps of <= 7x5 pixels) as ONE kernel:
//   expand 1x1 (tcgen05) -> BN -> swish -> depthwise kxk -> BN -> swish -> squeeze-excite (both FCs on tcgen05) ->
//   gate -> project 1x1 (tcgen05) -> BN (-> + skip), optionally several blocks back to back and the top conv + global
//   average pool at the end, with the expanded tensor never leaving the SM.
// Replaces, for these blocks, the six launches per block of the layer-by-layer schedule (expand GEMM, depthwise + pool,
// two SE GEMMs, gating pass, project GEMM) behind Keras' EfficientNetB0 block() (model defined at reference
// multilingual_kws/train_multilingual_embedding.py:66-83, cut at dense_2: embedding/transfer_learning.py:38-43;
// block structure: SURVEY.md App. B.1).
//
// Orientation.  Every contraction is computed TRANSPOSED: D^T[channels, rows] = W[channels, K] . X[rows, K]^T with the
// 16-bit weight matrix as the MMA's A operand (M = 128 output channels per instruction, streamed through a TMA ring) and
// the activations of G clips (rows = G * pixels, resident in shared memory) as the B operand (N = rows).  The
// accumulator therefore has one CHANNEL per TMEM lane and one (clip, pixel) per column: a thread that owns a lane sees
// every pixel of its channel for its clips in registers, so the depthwise convolution (tiny maps: the whole image and
// only the taps that overlap it), its BN + swish, the SE average pool, the gating and the skip connection are all
// thread-local — no shuffles, no shared-memory halo exchange.  What a thread produces is written 16 bits wide into the
// K-major SWIZZLE_128B layout the next MMA's B operand wants ([row][channel]).
//
// Per block and CTA (one CTA = G clips, one CTA per SM, 512 TMEM columns):
//   1. expand       for each chunk of 128 expanded channels: MMAs into one of two TMEM stages; the 16 compute warps
//                   (lane quarter = warp % 4; the four warps of a quarter split the clips) read their channel's pixels,
//                   apply bias + swish, the depthwise taps, bias + swish, write D (un-gated, 16-bit) to shared memory
//                   and the channel mean to the `pooled` operand.
//   2. SE reduce    s[G, se] = pooled . W1 (the one contraction with the clips on the accumulator's lanes: M = 128 rows
//                   of which 16 are real, N = se_pad, so several k-blocks of W1^T share one ring stage); bias + swish -> `s`.
//   3. SE expand    per chunk: gate^T[128, G] = W2^T . s^T; a thread applies sigmoid and scales exactly the D elements it
//                   wrote in step 1 (no cross-thread hazard).
//   4. project      out^T[cout, rows] = Wp . D^T; bias (+ residual read from the block's input tile in shared memory);
//                   the result becomes the next block's input tile (in place) or goes to global memory.
// Roles: warp 0 = TMA producer (weights only: it runs ahead through the ring regardless of the phase), warp 1 = MMA
// issuer + TMEM allocation, warps 2..17 = compute.  Hand-offs are mbarriers: TMA -> MMA (ring full / empty), MMA ->
// compute (tcgen05.commit), compute -> MMA (arrive after fence.proxy.async, since the B operands are written through the
// generic proxy and read by the tensor core through the async proxy).
#include "mbconv_fused.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.h"
#include "gemm_tcgen05.cuh"
#include "ptx.cuh"

namespace kws {

namespace {

constexpr int kFusedComputeWarps = 16;
constexpr int kFusedThreads = (4 + kFusedComputeWarps) * 32;   // 640: warp group 0 = TMA / MMA / 2 idle, groups 1..4 = compute
// Register split (setmaxnreg works per warp group, inside the CTA's own pool): the kernel is launched with
// 65536 / 640 -> 96 registers per thread = 61440; group 0 keeps 32, the four compute groups grow to
// (61440 - 128 * 32) / 512 = 112.
constexpr int kRegsControl = 32;
constexpr int kRegsCompute = 112;
constexpr uint32_t kStageBytes = 128 * 64 * 2;                 // one A tile: 128 channels x 64 k (SWIZZLE_128B)
constexpr int kMaxStages = 8;
constexpr int kFcN = 16;                                        // max clips per CTA; N of the excite MMAs
constexpr int kFc1Cols = 64;                                    // TMEM columns of the squeeze result (se_pad <= 64)
constexpr int kBoxRows = kFusedBoxRows;                        // rows per TMA box of a 128-row weight tile: the CTAs of a
                                                                // cluster (1, 2 or 4) each fetch 4 / C boxes and multicast them

struct FusedSmem {
  uint32_t x_off, d_off, pool_off, s_off, ring_off, bar_off, total;
  int stages;
  int fc_rows;             // rows per k-block of the pooled operand: 8 when G <= 8, else 16
};

// barrier slots (uint64_t each) at bar_off
enum {
  kBarFull = 0,                       // [kMaxStages] ring
  kBarEmpty = kMaxStages,             // [kMaxStages]
  kBarXFull = 2 * kMaxStages,         // block 0 input tile landed (TMA)
  kBarXReady,                         // next block's input tile written by the compute warps
  kBarExpFull,                        // [2] expand accumulator stage complete (tcgen05.commit)
  kBarExpEmpty = kBarExpFull + 2,     // [2] stage drained by the compute warps
  kBarPoolReady = kBarExpEmpty + 2,   // D + pooled operand written
  kBarFc1Full,
  kBarSReady,
  kBarFc2Full,
  kBarDReady,                         // D gated
  kBarProjFull,
  kBarCount
};

__host__ __device__ inline int round_up_i(int x, int a) { return (x + a - 1) / a * a; }

// mbarrier wait with a watchdog: a protocol bug traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();      // ~seconds: no legitimate wait in this kernel is longer than microseconds
  }
}

__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, uint32_t& r0) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// N consecutive columns (N a multiple of 4) of this thread's lane, register indices compile-time
template <int N, int DONE = 0>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]) {
  static_assert(N % 4 == 0, "column count must be a multiple of 4");
  if constexpr (N - DONE >= 16) {
    tmem_ld_x16(taddr + DONE, &r[DONE]);
    tmem_ld_cols<N, DONE + 16>(taddr, r);
  } else if constexpr (N - DONE >= 8) {
    tmem_ld_x8(taddr + DONE, &r[DONE]);
    tmem_ld_cols<N, DONE + 8>(taddr, r);
  } else if constexpr (N - DONE >= 4) {
    tmem_ld_x4(taddr + DONE, &r[DONE]);
    tmem_ld_cols<N, DONE + 4>(taddr, r);
  }
}

__device__ __forceinline__ uint16_t to_h16(float v, int bf) { return (uint16_t)(ptx::pack_h2(v, 0.0f, bf) & 0xFFFFu); }
__device__ __forceinline__ float from_h16(uint16_t h, int bf) { return ptx::unpack_h2((uint32_t)h, bf).x; }

// depthwise geometries of EfficientNet-B0's tail at 49x40 input (same numbering as dwse_kernel's tiny-map variants)
template <int GEOM> struct Geom;
template <> struct Geom<1> { static constexpr int K = 3, S = 2, H = 7, W = 5, PT = 1, PL = 1, HO = 4, WO = 3; };   // 4a
template <> struct Geom<2> { static constexpr int K = 3, S = 1, H = 4, W = 3, PT = 1, PL = 1, HO = 4, WO = 3; };   // 4b 4c
template <> struct Geom<3> { static constexpr int K = 5, S = 1, H = 4, W = 3, PT = 2, PL = 2, HO = 4, WO = 3; };   // 5a 5b 5c
template <> struct Geom<4> { static constexpr int K = 5, S = 2, H = 4, W = 3, PT = 1, PL = 2, HO = 2, WO = 2; };   // 6a
template <> struct Geom<5> { static constexpr int K = 5, S = 1, H = 2, W = 2, PT = 2, PL = 2, HO = 2, WO = 2; };   // 6b 6c 6d
template <> struct Geom<6> { static constexpr int K = 3, S = 1, H = 2, W = 2, PT = 1, PL = 1, HO = 2, WO = 2; };   // 7a
// pseudo-geometry of the top conv: 2x2 map, no depthwise (pool_out)
template <> struct Geom<7> { static constexpr int K = 1, S = 1, H = 2, W = 2, PT = 0, PL = 0, HO = 2, WO = 2; };

struct BlockDims {
  int pin, pout, npad_in, npad_out, rpad_out, nchunk, nout, kb_in, kb_exp;
  uint32_t col_exp1, col_fc1, col_fc2, col_proj;   // expand stage 0 starts at column 0; col_exp1 = 0: single stage
};
__host__ __device__ inline BlockDims block_dims(int cin, int cexp, int cout, int pin, int pout, int pool_out, int G) {
  BlockDims d;
  d.pin = pin; d.pout = pout;
  d.npad_in = round_up_i(G * pin, 16);
  d.npad_out = round_up_i(G * pout, 16);
  d.rpad_out = round_up_i(G * pout, 8);
  d.nchunk = (cexp + 127) / 128;
  d.nout = pool_out ? 0 : (cout + 127) / 128;
  d.kb_in = (cin + 63) / 64;
  d.kb_exp = (cexp + 63) / 64;
  const int two = (2 * d.npad_in + (pool_out ? 0 : kFc1Cols + kFcN * d.nchunk + d.nout * d.npad_out)) <= 512;
  d.col_exp1 = two ? (uint32_t)d.npad_in : 0u;
  d.col_fc1 = (uint32_t)((two ? 2 : 1) * d.npad_in);
  d.col_fc2 = d.col_fc1 + kFc1Cols;
  d.col_proj = d.col_fc2 + (uint32_t)(kFcN * d.nchunk);
  return d;
}
__host__ __device__ inline int block_tmem_cols(const BlockDims& d, int pool_out) {
  return pool_out ? (int)d.col_fc1 : (int)d.col_proj + d.nout * d.npad_out;
}

struct KernelArgs {
  const FusedBlockDev* blocks;
  int nblocks, batch, G, bf16;
  FusedSmem L;
  uint16_t* out;
  long long* dbg;          // optional phase timestamps of CTA 0 (tools only; nullptr in production)
};

// shared-memory accesses by 32-bit shared-space address (the dynamic smem pointer is re-aligned by hand, which makes
// the compiler fall back to generic LD / ST with 64-bit address arithmetic otherwise)
__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint16_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// element-wise product of two packed 16-bit pairs
__device__ __forceinline__ uint32_t mul_h2(uint32_t a, uint32_t b, int bf) {
  uint32_t r;
  if (bf) asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  else asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// named barrier over the compute warps only (barrier 0 is __syncthreads)
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kFusedComputeWarps * 32) : "memory"); }

// offset of (row, channel-in-k-block cc): row * 128 + swizzled 16-byte piece + position inside the piece
__device__ __forceinline__ uint32_t sw_off(int row, int cchunk, int clow2) {
  return (uint32_t)(row * 128 + (((cchunk ^ row) & 7) << 4) + clow2);
}

// ---- cluster helpers: weight tiles are fetched once per cluster and multicast into every CTA's ring
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* desc, uint64_t* bar, void* dst_smem, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(ptx::smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   ptx::smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// ------------------------------------------------------------------------------------------------ compute warps
struct ComputeCtx {
  uint32_t smem;           // shared-space address of the (1024-byte aligned) dynamic shared memory
  uint64_t* bars;
  uint32_t tmem_base;      // with this warp's lane quarter in the lane field
  int sub, chl;            // which of the quarter's 4 warps; channel within a 128-chunk (lane quarter * 32 + lane)
  int g0, gn, G, bf16, lane, cw;
  uint32_t exp_use0, exp_use1;
};

template <int KK>
struct DwWeights {
  float w[KK];
  float be, bd;
};

// One block, compile-time depthwise geometry.  `last` blocks store to global memory, others into the X tile.
template <int GEOM>
__device__ __forceinline__ void compute_block(ComputeCtx& cx, const KernelArgs& a, const FusedBlockDev* __restrict__ blk,
                                              int bi, bool first, bool last, int next_npad_in) {
  using Gm = Geom<GEOM>;
  constexpr int K = Gm::K, S = Gm::S, H = Gm::H, W = Gm::W, PT = Gm::PT, PL = Gm::PL, HO = Gm::HO, WO = Gm::WO;
  constexpr int PIN = H * W, POUT = HO * WO;
  constexpr int PIN4 = (PIN + 3) & ~3, POUT4 = (POUT + 3) & ~3;
  constexpr int NB = 1;     // clips per iteration (2: both TMEM loads share one wait — measured slower: register spills)
  const int cexp = blk->cexp, cout = blk->cout;
  const BlockDims d = block_dims(blk->cin, cexp, cout, PIN, POUT, blk->pool_out, cx.G);
  const uint32_t par = (uint32_t)bi & 1u;
  const uint32_t s_d = cx.smem + a.L.d_off, s_pool = cx.smem + a.L.pool_off, s_s = cx.smem + a.L.s_off,
                 s_x = cx.smem + a.L.x_off;
  const uint32_t pool_kb = (uint32_t)a.L.fc_rows * 128u;
  const int bf = cx.bf16;
  uint64_t* const bars = cx.bars;
  const uint32_t tmem_base = cx.tmem_base;
  uint32_t exp_use0 = cx.exp_use0, exp_use1 = cx.exp_use1;
  long long* const dbg = (a.dbg && blockIdx.x == 0 && cx.cw == 0 && cx.lane == 0) ? a.dbg + bi * 8 : nullptr;
  if (dbg) dbg[0] = clock64();

  // ---- 1. expand epilogue + depthwise + pool, chunk by chunk
  auto load_w = [&](int j, DwWeights<K * K>& Wt) {
    const int c = j * 128 + cx.chl;
    Wt.be = 0.0f; Wt.bd = 0.0f;
    if (c < cexp) {
      Wt.be = __ldg(blk->b_exp + c);
      if (GEOM != 7) {
        Wt.bd = __ldg(blk->b_dw + c);
#pragma unroll
        for (int t = 0; t < K * K; ++t) Wt.w[t] = __ldg(blk->w_dw + (size_t)t * cexp + c);
      }
    }
  };
  auto process = [&](int j, const DwWeights<K * K>& Wt) {
    const int c = j * 128 + cx.chl;
    const bool valid = c < cexp;
    const int st = (d.col_exp1 != 0) ? (j & 1) : 0;
    const uint32_t col_st = st ? d.col_exp1 : 0u;
    bar_wait(bars + kBarExpFull + st, (st ? exp_use1 : exp_use0) & 1u);
    if (st) ++exp_use1; else ++exp_use0;
    ptx::tc_fence_after();
    const int kb = c >> 6, cchunk = (c & 63) >> 3, clow2 = (c & 7) * 2;
    const uint32_t dk = s_d + (uint32_t)(kb * d.rpad_out * 128);
    // one clip of this thread's channel: activation of the expansion, depthwise taps, activation, D + pooled mean
    auto clip = [&](const uint32_t (&raw)[PIN4], int g) {
      float x[PIN];
#pragma unroll
      for (int q = 0; q < PIN; ++q) x[q] = ptx::swish_f(__uint_as_float(raw[q]) + Wt.be);
      if constexpr (GEOM == 7) {
        // top conv: swish + global average pool, straight to global memory [clip][cexp]
        const float m = 0.25f * ((x[0] + x[1]) + (x[2] + x[3]));
        a.out[(size_t)(cx.g0 + g) * cexp + c] = to_h16(m, bf);
      } else {
        float sum = 0.0f;
#pragma unroll
        for (int ho = 0; ho < HO; ++ho)
#pragma unroll
          for (int wo = 0; wo < WO; ++wo) {
            float acc = Wt.bd;
#pragma unroll
            for (int kh = 0; kh < K; ++kh)
#pragma unroll
              for (int kw = 0; kw < K; ++kw) {
                const int r = ho * S + kh - PT, cl = wo * S + kw - PL;       // compile-time after unrolling
                if (r >= 0 && r < H && cl >= 0 && cl < W) acc = fmaf(x[r * W + cl], Wt.w[kh * K + kw], acc);
              }
            acc = ptx::swish_f(acc);
            sum += acc;
            sts_u16(dk + sw_off(g * POUT + ho * WO + wo, cchunk, clow2), to_h16(acc, bf));
          }
        sts_u16(s_pool + (uint32_t)kb * pool_kb + sw_off(g, cchunk, clow2), to_h16(sum * (1.0f / (float)POUT), bf));
      }
    };
    for (int g = cx.sub; g < cx.gn; g += 4 * NB) {
      uint32_t raw0[PIN4], raw1[NB == 2 ? PIN4 : 4];
      const bool two = NB == 2 && g + 4 < cx.gn;        // warp-uniform
      tmem_ld_cols<PIN4>(tmem_base + col_st + (uint32_t)(g * PIN), raw0);
      if constexpr (NB == 2) {
        if (two) tmem_ld_cols<PIN4>(tmem_base + col_st + (uint32_t)((g + 4) * PIN), raw1);
      }
      ptx::tmem_ld_wait();
      if (valid) {
        clip(raw0, g);
        if constexpr (NB == 2) {
          if (two) clip(raw1, g + 4);
        }
      }
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (cx.lane == 0) ptx::mbar_arrive(bars + kBarExpEmpty + st);
  };
  {
    // one register set: the constants of chunk j + 1 are requested as soon as chunk j's last clip is done, so their L2
    // round trip overlaps the end-of-chunk hand-off, the next accumulator wait, its TMEM load and the first swishes
    DwWeights<K * K> wt;
    load_w(0, wt);
    for (int j = 0; j < d.nchunk; ++j) {
      process(j, wt);
      if (j + 1 < d.nchunk) load_w(j + 1, wt);
    }
  }
  cx.exp_use0 = exp_use0; cx.exp_use1 = exp_use1;
  if constexpr (GEOM == 7) return;

  ptx::fence_proxy_async();                             // D and pooled were written through the generic proxy
  __syncwarp();
  if (cx.lane == 0) ptx::mbar_arrive(bars + kBarPoolReady);
  if (dbg) dbg[1] = clock64();

  // ---- 2. squeeze: s[g][j] = swish(pooled . W1 + b1)
  bar_wait(bars + kBarFc1Full, par);
  ptx::tc_fence_after();
  if (dbg) dbg[2] = clock64();
  if (cx.cw == 0) {
    // accumulator: lane = clip (16 rows), column = squeeze unit.  Lane g writes row g of the `s` operand.
    const int g = cx.lane & 15;
    for (int j0 = 0; j0 < blk->se_pad; j0 += 16) {
      uint32_t raw[16];
      tmem_ld_x16(tmem_base + d.col_fc1 + (uint32_t)j0, raw);     // cw 0 is a quarter-0 warp: lanes 0..31 of TMEM
      ptx::tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + 2 * u;
        const float v0 = j < blk->se ? ptx::swish_f(__uint_as_float(raw[2 * u]) + __ldg(blk->b_se1 + j)) : 0.0f;
        const float v1 = j + 1 < blk->se ? ptx::swish_f(__uint_as_float(raw[2 * u + 1]) + __ldg(blk->b_se1 + j + 1)) : 0.0f;
        pk[u] = ptx::pack_h2(v0, v1, bf);
      }
      if (cx.lane < 16) {
        sts_v4(s_s + sw_off(g, j0 >> 3, 0), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        sts_v4(s_s + sw_off(g, (j0 >> 3) + 1, 0), make_uint4(pk[4], pk[5], pk[6], pk[7]));
      }
    }
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncwarp();
  if (cx.lane == 0) ptx::mbar_arrive(bars + kBarSReady);

  // ---- 3. excite: gate[g][c] = sigmoid(W2^T s + b2) goes to a 16-bit table (the pooled operand's memory: the squeeze
  // MMAs have completed), then all compute threads scale D in place with 16-byte accesses (8 channels of one row)
  bar_wait(bars + kBarFc2Full, par);
  ptx::tc_fence_after();
  if (dbg) dbg[3] = clock64();
  const uint32_t s_gate = s_pool;                       // [G][cexp] 16-bit, plain row-major
  for (int j = 0; j < d.nchunk; ++j) {
    const int c = j * 128 + cx.chl;
    const bool valid = c < cexp;
    const float b2 = valid ? __ldg(blk->b_se2 + c) : 0.0f;
    uint32_t raw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)                          // this warp's clips sub, sub + 4, ...: all loads, one wait
      if (cx.sub + 4 * i < cx.gn) tmem_ld_x1(tmem_base + d.col_fc2 + (uint32_t)(j * kFcN + cx.sub + 4 * i), raw[i]);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (valid && cx.sub + 4 * i < cx.gn)
        sts_u16(s_gate + (uint32_t)((cx.sub + 4 * i) * cexp + c) * 2u, to_h16(ptx::sigmoid_f(__uint_as_float(raw[i]) + b2), bf));
  }
  ptx::tc_fence_before();
  compute_sync();
  {
    const int nvr = cexp >> 3;                          // 16-byte pieces per row
    const int rows = cx.gn * POUT;
    for (int row = cx.cw; row < rows; row += kFusedComputeWarps) {
      const int g = row / POUT;
      for (int v = cx.lane; v < nvr; v += 32) {
        const uint32_t addr = s_d + (uint32_t)((v >> 3) * d.rpad_out * 128) + sw_off(row, v & 7, 0);
        uint4 dv = lds_v4(addr);
        const uint4 gv = lds_v4(s_gate + (uint32_t)(g * cexp + v * 8) * 2u);
        dv.x = mul_h2(dv.x, gv.x, bf); dv.y = mul_h2(dv.y, gv.y, bf);
        dv.z = mul_h2(dv.z, gv.z, bf); dv.w = mul_h2(dv.w, gv.w, bf);
        sts_v4(addr, dv);
      }
    }
  }
  ptx::fence_proxy_async();
  __syncwarp();
  if (cx.lane == 0) ptx::mbar_arrive(bars + kBarDReady);
  if (dbg) dbg[4] = clock64();

  // ---- 4. project epilogue: bias (+ skip) -> next block's input tile, or global memory
  bar_wait(bars + kBarProjFull, par);
  ptx::tc_fence_after();
  if (dbg) dbg[5] = clock64();
  if (first && blk->residual) bar_wait(bars + kBarXFull, 0);     // the TMA-written input tile is visible to this thread
  for (int o = 0; o < d.nout; ++o) {
    const int c = o * 128 + cx.chl;
    const bool valid = c < cout;
    const float bp = valid ? __ldg(blk->b_proj + c) : 0.0f;
    const int kb = c >> 6, cchunk = (c & 63) >> 3, clow2 = (c & 7) * 2;
    const uint32_t xin = s_x + (uint32_t)(kb * d.npad_in * 128), xout = s_x + (uint32_t)(kb * next_npad_in * 128);
    auto clip_out = [&](const uint32_t (&raw)[POUT4], int g) {
#pragma unroll
      for (int p = 0; p < POUT; ++p) {
        const int row = g * POUT + p;
        float v = __uint_as_float(raw[p]) + bp;
        if (blk->residual)      // the block's input tile: same rows, same channel (cin == cout, pin == pout)
          v += from_h16(lds_u16(xin + sw_off(row, cchunk, clow2)), bf);
        const uint16_t h = to_h16(v, bf);
        if (last) a.out[((size_t)(cx.g0 + g) * POUT + p) * cout + c] = h;
        else sts_u16(xout + sw_off(row, cchunk, clow2), h);
      }
    };
    for (int g = cx.sub; g < cx.gn; g += 4) {
      uint32_t raw0[POUT4];
      tmem_ld_cols<POUT4>(tmem_base + d.col_proj + (uint32_t)(o * d.npad_out + g * POUT), raw0);
      ptx::tmem_ld_wait();
      if (valid) clip_out(raw0, g);
    }
  }
  if (!last) {
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncwarp();
    if (cx.lane == 0) ptx::mbar_arrive(bars + kBarXReady);
  }
  if (dbg) dbg[6] = clock64();
}

__global__ void __launch_bounds__(kFusedThreads, 1)
mbconv_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const KernelArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.L.bar_off);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + kBarCount);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = a.L.stages;
  const int g0 = blockIdx.x * a.G;
  const int gn = max(0, min(a.G, a.batch - g0));        // 0 for the padding CTAs of the last cluster
  const int csize = (int)cluster_nctarank(), crank = (int)cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmap_x);
    ptx::tma_prefetch_desc(&a.blocks[0].tm_exp);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kMaxStages; ++s) {
        ptx::mbar_init(bars + kBarFull + s, 1);
        ptx::mbar_init(bars + kBarEmpty + s, (uint32_t)csize);   // one (multicast) commit from every CTA of the cluster
      }
      ptx::mbar_init(bars + kBarXFull, 1);
      ptx::mbar_init(bars + kBarXReady, kFusedComputeWarps);
      for (int s = 0; s < 2; ++s) {
        ptx::mbar_init(bars + kBarExpFull + s, 1);
        ptx::mbar_init(bars + kBarExpEmpty + s, kFusedComputeWarps);
      }
      ptx::mbar_init(bars + kBarPoolReady, kFusedComputeWarps);
      ptx::mbar_init(bars + kBarFc1Full, 1);
      ptx::mbar_init(bars + kBarSReady, kFusedComputeWarps);
      ptx::mbar_init(bars + kBarFc2Full, 1);
      ptx::mbar_init(bars + kBarDReady, kFusedComputeWarps);
      ptx::mbar_init(bars + kBarProjFull, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr_smem, 512);
  }
  // clear the pooled / squeeze operands once (rows without a clip only feed accumulator lanes nobody reads)
  for (uint32_t i = threadIdx.x * 16u; i < a.L.ring_off - a.L.pool_off; i += kFusedThreads * 16u)
    *reinterpret_cast<uint4*>(smem + a.L.pool_off + i) = make_uint4(0u, 0u, 0u, 0u);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();                    // every CTA's barriers exist before a peer signals them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();              // everything above touched only constants; activations are valid from here

  if (warp < 4) {
    // warp group 0: TMA producer, MMA issuer (one lane each) and two idle warps; it hands most of its registers to the
    // compute warp groups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsControl));
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      {
        const FusedBlockDev* b0 = a.blocks;
        const BlockDims d = block_dims(b0->cin, b0->cexp, b0->cout, b0->pin, b0->pout, b0->pool_out, a.G);
        ptx::mbar_expect_tx(bars + kBarXFull, (uint32_t)(d.kb_in * d.npad_in * 128));
        for (int kb = 0; kb < d.kb_in; ++kb)
          ptx::tma_load_2d(&tmap_x, bars + kBarXFull, smem + a.L.x_off + (size_t)kb * d.npad_in * 128, kb * 64, g0 * b0->pin);
      }
      // A ring stage is filled by all CTAs of the cluster together: this CTA waits until every CTA has released the
      // stage (its own empty barrier counts one commit per CTA), arms its own full barrier with the whole stage's
      // bytes and fetches its share of the boxes, which TMA writes into the same stage of every CTA in the cluster.
      auto acquire = [&](uint32_t bytes) -> uint8_t* {
        bar_wait(bars + kBarEmpty + stage, phase ^ 1u);
        ptx::mbar_expect_tx(bars + kBarFull + stage, bytes);
        return smem + a.L.ring_off + (size_t)stage * kStageBytes;
      };
      auto advance = [&]() { if (++stage == stages) { stage = 0; phase ^= 1u; } };
      auto load = [&](const CUtensorMap* tm, uint8_t* dst, int c0, int c1) {
        if (csize > 1) tma_load_2d_mc(tm, bars + kBarFull + stage, dst, c0, c1, cmask);
        else ptx::tma_load_2d(tm, bars + kBarFull + stage, dst, c0, c1);
      };
      // a 128-row x 64-column weight tile = four boxes of 32 rows; box q belongs to CTA q % csize
      auto load_tile = [&](const CUtensorMap* tm, int c0, int row0) {
        uint8_t* dst = acquire(kStageBytes);
        for (int q = crank; q < 128 / kBoxRows; q += csize)
          load(tm, dst + (size_t)q * (kBoxRows * 128), c0, row0 + q * kBoxRows);
        advance();
      };
      for (int bi = 0; bi < a.nblocks; ++bi) {
        const FusedBlockDev* blk = a.blocks + bi;
        const BlockDims d = block_dims(blk->cin, blk->cexp, blk->cout, blk->pin, blk->pout, blk->pool_out, a.G);
        for (int j = 0; j < d.nchunk; ++j)
          for (int kb = 0; kb < d.kb_in; ++kb) load_tile(&blk->tm_exp, kb * 64, j * 128);
        if (blk->pool_out) continue;
        // squeeze weights: se_pad rows per k-block, several k-blocks share one ring stage (one box each)
        const uint32_t sub_bytes = (uint32_t)blk->se_pad * 128u;
        const int pack = (int)(kStageBytes / sub_bytes);
        for (int kb0 = 0; kb0 < d.kb_exp; kb0 += pack) {
          const int n = min(pack, d.kb_exp - kb0);
          uint8_t* dst = acquire(sub_bytes * (uint32_t)n);
          for (int i = crank; i < n; i += csize) load(&blk->tm_se1, dst + (size_t)i * sub_bytes, (kb0 + i) * 64, 0);
          advance();
        }
        for (int j = 0; j < d.nchunk; ++j) load_tile(&blk->tm_se2, 0, j * 128);
        for (int o = 0; o < d.nout; ++o)
          for (int kb = 0; kb < d.kb_exp; ++kb) load_tile(&blk->tm_proj, kb * 64, o * 128);
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer =====================
      int stage = 0;
      uint32_t phase = 0;
      uint32_t exp_use0 = 0u, exp_use1 = 0u;
      const int fmt = a.bf16 ? 1 : 0;
      const uint32_t ring = ptx::smem_u32(smem + a.L.ring_off);
      const uint32_t pool_kb = (uint32_t)a.L.fc_rows * 128u;
      long long* const idbg = (a.dbg && blockIdx.x == 0) ? a.dbg : nullptr;
      long long starve = 0;                   // debug: cycles this thread waited for TMA data
      auto wait_stage = [&]() -> uint32_t {
        const long long t0 = idbg ? clock64() : 0;
        bar_wait(bars + kBarFull + stage, phase);
        if (idbg) starve += clock64() - t0;
        ptx::tc_fence_after();
        return ring + (uint32_t)stage * kStageBytes;
      };
      auto release_stage = [&]() {            // the stage may be refilled once the MMAs of EVERY CTA have read it
        if (csize > 1) tc_commit_mc(bars + kBarEmpty + stage, cmask);
        else ptx::tc_commit(bars + kBarEmpty + stage);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      };
      // `steps` MMAs of K = 16: A = 128 rows at a_addr, B = N rows at b_addr (both K-major SWIZZLE_128B k-blocks)
      auto mmas = [&](uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, int steps, uint32_t idesc, bool first_kb) {
        const uint64_t a_desc = ptx::umma_desc_kmajor(a_addr, 128);
        const uint64_t b_desc = ptx::umma_desc_kmajor(b_addr, 128);
        for (int k = 0; k < steps; ++k)
          ptx::tc_mma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                          (uint32_t)(!(first_kb && k == 0)));
      };
      for (int bi = 0; bi < a.nblocks; ++bi) {
        const FusedBlockDev* blk = a.blocks + bi;
        const BlockDims d = block_dims(blk->cin, blk->cexp, blk->cout, blk->pin, blk->pout, blk->pool_out, a.G);
        const uint32_t par = (uint32_t)bi & 1u;
        const uint32_t idesc_in = ptx::umma_idesc_h16_f32(128, d.npad_in, fmt);
        const uint32_t idesc_fc = ptx::umma_idesc_h16_f32(128, kFcN, fmt);
        const uint32_t idesc_out = ptx::umma_idesc_h16_f32(128, d.npad_out, fmt);
        const uint32_t idesc_se = ptx::umma_idesc_h16_f32(128, blk->se_pad > 0 ? blk->se_pad : 16, fmt);
        const uint32_t x_addr = ptx::smem_u32(smem + a.L.x_off);
        if (bi == 0) bar_wait(bars + kBarXFull, 0);
        else bar_wait(bars + kBarXReady, ((uint32_t)(bi - 1)) & 1u);
        ptx::tc_fence_after();
        // 1. expand
        for (int j = 0; j < d.nchunk; ++j) {
          const int st = d.col_exp1 != 0 ? (j & 1) : 0;
          bar_wait(bars + kBarExpEmpty + st, ((st ? exp_use1 : exp_use0) & 1u) ^ 1u);
          if (st) ++exp_use1; else ++exp_use0;
          ptx::tc_fence_after();
          for (int kb = 0; kb < d.kb_in; ++kb) {
            const int steps = min(4, (blk->cin - kb * 64 + 15) >> 4);
            const uint32_t sa = wait_stage();
            mmas(tmem_base + (st ? d.col_exp1 : 0u), sa, x_addr + (uint32_t)(kb * d.npad_in * 128), steps, idesc_in, kb == 0);
            release_stage();
          }
          ptx::tc_commit(bars + kBarExpFull + st);
        }
        if (blk->pool_out) continue;
        // 2. squeeze
        bar_wait(bars + kBarPoolReady, par);
        ptx::tc_fence_after();
        const uint32_t pool_addr = ptx::smem_u32(smem + a.L.pool_off);
        const uint32_t sub_bytes = (uint32_t)blk->se_pad * 128u;
        const int pack = (int)(kStageBytes / sub_bytes);
        for (int kb0 = 0; kb0 < d.kb_exp; kb0 += pack) {
          const int n = min(pack, d.kb_exp - kb0);
          const uint32_t sa = wait_stage();
          for (int i = 0; i < n; ++i) {
            const int kb = kb0 + i;
            const int steps = min(4, (blk->cexp - kb * 64 + 15) >> 4);
            // A = pooled means (clips on the accumulator's lanes: <= 16 real rows, the rest of the 128 reads whatever
            // follows in shared memory and lands in lanes nobody reads), B = se_pad rows of W1^T
            mmas(tmem_base + d.col_fc1, pool_addr + (uint32_t)kb * pool_kb, sa + (uint32_t)i * sub_bytes, steps, idesc_se,
                 kb == 0);
          }
          release_stage();
        }
        ptx::tc_commit(bars + kBarFc1Full);
        // 3. excite
        bar_wait(bars + kBarSReady, par);
        ptx::tc_fence_after();
        const uint32_t s_addr = ptx::smem_u32(smem + a.L.s_off);
        for (int j = 0; j < d.nchunk; ++j) {
          const uint32_t sa = wait_stage();
          mmas(tmem_base + d.col_fc2 + (uint32_t)(j * kFcN), sa, s_addr, blk->se_pad >> 4, idesc_fc, true);
          release_stage();
        }
        ptx::tc_commit(bars + kBarFc2Full);
        // 4. project
        bar_wait(bars + kBarDReady, par);
        ptx::tc_fence_after();
        const uint32_t d_addr = ptx::smem_u32(smem + a.L.d_off);
        for (int o = 0; o < d.nout; ++o)
          for (int kb = 0; kb < d.kb_exp; ++kb) {
            const int steps = min(4, (blk->cexp - kb * 64 + 15) >> 4);
            const uint32_t sa = wait_stage();
            mmas(tmem_base + d.col_proj + (uint32_t)(o * d.npad_out), sa, d_addr + (uint32_t)(kb * d.rpad_out * 128), steps,
                 idesc_out, kb == 0);
            release_stage();
          }
        ptx::tc_commit(bars + kBarProjFull);
        if (idbg) { idbg[bi * 8 + 7] = starve; starve = 0; }
      }
    }
    __syncwarp();
  } else {
    // ===================== compute warps (warp groups 1..4) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
    ComputeCtx cx;
    cx.smem = ptx::smem_u32(smem); cx.bars = bars;
    const int q = warp & 3;
    cx.tmem_base = tmem_base + ((uint32_t)(q * 32) << 16);
    cx.cw = warp - 4;
    cx.sub = cx.cw >> 2;
    cx.chl = q * 32 + lane;
    cx.g0 = g0; cx.gn = gn; cx.G = a.G; cx.bf16 = a.bf16; cx.lane = lane;
    cx.exp_use0 = cx.exp_use1 = 0u;
    for (int bi = 0; bi < a.nblocks; ++bi) {
      const FusedBlockDev* blk = a.blocks + bi;
      const bool first = bi == 0, last = bi + 1 == a.nblocks;   // only the last block of the launch stores to global
      const int next_npad_in = bi + 1 < a.nblocks ? round_up_i(a.G * a.blocks[bi + 1].pin, 16) : 0;
      switch (blk->geom) {
        case 1: compute_block<1>(cx, a, blk, bi, first, last, next_npad_in); break;
        case 2: compute_block<2>(cx, a, blk, bi, first, last, next_npad_in); break;
        case 3: compute_block<3>(cx, a, blk, bi, first, last, next_npad_in); break;
        case 4: compute_block<4>(cx, a, blk, bi, first, last, next_npad_in); break;
        case 5: compute_block<5>(cx, a, blk, bi, first, last, next_npad_in); break;
        case 6: compute_block<6>(cx, a, blk, bi, first, last, next_npad_in); break;
        default: compute_block<7>(cx, a, blk, bi, first, last, next_npad_in); break;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();                    // no CTA leaves while a peer may still signal its barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

FusedSmem fused_smem(const FusedBlockInfo* blocks, int nblocks, int G, int max_smem) {
  FusedSmem L;
  L.fc_rows = G <= 8 ? 8 : 16;
  uint32_t x_bytes = 0, d_bytes = 0, pool_bytes = 0;
  for (int i = 0; i < nblocks; ++i) {
    const FusedBlockInfo& b = blocks[i];
    const BlockDims d = block_dims(b.cin, b.cexp, b.cout, b.pin, b.pout, b.pool_out, G);
    x_bytes = max(x_bytes, (uint32_t)(d.kb_in * d.npad_in * 128));
    if (!b.pool_out) {
      d_bytes = max(d_bytes, (uint32_t)(d.kb_exp * d.rpad_out * 128 + 1024));   // + one 8-row group: N is padded to 16
      pool_bytes = max(pool_bytes, (uint32_t)(d.kb_exp * L.fc_rows * 128));    // also holds the [G][cexp] gate table
    }
  }
  pool_bytes = (pool_bytes + 1023u) & ~1023u;
  L.x_off = 0;
  L.d_off = x_bytes;
  L.pool_off = L.d_off + d_bytes;
  L.s_off = L.pool_off + pool_bytes;
  L.ring_off = L.s_off + (d_bytes ? kFcN * 128 : 0);
  const uint32_t fixed = L.ring_off + 8 * (kBarCount + 2) + 1024 /* alignment slack */;
  int stages = ((uint32_t)max_smem > fixed) ? (int)(((uint32_t)max_smem - fixed) / kStageBytes) : 0;
  if (stages > kMaxStages) stages = kMaxStages;
  L.stages = stages;
  L.bar_off = L.ring_off + (uint32_t)stages * kStageBytes;
  L.total = L.bar_off + 8 * (kBarCount + 2) + 1024;
  return L;
}

bool fused_fits(const FusedBlockInfo* blocks, int nblocks, int G, int max_smem) {
  if (G < 1 || G > kFcN) return false;
  for (int i = 0; i < nblocks; ++i) {
    const FusedBlockInfo& b = blocks[i];
    const BlockDims d = block_dims(b.cin, b.cexp, b.cout, b.pin, b.pout, b.pool_out, G);
    if (d.npad_in > 256 || d.npad_out > 256) return false;
    if (block_tmem_cols(d, b.pool_out) > 512) return false;
    if (b.cin % 8 || b.cexp % 16 || b.cout % 8) return false;
    if (!b.pool_out && (b.se_pad % 16 || b.se_pad > 64)) return false;
    if (b.pool_out && i + 1 != nblocks) return false;
    if (i > 0 && (blocks[i - 1].cout != b.cin || blocks[i - 1].pout != b.pin)) return false;
    if (b.residual && (b.cin != b.cout || b.pin != b.pout)) return false;
  }
  return fused_smem(blocks, nblocks, G, max_smem).stages >= 2;
}

}  // namespace

int fused_geom_id(int k, int s, int h, int w, int pad_top, int pad_left) {
  auto is = [&](int k_, int s_, int h_, int w_, int pt, int pl) {
    return k == k_ && s == s_ && h == h_ && w == w_ && pad_top == pt && pad_left == pl;
  };
  if (is(3, 2, 7, 5, 1, 1)) return 1;
  if (is(3, 1, 4, 3, 1, 1)) return 2;
  if (is(5, 1, 4, 3, 2, 2)) return 3;
  if (is(5, 2, 4, 3, 1, 2)) return 4;
  if (is(5, 1, 2, 2, 2, 2)) return 5;
  if (is(3, 1, 2, 2, 1, 1)) return 6;
  return 0;
}

int fused_max_group(const FusedBlockInfo* blocks, int nblocks, int max_smem) {
  for (int G = kFcN; G >= 1; --G)
    if (fused_fits(blocks, nblocks, G, max_smem)) return G;
  return 0;
}

int launch_mbconv_fused(const void* d_x, int batch, const FusedBlockDev* d_blocks, const FusedBlockInfo* h_blocks,
                        int nblocks, void* d_out, int bf16, int sm_count, int max_smem, cudaStream_t st) {
  if (batch == 0) return KWS_OK;
  KWS_REQUIRE(nblocks >= 1 && d_blocks && h_blocks && d_x && d_out, "mbconv_fused: bad argument");
  const int gmax = fused_max_group(h_blocks, nblocks, max_smem);
  KWS_REQUIRE(gmax >= 1, "mbconv_fused: block does not fit shared memory / TMEM");
  // KWS_FUSED_CLUSTER = 2 / 4: clusters of CTAs share every weight tile (one L2 read per cluster, multicast into all
  // rings).  Measured on the B200: identical phase times for 1, 2 and 4 — the weight stream is not what bounds the
  // kernel (its ~40 dependent, latency-bound steps per block are), so the default stays 1 (all 148 SMs;
  // clusters of 4 fit only 132).
  static const int cluster_env = [] { const char* e = getenv("KWS_FUSED_CLUSTER"); return e ? atoi(e) : 0; }();
  int C = (cluster_env == 2 || cluster_env == 4) ? cluster_env : 1;
  const int slots = C == 4 ? (sm_count * 132) / 148 : sm_count;
  // one CTA per SM when the batch allows it: G = ceil(batch / resident CTAs), capped by what fits
  int G = (batch + slots - 1) / (slots > 0 ? slots : 1);
  if (G > gmax) G = gmax;
  if (G < 1) G = 1;
  int grid = (batch + G - 1) / G;
  if (grid < 2 * C) C = 1;                              // tiny batches: nothing to share
  grid = (grid + C - 1) / C * C;                        // padding CTAs run the weight pipeline without clips
  KernelArgs a;
  a.blocks = d_blocks; a.nblocks = nblocks; a.batch = batch; a.G = G; a.bf16 = bf16;
  a.L = fused_smem(h_blocks, nblocks, G, max_smem);
  a.out = static_cast<uint16_t*>(d_out);
  a.dbg = nullptr;
  static const bool debug = [] { const char* e = getenv("KWS_FUSED_DEBUG"); return e && atoi(e); }();
  static long long* d_dbg = nullptr;
  if (debug) {                                          // tools only: synchronises the stream and prints phase clocks
    if (!d_dbg) KWS_CUDA_CHECK(cudaMalloc(&d_dbg, sizeof(long long) * 8 * 32));
    KWS_CUDA_CHECK(cudaMemsetAsync(d_dbg, 0, sizeof(long long) * 8 * 32, st));
    a.dbg = d_dbg;
  }
  const FusedBlockInfo& b0 = h_blocks[0];
  const BlockDims d0 = block_dims(b0.cin, b0.cexp, b0.cout, b0.pin, b0.pout, b0.pool_out, G);
  CUtensorMap tx;
  int rc = make_tmap_h16(&tx, d_x, (uint64_t)batch * b0.pin, (uint64_t)b0.cin, (uint32_t)d0.npad_in, bf16, 64);
  if (rc != KWS_OK) return rc;
  KWS_CUDA_CHECK(cudaFuncSetAttribute(mbconv_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.L.total));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kFusedThreads); cfg.dynamicSmemBytes = (size_t)a.L.total; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (C > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)C; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  KWS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mbconv_fused_kernel, tx, a));
  if (debug) {
    long long h[8 * 32];
    KWS_CUDA_CHECK(cudaStreamSynchronize(st));
    KWS_CUDA_CHECK(cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
    for (int b = 0; b < nblocks && b < 32; ++b) {
      const long long* t = h + b * 8;
      fprintf(stderr, "fused blk %d (cexp %d pin %d G %d grid %d cluster %d stages %d smem %u): dw %lld | fc1 wait %lld | "
              "s+fc2 wait %lld | gate %lld | proj wait %lld | epi %lld | TMA-starved %lld (cycles, CTA 0)\n", b, h_blocks[b].cexp,
              h_blocks[b].pin, G, grid, C, a.L.stages, a.L.total, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3],
              t[5] - t[4], t[6] - t[5], t[7]);
    }
  }
  return KWS_OK;
}

}  // namespace kws
