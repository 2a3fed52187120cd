// gemm_tcgen05.cu — see gemm_tcgen05.cuh for the design.
#include "gemm_tcgen05.cuh"

#include <cuda_runtime.h>

#include "common.h"
#include "ptx.cuh"

namespace kws {

namespace {


struct SmemLayout {
  uint32_t stage_bytes, bar_off, bias_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(int block_n, int block_k, int stages, int N) {
  SmemLayout L;
  L.stage_bytes = (uint32_t)(kGemmBlockM + block_n) * block_k * 2;     // A tile then W tile, both multiples of 512 B
  L.bar_off = (L.stage_bytes * stages + 1023) & ~1023u;
  L.bias_off = L.bar_off + 8 * (2 * kGemmMaxStages + 4) + 16;          // float s_bias[N rounded up to 16]
  L.total = L.bias_off + (uint32_t)((N + 15) & ~15) * 4;
  return L;
}

// swish / sigmoid: one tanh.approx per element (ptx::swish_f / sigmoid_f; -DKWS_SWISH_EXACT: EX2 + RCP).  The tanh form
// takes h = 0.5 (acc + bias): the bias vector is stored pre-halved in shared memory so h is ONE fma(acc, 0.5, bias/2).
__device__ __forceinline__ bool act_halves_bias(int act) {
#ifndef KWS_SWISH_EXACT
  return act == kActSwish || act == kActSigmoid;
#else
  (void)act;
  return false;
#endif
}
__device__ __forceinline__ float bias_act(float acc, float b, int act) {      // b = bias, or bias / 2 (see above)
#ifndef KWS_SWISH_EXACT
  if (act == kActSwish) { const float h = fmaf(acc, 0.5f, b); float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h)); return fmaf(h, t, h); }
  if (act == kActSigmoid) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(fmaf(acc, 0.5f, b))); return fmaf(0.5f, t, 0.5f); }
#endif
  const float x = acc + b;
  if (act == kActSwish) return ptx::swish_f(x);
  if (act == kActSigmoid) return ptx::sigmoid_f(x);
  if (act == kActRelu) return fmaxf(x, 0.0f);
  if (act == kActSelu) return x > 0.0f ? 1.0507009873554805f * x : 1.7580993408473766f * (__expf(x) - 1.0f);
  return x;
}

// 256-bit global accesses (one full 32-byte sector per thread and instruction)
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// One 16-column chunk of one accumulator row (this thread's row, columns n0 .. n0+15 of the output): bias, activation,
// residual, optional 4-row mean, store.  Every thread owns 32 contiguous output bytes (64 for fp32), so the chunk is
// written by one 256-bit store per thread: full sectors, no staging tile, no synchronisation between epilogue warps.
// ACT / RES / GAP / F32 / BF are compile-time for the configurations the embedding tower uses (-1 = read `ep`).
template <int ACT, int RES, int GAP, int F32, int BF>
__device__ __forceinline__ void epilogue_chunk(const GemmShape& sh, const GemmEpilogue& ep, const uint32_t (&r)[16],
                                               uint32_t s_bias_addr, int n0, int row, bool row_ok, int lane) {
  const int act = ACT >= 0 ? ACT : ep.act;
  const bool has_res = RES >= 0 ? (RES != 0) : (ep.residual != nullptr);
  const bool gap4 = GAP >= 0 ? (GAP != 0) : (ep.gap4 != 0);
  const bool out_f32 = F32 >= 0 ? (F32 != 0) : (ep.out_f32 != 0);
  const int bf16 = BF >= 0 ? BF : ep.bf16;
  const bool full = n0 + 16 <= sh.N;                 // N is a multiple of 8: a chunk holds 16 or 8 valid columns
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
  {                                                   // bias: the whole vector sits in smem (zeros without a bias);
#pragma unroll                                        // warp-uniform 16-byte shared loads (broadcast)
    for (int q = 0; q < 4; ++q) {
      float4 b;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(s_bias_addr + (uint32_t)(n0 + 4 * q) * 4u));
      v[4 * q] = bias_act(v[4 * q], b.x, act); v[4 * q + 1] = bias_act(v[4 * q + 1], b.y, act);
      v[4 * q + 2] = bias_act(v[4 * q + 2], b.z, act); v[4 * q + 3] = bias_act(v[4 * q + 3], b.w, act);
    }
  }
  // residual and 16-bit output rows are 32-byte aligned when the row pitch is a multiple of 16 elements
  if (has_res && row_ok) {          // plain (coherent) loads: the residual may alias the output (in-place skip); each
    uint32_t w[8];                  // thread reads exactly the bytes it overwrites below
    const uint16_t* rp = static_cast<const uint16_t*>(ep.residual) + (size_t)row * ep.ldr + n0;
    if (full && (ep.ldr & 15) == 0) {
      ldg256(rp, w);
    } else {
      const uint4 lo = *reinterpret_cast<const uint4*>(rp);
      uint4 hi = make_uint4(0u, 0u, 0u, 0u);
      if (full) hi = *reinterpret_cast<const uint4*>(rp + 8);
      w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w; w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = ptx::unpack_h2(w[j], bf16);
      v[2 * j] += f.x;
      v[2 * j + 1] += f.y;
    }
  }
  int orow = row;
  bool store = row_ok;
  if (gap4) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v[j] += __shfl_xor_sync(0xffffffffu, v[j], 1);
      v[j] += __shfl_xor_sync(0xffffffffu, v[j], 2);
      v[j] *= 0.25f;
    }
    orow = row >> 2;
    store = row_ok && ((lane & 3) == 0);
  }
  if (!store) return;
  if (out_f32) {
    float* o = reinterpret_cast<float*>(ep.out) + (size_t)orow * ep.ldo + n0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q >= 2 && !full) break;
      *reinterpret_cast<float4*>(o + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  } else {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = ptx::pack_h2(v[2 * j], v[2 * j + 1], bf16);
    uint16_t* o = reinterpret_cast<uint16_t*>(ep.out) + (size_t)orow * ep.ldo + n0;
    if (full && (ep.ldo & 15) == 0) {
      stg256(o, pk);
    } else {
      *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (full) *reinterpret_cast<uint4*>(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
}

template <int kMinBlocks, int ACT, int RES, int GAP, int F32, int BF>
__global__ void __launch_bounds__(kGemmThreads, kMinBlocks)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmShape sh, const GemmEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const SmemLayout L = smem_layout(sh.block_n, sh.block_k, sh.stages, sh.N);
  const uint32_t a_bytes = (uint32_t)kGemmBlockM * sh.block_k * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty_bar = full_bar + kGemmMaxStages;
  uint64_t* tmem_full = empty_bar + kGemmMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias_off);
  // the bias vector (a constant of the layer: safe to read before the dependency wait below) goes to smem once
  {
    const float bscale = act_halves_bias(ACT >= 0 ? ACT : ep.act) ? 0.5f : 1.0f;
    for (int i = threadIdx.x; i < ((sh.N + 15) & ~15); i += kGemmThreads)
      s_bias[i] = (ep.bias && i < sh.N) ? bscale * __ldg(ep.bias + i) : 0.0f;
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int block_n = sh.block_n, stages = sh.stages;
  const int num_kb = (sh.K + sh.block_k - 1) / sh.block_k;
  const int k16_total = (sh.K + 15) / 16;
  const int total_tiles = sh.m_tiles * sh.n_tiles;
  const int acc_stages = sh.acc_stages;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(acc_stages * block_n)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmap_a);
    ptx::tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        ptx::mbar_init(full_bar + s, 1);
        ptx::mbar_init(empty_bar + s, 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(tmem_full + a, 1);
        ptx::mbar_init(tmem_empty + a, 8);   // one arrive per epilogue warp
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr_smem, tmem_cols);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();              // the set-up above overlapped the predecessor's tail; operands are valid from here

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_t = tile / sh.n_tiles, n_t = tile - m_t * sh.n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * L.stage_bytes;
          ptx::mbar_expect_tx(full_bar + stage, L.stage_bytes);
          ptx::tma_load_2d(&tmap_a, full_bar + stage, sa, kb * sh.block_k, m_t * kGemmBlockM);
          ptx::tma_load_2d(&tmap_b, full_bar + stage, sa + a_bytes, kb * sh.block_k, n_t * block_n);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_h16_f32(kGemmBlockM, block_n, ep.bf16 ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * block_n);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar + stage, phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * L.stage_bytes);
          const uint64_t a_desc = ptx::umma_desc_kmajor(sa, (uint32_t)sh.block_k * 2);
          const uint64_t b_desc = ptx::umma_desc_kmajor(sa + a_bytes, (uint32_t)sh.block_k * 2);
          const int k16_per_kb = sh.block_k >> 4;
          const int ksteps = min(k16_per_kb, k16_total - kb * k16_per_kb);
          for (int k = 0; k < ksteps; ++k)   // +32 B along K inside the 128 B swizzle atom = +2 in the (addr >> 4) field
            ptx::tc_mma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          ptx::tc_commit(empty_bar + stage);              // smem slot free once these MMAs retire
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        ptx::tc_commit(tmem_full + acc);                  // accumulator complete
        if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    // The eight warps never synchronise with each other: each waits for the accumulator stage, converts its
    // 16-column chunks (quarter q = its 32 TMEM lanes = 32 output rows; the two warps of a quarter take alternate
    // chunks) and releases the stage.
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                     // which of the two warps of this quarter (chunk parity)
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t s_bias_addr = ptx::smem_u32(s_bias);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_t = tile / sh.n_tiles, n_t = tile - m_t * sh.n_tiles;
      const int row = m_t * kGemmBlockM + q * 32 + lane;
      const bool row_ok = row < sh.M;
      const int n_tile0 = n_t * block_n;
      const int n_lim = min(block_n, sh.N - n_tile0);     // valid columns of this tile
      ptx::mbar_wait(tmem_full + acc, acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * block_n);
      // software-pipelined: the TMEM load of chunk i+1 is in flight while chunk i is processed
      uint32_t ra[16], rb[16];
      int c0 = half * 16;
      if (c0 < n_lim) {
        ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, ra);
        for (;;) {
          ptx::tmem_ld_wait();
          const int c1 = c0 + 32;
          if (c1 < n_lim) ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c1, rb);
          epilogue_chunk<ACT, RES, GAP, F32, BF>(sh, ep, ra, s_bias_addr, n_tile0 + c0, row, row_ok, lane);
          if (c1 >= n_lim) break;
          ptx::tmem_ld_wait();
          c0 = c1 + 32;
          if (c0 < n_lim) ptx::tmem_ld_32x32b_x16(taddr + (uint32_t)c0, ra);
          epilogue_chunk<ACT, RES, GAP, F32, BF>(sh, ep, rb, s_bias_addr, n_tile0 + c1, row, row_ok, lane);
          if (c0 >= n_lim) break;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty + acc);   // TMEM stage may be overwritten by the next-but-one tile
      if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tile width by a small cost model: a CTA streams num_kb k-blocks of (16 KB A + bn*128 B W) per tile at roughly
// 60 KB/us and spends ~0.01 us per output column in the epilogue; tiles are spread over the SMs in waves.
// Few-tile problems (late layers, dense tower) get narrow tiles so all SMs work; big-M problems get the widest
// tile with little column padding.
int pick_block_k(int K) { return K <= 16 ? 16 : (K <= 32 ? 32 : 64); }

int pick_block_n(int N, int K, int m_tiles, int sm_count) {
  const int num_kb = (K + kGemmBlockK - 1) / kGemmBlockK;
  int best = 16;
  double best_cost = 1e30;
  for (int bn = 16; bn <= 256; bn += 16) {
    const int n_tiles = (N + bn - 1) / bn;
    const long long tiles = (long long)n_tiles * m_tiles;
    const long long waves = (tiles + sm_count - 1) / sm_count;
    const double load_us = num_kb * (16.0 + bn * 0.128) / 60.0;
    const double epi_us = bn * 0.01;
    const double tile_us = (load_us > epi_us ? load_us : epi_us) + 0.8;
    const double cost = waves * tile_us;
    if (cost < best_cost * 0.97 || (cost < best_cost && bn > best)) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace

int make_tmap_h16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, int bf16,
                  uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is unavailable (CUDA driver too old or no device)");
    return KWS_ERR_CUDA;
  }
  KWS_REQUIRE(cols % 8 == 0, "tensor map: inner dimension %llu must be a multiple of 8 elements", (unsigned long long)cols);
  KWS_REQUIRE(((uintptr_t)base & 15) == 0, "tensor map: base must be 16-byte aligned");
  KWS_REQUIRE(box_rows >= 1 && box_rows <= 256, "tensor map: box rows %u out of range", box_rows);
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * 2};
  KWS_REQUIRE(box_cols == 16 || box_cols == 32 || box_cols == 64, "tensor map: box columns must be 16, 32 or 64");
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, box_rows};
  const CUtensorMapSwizzle swz = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu box_rows %u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return KWS_ERR_CUDA;
  }
  return KWS_OK;
}

int gemm_h16(const void* a, const void* w, int M, int N, int K, int block_n, const GemmEpilogue& ep_in, int sm_count,
             cudaStream_t stream) {
  KWS_REQUIRE(N % 8 == 0 && K % 8 == 0 && N > 0 && K > 0, "gemm: N (%d) and K (%d) must be positive multiples of 8", N, K);
  KWS_REQUIRE(!ep_in.gap4 || M % 4 == 0, "gemm: gap4 epilogue needs M %% 4 == 0");
  if (M == 0) return KWS_OK;
  GemmEpilogue ep = ep_in;
  GemmShape sh;
  sh.M = M; sh.N = N; sh.K = K;
  sh.m_tiles = (M + kGemmBlockM - 1) / kGemmBlockM;
  // Every epilogue thread reads its residual bytes and then overwrites exactly those bytes, so the in-place skip
  // connection (residual == out) is race-free.
  sh.block_n = block_n > 0 ? block_n : pick_block_n(N, K, sh.m_tiles, sm_count);
  KWS_REQUIRE(sh.block_n % 16 == 0 && sh.block_n >= 16 && sh.block_n <= 256, "gemm: bad block_n %d", sh.block_n);
  sh.n_tiles = (N + sh.block_n - 1) / sh.block_n;
  sh.block_k = pick_block_k(K);
  const int num_kb = (K + sh.block_k - 1) / sh.block_k;
  const int tiles_per_cta = (sh.m_tiles * sh.n_tiles + sm_count - 1) / sm_count;
  // Stage count = how many k-blocks (for single-k-block layers: how many TILES) may be in flight per CTA.  Memory-
  // bound small-K layers want as many as fit in ~half the SM's shared memory (so two CTAs co-reside).
  // Many-tile problems (the early layers: M = 10^5 rows, small K and N) are bound by the per-tile latency chain
  // (TMEM load -> epilogue math -> store), not by operand bandwidth: they run two CTAs per SM, each
  // with half the shared memory and 256 TMEM columns (one accumulator stage when block_n > 128).
  const int tiles_total = sh.m_tiles * sh.n_tiles;
  const bool many_tiles = tiles_total >= 2 * sm_count;
  // One-tile-per-CTA launches (the whole tail of the network: <= 148 tiles) are latency chains, not bandwidth problems:
  // with at most half of the SM's shared memory (and <= 256 TMEM columns) the NEXT kernel's CTA fits beside this one, so
  // its set-up (barriers, TMEM allocation, bias staging, tensor-map fetch) overlaps this kernel under programmatic
  // dependent launch instead of starting when this CTA exits.  KWS_GEMM_BIG_SMEM=1 restores the old rule (A/B).
  static const bool big_smem = [] { const char* e = getenv("KWS_GEMM_BIG_SMEM"); return e && atoi(e); }();
  const bool one_tile = !big_smem && tiles_total <= sm_count && sh.block_n <= 128;
  const size_t budget = ((num_kb <= 2 && sh.block_n <= 128) || (many_tiles && num_kb <= 4) || one_tile) ? 110 * 1024 : 220 * 1024;
  sh.acc_stages = 2;
  int stages = kGemmMaxStages;
  while (stages > 2 && smem_layout(sh.block_n, sh.block_k, stages, N).total + 1024 > budget) --stages;
  const int useful = num_kb * (tiles_per_cta < 8 ? tiles_per_cta : 8) + 1;
  if (stages > useful) stages = useful;
  if (stages < 2) stages = 2;
  sh.stages = stages;
  const size_t smem = smem_layout(sh.block_n, sh.block_k, sh.stages, N).total + 1024;
  KWS_REQUIRE(smem <= 227 * 1024, "gemm: tile configuration does not fit shared memory");

  CUtensorMap ta, tb;
  int rc = make_tmap_h16(&ta, a, (uint64_t)M, (uint64_t)K, kGemmBlockM, ep.bf16, (uint32_t)sh.block_k);
  if (rc != KWS_OK) return rc;
  rc = make_tmap_h16(&tb, w, (uint64_t)N, (uint64_t)K, (uint32_t)sh.block_n, ep.bf16, (uint32_t)sh.block_k);
  if (rc != KWS_OK) return rc;
  const int tiles = sh.m_tiles * sh.n_tiles;
  // Memory/latency-bound shapes (small K, many tiles): two co-resident CTAs per SM hide each other's TMA, TMEM and
  // store latencies.  Needs half the shared memory and at most 256 TMEM columns per CTA.
  const bool two = smem <= 112 * 1024 && tiles >= 2 * sm_count;
  if (two && sh.block_n > 128) sh.acc_stages = 1;
  // compile-time specialised epilogues for the (activation, residual, gap) combinations of the tower in its default
  // fp16 / 16-bit-output form; anything else (bf16 storage, fp32 output, selu) runs the fully run-time variant
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const GemmShape, const GemmEpilogue);
  KernelFn k1 = gemm_tcgen05_kernel<1, -1, -1, -1, -1, -1>, k2 = gemm_tcgen05_kernel<2, -1, -1, -1, -1, -1>;
  const int res = ep.residual ? 1 : 0;
#define KWS_GEMM_VARIANT(A, R, G)                                                \
  if (ep.act == A && res == R && ep.gap4 == G && !ep.out_f32 && !ep.bf16) {      \
    k1 = gemm_tcgen05_kernel<1, A, R, G, 0, 0>;                                  \
    k2 = gemm_tcgen05_kernel<2, A, R, G, 0, 0>;                                  \
  }
  KWS_GEMM_VARIANT(kActSwish, 0, 0)
  KWS_GEMM_VARIANT(kActNone, 0, 0)
  KWS_GEMM_VARIANT(kActNone, 1, 0)
  KWS_GEMM_VARIANT(kActRelu, 0, 0)
  KWS_GEMM_VARIANT(kActSigmoid, 0, 0)
  KWS_GEMM_VARIANT(kActSwish, 0, 1)
#undef KWS_GEMM_VARIANT
  KernelFn kern = two ? k2 : k1;
  KWS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = two ? 2 : 1;
  const int grid = tiles < per_sm * sm_count ? tiles : per_sm * sm_count;
  KWS_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(kGemmThreads), smem, stream, ta, tb, sh, ep));
  return KWS_OK;
}

}  // namespace kws
