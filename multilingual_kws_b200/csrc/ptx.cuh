// ptx.cuh — inline-PTX wrappers for the sm_100a features the kernels use: mbarrier, TMA (bulk and
// tensor copies), tcgen05 (TMEM alloc, MMA, commit, ld) and the matching fences.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>

namespace kws {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may begin while its
// predecessor on the stream is still running; pdl_wait() blocks until the predecessor grid has completed and its
// writes are visible, pdl_launch_dependents() lets the successor's CTAs be scheduled once every CTA of this grid has
// issued it (or exited).  Everything before pdl_wait() must touch only constants (weights, tensor maps, barriers).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* desc, uint64_t* bar, void* dst_smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16/fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread i = TMEM lane base+i)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30) (ignored
// for swizzled K-major, canonical 1), SBO>>4 [32,46) = bytes between 8-row groups, version [46,48) = 1 on sm_100,
// layout [61,64) = 2 / 4 / 6 (SWIZZLE_128B / 64B / 32B) for tiles whose rows are 128 / 64 / 32 B wide: SBO = 8 rows * row_bytes.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t row_bytes) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
  return ((uint64_t)hi << 32) | lo;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a format [7,10), b format [10,13),
// a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
// fmt: 0 = F16, 1 = BF16 (F16F32Format) for both A and B.
__device__ __forceinline__ uint32_t umma_idesc_h16_f32(int m, int n, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// ---- swish / sigmoid.  Default: sigmoid(x) = 0.5 tanh(x / 2) + 0.5 with the hardware tanh (one MUFU op, relative error
// 2^-11 = the size of the 16-bit rounding the result gets anyway).  -DKWS_SWISH_EXACT selects 1 / (1 + 2^(-x log2 e))
// as EX2 + RCP (two MUFU ops, relative error ~2^-22): measured on the B200 it changes the embedding's cosine against the
// fp32 oracle in the sixth decimal (0.999914 vs 0.999909 on the trained-like net) and costs 6 % of the forward pass
// (profiles/README.md, round 2), so it is not the default.
__device__ __forceinline__ float sigmoid_f(float x) {
#ifndef KWS_SWISH_EXACT
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(0.5f * x));
  return fmaf(0.5f, y, 0.5f);
#else
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
#endif
}
__device__ __forceinline__ float swish_f(float x) {
#ifndef KWS_SWISH_EXACT
  const float h = 0.5f * x;
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(h));
  return fmaf(h, y, h);
#else
  return x * sigmoid_f(x);
#endif
}

// ---- 16-bit storage helpers: two values per 32-bit word; bf = 1 -> bfloat16, 0 -> IEEE half (saturating)
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi, int bf) {
  uint32_t r;
  if (bf) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t w, int bf) {
  if (bf) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
  float2 r;
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "r"(w));
  return r;
}

}  // namespace ptx

// Launch with programmatic stream serialization (see ptx::pdl_wait).  Every kernel launched through this helper
// executes ptx::pdl_wait() on all of its threads before it reads activations or writes anything.  KWS_NO_PDL=1
// turns the attribute off (plain stream order) for A/B measurements.
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("KWS_NO_PDL"); return !(e && atoi(e)); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
}  // namespace kws
