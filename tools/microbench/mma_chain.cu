// mma_chain.cu — cost of one tcgen05.mma (kind::f16, M = 128, cta_group::1, both operands in shared memory, K = 16) as a
// function of N and of the number of independent TMEM accumulators it is spread over.
//
// The first version of this file (round 2) computed the accumulator index as `i % chains` with a run-time `chains`: the
// integer-division sequence (I2F / MUFU.RCP / F2I / IMAD.HI ..., ~35 dependent instructions on the ONE issuing thread) cost
// 148 cycles per iteration and that — not the tensor core — was what it reported for every N.  Here the issuing loop is
// four instructions per MMA (accumulator index by mask, descriptors precomputed, loop unrolled by 8).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multilingual_kws_b200/csrc -o /tmp/mma_chain tools/microbench/mma_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace kws;

template <int kChains>
__global__ void __launch_bounds__(128, 1) chain_kernel(int n, int count, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h
  if (threadIdx.x < 32) ptx::tmem_alloc(&tmem_ptr, 512);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_h16_f32(128, n, 0);
    uint64_t a_desc[4], b_desc[4];
    uint32_t acc[kChains];
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // the four K = 16 slices of one 128-byte swizzled row, as a GEMM k-step walks them
      a_desc[k] = ptx::umma_desc_kmajor(ptx::smem_u32(smem), 128) + (uint64_t)(2 * k);
      b_desc[k] = ptx::umma_desc_kmajor(ptx::smem_u32(smem + 16384), 128) + (uint64_t)(2 * k);
    }
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = tmem + (uint32_t)(c * n);
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < count; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          ptx::tc_mma_f16(acc[u % kChains], a_desc[u & 3], b_desc[u & 3], idesc, (i | u) >= kChains);
      }
      const long long t1 = clock64();
      ptx::tc_commit(&bar);
      ptx::mbar_wait(&bar, (uint32_t)rep & 1u);
      const long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

template <int kChains>
static void run(int n, int count, long long* d) {
  if (kChains * n > 512) return;
  cudaFuncSetAttribute(chain_kernel<kChains>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  chain_kernel<kChains><<<1, 128, 100 * 1024>>>(n, count, d);
  long long h[4];
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("N=%3d accumulators=%d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (tensor floor 128*N/256 = %d)%s\n", n, kChains,
         (double)h[2] / count, (double)h[3] / count, n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int count = 1024;
  for (int n : {16, 32, 48, 64, 96, 128, 192, 256}) {
    run<1>(n, count, d);
    run<2>(n, count, d);
    run<4>(n, count, d);
  }
  return 0;
}
