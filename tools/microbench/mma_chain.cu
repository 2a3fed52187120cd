// mma_chain.cu — how long does a tcgen05.mma (kind::f16, M = 128, cta_group::1, both operands in shared memory) take when
// consecutive instructions accumulate into the SAME TMEM tile, versus round-robin over several independent tiles?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multilingual_kws_b200/csrc -o /tmp/mma_chain tools/microbench/mma_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace kws;

__global__ void __launch_bounds__(128, 1) chain_kernel(int n, int chains, int count, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h
  if (threadIdx.x < 32) ptx::tmem_alloc(&tmem_ptr, 512);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_h16_f32(128, n, 0);
    const uint64_t a_desc = ptx::umma_desc_kmajor(ptx::smem_u32(smem), 128);
    const uint64_t b_desc = ptx::umma_desc_kmajor(ptx::smem_u32(smem + 16384), 128);
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < count; ++i)
        ptx::tc_mma_f16(tmem + (uint32_t)((i % chains) * n), a_desc + (uint64_t)(2 * (i & 3)), b_desc + (uint64_t)(2 * (i & 3)), idesc, i >= chains);
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

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int count = 256;
  for (int n : {16, 32, 48, 96, 128, 256})
    for (int chains : {1, 2, 4, 8}) {
      if (chains * n > 512) continue;
      chain_kernel<<<1, 128, 100 * 1024>>>(n, chains, count, d);
      long long h[4];
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("N=%3d chains=%d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (ideal N/2 = %d)%s\n", n, chains, (double)h[2] / count,
             (double)h[3] / count, n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
