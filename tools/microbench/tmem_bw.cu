// TMEM read-bandwidth probe: W warps per CTA (1 CTA per SM) issue tcgen05.ld.32x32b.x32 back to back.
// Prints bytes per clock per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cm3p_b200/csrc
//   tools/microbench/tmem_bw.cu -o tmem_bw && ./tmem_bw
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"

template <int UNROLL>
__global__ void __launch_bounds__(512, 1) probe(int iters, unsigned* sink, long long* cycles, int active_warps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    ptx::tmem_alloc(&slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 32 * UNROLL % 512;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < active_warps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t r[UNROLL][32];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) ptx::tmem_ld_32x32b_x32(base + (u * 32) % 128, r[u]);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= r[u][k];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(slot, 512);
  }
}

int main() {
  unsigned* sink;
  long long* cyc;
  cudaMalloc(&sink, 148 * 512 * 4);
  cudaMallocManaged(&cyc, 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    probe<2><<<148, 512>>>(iters, sink, cyc, warps);
    cudaDeviceSynchronize();
    probe<2><<<148, 512>>>(iters, sink, cyc, warps);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 1; }
    const double bytes = double(iters) * warps * 2 * 4096.0;
    printf("warps %2d: %lld clk, %.1f B/clk/SM\n", warps, cyc[0], bytes / double(cyc[0]));
  }
  return 0;
}
