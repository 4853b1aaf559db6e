// Microbenchmark 3: MMA issue cost when the issuing warp runs in warp-uniform control flow (warp index and TMEM base
// broadcast with shfl so the compiler can keep descriptor arithmetic on the uniform datapath) and issues under elect.sync.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../vla_adapter_b200/csrc/common.cuh"
using namespace vla;

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

template <int N, int SYNC_EACH>
__global__ void __launch_bounds__(128, 1) k(int count, int stages, long long* cyc) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t sb = (smem_u32(raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); fence_proxy_async(); }
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); tc_fence_before(); }
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t tm = __shfl_sync(0xffffffffu, slot, 0);
    const uint32_t idesc = make_idesc_bf16(128, N);
    long long t0 = clock64();
    int stage = 0;
    for (int i = 0; i < count; i += 4) {
      const uint32_t sa = sb + stage * 4096, sbb = sb + 16384 + stage * 4096;  // runtime stage, like a real ring
      if (SYNC_EACH) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (elect_one()) umma_bf16(tm + 256, mk_desc(sa + kk * 32, 1024, 2), mk_desc(sbb + kk * 32, 1024, 2), idesc, 1u);
          __syncwarp();
        }
      } else {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tm + 256, mk_desc(sa + kk * 32, 1024, 2), mk_desc(sbb + kk * 32, 1024, 2), idesc, 1u);
        }
        __syncwarp();
      }
      if (++stage == stages) stage = 0;
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int N, int SYNC_EACH>
void run(long long* cyc) {
  long long h[2];
  cudaFuncSetAttribute(k<N, SYNC_EACH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int count = 512;
  k<N, SYNC_EACH><<<148, 128, 64 * 1024>>>(count, 3, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d %s: issue %6.1f cyc/MMA, issue+complete %6.1f cyc/MMA -> %.0f flop/clk/SM (%s)\n", N,
         SYNC_EACH ? "elect per MMA " : "elect per 4 MMA", double(h[0]) / count, double(h[1]) / count,
         2.0 * 128 * N * 16 * count / h[1], cudaGetErrorString(e));
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 16);
  run<256, 0>(cyc); run<256, 1>(cyc); run<128, 0>(cyc); run<128, 1>(cyc); run<64, 0>(cyc); run<64, 1>(cyc); run<16, 0>(cyc);
  return 0;
}
