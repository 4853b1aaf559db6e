// Microbenchmark 2: is the ~96-cycle tcgen05.mma floor an issue-side cost, a same-accumulator dependency, or the shape?
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

// NACC: number of distinct accumulators cycled through; N: MMA N; UNROLL fixed 8
template <int N, int NACC, int M>
__global__ void __launch_bounds__(128, 1) k(int count, long long* cyc) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t sb = (smem_u32(raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); fence_proxy_async(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); tc_fence_before(); }
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N);
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { ad[q] = mk_desc(sb + q * 32, 1024, 2); bd[q] = mk_desc(sb + 16384 + q * 32, 1024, 2); }
    long long t0 = clock64();
    for (int i = 0; i < count; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) umma_bf16(tm + (u % NACC) * 128, ad[u & 3], bd[u & 3], idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, int NACC, int M>
void run(long long* cyc) {
  long long h[2];
  cudaFuncSetAttribute(k<N, NACC, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int count = 512;
  k<N, NACC, M><<<148, 128, 64 * 1024>>>(count, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
  printf("M=%3d N=%3d accumulators=%d: issue %6.1f cyc/MMA, issue+complete %6.1f cyc/MMA  -> %.0f flop/clk/SM (%s)\n", M, N, NACC,
         double(h[0]) / count, double(h[1]) / count, 2.0 * M * N * 16 * count / h[1], cudaGetErrorString(e));
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 16);
  run<256, 1, 128>(cyc); run<256, 2, 128>(cyc);
  run<128, 1, 128>(cyc); run<128, 2, 128>(cyc); run<128, 4, 128>(cyc);
  run<64, 1, 128>(cyc); run<64, 2, 128>(cyc); run<64, 4, 128>(cyc);
  run<32, 1, 128>(cyc); run<32, 4, 128>(cyc);
  run<16, 1, 128>(cyc); run<16, 4, 128>(cyc);
  run<128, 1, 64>(cyc); run<128, 4, 64>(cyc); run<64, 4, 64>(cyc);
  return 0;
}
