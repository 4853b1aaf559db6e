// Microbenchmark: tcgen05.mma issue+execution time per instruction for the operand forms the attention kernel uses.
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: bit0 = A from TMEM, bit1 = B MN-major, bit2 = 32B swizzle; n = N
template <int ELECT>
__global__ void __launch_bounds__(128, 1) k(int mode, int n, int count, long long* cyc) {
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
  if (ELECT ? threadIdx.x < 32 : threadIdx.x == 0) {
    const bool ts = mode & 1, mn = mode & 2, s32 = mode & 4;
    const uint32_t idesc = make_idesc_bf16(128, n) | (mn ? (1u << 16) : 0u);
    const uint32_t layout = s32 ? 6u : 2u, sbo = s32 ? 256u : 1024u;
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t step = mn ? (s32 ? 512u : 2048u) : 32u;
      const uint64_t bd = mk_desc(sb + 16384 + (i & 3) * step, sbo, layout);
      if (!ELECT || elect_one()) {
        if (ts) mma_ts(tm + 256, tm + (i & 7) * 8, bd, idesc, i != 0);
        else umma_bf16(tm + 256, mk_desc(sb + (i & 3) * 32, sbo, layout), bd, idesc, i != 0);
      }
      if (ELECT) __syncwarp();
    }
    long long t1 = clock64();
    if (!ELECT || elect_one()) umma_commit(smem_u32(&bar));
    if (ELECT) __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* cyc; long long h[2];
  cudaMalloc(&cyc, 16);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct { const char* name; int mode, n; } cases[] = {
      {"SS K-major sw128 N=256", 0, 256}, {"SS K-major sw128 N=128", 0, 128}, {"SS K-major sw128 N=64", 0, 64},
      {"SS K-major sw32  N=128", 4, 128}, {"TS B K-major sw128 N=64", 1, 64}, {"TS B MN-major sw128 N=64", 3, 64},
      {"TS B MN-major sw128 N=128", 3, 128}, {"TS B MN-major sw32 N=16", 7, 16}, {"SS B MN-major sw128 N=64", 2, 64}};
  for (auto& c : cases) {
    for (int el = 0; el < 2; ++el) {
      const int count = 256;
      if (el) k<1><<<148, 128, 64 * 1024>>>(c.mode, c.n, count, cyc);
      else k<0><<<148, 128, 64 * 1024>>>(c.mode, c.n, count, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("%-28s %s: issue %7.1f cyc/MMA, issue+complete %7.1f cyc/MMA (%s)\n", c.name, el ? "elect.sync" : "lane==0   ",
             double(h[0]) / count, double(h[1]) / count, cudaGetErrorString(e));
    }
  }
  return 0;
}
