// Microbenchmark: the softmax exp-phase instruction mix (FFMA2 scale/shift, MUFU.EX2, FADD2 row sum, bf16 pack) per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, int iters, long long* cyc) {
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = -0.01f * (threadIdx.x + i);
  const float2 sl = make_float2(0.18f, 0.18f), nb = make_float2(-0.3f, -0.3f);
  float2 sum = make_float2(0.f, 0.f);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 t = __ffma2_rn(make_float2(v[c * 32 + 2 * i], v[c * 32 + 2 * i + 1]), sl, nb);
        if (MODE != 2) { t.x = ex2f(t.x); t.y = ex2f(t.y); }
        v[c * 32 + 2 * i] = t.x; v[c * 32 + 2 * i + 1] = t.y;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 e = make_float2(v[c * 32 + 2 * i], v[c * 32 + 2 * i + 1]);
        if (MODE != 1) { sum = __fadd2_rn(sum, e); pk[i] = pack_bf16(e.x, e.y); } else pk[i] = __float_as_uint(e.x);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= pk[i];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(sum.x + sum.y);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  uint32_t* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 200;
  k<MODE><<<148, threads>>>(out, iters, cyc);
  k<MODE><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-26s warps/SM=%2d: %7.1f cycles per 128-element row, %5.2f elements/clk/SM\n", name, threads / 32, double(h) / iters,
         128.0 * threads * iters / h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int t : {128, 256, 512}) {
    run<0>("full mix", t); run<1>("ffma2 + ex2 only", t); run<2>("no ex2 (fma/add/pack)", t);
  }
  return 0;
}
