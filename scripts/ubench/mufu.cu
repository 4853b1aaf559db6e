// Microbenchmark: MUFU.EX2 throughput (f32, f16x2, bf16x2) and FMA-pipe polynomial exp2 per SM on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[8];
  unsigned int h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0x3c003c00u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) {  // Cody-Waite + degree-3 polynomial on the FMA pipe
        float x = a[i];
        float fl = floorf(x);
        float f = x - fl;
        float p = fmaf(f, 0.0555041086f, 0.2402265069f);
        p = fmaf(p, f, 0.6931471806f);
        p = fmaf(p, f, 1.0f);
        a[i] = __int_as_float(__float_as_int(p) + (static_cast<int>(fl) << 23)) * -0.5f;
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE><<<148, threads>>>(out, iters, cyc);
  k<MODE><<<148, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double ops = double(iters) * 8 * threads * ((MODE == 1 || MODE == 2) ? 2 : 1);
  printf("%-10s threads/SM=%4d: %8lld cycles, %.2f results/clk/SM\n", name, threads, h, ops / h);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int t : {128, 256, 512, 1024}) {
    run<0>("ex2.f32", t); run<1>("ex2.f16x2", t); run<2>("ex2.bf16x2", t); run<3>("poly3", t);
  }
  return 0;
}
