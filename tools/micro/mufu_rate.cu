// Microbenchmark: per-SM throughput of ex2.approx f32 vs f16x2 vs bf16x2 (elements per clock per SM).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float a[8];
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b801u + i + threadIdx.x; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("{.reg .b16 l, u; mov.b32 {l, u}, %0; ex2.approx.f16 l, l; mov.b32 %0, {l, u};}" : "+r"(h[i]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[MODE] = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 1 << 22); cudaMallocManaged(&clk, 64);
  const int iters = 4096, threads = 512;
  k<0><<<148, threads>>>(out, clk, iters); k<1><<<148, threads>>>(out, clk, iters); k<2><<<148, threads>>>(out, clk, iters); k<3><<<148, threads>>>(out, clk, iters);
  cudaDeviceSynchronize();
  const char* nm[4] = {"f32", "f16x2", "bf16x2", "f16"};
  int el[4] = {1, 2, 2, 1};
  for (int m = 0; m < 4; ++m)
    printf("%-7s %lld clk: %.2f instr-lanes/clk/SM, %.2f elements/clk/SM\n", nm[m], clk[m], (double)iters * 8 * threads / clk[m], (double)iters * 8 * threads * el[m] / clk[m]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
