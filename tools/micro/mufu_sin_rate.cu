// Microbenchmark: per-SM throughput of sin.approx.f32 (FMUL.RZ + MUFU.SIN) against ex2.approx.f32 (MUFU.EX2), and of a
// mix of MUFU.SIN with independent FFMA work (results per clock per SM).  Decides whether Activation1d's 2 sines per
// output run at the 16 results/clk/SM measured for ex2.
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float a[8], b[8];
  for (int i = 0; i < 8; ++i) { a[i] = 0.001f * (threadIdx.x + i); b[i] = 1.0f + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 2) asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 3) {
        asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(b[i]));
        asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(b[i]));
      }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[MODE] = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 1 << 22); cudaMallocManaged(&clk, 64);
  const int iters = 4096, threads = 512;
  k<0><<<148, threads>>>(out, clk, iters); k<1><<<148, threads>>>(out, clk, iters); k<2><<<148, threads>>>(out, clk, iters); k<3><<<148, threads>>>(out, clk, iters);
  cudaDeviceSynchronize();
  const char* nm[4] = {"ex2", "sin", "cos", "sin+2fma"};
  for (int m = 0; m < 4; ++m)
    printf("%-9s %lld clk: %.2f results/clk/SM\n", nm[m], clk[m], (double)iters * 8 * threads / clk[m]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
