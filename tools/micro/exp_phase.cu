// Microbenchmark: clocks for the softmax exp phase of one 48-key tile per thread (FADD2 + 2 MUFU.EX2 + F2FP pack per
// key pair) with W warps per scheduler, and variants (no pack / no ex2 / poly share).
#include <cstdio>
#include <cuda_bf16.h>
#include <cstdint>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packbf(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
template <int MODE>
__global__ void k(const float* in, uint32_t* out, long long* clk, int iters) {
  float s[48];
  for (int i = 0; i < 48; ++i) s[i] = in[(threadIdx.x * 48 + i) & 4095];
  uint32_t acc = 0;
  float m = in[threadIdx.x & 1023];
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[24];
    const float2 nm = make_float2(-m, -m);
#pragma unroll
    for (int e = 0; e < 48; e += 2) {
      float2 d = __fadd2_rn(make_float2(s[e], s[e + 1]), nm);
      if (MODE == 0) pk[e >> 1] = packbf(ex2(d.x), ex2(d.y));
      if (MODE == 1) pk[e >> 1] = __float_as_uint(ex2(d.x)) ^ __float_as_uint(ex2(d.y));   // no pack
      if (MODE == 2) pk[e >> 1] = packbf(d.x, d.y);                                        // no ex2
    }
#pragma unroll
    for (int e = 0; e < 24; ++e) acc ^= pk[e];
    m += __uint_as_float(acc & 1);   // loop-carried dependence
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  float* in; uint32_t* out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, 1 << 22); cudaMallocManaged(&clk, 64);
  const int iters = 1000;
  for (int w = 1; w <= 4; ++w) {
    k<0><<<148, 128 * w>>>(in, out, clk, iters); cudaDeviceSynchronize(); long long a = clk[0];
    k<1><<<148, 128 * w>>>(in, out, clk, iters); cudaDeviceSynchronize(); long long b = clk[0];
    k<2><<<148, 128 * w>>>(in, out, clk, iters); cudaDeviceSynchronize(); long long c = clk[0];
    printf("%d warps/scheduler: clocks per 48-key tile: full %.0f  no-pack %.0f  no-ex2 %.0f\n", w, (double)a / iters, (double)b / iters, (double)c / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
