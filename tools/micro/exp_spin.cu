// Microbenchmark: does a warp spinning on mbarrier.try_wait slow down the exp stream of another warp on the same scheduler?
#include <cstdio>
#include <cuda_bf16.h>
#include <cstdint>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packbf(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
// MODE 0: spinners use try_wait; 1: test_wait; 2: try_wait + nanosleep(20); 3: spinners idle (exit)
template <int MODE>
__global__ void k(const float* in, uint32_t* out, long long* clk, int iters, int nspin) {
  __shared__ uint64_t bar;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar))); done = 0; }
  __syncthreads();
  const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&bar);
  if (warp >= 4) {            // spinner warps: warp 4.. share schedulers with warps 0..3
    if (MODE == 3 || warp >= 4 + 4 * nspin) return;
    while (!done) {
      uint32_t ok;
      if (MODE == 1) asm volatile("{.reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], 0; selp.b32 %0,1,0,P;}" : "=r"(ok) : "r"(ba) : "memory");
      else asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.b32 %0,1,0,P;}" : "=r"(ok) : "r"(ba) : "memory");
      if (MODE == 2) __nanosleep(20);
    }
    return;
  }
  float s[48];
  for (int i = 0; i < 48; ++i) s[i] = in[(threadIdx.x * 48 + i) & 4095];
  uint32_t acc = 0;
  float m = in[threadIdx.x & 1023];
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[24];
    const float2 nm = make_float2(-m, -m);
#pragma unroll
    for (int e = 0; e < 48; e += 2) {
      float2 d = __fadd2_rn(make_float2(s[e], s[e + 1]), nm);
      pk[e >> 1] = packbf(ex2(d.x), ex2(d.y));
    }
#pragma unroll
    for (int e = 0; e < 24; ++e) acc ^= pk[e];
    m += __uint_as_float(acc & 1);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  __threadfence_block();
  done = 1;
}
int main() {
  float* in; uint32_t* out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4); cudaMalloc(&out, 1 << 22); cudaMallocManaged(&clk, 64);
  const int iters = 1000;
  for (int ns = 1; ns <= 2; ++ns) {
    k<0><<<148, 384>>>(in, out, clk, iters, ns); cudaDeviceSynchronize(); long long a = clk[0];
    k<1><<<148, 384>>>(in, out, clk, iters, ns); cudaDeviceSynchronize(); long long b = clk[0];
    k<2><<<148, 384>>>(in, out, clk, iters, ns); cudaDeviceSynchronize(); long long c = clk[0];
    k<3><<<148, 384>>>(in, out, clk, iters, ns); cudaDeviceSynchronize(); long long d = clk[0];
    printf("1 exp warp + %d spinner(s) per scheduler: clocks per 48-key tile: try_wait %.0f  test_wait %.0f  try_wait+nanosleep %.0f  none %.0f\n", ns, (double)a / iters, (double)b / iters, (double)c / iters, (double)d / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
