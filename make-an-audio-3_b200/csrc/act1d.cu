// Fused anti-aliased periodic activation of BigVGAN (one pass over HBM):
//   replicate-pad -> x2 polyphase up-sampling (12-tap Kaiser-sinc) -> Snake / SnakeBeta -> replicate-pad ->
//   12-tap low-pass, stride 2.
// Replaces Activation1d.forward (vocoder/bigvgan/alias_free_torch/act.py:23-28) = UpSample1d (resample.py:25-33) ->
// SnakeBeta/Snake (activations.py:48-59,107-119) -> DownSample1d/LowPassFilter1d (resample.py:46-49,
// filter.py:86-95), which materialise five intermediates of twice the size.
//
// Data is channels-last [B, T, C].  A block stages (TB + 10) time rows x CT channels in shared memory with 16-byte
// loads (row index clamped = the replicate padding of x); each thread owns one channel pair and kTT consecutive
// outputs and slides a 12-entry window of activated 2x-rate samples through registers.
//
// Index algebra (derived from the reference's pad/crop constants 5/5, 15/15, 5/6):
//   u[2q]   = 2 * sum_{k<6} f[11-2k] * x[clamp(q-3+k)]        u[2q+1] = 2 * sum_{k<6} f[10-2k] * x[clamp(q-2+k)]
//   s[m]    = u[m] + sin^2(a u[m]) / (b + 1e-9)
//   out[t]  = sum_{j<12} f[j] * s[clamp(2t-5+j, 0, 2T-1)]
#include "host_common.h"
#include "ptx.cuh"

namespace ma3 {

__constant__ float c_fdn[12];  // f
__constant__ float c_fup[12];  // 2 f

constexpr int kTT = 16;
constexpr int kActThreads = 256;

template <typename T> struct Ld2;
template <> struct Ld2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
};
template <> struct Ld2<__half> {
  static __device__ __forceinline__ float2 ld(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
};
template <> struct Ld2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
  }
};
template <typename T> struct St2;
template <> struct St2<float> {
  static __device__ __forceinline__ void st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
};
template <> struct St2<__half> {
  static __device__ __forceinline__ void st(__half* p, float2 v) { *reinterpret_cast<__half2*>(p) = __float22half2_rn(v); }
};
template <> struct St2<__nv_bfloat16> {
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float2 v) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __float22bfloat162_rn(v);
  }
};

__device__ __forceinline__ float snake1(float u, float a, float inv_b) {
  const float sn = __sinf(u * a);
  return fmaf(inv_b * sn, sn, u);
}

// activated 2x-rate sample with local index n (m = 2*t0 - 5 + n) from the thread's window xw[0 .. kTT+9]
template <int n>
__device__ __forceinline__ float2 s_local(const float2 (&xw)[kTT + 10], float2 a, float2 ib) {
  float2 u = make_float2(0.f, 0.f);
  if constexpr (n & 1) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      u.x = fmaf(c_fup[11 - 2 * k], xw[(n - 1) / 2 + k].x, u.x);
      u.y = fmaf(c_fup[11 - 2 * k], xw[(n - 1) / 2 + k].y, u.y);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      u.x = fmaf(c_fup[10 - 2 * k], xw[n / 2 + k].x, u.x);
      u.y = fmaf(c_fup[10 - 2 * k], xw[n / 2 + k].y, u.y);
    }
  }
  return make_float2(snake1(u.x, a.x, ib.x), snake1(u.y, a.y, ib.y));
}

template <int i, typename TOut>
__device__ __forceinline__ void slide(const float2 (&xw)[kTT + 10], float2 (&S)[12], float2 a, float2 ib, TOut* orow,
                                      long long ostride, int valid) {
  if constexpr (i < kTT) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      acc.x = fmaf(c_fdn[j], S[j].x, acc.x);
      acc.y = fmaf(c_fdn[j], S[j].y, acc.y);
    }
    if (i < valid) St2<TOut>::st(orow + (long long)i * ostride, acc);
    if constexpr (i + 1 < kTT) {
#pragma unroll
      for (int j = 0; j < 10; ++j) S[j] = S[j + 2];
      S[10] = s_local<2 * i + 12>(xw, a, ib);
      S[11] = s_local<2 * i + 13>(xw, a, ib);
      slide<i + 1>(xw, S, a, ib, orow, ostride, valid);
    }
  }
}

template <int n>
__device__ __forceinline__ void fill_window(const float2 (&xw)[kTT + 10], float2 (&S)[12], float2 a, float2 ib) {
  if constexpr (n < 12) {
    S[n] = s_local<n>(xw, a, ib);
    fill_window<n + 1>(xw, S, a, ib);
  }
}

// sequence-edge threads: the 2x-rate index is clamped (replicate padding of the activated signal)
template <typename TIn, typename TOut>
__device__ __noinline__ void edge_path(const TIn* srow, int CT, float2 a, float2 ib, TOut* orow, long long ostride,
                                       int valid, int n_lo, int n_hi) {
  for (int i = 0; i < valid; ++i) {
    float2 acc = make_float2(0.f, 0.f);
    for (int j = 0; j < 12; ++j) {
      int n = 2 * i + j;
      n = n < n_lo ? n_lo : (n > n_hi ? n_hi : n);
      const int base = (n & 1) ? (n - 1) / 2 : n / 2;
      const int f0 = (n & 1) ? 11 : 10;
      float2 u = make_float2(0.f, 0.f);
      for (int k = 0; k < 6; ++k) {
        const float2 xv = Ld2<TIn>::ld(srow + (long long)(base + k) * CT);
        u.x = fmaf(c_fup[f0 - 2 * k], xv.x, u.x);
        u.y = fmaf(c_fup[f0 - 2 * k], xv.y, u.y);
      }
      acc.x = fmaf(c_fdn[j], snake1(u.x, a.x, ib.x), acc.x);
      acc.y = fmaf(c_fdn[j], snake1(u.y, a.y, ib.y), acc.y);
    }
    St2<TOut>::st(orow + (long long)i * ostride, acc);
  }
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kActThreads) act1d_kernel(const TIn* __restrict__ x, TOut* __restrict__ out,
                                                            const float* __restrict__ alpha,
                                                            const float* __restrict__ beta, int T, int C, int CT,
                                                            int tiles_c, int logscale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  TIn* tile = reinterpret_cast<TIn*>(smem_raw);
  pdl_launch_dependents();
  pdl_wait();
  const int CP = CT >> 1;                 // channel pairs per tile
  const int groups = kActThreads / CP;    // time groups per block
  const int TB = groups * kTT;
  const int tile_c = blockIdx.x % tiles_c, tile_t = blockIdx.x / tiles_c;
  const int b = blockIdx.y;
  const int c0 = tile_c * CT, tb0 = tile_t * TB;
  const TIn* xb = x + (long long)b * T * C;

  // stage (TB + 10) x CT with 16-byte vectors; row index clamped to [0, T-1]
  constexpr int kVecElems = 16 / sizeof(TIn);
  const int vpr = CT / kVecElems;
  const int rows = TB + 10;
  for (int idx = threadIdx.x; idx < rows * vpr; idx += kActThreads) {
    const int r = idx / vpr, v = idx - r * vpr;
    int gr = tb0 - 5 + r;
    gr = gr < 0 ? 0 : (gr > T - 1 ? T - 1 : gr);
    const uint4 val = *reinterpret_cast<const uint4*>(xb + (long long)gr * C + c0 + v * kVecElems);
    *reinterpret_cast<uint4*>(tile + (long long)r * CT + v * kVecElems) = val;
  }
  __syncthreads();

  const int cp = threadIdx.x % CP, tg = threadIdx.x / CP;
  const int t0 = tb0 + tg * kTT;
  if (t0 >= T) return;
  const int c = c0 + 2 * cp;
  float2 a = make_float2(alpha[c], alpha[c + 1]);
  float2 bb = beta ? make_float2(beta[c], beta[c + 1]) : a;
  if (logscale) {
    a = make_float2(__expf(a.x), __expf(a.y));
    bb = make_float2(__expf(bb.x), __expf(bb.y));
  }
  const float2 ib = make_float2(1.f / (bb.x + 1e-9f), 1.f / (bb.y + 1e-9f));
  const TIn* srow = tile + (long long)(tg * kTT) * CT + 2 * cp;  // window row 0 <-> x[t0 - 5]
  TOut* orow = out + ((long long)b * T + t0) * C + c;
  const int valid = min(kTT, T - t0);
  const int n_lo = max(0, 5 - 2 * t0);
  const int n_hi = min(2 * kTT + 9, 2 * (T - t0) + 4);
  if (n_lo > 0 || n_hi < 2 * kTT + 9) {
    edge_path<TIn, TOut>(srow, CT, a, ib, orow, C, valid, n_lo, n_hi);
    return;
  }
  float2 xw[kTT + 10];
#pragma unroll
  for (int j = 0; j < kTT + 10; ++j) xw[j] = Ld2<TIn>::ld(srow + (long long)j * CT);
  float2 S[12];
  fill_window<0>(xw, S, a, ib);
  slide<0, TOut>(xw, S, a, ib, orow, C, valid);
}

static bool g_filter_set = false;

}  // namespace ma3

using namespace ma3;

extern "C" {

// 12 taps of the Kaiser-sinc low-pass (filter.py:28-57 with cutoff 0.25, half-width 0.3); set once per process.
int ma3_act1d_set_filter(const float* taps12, void* stream) {
  MA3_REQUIRE(taps12, "act1d_set_filter: null taps");
  float up[12];
  for (int i = 0; i < 12; ++i) up[i] = 2.f * taps12[i];
  cudaError_t e = cudaMemcpyToSymbolAsync(c_fdn, taps12, sizeof(float) * 12, 0, cudaMemcpyHostToDevice,
                                          reinterpret_cast<cudaStream_t>(stream));
  if (e == cudaSuccess)
    e = cudaMemcpyToSymbolAsync(c_fup, up, sizeof(float) * 12, 0, cudaMemcpyHostToDevice,
                                reinterpret_cast<cudaStream_t>(stream));
  if (e == cudaSuccess) e = cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) MA3_FAIL((int)e, "act1d_set_filter: %s", cudaGetErrorString(e));
  g_filter_set = true;
  return 0;
}

// x [B, T, C] (f32 / f16 / bf16) -> out [B, T, C] (f16 / bf16); alpha, beta [C] fp32 (beta NULL = Snake).
int ma3_act1d(const void* x, int in_dtype, void* out, int out_dtype, const float* alpha, const float* beta, int B,
              int T, int C, int logscale, void* stream) {
  MA3_REQUIRE(g_filter_set, "act1d: call ma3_act1d_set_filter first");
  MA3_REQUIRE(x && out && alpha && B > 0 && T > 0, "act1d: null pointer or empty");
  MA3_REQUIRE(C % 16 == 0, "act1d: C=%d must be a multiple of 16 (pad channels)", C);
  MA3_REQUIRE(aligned16(x) && aligned16(out), "act1d: pointers must be 16-byte aligned");
  const int CT = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16);
  const int groups = kActThreads / (CT / 2);
  const int TB = groups * kTT;
  const int tiles_c = C / CT, tiles_t = (T + TB - 1) / TB;
  dim3 grid((unsigned)(tiles_c * tiles_t), (unsigned)B);
  const size_t smem = (size_t)(TB + 10) * CT * dtype_bytes(in_dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t le = cudaSuccess;
#define ACT_CASE(TI, TO) \
  le = launch_pdl(act1d_kernel<TI, TO>, grid, dim3(kActThreads), smem, st, 1, (const TI*)x, (TO*)out, alpha, beta, T, C, CT, tiles_c, logscale)
  if (in_dtype == MA3_F16 && out_dtype == MA3_F16) ACT_CASE(__half, __half);
  else if (in_dtype == MA3_F32 && out_dtype == MA3_F16) ACT_CASE(float, __half);
  else if (in_dtype == MA3_BF16 && out_dtype == MA3_BF16) ACT_CASE(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == MA3_F32 && out_dtype == MA3_BF16) ACT_CASE(float, __nv_bfloat16);
  else if (in_dtype == MA3_F32 && out_dtype == MA3_F32) ACT_CASE(float, float);
  else MA3_FAIL(MA3_EINVAL, "act1d: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
#undef ACT_CASE
  if (le != cudaSuccess) MA3_FAIL((int)le, "act1d launch: %s", cudaGetErrorString(le));
  MA3_LAUNCH_CHECK("act1d");
  return 0;
}

}  // extern "C"
