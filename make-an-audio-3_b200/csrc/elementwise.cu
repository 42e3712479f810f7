// HBM-bound kernels of the sampling path: adaLN-modulated RMSNorm, final layer (+ fused CFG combine and Euler
// update), latent projection, conditioning pool + LayerNorm, timestep embedding, GroupNorm+swish, row softmax,
// layout changes.  All are judged by achieved HBM GB/s; see include/ma3_b200.h for the reference lines.
#include "host_common.h"
#include "ptx.cuh"

#include <stdlib.h>
#include <type_traits>

namespace ma3 {

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float to(float v) { return v; }
  static __device__ __forceinline__ float from(float v) { return v; }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ __nv_bfloat16 to(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float from(__nv_bfloat16 v) { return __bfloat162float(v); }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ __half to(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float from(__half v) { return __half2float(v); }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = lane < nw ? red[lane] : 0.f;
  return warp_sum(t);
}

// ---------------------------------------------------------------------------------------- rmsnorm + modulate
// out[m, :] = rms(x[m, :]) * w * (1 + scale[sample]) + shift[sample]     (one warp per row, row cached in registers)
constexpr int kMaxVecPerLane = 16;  // D <= 2048

// grid = (row chunks of kRmsRows within a sample, samples).  The block first combines the per-sample modulation with
// the norm weight into shared memory (a = w (1 + scale), b = shift), so the row loop touches global memory only for x
// and the output; each warp owns two rows and issues both rows' loads before reducing (memory-level parallelism).
constexpr int kRmsRows = 16;

template <typename TOut, int kVec>
__global__ void __launch_bounds__(256) rmsnorm_modulate_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ mod, long long mod_ld,
                                                               int shift_off, int scale_off, int rows_per_sample,
                                                               TOut* __restrict__ out, int M, int D, float eps) {
  extern __shared__ __align__(16) float ab[];  // a[D] | b[D]
  pdl_launch_dependents();
  pdl_wait();
  float* sa = ab;
  float* sb = ab + D;
  const int sample = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = D >> 2;
  const int row_in_sample = blockIdx.x * kRmsRows + warp * 2;
  const long long row0 = (long long)sample * rows_per_sample + row_in_sample;
  const bool ok0 = row_in_sample < rows_per_sample && row0 < M;
  const bool ok1 = row_in_sample + 1 < rows_per_sample && row0 + 1 < M;
  // issue this warp's row loads first: they are the long-latency (HBM) part
  float4 v0[kVec], v1[kVec];
  const float4* x0 = reinterpret_cast<const float4*>(x + row0 * D);
  const float4* x1 = reinterpret_cast<const float4*>(x + (row0 + 1) * D);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int j = i * 32 + lane;
    v0[i] = (ok0 && j < nvec) ? x0[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    v1[i] = (ok1 && j < nvec) ? x1[j] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float* sc = mod ? mod + (long long)sample * mod_ld + scale_off : nullptr;
  const float* sh = mod ? mod + (long long)sample * mod_ld + shift_off : nullptr;
  for (int j = threadIdx.x; j < nvec; j += blockDim.x) {
    float4 a = w ? reinterpret_cast<const float4*>(w)[j] : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mod) {
      const float4 s1 = reinterpret_cast<const float4*>(sc)[j];
      b = reinterpret_cast<const float4*>(sh)[j];
      a.x *= 1.f + s1.x; a.y *= 1.f + s1.y; a.z *= 1.f + s1.z; a.w *= 1.f + s1.w;
    }
    reinterpret_cast<float4*>(sa)[j] = a;
    reinterpret_cast<float4*>(sb)[j] = b;
  }
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    s0 += v0[i].x * v0[i].x + v0[i].y * v0[i].y + v0[i].z * v0[i].z + v0[i].w * v0[i].w;
    s1 += v1[i].x * v1[i].x + v1[i].y * v1[i].y + v1[i].z * v1[i].z + v1[i].w * v1[i].w;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  const float r0 = rsqrtf(s0 / (float)D + eps), r1 = rsqrtf(s1 / (float)D + eps);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int j = i * 32 + lane;
    if (j < nvec) {
      const float4 a = reinterpret_cast<const float4*>(sa)[j];
      const float4 b = reinterpret_cast<const float4*>(sb)[j];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (!(rr ? ok1 : ok0)) continue;
        const float4 xv = rr ? v1[i] : v0[i];
        const float r = rr ? r1 : r0;
        float4 o;
        o.x = fmaf(xv.x * r, a.x, b.x); o.y = fmaf(xv.y * r, a.y, b.y);
        o.z = fmaf(xv.z * r, a.z, b.z); o.w = fmaf(xv.w * r, a.w, b.w);
        TOut* op = out + (row0 + rr) * D + j * 4;
        if constexpr (sizeof(TOut) == 2) {
          uint2 u;
          if constexpr (std::is_same<TOut, __nv_bfloat16>::value) {
            u.x = pack_bf16(o.x, o.y); u.y = pack_bf16(o.z, o.w);
          } else {
            u.x = pack_f16(o.x, o.y); u.y = pack_f16(o.z, o.w);
          }
          *reinterpret_cast<uint2*>(op) = u;
        } else {
          *reinterpret_cast<float4*>(op) = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------- final layer
// LayerNorm(no affine, eps) -> modulate -> Linear(D -> Cout) ; output transposed to [N, Cout, T].
// kCfg: rows come in (uncond n, cond n+B) pairs; the guided velocity v = vu + s (vc - vu) is formed in-kernel and
// the Euler update x <- x + dt v applied (cfm1_audio.py:154-161 + torchdyn Euler), so no velocity round trip.
constexpr int kMaxCout = 32;

// One warp normalises a row (kept in registers) and takes its Cout dot products against W staged in shared memory.
__device__ __forceinline__ void final_row(const float* __restrict__ hrow, const float* __restrict__ sc,
                                          const float* __restrict__ sh, const float* __restrict__ sW,
                                          const float* __restrict__ bias, int D, int Cout, float eps, int lane,
                                          float& result) {
  // returns in `result` the output channel `lane` (valid for lane < Cout)
  constexpr int kMaxPerLane = 64;  // D <= 2048
  float xn[kMaxPerLane];
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int j = i * 32 + lane;
    xn[i] = j < D ? hrow[j] : 0.f;
    s += xn[i];
    s2 += xn[i] * xn[i];
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  const float mean = s / (float)D;
  const float rstd = rsqrtf(fmaxf(s2 / (float)D - mean * mean, 0.f) + eps);
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    const int j = i * 32 + lane;
    xn[i] = j < D ? (xn[i] - mean) * rstd * (1.f + sc[j]) + sh[j] : 0.f;
  }
  float mine = 0.f;
  for (int c = 0; c < Cout; ++c) {
    const float* wr = sW + (long long)c * D;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int j = i * 32 + lane;
      if (j < D) acc = fmaf(xn[i], wr[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == c) mine = acc + bias[c];
  }
  result = mine;
}

// Fast path for D = 128 * kV4: one warp takes TWO rows at a time (the uncond / cond pair in CFG mode), keeps both
// normalised rows in registers as float4 pieces and walks W once for both: 2 * Cout independent accumulators per lane
// (instruction-level parallelism instead of one dependent FMA chain per channel), then one "transposing" butterfly per
// row that leaves lane c holding output channel c (31 shuffles for up to 32 channels instead of 5 per channel).
template <int kV4>
__device__ __forceinline__ void norm_row4(const float* __restrict__ hrow, const float* __restrict__ sc,
                                          const float* __restrict__ sh, int D, float eps, int lane, float4 (&xn)[kV4]) {
  float s = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    xn[i] = reinterpret_cast<const float4*>(hrow)[i * 32 + lane];
    s += (xn[i].x + xn[i].y) + (xn[i].z + xn[i].w);
    s2 += (xn[i].x * xn[i].x + xn[i].y * xn[i].y) + (xn[i].z * xn[i].z + xn[i].w * xn[i].w);
  }
  s = warp_sum(s);
  s2 = warp_sum(s2);
  const float mean = s / (float)D;
  const float rstd = rsqrtf(fmaxf(s2 / (float)D - mean * mean, 0.f) + eps);
#pragma unroll
  for (int i = 0; i < kV4; ++i) {
    const float4 a = reinterpret_cast<const float4*>(sc)[i * 32 + lane];
    const float4 b = reinterpret_cast<const float4*>(sh)[i * 32 + lane];
    xn[i].x = (xn[i].x - mean) * rstd * (1.f + a.x) + b.x;
    xn[i].y = (xn[i].y - mean) * rstd * (1.f + a.y) + b.y;
    xn[i].z = (xn[i].z - mean) * rstd * (1.f + a.z) + b.z;
    xn[i].w = (xn[i].w - mean) * rstd * (1.f + a.w) + b.w;
  }
}

// v[0..32) per lane -> lane l returns sum over the warp of v[l]
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = up ? v[k] : v[k + off];
      const float keep = up ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int kV4>
__device__ __forceinline__ void final_pair(const float* __restrict__ h0, const float* __restrict__ h1,
                                           const float* __restrict__ mod0, const float* __restrict__ mod1, int shift_off,
                                           int scale_off, const float* __restrict__ sW, const float* __restrict__ bias,
                                           int D, int Cout, float eps, int lane, float& r0, float& r1) {
  float4 x0[kV4], x1[kV4];
  norm_row4<kV4>(h0, mod0 + scale_off, mod0 + shift_off, D, eps, lane, x0);
  norm_row4<kV4>(h1, mod1 + scale_off, mod1 + shift_off, D, eps, lane, x1);
  float a0[32], a1[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
#pragma unroll
  for (int c = 0; c < kMaxCout; ++c) {
    if (c < Cout) {
      const float4* wr = reinterpret_cast<const float4*>(sW + (long long)c * D);
#pragma unroll
      for (int i = 0; i < kV4; ++i) {
        const float4 w = wr[i * 32 + lane];
        a0[c] = fmaf(x0[i].x, w.x, fmaf(x0[i].y, w.y, fmaf(x0[i].z, w.z, fmaf(x0[i].w, w.w, a0[c]))));
        a1[c] = fmaf(x1[i].x, w.x, fmaf(x1[i].y, w.y, fmaf(x1[i].z, w.z, fmaf(x1[i].w, w.w, a1[c]))));
      }
    }
  }
  const float b = lane < Cout ? bias[lane] : 0.f;
  r0 = warp_transpose_sum32(a0, lane) + b;
  r1 = warp_transpose_sum32(a1, lane) + b;
}

template <bool kCfg, int kV4>
__global__ void __launch_bounds__(256) final_layer_fast_kernel(const float* __restrict__ h, const float* __restrict__ mod,
                                                               long long mod_ld, int shift_off, int scale_off,
                                                               const float* __restrict__ W, const float* __restrict__ bias,
                                                               int N, int T, int D, int Cout, float eps,
                                                               float* __restrict__ v_out, const float* __restrict__ x_in,
                                                               float* __restrict__ x_out, float dt, float guidance) {
  extern __shared__ __align__(16) float sW[];  // [Cout, D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < (Cout * D) >> 2; i += blockDim.x)
    reinterpret_cast<float4*>(sW)[i] = reinterpret_cast<const float4*>(W)[i];
  __syncthreads();
  if constexpr (!kCfg) {
    const int pairs = (N * T + 1) >> 1;
    for (int pid = blockIdx.x * nwarp + warp; pid < pairs; pid += gridDim.x * nwarp) {
      const int w0 = 2 * pid, w1 = min(2 * pid + 1, N * T - 1);
      const int n0 = w0 / T, n1 = w1 / T;
      float r0, r1;
      final_pair<kV4>(h + (long long)w0 * D, h + (long long)w1 * D, mod + (long long)n0 * mod_ld, mod + (long long)n1 * mod_ld,
                      shift_off, scale_off, sW, bias, D, Cout, eps, lane, r0, r1);
      if (lane < Cout) {
        v_out[((long long)n0 * Cout + lane) * T + (w0 - n0 * T)] = r0;
        if (2 * pid + 1 < N * T) v_out[((long long)n1 * Cout + lane) * T + (w1 - n1 * T)] = r1;
      }
    }
  } else {
    const int B = N >> 1;
    for (int wid = blockIdx.x * nwarp + warp; wid < B * T; wid += gridDim.x * nwarp) {
      const int b = wid / T, t = wid - b * T;
      const int nc = b + B;
      float ru, rc;
      final_pair<kV4>(h + (long long)wid * D, h + ((long long)nc * T + t) * D, mod + (long long)b * mod_ld,
                      mod + (long long)nc * mod_ld, shift_off, scale_off, sW, bias, D, Cout, eps, lane, ru, rc);
      if (lane < Cout) {
        const float vg = ru + guidance * (rc - ru);
        const long long idx = ((long long)b * Cout + lane) * T + t;
        if (v_out) v_out[idx] = vg;
        if (x_out) x_out[idx] = x_in[idx] + dt * vg;
      }
    }
  }
}

template <bool kCfg>
__global__ void __launch_bounds__(256) final_layer_kernel(const float* __restrict__ h, const float* __restrict__ mod,
                                                          long long mod_ld, int shift_off, int scale_off,
                                                          const float* __restrict__ W, const float* __restrict__ bias,
                                                          int N, int T, int D, int Cout, float eps,
                                                          float* __restrict__ v_out,  // [N or B, Cout, T] (nullable if kCfg)
                                                          const float* __restrict__ x_in, float* __restrict__ x_out,
                                                          float dt, float guidance) {
  extern __shared__ __align__(16) float sW[];  // [Cout, D]: read once per block instead of once per row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < (Cout * D) >> 2; i += blockDim.x)
    reinterpret_cast<float4*>(sW)[i] = reinterpret_cast<const float4*>(W)[i];
  __syncthreads();
  if constexpr (!kCfg) {
    for (int wid = blockIdx.x * nwarp + warp; wid < N * T; wid += gridDim.x * nwarp) {
      const int n = wid / T, t = wid - n * T;
      float r;
      final_row(h + (long long)wid * D, mod + (long long)n * mod_ld + scale_off, mod + (long long)n * mod_ld + shift_off,
                sW, bias, D, Cout, eps, lane, r);
      if (lane < Cout) v_out[((long long)n * Cout + lane) * T + t] = r;
    }
  } else {
    const int B = N >> 1;
    for (int wid = blockIdx.x * nwarp + warp; wid < B * T; wid += gridDim.x * nwarp) {
      const int b = wid / T, t = wid - b * T;
      float ru, rc;
      final_row(h + (long long)wid * D, mod + (long long)b * mod_ld + scale_off, mod + (long long)b * mod_ld + shift_off,
                sW, bias, D, Cout, eps, lane, ru);
      const int nc = b + B;
      final_row(h + ((long long)nc * T + t) * D, mod + (long long)nc * mod_ld + scale_off,
                mod + (long long)nc * mod_ld + shift_off, sW, bias, D, Cout, eps, lane, rc);
      if (lane < Cout) {
        const float vg = ru + guidance * (rc - ru);
        const long long idx = ((long long)b * Cout + lane) * T + t;
        if (v_out) v_out[idx] = vg;
        if (x_out) x_out[idx] = x_in[idx] + dt * vg;
      }
    }
  }
}

// x <- x + dt * (vu + s (vc - vu)) as a stand-alone elementwise kernel (drop-in forward() path).
__global__ void cfg_euler_kernel(const float* __restrict__ v, const float* __restrict__ x, float* __restrict__ out,
                                 long long half_elems, float dt, float guidance, int cfg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half_elems) return;
  float vg;
  if (cfg) {
    const float vu = v[i], vc = v[i + half_elems];
    vg = vu + guidance * (vc - vu);
  } else {
    vg = v[i];
  }
  out[i] = x[i] + dt * vg;
}

// ---------------------------------------------------------------------------------------- proj_in
// h[n*T + t, d] = sum_c x[n, c, t] * Wt[c, d] + b[d]     (flag_large_dit.py:186-187); x row n % x_batch (CFG halves
// share x).
__global__ void __launch_bounds__(512) proj_in_kernel(const float* __restrict__ x, const float* __restrict__ Wt,
                                                      const float* __restrict__ b, float* __restrict__ h, int N, int xB,
                                                      int C, int T, int D) {
  // one block per output row; Wt is the transposed weight [C, D] so every thread streams contiguous float4
  __shared__ float sx[64];
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  const int t = row % T, n = (row / T) % xB;
  if (threadIdx.x < C) sx[threadIdx.x] = x[((long long)n * C + threadIdx.x) * T + t];
  __syncthreads();
  const int nvec = D >> 2;
  for (int j = threadIdx.x; j < nvec; j += blockDim.x) {
    float4 acc = reinterpret_cast<const float4*>(b)[j];
    for (int c = 0; c < C; ++c) {
      const float4 w = reinterpret_cast<const float4*>(Wt + (long long)c * D)[j];
      const float xv = sx[c];
      acc.x = fmaf(xv, w.x, acc.x); acc.y = fmaf(xv, w.y, acc.y); acc.z = fmaf(xv, w.z, acc.z); acc.w = fmaf(xv, w.w, acc.w);
    }
    reinterpret_cast<float4*>(h + (long long)row * D)[j] = acc;
  }
}

// ---------------------------------------------------------------------------------------- timestep embedding
// out[m, 0:half] = cos(t f_i), out[m, half:] = sin(t f_i), f_i = exp(-ln(max_period) i / half)
// (flag_large_dit_moe.py:110-127).  Accurate sincosf: arguments reach 1000 rad.
template <typename TOut>
__global__ void timestep_embed_kernel(const long long* __restrict__ t, TOut* __restrict__ out, int M, int dim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (i >= M * half) return;
  const int m = i / half, j = i - m * half;
  const float f = expf(-logf(10000.f) * (float)j / (float)half);
  const float a = (float)t[m] * f;
  float s, c;
  sincosf(a, &s, &c);
  out[(long long)m * dim + j] = Cvt<TOut>::to(c);
  out[(long long)m * dim + half + j] = Cvt<TOut>::to(s);
}

// ---------------------------------------------------------------------------------------- pool + LayerNorm
// out[n, :] = LayerNorm_affine(mean_L(ctx[n, :, :]))    (flag_large_dit.py:193-198); one block per sample.
// TIn: fp32 context, or the 16-bit embedded context of the video variant.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) pool_layernorm_kernel(const TIn* __restrict__ ctx, const float* __restrict__ w,
                                                             const float* __restrict__ b, TOut* __restrict__ out,
                                                             int L, int Cd, float eps) {
  extern __shared__ float pooled[];
  __shared__ float red[32];
  const int n = blockIdx.x;
  const TIn* base = ctx + (long long)n * L * Cd;
  float s = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < Cd; c += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc += Cvt<TIn>::from(base[(long long)l * Cd + c]);
    acc /= (float)L;
    pooled[c] = acc;
    s += acc;
  }
  const float mean = block_sum(s, red) / (float)Cd;
  for (int c = threadIdx.x; c < Cd; c += blockDim.x) {
    const float d = pooled[c] - mean;
    s2 += d * d;
  }
  const float var = block_sum(s2, red) / (float)Cd;
  const float rstd = rsqrtf(var + eps);
  for (int c = threadIdx.x; c < Cd; c += blockDim.x)
    out[(long long)n * Cd + c] = Cvt<TOut>::to((pooled[c] - mean) * rstd * w[c] + b[c]);
}

// Row LayerNorm with affine (the ConditionEmbedder's trailing nn.LayerNorm, flag_large_dit_moe.py:151-162).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const TIn* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, TOut* __restrict__ out,
                                                             int M, int D, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const TIn* xr = x + (long long)row * D;
  float s = 0.f;
  for (int j = lane; j < D; j += 32) s += Cvt<TIn>::from(xr[j]);
  const float mean = warp_sum(s) / (float)D;
  float s2 = 0.f;
  for (int j = lane; j < D; j += 32) {
    const float d = Cvt<TIn>::from(xr[j]) - mean;
    s2 += d * d;
  }
  const float rstd = rsqrtf(warp_sum(s2) / (float)D + eps);
  for (int j = lane; j < D; j += 32)
    out[(long long)row * D + j] = Cvt<TOut>::to((Cvt<TIn>::from(xr[j]) - mean) * rstd * w[j] + b[j]);
}

// ---------------------------------------------------------------------------------------- GroupNorm (+ swish)
// channels-last x[B, T, C]; one block per (b, group).  autoencoder1d.py:169-175 (eps 1e-6, affine, swish = x*sigmoid).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(512) groupnorm_swish_kernel(const TIn* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ b, TOut* __restrict__ out,
                                                              int T, int C, int groups, float eps, int swish) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const int cg = C / groups;
  const int bidx = blockIdx.x / groups, g = blockIdx.x - bidx * groups;
  const TIn* xb = x + (long long)bidx * T * C + g * cg;
  TOut* ob = out + (long long)bidx * T * C + g * cg;
  const int n = T * cg;
  float s = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int t = i / cg, c = i - t * cg;
    const float v = Cvt<TIn>::from(xb[(long long)t * C + c]);
    s += v;
    s2 += v * v;
  }
  s = block_sum(s, red);
  s2 = block_sum(s2, red);
  const float mean = s / (float)n;
  const float rstd = rsqrtf(fmaxf(s2 / (float)n - mean * mean, 0.f) + eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int t = i / cg, c = i - t * cg;
    float v = (Cvt<TIn>::from(xb[(long long)t * C + c]) - mean) * rstd * w[g * cg + c] + b[g * cg + c];
    if (swish) v = v / (1.f + __expf(-v));
    ob[(long long)t * C + c] = Cvt<TOut>::to(v);
  }
}

// ---------------------------------------------------------------------------------------- row softmax
// P[z, i, 0:n] = softmax(scale * S[z, i, 0:n]); P[z, i, n:ld_out] = 0     (autoencoder1d.py:265-270)
template <typename TOut>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, TOut* __restrict__ P, int rows,
                                                           int n, long long ld_in, long long ld_out, float scale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* s = S + (long long)row * ld_in;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, s[j] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) sum += __expf(s[j] * scale - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  TOut* p = P + (long long)row * ld_out;
  for (int j = lane; j < ld_out; j += 32) p[j] = Cvt<TOut>::to(j < n ? __expf(s[j] * scale - mx) * inv : 0.f);
}

// ---------------------------------------------------------------------------------------- layout changes
// [B, C, T] fp32 -> [B, T, Cp] 16-bit, scaled, zero-padded channels (latent z / mel in)
template <typename TOut>
__global__ void nct_to_ntc_kernel(const float* __restrict__ x, TOut* __restrict__ out, int B, int C, int T, int Cp,
                                  float scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T * Cp) return;
  const int c = (int)(i % Cp);
  const long long r = i / Cp;
  const int t = (int)(r % T), b = (int)(r / T);
  out[i] = Cvt<TOut>::to(c < C ? x[((long long)b * C + c) * T + t] * scale : 0.f);
}

// [B, T, ld] (first C channels) -> [B, C, T] fp32 (mel out)
template <typename TIn>
__global__ void ntc_to_nct_kernel(const TIn* __restrict__ x, float* __restrict__ out, int B, int C, int T, long long ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * T) return;
  const int t = (int)(i % T);
  const long long r = i / T;
  const int c = (int)(r % C), b = (int)(r / C);
  out[i] = Cvt<TIn>::from(x[((long long)b * T + t) * ld + c]);
}

// nearest x2 along T on channels-last 16-bit data (autoencoder1d.py:291-292), 16-byte vectors
__global__ void upsample2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, long long rows, int vec_per_row) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 2 * vec_per_row) return;
  const int v = (int)(i % vec_per_row);
  const long long orow = i / vec_per_row;
  out[i] = x[(orow >> 1) * vec_per_row + v];
}

template <typename TIn, typename TOut>
__global__ void cast_kernel(const TIn* __restrict__ x, TOut* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = Cvt<TOut>::to(Cvt<TIn>::from(x[i]));
}

}  // namespace ma3

using namespace ma3;

#define ST(s) reinterpret_cast<cudaStream_t>(s)

static inline unsigned nblk(long long n, int per) { return (unsigned)((n + per - 1) / per); }

extern "C" {

int ma3_rmsnorm_modulate(const float* x, const float* w, const float* mod, int64_t mod_ld, int shift_off,
                         int scale_off, int rows_per_sample, void* out, int out_dtype, int M, int D, float eps,
                         void* stream) {
  MA3_REQUIRE(x && out && M > 0, "rmsnorm_modulate: null pointer or empty");
  MA3_REQUIRE(D % 4 == 0 && D <= 128 * kMaxVecPerLane, "rmsnorm_modulate: D=%d must be a multiple of 4 and <= 2048", D);
  MA3_REQUIRE(!mod || (rows_per_sample > 0 && mod_ld % 4 == 0 && shift_off % 4 == 0 && scale_off % 4 == 0),
              "rmsnorm_modulate: modulation offsets must be multiples of 4");
  MA3_REQUIRE(aligned16(x) && aligned16(out) && (!w || aligned16(w)) && (!mod || aligned16(mod)),
              "rmsnorm_modulate: pointers must be 16-byte aligned");
  const int rps = mod ? rows_per_sample : M;   // without modulation all rows form one "sample"
  MA3_REQUIRE(!mod || M % rows_per_sample == 0, "rmsnorm_modulate: M must be samples * rows_per_sample");
  const dim3 grid((unsigned)((rps + kRmsRows - 1) / kRmsRows), (unsigned)(M / rps));
  const size_t smem = 2 * (size_t)D * sizeof(float);
  const int nv = (D / 4 + 31) / 32;  // float4 per lane
#define RMS_LAUNCH(TO, KV)                                                                                         \
  launch_pdl(rmsnorm_modulate_kernel<TO, KV>, grid, dim3(256), smem, ST(stream), 1, x, w, mod, (long long)mod_ld,  \
             shift_off, scale_off, rps, (TO*)out, M, D, eps)
#define RMS_DISPATCH(TO)                 \
  do {                                   \
    if (nv <= 6) RMS_LAUNCH(TO, 6);      \
    else if (nv <= 9) RMS_LAUNCH(TO, 9); \
    else if (nv <= 12) RMS_LAUNCH(TO, 12); \
    else RMS_LAUNCH(TO, 16);             \
  } while (0)
  if (out_dtype == MA3_BF16) RMS_DISPATCH(__nv_bfloat16);
  else if (out_dtype == MA3_F16) RMS_DISPATCH(__half);
  else RMS_DISPATCH(float);
#undef RMS_DISPATCH
#undef RMS_LAUNCH
  MA3_LAUNCH_CHECK("rmsnorm_modulate");
  return 0;
}

int ma3_final_layer(const float* h, const float* mod, int64_t mod_ld, int shift_off, int scale_off, const float* W,
                    const float* bias, int N, int T, int D, int Cout, float eps, float* v_out, void* stream) {
  MA3_REQUIRE(h && mod && W && bias && v_out, "final_layer: null pointer");
  MA3_REQUIRE(Cout <= kMaxCout && N > 0 && T > 0 && D <= 2048, "final_layer: Cout=%d must be <= 32, D <= 2048", Cout);
  MA3_REQUIRE((Cout * D) % 4 == 0 && aligned16(W), "final_layer: W must be 16-byte aligned");
  const size_t smem = (size_t)Cout * D * sizeof(float);
  {
    static DeviceOnce configured;
    if (configured.pending()) {
      cudaFuncSetAttribute(final_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      configured.mark();
    }
  }
  MA3_REQUIRE(smem <= 200 * 1024, "final_layer: Cout * D too large for shared memory");
  const unsigned fgrid = (unsigned)min((long long)num_sms(), ((long long)N * T + 7) / 8);
#define FL_FAST(KV4)                                                                                                   \
  do {                                                                                                                 \
    static DeviceOnce cfgd;                                                                                                \
    if (cfgd.pending()) {                                                                                                       \
      cudaFuncSetAttribute(final_layer_fast_kernel<false, KV4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      cfgd.mark();                                                                                                     \
    }                                                                                                                  \
    launch_pdl(final_layer_fast_kernel<false, KV4>, dim3(fgrid), dim3(256), smem, ST(stream), 1, h, mod, (long long)mod_ld, \
               shift_off, scale_off, W, bias, N, T, D, Cout, eps, v_out, (const float*)nullptr, (float*)nullptr, 0.f, 0.f); \
  } while (0)
  const bool fast = aligned16(h) && aligned16(mod) && mod_ld % 4 == 0 && shift_off % 4 == 0 && scale_off % 4 == 0;
  if (fast && D == 768) FL_FAST(6);
  else if (fast && D == 1152) FL_FAST(9);
  else if (fast && D == 1536) FL_FAST(12);
  else
  launch_pdl(final_layer_kernel<false>, dim3(fgrid), dim3(256), smem, ST(stream), 1, h, mod,
             (long long)mod_ld, shift_off, scale_off, W, bias, N, T, D, Cout, eps, v_out, (const float*)nullptr,
             (float*)nullptr, 0.f, 0.f);
#undef FL_FAST
  MA3_LAUNCH_CHECK("final_layer");
  return 0;
}

int ma3_final_layer_cfg_euler(const float* h, const float* mod, int64_t mod_ld, int shift_off, int scale_off,
                              const float* W, const float* bias, int N, int T, int D, int Cout, float eps,
                              float guidance, float dt, const float* x_in, float* x_out, float* v_out,
                              void* stream) {
  MA3_REQUIRE(h && mod && W && bias && x_in && x_out, "final_layer_cfg_euler: null pointer");
  MA3_REQUIRE(Cout <= kMaxCout && N > 0 && N % 2 == 0 && T > 0 && D <= 2048, "final_layer_cfg_euler: N must be even, Cout <= 32, D <= 2048");
  MA3_REQUIRE((Cout * D) % 4 == 0 && aligned16(W), "final_layer_cfg_euler: W must be 16-byte aligned");
  const size_t smem = (size_t)Cout * D * sizeof(float);
  {
    static DeviceOnce configured;
    if (configured.pending()) {
      cudaFuncSetAttribute(final_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      configured.mark();
    }
  }
  MA3_REQUIRE(smem <= 200 * 1024, "final_layer_cfg_euler: Cout * D too large for shared memory");
  const unsigned fgrid = (unsigned)min((long long)num_sms(), ((long long)(N / 2) * T + 7) / 8);
#define FL_FAST(KV4)                                                                                                   \
  do {                                                                                                                 \
    static DeviceOnce cfgd;                                                                                                \
    if (cfgd.pending()) {                                                                                                       \
      cudaFuncSetAttribute(final_layer_fast_kernel<true, KV4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      cfgd.mark();                                                                                                     \
    }                                                                                                                  \
    launch_pdl(final_layer_fast_kernel<true, KV4>, dim3(fgrid), dim3(256), smem, ST(stream), 1, h, mod, (long long)mod_ld, \
               shift_off, scale_off, W, bias, N, T, D, Cout, eps, v_out, x_in, x_out, dt, guidance);                   \
  } while (0)
  const bool fast = aligned16(h) && aligned16(mod) && mod_ld % 4 == 0 && shift_off % 4 == 0 && scale_off % 4 == 0;
  if (fast && D == 768) FL_FAST(6);
  else if (fast && D == 1152) FL_FAST(9);
  else if (fast && D == 1536) FL_FAST(12);
  else
  launch_pdl(final_layer_kernel<true>, dim3(fgrid), dim3(256), smem, ST(stream), 1, h, mod,
             (long long)mod_ld, shift_off, scale_off, W, bias, N, T, D, Cout, eps, v_out, x_in, x_out, dt, guidance);
#undef FL_FAST
  MA3_LAUNCH_CHECK("final_layer_cfg_euler");
  return 0;
}

int ma3_cfg_euler_update(const float* v, const float* x, float* out, int64_t elems, float dt, float guidance, int cfg,
                         void* stream) {
  MA3_REQUIRE(v && x && out && elems > 0, "cfg_euler_update: null pointer or empty");
  cfg_euler_kernel<<<nblk(elems, 256), 256, 0, ST(stream)>>>(v, x, out, elems, dt, guidance, cfg);
  MA3_LAUNCH_CHECK("cfg_euler_update");
  return 0;
}

int ma3_proj_in(const float* x, const float* W, const float* b, float* h, int N, int x_batch, int C, int T, int D,
                void* stream) {
  MA3_REQUIRE(x && W && b && h && N > 0 && x_batch > 0, "proj_in: null pointer or empty");
  MA3_REQUIRE(C <= 64 && D % 4 == 0 && aligned16(W) && aligned16(b) && aligned16(h), "proj_in: C <= 64, D %% 4 == 0, aligned");
  int threads = ((D / 4 + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  if (threads < 64) threads = 64;
  launch_pdl(proj_in_kernel, dim3((unsigned)(N * T)), dim3(threads), 0, ST(stream), 1, x, W, b, h, N, x_batch, C, T, D);
  MA3_LAUNCH_CHECK("proj_in");
  return 0;
}

int ma3_timestep_embed(const int64_t* t, void* out, int out_dtype, int M, int dim, void* stream) {
  MA3_REQUIRE(t && out && M > 0 && dim % 2 == 0, "timestep_embed: bad arguments");
  const unsigned grid = nblk((long long)M * (dim / 2), 128);
  if (out_dtype == MA3_BF16)
    timestep_embed_kernel<__nv_bfloat16><<<grid, 128, 0, ST(stream)>>>((const long long*)t, (__nv_bfloat16*)out, M, dim);
  else if (out_dtype == MA3_F32)
    timestep_embed_kernel<float><<<grid, 128, 0, ST(stream)>>>((const long long*)t, (float*)out, M, dim);
  else
    MA3_FAIL(MA3_EINVAL, "timestep_embed: out dtype must be bf16 or f32");
  MA3_LAUNCH_CHECK("timestep_embed");
  return 0;
}

int ma3_pool_layernorm(const void* ctx, int in_dtype, const float* w, const float* b, void* out, int out_dtype, int N,
                       int L, int Cd, float eps, void* stream) {
  MA3_REQUIRE(ctx && w && b && out && N > 0 && L > 0 && Cd > 0, "pool_layernorm: bad arguments");
  MA3_REQUIRE((out_dtype == MA3_BF16 || (out_dtype == MA3_F32 && in_dtype == MA3_F32)) &&
                  (in_dtype == MA3_F32 || in_dtype == MA3_BF16),
              "pool_layernorm: in f32/bf16, out bf16 (or f32 from f32)");
  const size_t smem = (size_t)Cd * sizeof(float);
  if (in_dtype == MA3_F32 && out_dtype == MA3_F32)
    pool_layernorm_kernel<float, float><<<N, 256, smem, ST(stream)>>>((const float*)ctx, w, b, (float*)out, L, Cd, eps);
  else if (in_dtype == MA3_F32)
    pool_layernorm_kernel<float, __nv_bfloat16><<<N, 256, smem, ST(stream)>>>((const float*)ctx, w, b,
                                                                               (__nv_bfloat16*)out, L, Cd, eps);
  else
    pool_layernorm_kernel<__nv_bfloat16, __nv_bfloat16><<<N, 256, smem, ST(stream)>>>(
        (const __nv_bfloat16*)ctx, w, b, (__nv_bfloat16*)out, L, Cd, eps);
  MA3_LAUNCH_CHECK("pool_layernorm");
  return 0;
}

int ma3_layernorm_rows(const void* x, int in_dtype, const float* w, const float* b, void* out, int out_dtype, int M,
                       int D, float eps, void* stream) {
  MA3_REQUIRE(x && w && b && out && M > 0, "layernorm_rows: bad arguments");
  MA3_REQUIRE(in_dtype == MA3_F32 && (out_dtype == MA3_F32 || out_dtype == MA3_BF16), "layernorm_rows: f32 in");
  if (out_dtype == MA3_F32)
    layernorm_rows_kernel<float, float><<<nblk(M, 8), 256, 0, ST(stream)>>>((const float*)x, w, b, (float*)out, M, D, eps);
  else
    layernorm_rows_kernel<float, __nv_bfloat16><<<nblk(M, 8), 256, 0, ST(stream)>>>((const float*)x, w, b,
                                                                                   (__nv_bfloat16*)out, M, D, eps);
  MA3_LAUNCH_CHECK("layernorm_rows");
  return 0;
}

int ma3_groupnorm_swish(const void* x, int in_dtype, const float* w, const float* b, void* out, int out_dtype, int B,
                        int T, int C, int groups, float eps, int swish, void* stream) {
  MA3_REQUIRE(x && w && b && out && B > 0 && T > 0 && C % groups == 0, "groupnorm_swish: bad arguments");
  const unsigned grid = (unsigned)(B * groups);
#define GN_CASE(TI, TO) \
  launch_pdl(groupnorm_swish_kernel<TI, TO>, dim3(grid), dim3(512), 0, ST(stream), 1, (const TI*)x, w, b, (TO*)out, T, C, \
             groups, eps, swish)
  if (in_dtype == MA3_BF16 && out_dtype == MA3_BF16) GN_CASE(__nv_bfloat16, __nv_bfloat16);
  else if (in_dtype == MA3_F32 && out_dtype == MA3_BF16) GN_CASE(float, __nv_bfloat16);
  else if (in_dtype == MA3_F16 && out_dtype == MA3_F16) GN_CASE(__half, __half);
  else if (in_dtype == MA3_F32 && out_dtype == MA3_F16) GN_CASE(float, __half);
  else MA3_FAIL(MA3_EINVAL, "groupnorm_swish: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
#undef GN_CASE
  MA3_LAUNCH_CHECK("groupnorm_swish");
  return 0;
}

int ma3_softmax_rows(const float* S, void* P, int out_dtype, int rows, int n, int64_t ld_in, int64_t ld_out,
                     float scale, void* stream) {
  MA3_REQUIRE(S && P && rows > 0 && n > 0 && ld_out >= n, "softmax_rows: bad arguments");
  if (out_dtype == MA3_BF16)
    softmax_rows_kernel<__nv_bfloat16><<<nblk(rows, 8), 256, 0, ST(stream)>>>(S, (__nv_bfloat16*)P, rows, n, ld_in, ld_out, scale);
  else if (out_dtype == MA3_F16)
    softmax_rows_kernel<__half><<<nblk(rows, 8), 256, 0, ST(stream)>>>(S, (__half*)P, rows, n, ld_in, ld_out, scale);
  else
    MA3_FAIL(MA3_EINVAL, "softmax_rows: 16-bit output only");
  MA3_LAUNCH_CHECK("softmax_rows");
  return 0;
}

int ma3_nct_to_ntc(const float* x, void* out, int out_dtype, int B, int C, int T, int Cp, float scale, void* stream) {
  MA3_REQUIRE(x && out && B > 0 && Cp >= C, "nct_to_ntc: bad arguments");
  const unsigned grid = nblk((long long)B * T * Cp, 256);
  if (out_dtype == MA3_BF16)
    nct_to_ntc_kernel<__nv_bfloat16><<<grid, 256, 0, ST(stream)>>>(x, (__nv_bfloat16*)out, B, C, T, Cp, scale);
  else if (out_dtype == MA3_F16)
    nct_to_ntc_kernel<__half><<<grid, 256, 0, ST(stream)>>>(x, (__half*)out, B, C, T, Cp, scale);
  else
    MA3_FAIL(MA3_EINVAL, "nct_to_ntc: 16-bit output only");
  MA3_LAUNCH_CHECK("nct_to_ntc");
  return 0;
}

int ma3_ntc_to_nct(const void* x, int in_dtype, float* out, int B, int C, int T, int64_t ld, void* stream) {
  MA3_REQUIRE(x && out && B > 0 && ld >= C, "ntc_to_nct: bad arguments");
  const unsigned grid = nblk((long long)B * C * T, 256);
  if (in_dtype == MA3_BF16)
    ntc_to_nct_kernel<__nv_bfloat16><<<grid, 256, 0, ST(stream)>>>((const __nv_bfloat16*)x, out, B, C, T, ld);
  else if (in_dtype == MA3_F16)
    ntc_to_nct_kernel<__half><<<grid, 256, 0, ST(stream)>>>((const __half*)x, out, B, C, T, ld);
  else
    ntc_to_nct_kernel<float><<<grid, 256, 0, ST(stream)>>>((const float*)x, out, B, C, T, ld);
  MA3_LAUNCH_CHECK("ntc_to_nct");
  return 0;
}

int ma3_upsample_nearest2(const void* x, void* out, int64_t rows, int C, void* stream) {
  MA3_REQUIRE(x && out && rows > 0 && C % 8 == 0 && aligned16(x) && aligned16(out), "upsample_nearest2: C %% 8, aligned");
  upsample2_kernel<<<nblk(rows * 2 * (C / 8), 256), 256, 0, ST(stream)>>>((const uint4*)x, (uint4*)out, rows, C / 8);
  MA3_LAUNCH_CHECK("upsample_nearest2");
  return 0;
}

int ma3_cast(const void* x, int in_dtype, void* out, int out_dtype, int64_t n, void* stream) {
  MA3_REQUIRE(x && out && n > 0, "cast: bad arguments");
  const unsigned grid = nblk(n, 256);
#define CAST_CASE(TI, TO) cast_kernel<TI, TO><<<grid, 256, 0, ST(stream)>>>((const TI*)x, (TO*)out, n)
  if (in_dtype == MA3_F32 && out_dtype == MA3_BF16) CAST_CASE(float, __nv_bfloat16);
  else if (in_dtype == MA3_F32 && out_dtype == MA3_F16) CAST_CASE(float, __half);
  else if (in_dtype == MA3_BF16 && out_dtype == MA3_F32) CAST_CASE(__nv_bfloat16, float);
  else if (in_dtype == MA3_F16 && out_dtype == MA3_F32) CAST_CASE(__half, float);
  else MA3_FAIL(MA3_EINVAL, "cast: unsupported dtype pair");
#undef CAST_CASE
  MA3_LAUNCH_CHECK("cast");
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------- embedding lookup
namespace ma3 {
__global__ void embed_rows_kernel(const float* __restrict__ table, long long vocab, const long long* __restrict__ ids,
                                  const float* __restrict__ pos, const float* __restrict__ type0, float* __restrict__ out,
                                  int M, int T, int D) {
  const int m = blockIdx.x;
  long long id = ids[m];
  if (id < 0 || id >= vocab) id = 0;   // clamped: the host wrapper validates the ids (they come from a tokenizer)
  const float4* src = reinterpret_cast<const float4*>(table + id * D);
  const float4* ps = pos ? reinterpret_cast<const float4*>(pos + (long long)(m % T) * D) : nullptr;
  const float4* ty = type0 ? reinterpret_cast<const float4*>(type0) : nullptr;
  float4* dst = reinterpret_cast<float4*>(out + (long long)m * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    float4 v = src[i];
    if (ps) { const float4 q = ps[i]; v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
    if (ty) { const float4 q = ty[i]; v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
    dst[i] = v;
  }
}
}  // namespace ma3

extern "C" int ma3_embed_rows(const float* table, int64_t vocab, const int64_t* ids, const float* pos, const float* type0,
                              float* out, int M, int T, int D, void* stream) {
  MA3_REQUIRE(table && ids && out && M > 0 && T > 0 && D > 0 && D % 4 == 0 && vocab > 0, "embed_rows: bad arguments");
  ma3::embed_rows_kernel<<<(unsigned)M, 128, 0, ST(stream)>>>(table, (long long)vocab, (const long long*)ids, pos, type0, out,
                                                             M, T, D);
  MA3_LAUNCH_CHECK("embed_rows");
  return 0;
}

// ---------------------------------------------------------------------------------------- adaLN input
// out[s*N + n, :] = silu(temb[s*ts_s + n*ts_n, :] + cap[n, :])  (flag_large_dit.py:200 then the SiLU of :50-51);
// sampler: one timestep row per step (ts_s = 1, ts_n = 0); drop-in forward(): one row per sample (ts_s = 0, ts_n = 1).
namespace ma3 {
template <typename TOut>
__global__ void adaln_input_kernel(const float* __restrict__ temb, const float* __restrict__ cap, TOut* __restrict__ out,
                                   int S, int N, int D, int ts_s, int ts_n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)S * N * D) return;
  const int d = (int)(i % D);
  const long long r = i / D;
  const int n = (int)(r % N), s = (int)(r / N);
  const float a = temb[(long long)(s * ts_s + n * ts_n) * D + d] + cap[(long long)n * D + d];
  out[i] = Cvt<TOut>::to(a / (1.f + __expf(-a)));
}
}  // namespace ma3

// ---------------------------------------------------------------------------------------- QK-norm + RoPE + scatter
// qk_norm=True (flag_large_dit_moe.py:199-207,345-352): q and k are LayerNorm-ed over the FULL model dim (all heads,
// affine, eps 1e-5) before the rotary embedding.  A GEMM tile cannot see a whole row, so this variant stores the raw
// projections (fp32 [M, sections * D]) and this kernel finishes the job in one pass per row: LayerNorm(q), LayerNorm(k),
// RoPE, the softmax scale on q, and the scatter into the attention layouts q, k [sample, head, t, hd_pad] and
// V^T [sample, head, hd_pad, t_pad].  first_section = 1: the row holds k | v only (cross-attention K/V with ky_norm).
// No shipped config enables qk_norm; the path is functional, HBM-bound and not tuned further.
namespace ma3 {
template <typename T16>
__global__ void __launch_bounds__(256) qknorm_rope_kernel(const float* __restrict__ x, long long ld, int first_section,
                                                           const float* __restrict__ qw, const float* __restrict__ qb,
                                                           const float* __restrict__ kw, const float* __restrict__ kb,
                                                           float eps, const float* __restrict__ rope, T16* __restrict__ q_out,
                                                           T16* __restrict__ k_out, T16* __restrict__ vt_out, int tokens,
                                                           int tokens_pad, int D, int hd, int hdp, float q_scale) {
  __shared__ float red[64];
  const int m = blockIdx.x, sample = m / tokens, t = m - sample * tokens;
  const int H = D / hd, half = hd >> 1;
  const float* row = x + (long long)m * ld;
  for (int sec = first_section; sec < 3; ++sec) {
    const float* src = row + (long long)(sec - first_section) * D;
    float mean = 0.f, rstd = 1.f;
    const float* w = sec == 0 ? qw : kw;
    const float* b = sec == 0 ? qb : kb;
    const bool norm = sec < 2 && w != nullptr;
    if (norm) {
      float s = 0.f;
      for (int i = threadIdx.x; i < D; i += blockDim.x) s += src[i];
      mean = block_sum(s, red) / (float)D;
      float v = 0.f;
      for (int i = threadIdx.x; i < D; i += blockDim.x) { const float d0 = src[i] - mean; v += d0 * d0; }
      rstd = rsqrtf(block_sum(v, red) / (float)D + eps);
    }
    for (int p = threadIdx.x; p < (D >> 1); p += blockDim.x) {
      const int c = 2 * p, h = c / hd, d = c - h * hd;
      float a0 = src[c], a1 = src[c + 1];
      if (norm) {
        a0 = (a0 - mean) * rstd * w[c] + b[c];
        a1 = (a1 - mean) * rstd * w[c + 1] + b[c + 1];
      }
      if (sec < 2) {
        if (rope != nullptr) {
          const float2 cs = *reinterpret_cast<const float2*>(rope + ((long long)t * half + (d >> 1)) * 2);
          const float r0 = a0 * cs.x - a1 * cs.y, r1 = a0 * cs.y + a1 * cs.x;
          a0 = r0; a1 = r1;
        }
        if (sec == 0) { a0 *= q_scale; a1 *= q_scale; }
        T16* dst = (sec == 0 ? q_out : k_out) + (((long long)sample * H + h) * tokens + t) * hdp + d;
        dst[0] = Cvt<T16>::to(a0);
        dst[1] = Cvt<T16>::to(a1);
      } else {
        T16* dst = vt_out + (((long long)sample * H + h) * hdp + d) * tokens_pad + t;
        dst[0] = Cvt<T16>::to(a0);
        dst[tokens_pad] = Cvt<T16>::to(a1);
      }
    }
    __syncthreads();
  }
}
}  // namespace ma3

extern "C" int ma3_qknorm_rope(const float* x, int64_t ld, int first_section, const float* qw, const float* qb,
                               const float* kw, const float* kb, float eps, const float* rope, void* q_out, void* k_out,
                               void* vt_out, int dtype, int M, int tokens, int tokens_pad, int D, int hd, int hdp,
                               float q_scale, void* stream) {
  MA3_REQUIRE(x && k_out && vt_out && M > 0 && tokens > 0 && M % tokens == 0, "qknorm_rope: bad arguments");
  MA3_REQUIRE(first_section == 0 || first_section == 1, "qknorm_rope: first_section must be 0 or 1");
  MA3_REQUIRE(first_section == 1 || q_out != nullptr, "qknorm_rope: q_out required");
  MA3_REQUIRE(D % hd == 0 && hd % 2 == 0 && hdp >= hd && tokens_pad >= tokens && ld >= (3 - first_section) * (int64_t)D,
              "qknorm_rope: D %% hd == 0, hd even, hdp >= hd, tokens_pad >= tokens, ld >= sections * D");
  MA3_REQUIRE(dtype == MA3_BF16 || dtype == MA3_F16, "qknorm_rope: 16-bit outputs only");
  if (dtype == MA3_BF16)
    ma3::qknorm_rope_kernel<__nv_bfloat16><<<M, 256, 0, ST(stream)>>>(x, ld, first_section, qw, qb, kw, kb, eps, rope,
        (__nv_bfloat16*)q_out, (__nv_bfloat16*)k_out, (__nv_bfloat16*)vt_out, tokens, tokens_pad, D, hd, hdp, q_scale);
  else
    ma3::qknorm_rope_kernel<__half><<<M, 256, 0, ST(stream)>>>(x, ld, first_section, qw, qb, kw, kb, eps, rope,
        (__half*)q_out, (__half*)k_out, (__half*)vt_out, tokens, tokens_pad, D, hd, hdp, q_scale);
  MA3_LAUNCH_CHECK("qknorm_rope");
  return 0;
}

// ---------------------------------------------------------------------------------------- bf16 hi/lo split
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): out[i] = hi, out[n + i] = lo (two stacked operand copies).  A GEMM
// over the taps (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo) then carries ~16 mantissa bits through the bf16 tensor cores;
// used for the step-invariant conditioning path (timestep / caption embedders, adaLN modulation), whose rounding
// errors are systematic per (sample, channel) and are amplified by the guidance combine.
namespace ma3 {
// slice b of the input = columns [col0 + b * col_step, + cols) of the [rows, ld] fp32 matrix x;
// out[b] = [2 * rows, cols] bf16: hi rows, then lo rows
__global__ void split_bf16_kernel(const float* __restrict__ x, long long ld, int col0, int col_step, int rows, int cols,
                                  __nv_bfloat16* __restrict__ out, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % cols);
  const long long rb = i / cols;
  const int r = (int)(rb % rows), b = (int)(rb / rows);
  const float v = x[(long long)r * ld + col0 + (long long)b * col_step + c];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  __nv_bfloat16* ob = out + (long long)b * 2 * rows * cols;
  ob[(long long)r * cols + c] = hi;
  ob[(long long)(rows + r) * cols + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// tail[r, (2 i + j) * D + d] = norm_w[i][j][d] * (1 + mod[r, 6 D i + (j ? 4 D : D) + d]): the folded RMSNorm weight
// wn_s = w * (1 + scale_s) of both norms of every block, stored behind the modulation columns of the same row so that
// it shares the row pitch of the gate vectors (fused-RMSNorm producer, see ma3_gemm_t.norm_w).
__global__ void norm_weights_kernel(float* __restrict__ mod, long long ld, const float* __restrict__ nw, int depth, int D,
                                    int tail_off, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int d = (int)(i % D);
  const long long t = i / D;
  const int ij = (int)(t % (2 * depth));
  const long long r = t / (2 * depth);
  const int blk = ij >> 1, j = ij & 1;
  float* row = mod + r * ld;
  row[tail_off + (long long)ij * D + d] = nw[(long long)ij * D + d] * (1.f + row[6LL * D * blk + (j ? 4 : 1) * D + d]);
}
}  // namespace ma3

extern "C" int ma3_split_bf16(const float* x, int64_t ld, int col0, int col_step, int nb, int rows, int cols, void* out,
                              void* stream) {
  MA3_REQUIRE(x && out && nb > 0 && rows > 0 && cols > 0 && ld >= cols, "split_bf16: bad arguments");
  const long long total = (long long)nb * rows * cols;
  ma3::split_bf16_kernel<<<nblk(total, 256), 256, 0, ST(stream)>>>(x, ld, col0, col_step, rows, cols,
                                                                    (__nv_bfloat16*)out, total);
  MA3_LAUNCH_CHECK("split_bf16");
  return 0;
}

extern "C" int ma3_norm_weights(float* mod, int64_t ld, const float* norm_w, int rows, int depth, int D, int tail_off,
                                void* stream) {
  MA3_REQUIRE(mod && norm_w && rows > 0 && depth > 0 && D > 0 && tail_off >= 6 * D * depth &&
                  ld >= tail_off + 2LL * depth * D,
              "norm_weights: bad arguments");
  const long long total = (long long)rows * 2 * depth * D;
  ma3::norm_weights_kernel<<<nblk(total, 256), 256, 0, ST(stream)>>>(mod, ld, norm_w, depth, D, tail_off, total);
  MA3_LAUNCH_CHECK("norm_weights");
  return 0;
}

extern "C" int ma3_adaln_input(const float* temb, const float* cap, void* out, int out_dtype, int S, int N, int D,
                               int ts_s, int ts_n, void* stream) {
  MA3_REQUIRE(temb && cap && out && S > 0 && N > 0 && D > 0, "adaln_input: bad arguments");
  MA3_REQUIRE(out_dtype >= MA3_F32 && out_dtype <= MA3_F16, "adaln_input: bad output dtype");
  const unsigned grid = nblk((long long)S * N * D, 256);
  if (out_dtype == MA3_F32)
    ma3::adaln_input_kernel<float><<<grid, 256, 0, ST(stream)>>>(temb, cap, (float*)out, S, N, D, ts_s, ts_n);
  else if (out_dtype == MA3_BF16)
    ma3::adaln_input_kernel<__nv_bfloat16><<<grid, 256, 0, ST(stream)>>>(temb, cap, (__nv_bfloat16*)out, S, N, D, ts_s, ts_n);
  else
    ma3::adaln_input_kernel<__half><<<grid, 256, 0, ST(stream)>>>(temb, cap, (__half*)out, S, N, D, ts_s, ts_n);
  MA3_LAUNCH_CHECK("adaln_input");
  return 0;
}
