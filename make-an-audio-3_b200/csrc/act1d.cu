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

#include <stdlib.h>

namespace ma3 {

__constant__ float c_fdn[12];  // f
__constant__ float c_fup[12];  // 2 f

#ifndef MA3_ACT_TT
#define MA3_ACT_TT 16
#endif
#ifndef MA3_ACT_MINBLOCKS
#define MA3_ACT_MINBLOCKS 1
#endif
constexpr int kTT = MA3_ACT_TT;       // outputs per thread (window of kTT + 10 inputs lives in registers)
constexpr int kActThreads = 256;

// Shared-memory loads of a channel pair.  `asm volatile` on purpose: it pins the 26-entry input window in registers;
// left to itself ptxas prefers to re-load (and re-convert) every tap operand from shared memory, which made the kernel
// instruction-bound at ~59 instructions per element.
template <typename T> struct Ld2;
template <> struct Ld2<float> {
  static __device__ __forceinline__ float2 ld(const float* p) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(smem_u32(p)));
    return v;
  }
};
template <> struct Ld2<__half> {
  static __device__ __forceinline__ float2 ld(const __half* p) {
    uint32_t u;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(smem_u32(p)));
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
  }
};
template <> struct Ld2<__nv_bfloat16> {
  static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
    uint32_t u;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(smem_u32(p)));
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
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

// two channels at once: s = u + inv_b * sin^2(a u), packed fp32x2 arithmetic around the two MUFU.SIN
__device__ __forceinline__ float2 snake2(float2 u, float2 a, float2 ib) {
  const float2 arg = fmul2(u, a);
  const float2 sn = make_float2(__sinf(arg.x), __sinf(arg.y));
  return ffma2(fmul2(ib, sn), sn, u);
}

// activated 2x-rate sample with local index n (m = 2*t0 - 5 + n) from the thread's window xw[0 .. kTT+9];
// 6-tap polyphase branch as packed FFMA2 (one instruction per tap for both channels)
template <int n>
__device__ __forceinline__ float2 s_local(const float2 (&xw)[kTT + 10], float2 a, float2 ib) {
  float2 u = make_float2(0.f, 0.f);
  if constexpr (n & 1) {
#pragma unroll
    for (int k = 0; k < 6; ++k) u = ffma2(make_float2(c_fup[11 - 2 * k], c_fup[11 - 2 * k]), xw[(n - 1) / 2 + k], u);
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) u = ffma2(make_float2(c_fup[10 - 2 * k], c_fup[10 - 2 * k]), xw[n / 2 + k], u);
  }
  return snake2(u, a, ib);
}

// Sequence edges: the reference replicate-pads the ACTIVATED 2x-rate signal, i.e. local indices below n_lo / above
// n_hi take the value at n_lo / n_hi.  kEdge threads (first / last time group of a sequence) run the same unrolled
// code with two selects per sample instead of a slow generic path (which used to form the kernel's tail).
struct EdgeCtx {
  int n_lo, n_hi;
  float2 s_lo, s_hi;
};

template <int n, bool kEdge>
__device__ __forceinline__ float2 s_val(const float2 (&xw)[kTT + 10], float2 a, float2 ib, const EdgeCtx& ec) {
  float2 v = s_local<n>(xw, a, ib);
  if constexpr (kEdge) {
    if (n < ec.n_lo) v = ec.s_lo;
    else if (n > ec.n_hi) v = ec.s_hi;
  }
  return v;
}

template <int i, bool kEdge, typename TOut>
__device__ __forceinline__ void slide(const float2 (&xw)[kTT + 10], float2 (&S)[12], float2 a, float2 ib, TOut* orow,
                                      long long ostride, int valid, const EdgeCtx& ec) {
  if constexpr (i < kTT) {
    // the low-pass is symmetric (f[j] = f[11-j]): two independent 6-term FFMA2 chains
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      acc0 = ffma2(make_float2(c_fdn[j], c_fdn[j]), S[j], acc0);
      acc1 = ffma2(make_float2(c_fdn[11 - j], c_fdn[11 - j]), S[11 - j], acc1);
    }
    const float2 acc = make_float2(acc0.x + acc1.x, acc0.y + acc1.y);
    if (i < valid) St2<TOut>::st(orow + (long long)i * ostride, acc);
    if constexpr (i + 1 < kTT) {
#pragma unroll
      for (int j = 0; j < 10; ++j) S[j] = S[j + 2];
      S[10] = s_val<2 * i + 12, kEdge>(xw, a, ib, ec);
      S[11] = s_val<2 * i + 13, kEdge>(xw, a, ib, ec);
      slide<i + 1, kEdge, TOut>(xw, S, a, ib, orow, ostride, valid, ec);
    }
  }
}

template <int n, bool kEdge>
__device__ __forceinline__ void fill_window(const float2 (&xw)[kTT + 10], float2 (&S)[12], float2 a, float2 ib,
                                            const EdgeCtx& ec) {
  if constexpr (n < 12) {
    S[n] = s_val<n, kEdge>(xw, a, ib, ec);
    fill_window<n + 1, kEdge>(xw, S, a, ib, ec);
  }
}

// activated sample at a RUN-TIME local index (only the two clamp targets of an edge thread), straight from the tile
template <typename TIn>
__device__ __forceinline__ float2 s_at(const TIn* srow, int CT, int n, float2 a, float2 ib) {
  const int base = (n & 1) ? (n - 1) / 2 : n / 2;
  const int f0 = (n & 1) ? 11 : 10;
  float2 u = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float2 xv = Ld2<TIn>::ld(srow + (long long)(base + k) * CT);
    const float c = c_fup[f0 - 2 * k];
    u = ffma2(make_float2(c, c), xv, u);
  }
  return snake2(u, a, ib);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent blocks loop over (batch, time tile, channel tile) work items; the next item's (TB + 10) x CT input tile
// is fetched with cp.async into the other shared-memory buffer while the current one is being computed, so the
// HBM latency of the loads is never exposed (the one-tile-per-block version spent its time waiting on them).
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kActThreads, MA3_ACT_MINBLOCKS) act1d_kernel(const TIn* __restrict__ x, TOut* __restrict__ out,
                                                            const float* __restrict__ alpha,
                                                            const float* __restrict__ beta, int B, int T, int C, int CT,
                                                            int tiles_c, int tiles_t, int logscale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  pdl_launch_dependents();
  pdl_wait();
  const int CP = CT >> 1;                 // channel pairs per tile
  const int groups = kActThreads / CP;    // time groups per block
  const int TB = groups * kTT;
  const int rows = TB + 10;
  constexpr int kVecElems = 16 / sizeof(TIn);
  const int vpr = CT / kVecElems;
  const size_t buf_bytes = (size_t)rows * CT * sizeof(TIn);
  const int total = tiles_c * tiles_t * B;
  const int cp = threadIdx.x % CP, tg = threadIdx.x / CP;

  auto decode = [&](int item, int& b, int& tb0, int& c0) {
    const int tile_c = item % tiles_c;
    const int rest = item / tiles_c;
    c0 = tile_c * CT;
    tb0 = (rest % tiles_t) * TB;
    b = rest / tiles_t;
  };
  // stage (TB + 10) x CT with 16-byte async copies; row index clamped to [0, T-1] (= replicate padding of x)
  auto prefetch = [&](int item, int buf) {
    int b, tb0, c0;
    decode(item, b, tb0, c0);
    const TIn* xb = x + (long long)b * T * C;
    TIn* tile = reinterpret_cast<TIn*>(smem_raw + buf * buf_bytes);
    for (int idx = threadIdx.x; idx < rows * vpr; idx += kActThreads) {
      const int r = idx / vpr, v = idx - r * vpr;
      int gr = tb0 - 5 + r;
      gr = gr < 0 ? 0 : (gr > T - 1 ? T - 1 : gr);
      cp_async16(tile + (long long)r * CT + v * kVecElems, xb + (long long)gr * C + c0 + v * kVecElems);
    }
  };

  int it = 0;
  if ((int)blockIdx.x < total) prefetch(blockIdx.x, 0);
  cp_async_commit();
  for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
    const int buf = it & 1;
    const int next = item + gridDim.x;
    if (next < total) prefetch(next, buf ^ 1);
    cp_async_commit();
    int b, tb0, c0;
    decode(item, b, tb0, c0);
    // per-channel Snake parameters: issued before waiting on the tile so their latency overlaps too
    const int c = c0 + 2 * cp;
    float2 a = make_float2(alpha[c], alpha[c + 1]);
    float2 bb = beta ? make_float2(beta[c], beta[c + 1]) : a;
    cp_async_wait<1>();
    __syncthreads();
    const TIn* tile = reinterpret_cast<const TIn*>(smem_raw + buf * buf_bytes);
    const int t0 = tb0 + tg * kTT;
    if (t0 < T) {
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
      float2 xw[kTT + 10];
#pragma unroll
      for (int j = 0; j < kTT + 10; ++j) xw[j] = Ld2<TIn>::ld(srow + (long long)j * CT);
      float2 S[12];
      EdgeCtx ec;
      ec.n_lo = n_lo;
      ec.n_hi = n_hi;
      if (n_lo > 0 || n_hi < 2 * kTT + 9) {
        ec.s_lo = s_at<TIn>(srow, CT, n_lo, a, ib);
        ec.s_hi = s_at<TIn>(srow, CT, n_hi, a, ib);
        fill_window<0, true>(xw, S, a, ib, ec);
        slide<0, true, TOut>(xw, S, a, ib, orow, C, valid, ec);
      } else {
        fill_window<0, false>(xw, S, a, ib, ec);
        slide<0, false, TOut>(xw, S, a, ib, orow, C, valid, ec);
      }
    }
    __syncthreads();   // everyone is done with `buf` before the prefetch two items ahead overwrites it
  }
}

// ------------------------------------------------------------------------------------------------ tensor-core variant
// The fp32 kernel above is bound by the FMA pipe (24 filter FMAs per element against 4 bytes of HBM traffic).  This
// variant moves both 12-tap filters onto the tensor cores as small banded-Toeplitz products, which leaves the FP32 / SFU
// pipes with only the Snake non-linearity (fp16 in / fp16 out, T % 8 == 0):
//   stage 1  U^T[16 ch x 16 samples] = X^T[16 ch x 16 rows] . G^T      (2 x mma.m16n8k16; X^T via ldmatrix.trans)
//   snake    s = u + sin^2(a u) / b on the fp32 accumulator fragments, replicate padding of s at the sequence ends
//   stage 2  out^T[16 ch x 8] = [S_prev | S_cur]^T[16 ch x 32 samples] . F^T   (2 x mma; the accumulator fragment of
//            stage 1 is exactly the A fragment stage 2 needs, so s never leaves registers)
// and the result goes back to channels-last through stmatrix.trans + a coalesced 16-byte copy.  A warp owns 16 channels
// and a run of TS outputs; sample block j covers 2x-rate samples [2 tw0 - 6 + 16 j, +16) and needs input rows
// [tw0 - 6 + 8 j, +16); output group j (8 outputs from tw0 + 8 j) consumes blocks j and j + 1.
constexpr int kMmaTS = 64;                 // outputs per warp segment
constexpr int kMmaWarps = 8;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void stmatrix_x2_trans(uint32_t addr, uint32_t r0, uint32_t r1) {
  asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// up-sampling tap of 2x-rate sample i (block-local) on input row k (block-local): u[mb+i] = sum_k G[i][k] x[base+k]
__device__ __forceinline__ float up_coef(int i, int k) {
  const int kk = k - (i >> 1) - (i & 1);
  return (kk >= 0 && kk <= 5) ? c_fup[((i & 1) ? 10 : 11) - 2 * kk] : 0.f;
}
// low-pass tap of window sample k (0..31) for output n (0..7): out[t0+n] = sum_k F[k][n] s[2 t0 - 6 + k]
__device__ __forceinline__ float dn_coef(int k, int n) {
  const int j = k - 2 * n - 1;
  return (j >= 0 && j <= 11) ? c_fdn[j] : 0.f;
}

__global__ void __launch_bounds__(kMmaWarps * 32, 3) act1d_mma_kernel(const __half* __restrict__ x, __half* __restrict__ out,
                                                                      const float* __restrict__ alpha,
                                                                      const float* __restrict__ beta, int B, int T, int C,
                                                                      int CT, int tiles_c, int tiles_t, int logscale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  pdl_launch_dependents();
  pdl_wait();
  const int cgs = CT >> 4;                   // 16-channel groups per tile
  const int tsplit = kMmaWarps / cgs;        // time segments per tile
  const int TB = tsplit * kMmaTS;
  const int rows = TB + 16;
  const int pitch = CT * 2 + 16;             // bytes; +16 keeps the 8 rows of an ldmatrix / stmatrix on distinct banks
  const int vpr = CT >> 3;                   // 16-byte vectors per row
  const int vshift = 31 - __clz(vpr);
  const size_t in_bytes = (size_t)rows * pitch;
  uint8_t* os = smem_raw + 2 * in_bytes;     // output tile [TB][pitch]
  const int total = tiles_c * tiles_t * B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int cgi = warp % cgs, seg = warp / cgs;

  // constant B fragments: G^T (two 8-sample halves) and F^T (two 16-sample k-steps)
  uint32_t bu[2][2], bd[2][2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = 2 * tq + 8 * r;
      bu[hh][r] = pack_f16(up_coef(8 * hh + g, k), up_coef(8 * hh + g, k + 1));
      bd[hh][r] = pack_f16(dn_coef(16 * hh + k, g), dn_coef(16 * hh + k + 1, g));
    }
  }

  auto decode = [&](int item, int& b, int& tb0, int& c0) {
    const int tile_c = item % tiles_c;
    const int rest = item / tiles_c;
    c0 = tile_c * CT;
    tb0 = (rest % tiles_t) * TB;
    b = rest / tiles_t;
  };
  // stage rows [tb0 - 6, tb0 + TB + 10) x CT, row index clamped to [0, T-1] (= replicate padding of x)
  auto prefetch = [&](int item, int buf) {
    int b, tb0, c0;
    decode(item, b, tb0, c0);
    const __half* xb = x + (long long)b * T * C + c0;
    uint8_t* tile = smem_raw + buf * in_bytes;
    for (int idx = threadIdx.x; idx < rows * vpr; idx += kMmaWarps * 32) {
      const int r = idx >> vshift, v = idx & (vpr - 1);   // vpr is 2, 4 or 8
      int gr = tb0 - 6 + r;
      gr = gr < 0 ? 0 : (gr > T - 1 ? T - 1 : gr);
      cp_async16(tile + (size_t)r * pitch + v * 16, xb + (long long)gr * C + v * 8);
    }
  };

  int it = 0;
  if ((int)blockIdx.x < total) prefetch(blockIdx.x, 0);
  cp_async_commit();
  for (int item = blockIdx.x; item < total; item += gridDim.x, ++it) {
    const int buf = it & 1;
    const int next = item + gridDim.x;
    if (next < total) prefetch(next, buf ^ 1);
    cp_async_commit();
    int b, tb0, c0;
    decode(item, b, tb0, c0);
    const int cw = c0 + cgi * 16;            // first channel of this warp
    float a0 = alpha[cw + g], a1 = alpha[cw + g + 8];
    float b0 = beta ? beta[cw + g] : a0, b1 = beta ? beta[cw + g + 8] : a1;
    if (logscale) { a0 = __expf(a0); a1 = __expf(a1); b0 = __expf(b0); b1 = __expf(b1); }
    const float ib0 = 1.f / (b0 + 1e-9f), ib1 = 1.f / (b1 + 1e-9f);
    cp_async_wait<1>();
    __syncthreads();
    const int tw0 = tb0 + seg * kMmaTS;      // first output of this warp's segment
    if (tw0 < T) {
      const uint32_t xs = smem_u32(smem_raw + buf * in_bytes);
      const int ngroups = min(kMmaTS, T - tw0) >> 3;
      // ldmatrix row address of this lane for block 0: matrix mi = lane / 8 -> rows (mi / 2) * 8 + lane % 8, channels
      // (mi % 2) * 8 of the warp's group
      const uint32_t xaddr0 = xs + (uint32_t)((seg * kMmaTS + ((lane >> 4) << 3) + (lane & 7)) * pitch) +
                              (uint32_t)((cgi * 16 + ((lane >> 3) & 1) * 8) * 2);
      const uint32_t oaddr0 = smem_u32(os) + (uint32_t)((seg * kMmaTS + (lane & 7)) * pitch) +
                              (uint32_t)((cgi * 16 + ((lane >> 3) & 1) * 8) * 2);
      // one sample block: stage-1 MMAs, Snake on the accumulator fragments, replicate padding of the activated signal
      // at the sequence ends, packed to the fp16 A fragment stage 2 consumes
      auto sample_block = [&](int j, uint32_t (&sc)[4]) {
        uint32_t xa[4];
        ldmatrix_x4_trans(xa, xaddr0 + (uint32_t)(j * 8 * pitch));
        float ul[4] = {0.f, 0.f, 0.f, 0.f}, uh[4] = {0.f, 0.f, 0.f, 0.f};
        mma_f16(ul, xa, bu[0][0], bu[0][1]);
        mma_f16(uh, xa, bu[1][0], bu[1][1]);
        float sl[4], sh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {   // (low, high) sample of one channel as a packed fp32 pair around the two MUFU.SIN
          const float a = e < 2 ? a0 : a1, ib = e < 2 ? ib0 : ib1;
          const float2 u = make_float2(ul[e], uh[e]);
          const float2 arg = fmul2(u, make_float2(a, a));
          const float2 sn = make_float2(__sinf(arg.x), __sinf(arg.y));
          const float2 r = ffma2(fmul2(make_float2(ib, ib), sn), sn, u);
          sl[e] = r.x;
          sh[e] = r.y;
        }
        const int mb = 2 * tw0 - 6 + 16 * j;
        if (mb < 0) {             // only block 0 of a sequence: m = 0 is local sample 6 (lane tq = 3, low half, e = 0)
          const float v0 = __shfl_sync(0xffffffffu, sl[0], (lane & ~3) | 3);
          const float v1 = __shfl_sync(0xffffffffu, sl[2], (lane & ~3) | 3);
          if (tq < 3) { sl[0] = v0; sl[1] = v0; sl[2] = v1; sl[3] = v1; }
        }
        if (mb + 15 > 2 * T - 1) {   // last block of a sequence (T % 8 == 0): m = 2T-1 is local sample 5 (tq = 2, e = 1)
          const float v0 = __shfl_sync(0xffffffffu, sl[1], (lane & ~3) | 2);
          const float v1 = __shfl_sync(0xffffffffu, sl[3], (lane & ~3) | 2);
          if (tq == 3) { sl[0] = v0; sl[1] = v0; sl[2] = v1; sl[3] = v1; }
          sh[0] = v0; sh[1] = v0; sh[2] = v1; sh[3] = v1;
        }
        sc[0] = pack_f16(sl[0], sl[1]); sc[1] = pack_f16(sl[2], sl[3]);
        sc[2] = pack_f16(sh[0], sh[1]); sc[3] = pack_f16(sh[2], sh[3]);
      };
      // output group g (8 outputs from tw0 + 8 g) = low-pass over sample blocks g (window samples 0-15) and g + 1 (16-31)
      auto out_group = [&](int gidx, const uint32_t (&lo)[4], const uint32_t (&hi)[4]) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        mma_f16(o, lo, bd[0][0], bd[0][1]);
        mma_f16(o, hi, bd[1][0], bd[1][1]);
        stmatrix_x2_trans(oaddr0 + (uint32_t)(gidx * 8 * pitch), pack_f16(o[0], o[1]), pack_f16(o[2], o[3]));
      };
      uint32_t sp[4];
      sample_block(0, sp);
      int j = 1;
      for (; j + 1 <= ngroups; j += 2) {   // two blocks per trip: their chains are independent (instruction-level parallelism)
        uint32_t c1[4], c2[4];
        sample_block(j, c1);
        sample_block(j + 1, c2);
        out_group(j - 1, sp, c1);
        out_group(j, c1, c2);
#pragma unroll
        for (int e = 0; e < 4; ++e) sp[e] = c2[e];
      }
      if (j <= ngroups) {
        uint32_t c1[4];
        sample_block(j, c1);
        out_group(j - 1, sp, c1);
      }
    }
    __syncthreads();
    // coalesced copy of the output tile to channels-last global memory
    {
      __half* ob = out + ((long long)b * T + tb0) * C + c0;
      const int nrows = min(TB, T - tb0);
      for (int idx = threadIdx.x; idx < nrows * vpr; idx += kMmaWarps * 32) {
        const int r = idx >> vshift, v = idx & (vpr - 1);
        *reinterpret_cast<uint4*>(ob + (long long)r * C + v * 8) = *reinterpret_cast<const uint4*>(os + (size_t)r * pitch + v * 16);
      }
    }
  }
}

// ------------------------------------------------------------------------------------- tensor-core variant, TMA-staged
// Same mathematics as act1d_mma_kernel; what changed is everything around the MMAs, which is where that kernel spent
// two thirds of its issue slots (profiles/r02_ncu_full_act1d_summary.txt: 34.9 M warp-instructions, 0.63 IPC):
//   * tiles move with TMA in both directions (3-D tensor maps over [B, T, C], 128/64/32-byte swizzle = one tile row):
//     no per-thread staging loops, no index clamping (rows outside the sequence arrive as zeros and the three rows the
//     replicate padding of x really needs are patched by the one warp that reads them), no output copy loop;
//   * a warp owns 16 channels x 128 outputs (17 sample blocks for 16 output groups: 6 % redundant work, was 12.5 %);
//   * Snake runs on the accumulator register pairs as the MMA delivers them (two samples of one channel per packed
//     fp32 instruction: no register shuffling), and the sequence-end handling is peeled out of the steady-state loop.
// Per 16 channels x 16 samples the loop is 1 ldmatrix + 4 MMAs + 8 MUFU.SIN (+ 8 range-reduction FMULs) + 12 packed
// fp32 + 6 packs + 1 stmatrix; MUFU (16 results / clk / SM) then bounds the kernel at ~0.75 of the HBM time.
#ifndef MA3_ACT_ILP4
#define MA3_ACT_ILP4 1
#endif
#ifndef MA3_ACT_NBUF
#define MA3_ACT_NBUF 2   // input buffers of the in-place variant: 3 = loads issued two tiles ahead (experiment)
#endif
#ifndef MA3_ACT_KO
#define MA3_ACT_KO 0   // knock-out builds for diagnostics only (1 = no MUFU.SIN, 2 = no low-pass MMAs, 3 = no stmatrix)
#endif
constexpr int kTmaWarps = 4;
constexpr int kTmaSeg = 128;               // outputs per warp and tile

struct Act1dTmaParams {
  CUtensorMap tin, tout;                   // [B, T, C] fp16, box {CT, load rows | 128, 1}
  const float* alpha;
  const float* beta;
  int B, T, C, tiles_c, tiles_t, logscale;
  int seg;                                 // outputs per warp and tile (multiple of 8, <= kTmaSeg; host picks it so that
                                           // the tiles divide evenly over the persistent grid)
  int boxr;                                // rows of a load box (>= seg + 16)
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One sample block: 16 channels x 16 activated 2x-rate samples as the fp16 A fragment of the low-pass MMAs.
// EDGE 1: first block of a sequence (samples 0..5 lie before m = 0 and take s[0] = local sample 6);
// EDGE 2: last block of a sequence (samples 6..15 lie beyond m = 2T-1 = local sample 5).
template <int EDGE>
__device__ __forceinline__ void act_sample_block(uint32_t xaddr, const uint32_t (&bu)[2][2], float2 a0, float2 a1,
                                                 float2 ib0, float2 ib1, int lane, uint32_t (&sc)[4]) {
  uint32_t xa[4];
  ldmatrix_x4_trans(xa, xaddr);
  float ul[4] = {0.f, 0.f, 0.f, 0.f}, uh[4] = {0.f, 0.f, 0.f, 0.f};
  mma_f16(ul, xa, bu[0][0], bu[0][1]);     // samples 0..7:  [0],[1] = channel g, samples 2tq, 2tq+1; [2],[3] = channel g+8
  mma_f16(uh, xa, bu[1][0], bu[1][1]);     // samples 8..15
#if MA3_ACT_KO == 1   // diagnostics (tools/probe_act1d_ko.sh): no MUFU.SIN, same packed arithmetic around it; wrong results
  auto snake_ko = [](float2 u, float2 a, float2 ib) { const float2 g = fmul2(u, a); return ffma2(fmul2(ib, g), g, u); };
  float2 p0 = snake_ko(make_float2(ul[0], ul[1]), a0, ib0), p1 = snake_ko(make_float2(ul[2], ul[3]), a1, ib1);
  float2 p2 = snake_ko(make_float2(uh[0], uh[1]), a0, ib0), p3 = snake_ko(make_float2(uh[2], uh[3]), a1, ib1);
#else
  float2 p0 = snake2(make_float2(ul[0], ul[1]), a0, ib0);
  float2 p1 = snake2(make_float2(ul[2], ul[3]), a1, ib1);
  float2 p2 = snake2(make_float2(uh[0], uh[1]), a0, ib0);
  float2 p3 = snake2(make_float2(uh[2], uh[3]), a1, ib1);
#endif
  if constexpr (EDGE == 1) {
    const float v0 = __shfl_sync(0xffffffffu, p0.x, (lane & ~3) | 3);
    const float v1 = __shfl_sync(0xffffffffu, p1.x, (lane & ~3) | 3);
    if ((lane & 3) < 3) { p0 = make_float2(v0, v0); p1 = make_float2(v1, v1); }
  }
  if constexpr (EDGE == 2) {
    const float v0 = __shfl_sync(0xffffffffu, p0.y, (lane & ~3) | 2);
    const float v1 = __shfl_sync(0xffffffffu, p1.y, (lane & ~3) | 2);
    if ((lane & 3) == 3) { p0 = make_float2(v0, v0); p1 = make_float2(v1, v1); }
    p2 = make_float2(v0, v0);
    p3 = make_float2(v1, v1);
  }
  sc[0] = pack_f16(p0.x, p0.y);
  sc[1] = pack_f16(p1.x, p1.y);
  sc[2] = pack_f16(p2.x, p2.y);
  sc[3] = pack_f16(p3.x, p3.y);
}

__device__ __forceinline__ void act_out_group(uint32_t oaddr, const uint32_t (&lo)[4], const uint32_t (&hi)[4],
                                              const uint32_t (&bd)[2][2]) {
#if MA3_ACT_KO == 2   // diagnostics: no low-pass MMAs (the activated samples are stored as they are); wrong results
  stmatrix_x2_trans(oaddr, lo[0] ^ hi[2], lo[1] ^ hi[3]);
#elif MA3_ACT_KO == 3  // diagnostics: no stmatrix (one lane keeps the value alive through a never-taken store); wrong results
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  mma_f16(o, lo, bd[0][0], bd[0][1]);
  mma_f16(o, hi, bd[1][0], bd[1][1]);
  if (o[0] == 1.2345e30f) stmatrix_x2_trans(oaddr, pack_f16(o[0], o[1]), pack_f16(o[2], o[3]));
#else
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  mma_f16(o, lo, bd[0][0], bd[0][1]);
  mma_f16(o, hi, bd[1][0], bd[1][1]);
  stmatrix_x2_trans(oaddr, pack_f16(o[0], o[1]), pack_f16(o[2], o[3]));
#endif
}

// INPLACE: every 128-output segment has its own (outputs + 16 halo) row region of the input buffer and a warp writes
// output group g over rows 8g+8 .. 8g+15 of its region -- rows whose last readers (sample blocks g and g+1) are the very
// operands of that group, so the ordering is a true register dependency -- and the TMA store reads the tile back from
// the input buffer: no output tile, 37-42 KB instead of 54 KB per CTA, five CTAs (20 warps) per SM instead of four.
template <int CT, bool INPLACE>
__global__ void __launch_bounds__(kTmaWarps * 32, (INPLACE && MA3_ACT_NBUF == 2) ? 5 : 4) act1d_tma_kernel(const __grid_constant__ Act1dTmaParams p) {
  constexpr int CGS = CT / 16;                       // 16-channel groups per tile
  constexpr int TSPLIT = kTmaWarps / CGS;            // 128-output time segments per tile
  constexpr int SMAX = kTmaSeg;
  const int S = p.seg;
  const int TB = TSPLIT * S;
  constexpr uint32_t PITCH = CT * 2;                 // bytes per tile row = the swizzle span of the tensor maps
  constexpr int BOXR = CT == 16 ? 160 : 144;         // rows per load box: S + 16 halo rows, rounded up so that a box is
                                                     // a multiple of 1024 bytes (swizzle pattern period x 8 rows)
  constexpr int IN_ROWS = INPLACE ? TSPLIT * BOXR : (TSPLIT - 1) * SMAX + BOXR;
  constexpr uint32_t IN_BYTES = IN_ROWS * PITCH;
  constexpr uint32_t OUT_BYTES = INPLACE ? 0 : TSPLIT * SMAX * PITCH;
  constexpr uint32_t SWZ = CT == 64 ? 0x70u : (CT == 32 ? 0x30u : 0x10u);
  static_assert(IN_BYTES % 1024 == 0 && OUT_BYTES % 1024 == 0 && (SMAX * PITCH) % 1024 == 0, "tile alignment");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle patterns are functions of the address bits
  constexpr int NB = INPLACE ? MA3_ACT_NBUF : 2;                 // input buffers; loads run NB - 1 tiles ahead
  const uint32_t out_base = base + NB * IN_BYTES;
  const uint32_t bar0 = out_base + OUT_BYTES;                    // NB 8-byte "tile landed" barriers
  auto swz = [](uint32_t off) { return off ^ ((off >> 3) & SWZ); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int cgi = warp % CGS, seg = warp / CGS;
  const int total = p.tiles_c * p.tiles_t * p.B;
  const int T = p.T;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NB; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
    fence_barrier_init();
    prefetch_tmap(&p.tin);
    prefetch_tmap(&p.tout);
  }
  // constant B fragments: G^T (two 8-sample halves) and F^T (two 16-sample k-steps)
  uint32_t bu[2][2], bd[2][2];
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = 2 * tq + 8 * r;
      bu[hh][r] = pack_f16(up_coef(8 * hh + g, k), up_coef(8 * hh + g, k + 1));
      bd[hh][r] = pack_f16(dn_coef(16 * hh + k, g), dn_coef(16 * hh + k + 1, g));
    }
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();

  auto decode = [&](int item, int& b, int& tb0, int& c0) {
    const int tile_c = item % p.tiles_c;
    const int rest = item / p.tiles_c;
    c0 = tile_c * CT;
    tb0 = (rest % p.tiles_t) * TB;
    b = rest / p.tiles_t;
  };
  // rows [tb0 - 6, tb0 - 6 + IN_ROWS) x CT of sequence b; rows outside [0, T) arrive as zeros
  auto issue_load = [&](int item, int buf) {
    int b, tb0, c0;
    decode(item, b, tb0, c0);
    const uint32_t bar = bar0 + 8 * buf;
    mbar_arrive_expect_tx_u32(bar, (uint32_t)(TSPLIT * p.boxr) * PITCH);
#pragma unroll
    for (int i = 0; i < TSPLIT; ++i)
      tma_load_3d_u32(base + buf * IN_BYTES + i * ((INPLACE ? BOXR : S) * PITCH), &p.tin, bar, c0, tb0 - 6 + i * S, b);
  };
  const int seg_row = seg * (INPLACE ? BOXR : S);   // first buffer row of this warp's segment

  // Tiles of this block: item = blockIdx + i * grid.  The hardware hands consecutive block indices to different SMs, so
  // the total % grid blocks that run one tile more are spread one (or two) per SM; moving them to every (grid / rem)-th
  // block index instead measured 12 % slower (35.6 vs 31.8 us at C = 384, T = 9984, B = 8).
  const int grid = (int)gridDim.x, bid = (int)blockIdx.x;
  const int n_items = bid < total ? (total - bid + grid - 1) / grid : 0;
  auto item_of = [&](int i) { return bid + i * grid; };
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NB - 1; ++i)
      if (i < n_items) issue_load(item_of(i), i);
  }
  for (int it = 0; it < n_items; ++it) {
    const int buf = it % NB;
    const int item = item_of(it);
    if (tid == 0) {
      // the buffer of tile it - 1 was released by the barrier that closed the previous iteration; in place, that tile's
      // store still reads its outputs from it
      if constexpr (INPLACE) bulk_wait_read0();
      if (it + NB - 1 < n_items) {
        fence_proxy_async_smem();
        issue_load(item_of(it + NB - 1), (it + NB - 1) % NB);
      }
      if constexpr (!INPLACE) bulk_wait_read0();   // the previous tile's store has drained the output tile
    }
    int b, tb0, c0;
    decode(item, b, tb0, c0);
    const int cw = c0 + cgi * 16;              // first channel of this warp
    float al0 = p.alpha[cw + g], al1 = p.alpha[cw + g + 8];
    float be0 = p.beta ? p.beta[cw + g] : al0, be1 = p.beta ? p.beta[cw + g + 8] : al1;
    if (p.logscale) { al0 = __expf(al0); al1 = __expf(al1); be0 = __expf(be0); be1 = __expf(be1); }
    const float i0 = 1.f / (be0 + 1e-9f), i1 = 1.f / (be1 + 1e-9f);
    const float2 a0 = make_float2(al0, al0), a1 = make_float2(al1, al1);
    const float2 ib0 = make_float2(i0, i0), ib1 = make_float2(i1, i1);
    while (!mbar_try_wait(bar0 + 8 * buf, (uint32_t)((it / NB) & 1))) {
    }
    if constexpr (!INPLACE) __syncthreads();   // nobody writes the output tile before its previous store has read it
    const int tw0 = tb0 + seg * S;             // first output of this warp's segment
    if (tw0 < T) {
      const uint32_t in0 = base + buf * IN_BYTES;
      const int ng = min(S, T - tw0) >> 3;     // output groups of 8; sample blocks 0 .. ng
      const bool at_start = tw0 == 0, at_end = tw0 + 8 * ng == T;
      // replicate padding of x: the first / last sample block of a sequence reads three rows beyond it (t = -3..-1,
      // t = T..T+2).  Only this warp reads them in its 16 channels, so it patches them itself.
      if (at_start && lane < 6) {
        const uint32_t col = (uint32_t)(cgi * 32 + (lane & 1) * 16);
        sts128(in0 + swz((3 + (lane >> 1)) * PITCH + col), lds128(in0 + swz(6 * PITCH + col)));
      }
      if (at_end && lane >= 8 && lane < 14) {
        const int l = lane - 8;
        const uint32_t col = (uint32_t)(cgi * 32 + (l & 1) * 16);
        const uint32_t rlast = (uint32_t)(seg_row + T - 1 - tw0 + 6);
        sts128(in0 + swz((rlast + 1 + (l >> 1)) * PITCH + col), lds128(in0 + swz(rlast * PITCH + col)));
      }
      __syncwarp();
      // ldmatrix / stmatrix row addresses of this lane for block / group 0: matrix mi = lane / 8 -> rows (mi / 2) * 8 +
      // lane % 8, channels (mi % 2) * 8 of the warp's group.  Later blocks are 8 rows on: the swizzle term is unchanged.
      const uint32_t xaddr0 =
          in0 + swz((uint32_t)(seg_row + ((lane >> 4) << 3) + (lane & 7)) * PITCH + (uint32_t)(cgi * 32 + ((lane >> 3) & 1) * 16));
      const uint32_t oaddr0 =
          (INPLACE ? in0 : out_base) +
          swz((uint32_t)(seg_row + (INPLACE ? 8 : 0) + (lane & 7)) * PITCH + (uint32_t)(cgi * 32 + ((lane >> 3) & 1) * 16));
      uint32_t sp[4];
      if (at_start) act_sample_block<1>(xaddr0, bu, a0, a1, ib0, ib1, lane, sp);
      else act_sample_block<0>(xaddr0, bu, a0, a1, ib0, ib1, lane, sp);
      int j = 1;
#if MA3_ACT_ILP4
      {                                        // four blocks per trip: independent chains
#pragma unroll 1
        for (; j + 3 < ng; j += 4) {
          uint32_t c1[4], c2[4], c3[4], c4[4];
          act_sample_block<0>(xaddr0 + (uint32_t)j * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c1);
          act_sample_block<0>(xaddr0 + (uint32_t)(j + 1) * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c2);
          act_sample_block<0>(xaddr0 + (uint32_t)(j + 2) * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c3);
          act_sample_block<0>(xaddr0 + (uint32_t)(j + 3) * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c4);
          act_out_group(oaddr0 + (uint32_t)(j - 1) * (8 * PITCH), sp, c1, bd);
          act_out_group(oaddr0 + (uint32_t)j * (8 * PITCH), c1, c2, bd);
          act_out_group(oaddr0 + (uint32_t)(j + 1) * (8 * PITCH), c2, c3, bd);
          act_out_group(oaddr0 + (uint32_t)(j + 2) * (8 * PITCH), c3, c4, bd);
#pragma unroll
          for (int e = 0; e < 4; ++e) sp[e] = c4[e];
        }
      }
#endif
#pragma unroll 2
      for (; j + 1 < ng; j += 2) {             // blocks 1 .. ng-1 never touch a sequence end; two per trip for ILP
        uint32_t c1[4], c2[4];
        act_sample_block<0>(xaddr0 + (uint32_t)j * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c1);
        act_sample_block<0>(xaddr0 + (uint32_t)(j + 1) * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c2);
        act_out_group(oaddr0 + (uint32_t)(j - 1) * (8 * PITCH), sp, c1, bd);
        act_out_group(oaddr0 + (uint32_t)j * (8 * PITCH), c1, c2, bd);
#pragma unroll
        for (int e = 0; e < 4; ++e) sp[e] = c2[e];
      }
      if (j < ng) {
        uint32_t c1[4];
        act_sample_block<0>(xaddr0 + (uint32_t)j * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c1);
        act_out_group(oaddr0 + (uint32_t)(j - 1) * (8 * PITCH), sp, c1, bd);
#pragma unroll
        for (int e = 0; e < 4; ++e) sp[e] = c1[e];
        ++j;
      }
      {                                        // block ng closes the segment
        uint32_t c1[4];
        if (at_end) act_sample_block<2>(xaddr0 + (uint32_t)ng * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c1);
        else act_sample_block<0>(xaddr0 + (uint32_t)ng * (8 * PITCH), bu, a0, a1, ib0, ib1, lane, c1);
        act_out_group(oaddr0 + (uint32_t)(ng - 1) * (8 * PITCH), sp, c1, bd);
      }
    }
    fence_proxy_async_smem();                  // stmatrix results -> visible to the TMA store
    __syncthreads();                           // output tile complete, input buffer `buf` released
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < TSPLIT; ++i)
        if (tb0 + i * S < T)
          tma_store_3d(&p.tout, INPLACE ? base + buf * IN_BYTES + (uint32_t)(i * BOXR + 8) * PITCH : out_base + i * (S * PITCH), c0,
                       tb0 + i * S, b);
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait0();
}

static DeviceOnce g_filter_set;   // the taps live in __constant__ memory: one copy per device

}  // namespace ma3

using namespace ma3;

static int g_act_version = 0;   // 0 = default (TMA-staged tensor-core kernel where eligible), 1 = first-generation kernel

extern "C" {

int ma3_debug_set_act_version(int v) {
  g_act_version = v;
  return 0;
}

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
  g_filter_set.mark();
  return 0;
}

// x [B, T, C] (f32 / f16 / bf16) -> out [B, T, C] (f16 / bf16); alpha, beta [C] fp32 (beta NULL = Snake).
int ma3_act1d(const void* x, int in_dtype, void* out, int out_dtype, const float* alpha, const float* beta, int B,
              int T, int C, int logscale, void* stream) {
  MA3_REQUIRE(!g_filter_set.pending(), "act1d: call ma3_act1d_set_filter first (once per device)");
  MA3_REQUIRE(x && out && alpha && B > 0 && T > 0, "act1d: null pointer or empty");
  MA3_REQUIRE(C % 16 == 0, "act1d: C=%d must be a multiple of 16 (pad channels)", C);
  MA3_REQUIRE(aligned16(x) && aligned16(out), "act1d: pointers must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool no_mma = getenv("MA3_ACT_FP32") != nullptr && getenv("MA3_ACT_FP32")[0] == '1';
  if (!no_mma && g_act_version != 1 && in_dtype == MA3_F16 && out_dtype == MA3_F16 && T % 8 == 0) {
    // tensor-core variant, tiles staged by TMA
    const int CT = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16);
    const int tsplit = kTmaWarps / (CT / 16);
    const int boxr_max = CT == 16 ? 160 : 144, row_align = 1024 / (CT * 2);
    const long long grid_max = 4LL * num_sms();
    // outputs per warp and tile: a segment of S outputs costs S / 8 + 1 sample blocks and the slowest block of the
    // persistent grid runs ceil(tiles / grid) tiles, so a slightly shorter segment can remove a mostly idle last round
    int seg = kTmaSeg;
    static const bool seg_auto = getenv("MA3_ACT_SEG_AUTO") != nullptr && getenv("MA3_ACT_SEG_AUTO")[0] == '1';
    if (seg_auto) {   // measured slower at the BigVGAN shapes (33.1 vs 31.8 us at C=384): off by default
      long long best = -1;
      for (int sc = kTmaSeg; sc >= 96; sc -= 8) {
        const long long tiles = (long long)(C / CT) * ((T + tsplit * sc - 1) / (tsplit * sc)) * B;
        const long long cost = ((tiles + grid_max - 1) / grid_max) * (sc / 8 + 1);
        if (best < 0 || cost < best) { best = cost; seg = sc; }
      }
    }
    const int TB = tsplit * seg;
    int boxr = (seg + 16 + row_align - 1) / row_align * row_align;
    if (boxr > boxr_max) boxr = boxr_max;
    Act1dTmaParams p;
    memset(&p, 0, sizeof(p));
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)T, (uint64_t)B};
    const uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)T * C * 2};
    const uint32_t box_in[3] = {(uint32_t)CT, (uint32_t)boxr, 1}, box_out[3] = {(uint32_t)CT, (uint32_t)seg, 1};
    int rc = encode_tmap(&p.tin, x, 2, 3, dims, str, box_in, CT * 2);
    if (rc == 0) rc = encode_tmap(&p.tout, out, 2, 3, dims, str, box_out, CT * 2);
    if (rc != 0) return rc;
    p.alpha = alpha;
    p.beta = beta;
    p.B = B;
    p.T = T;
    p.C = C;
    p.tiles_c = C / CT;
    p.tiles_t = (T + TB - 1) / TB;
    p.logscale = logscale;
    p.seg = seg;
    p.boxr = boxr;
    const long long total = (long long)p.tiles_c * p.tiles_t * B;
    MA3_REQUIRE(total < (1ll << 31), "act1d: too many tiles");
    // in-place outputs (five CTAs per SM instead of four): measured equal (32.9 vs 32.5 us at C = 384, T = 9984, B = 8:
    // 4.2 instead of 3.7 warps per scheduler issue the same 0.48 instructions per clock), so the simpler layout is default
    static const bool inplace = getenv("MA3_ACT_INPLACE") != nullptr && getenv("MA3_ACT_INPLACE")[0] == '1';
    const size_t in_bytes = (size_t)(inplace ? tsplit * boxr_max : (tsplit - 1) * kTmaSeg + boxr_max) * CT * 2;
    // + output tile (not in place) + barriers + 1024-byte alignment slack
    const int nbuf = inplace ? MA3_ACT_NBUF : 2;
    const size_t smem = nbuf * in_bytes + (inplace ? 0 : (size_t)tsplit * kTmaSeg * CT * 2) + 32 + 1024;
    const long long grid_cap = ((inplace && nbuf == 2) ? 5LL : 4LL) * num_sms();
    long long gridl = grid_cap;
    if (gridl > total) gridl = total;
    cudaError_t le = cudaSuccess;
#define ACT_TMA_CASE(CTV, IP)                                                                                        \
  do {                                                                                                               \
    static DeviceOnce configured;                                                                                    \
    if (configured.pending()) {                                                                                      \
      cudaFuncSetAttribute(act1d_tma_kernel<CTV, IP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);       \
      configured.mark();                                                                                             \
    }                                                                                                                \
    le = launch_pdl(act1d_tma_kernel<CTV, IP>, dim3((unsigned)gridl), dim3(kTmaWarps * 32), smem, st, 1, p);         \
  } while (0)
    if (inplace) {
      if (CT == 64) ACT_TMA_CASE(64, true);
      else if (CT == 32) ACT_TMA_CASE(32, true);
      else ACT_TMA_CASE(16, true);
    } else {
      if (CT == 64) ACT_TMA_CASE(64, false);
      else if (CT == 32) ACT_TMA_CASE(32, false);
      else ACT_TMA_CASE(16, false);
    }
#undef ACT_TMA_CASE
    if (le != cudaSuccess) MA3_FAIL((int)le, "act1d launch: %s", cudaGetErrorString(le));
    MA3_LAUNCH_CHECK("act1d");
    return 0;
  }
  if (!no_mma && in_dtype == MA3_F16 && out_dtype == MA3_F16 && T % 8 == 0) {
    // first-generation tensor-core variant (cp.async staging); kept selectable: ma3_debug_set_act_version(1)
    const int CT = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16);
    const int TB = (kMmaWarps / (CT / 16)) * kMmaTS;
    const int tiles_c = C / CT, tiles_t = (T + TB - 1) / TB;
    const long long total = (long long)tiles_c * tiles_t * B;
    const int pitch = CT * 2 + 16;
    const size_t smem = 2 * (size_t)(TB + 16) * pitch + (size_t)TB * pitch;
    static DeviceOnce configured;
    if (configured.pending()) {
      cudaFuncSetAttribute(act1d_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      configured.mark();
    }
    long long gridl = 3LL * num_sms();
    if (gridl > total) gridl = total;
    cudaError_t le = launch_pdl(act1d_mma_kernel, dim3((unsigned)gridl), dim3(kMmaWarps * 32), smem, st, 1,
                                (const __half*)x, (__half*)out, alpha, beta, B, T, C, CT, tiles_c, tiles_t, logscale);
    if (le != cudaSuccess) MA3_FAIL((int)le, "act1d launch: %s", cudaGetErrorString(le));
    MA3_LAUNCH_CHECK("act1d");
    return 0;
  }
  const int CT = C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16);
  const int groups = kActThreads / (CT / 2);
  const int TB = groups * kTT;
  const int tiles_c = C / CT, tiles_t = (T + TB - 1) / TB;
  const long long total = (long long)tiles_c * tiles_t * B;
  const size_t smem = 2 * (size_t)(TB + 10) * CT * dtype_bytes(in_dtype);   // double-buffered input tile
  // persistent grid: two blocks per SM (registers) when there is enough work
  long long gridl = (long long)(MA3_ACT_MINBLOCKS > 2 ? MA3_ACT_MINBLOCKS : 2) * num_sms();
  if (gridl > total) gridl = total;
  dim3 grid((unsigned)gridl);
  cudaError_t le = cudaSuccess;
#define ACT_CASE(TI, TO)                                                                                            \
  do {                                                                                                              \
    static DeviceOnce configured;                                                                                 \
    if (configured.pending()) {                                                                                              \
      cudaFuncSetAttribute(act1d_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);          \
      configured.mark();                                                                                            \
    }                                                                                                               \
    le = launch_pdl(act1d_kernel<TI, TO>, grid, dim3(kActThreads), smem, st, 1, (const TI*)x, (TO*)out, alpha, beta, \
                    B, T, C, CT, tiles_c, tiles_t, logscale);                                                       \
  } while (0)
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
