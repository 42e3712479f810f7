// Flash attention for the Next-DiT block on tcgen05: self-attention over the T latent tokens plus gated
// cross-attention over the L context tokens in ONE launch, sharing the rotary-embedded Q tile:
//     out = softmax(q k^T) v  +  tanh(gate_h) * softmax(q ky^T) vy
// (flag_large_dit_moe.py:382-406; q is pre-multiplied by log2(e)/sqrt(hd) in the QKV GEMM epilogue, q and k already
// carry RoPE; the all-ones masks of the reference are dropped).
//
// CTA = (128-query tile, head, sample).  Warp 4 (one elected lane) feeds TMA and issues the MMAs
// S = Q K^T  and  O_tile = P V  into TMEM; warps 0-3 (thread <-> query row <-> TMEM lane) run the online softmax,
// write P (bf16) into shared memory in the 128B-swizzled K-major layout the tensor core reads, and accumulate the
// O tiles in registers with the usual rescaling.  K / V^T tiles stream through a 2-stage TMA ring.
#include "host_common.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace ma3 {

constexpr int kAttnThreads = 160;
#ifndef MA3_ATTN_POLY_PAIRS
#define MA3_ATTN_POLY_PAIRS 0
#endif

struct AttnParams {
  CUtensorMap tmQ, tmK, tmVt, tmKy, tmVyt;
  CUtensorMap tmVt2, tmVyt2;   // BKV = 80: keys 64..79 of a V^T tile (16 keys = 32-byte rows, 32B swizzle)
  int T, L, H, D;  // D = H * hd (row pitch of out)
  void* out;
  const float* gate;
  int dtype;
  long long* trace;  // diagnostics (CTA 0 records clock64 at pipeline events when non-null)
  int stagger;       // v3: clocks by which consecutive contexts start their first tile apart
};

__device__ __forceinline__ void attn_trace(const AttnParams& p, int it, int slot) {
  if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 16) p.trace[it * 16 + slot] = clock64();
}

template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t addr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld_n<32>(uint32_t addr, uint32_t* r) { tmem_ld32(addr, *reinterpret_cast<uint32_t(*)[32]>(r)); }
template <>
__device__ __forceinline__ void tmem_ld_n<16>(uint32_t addr, uint32_t* r) { tmem_ld16(addr, *reinterpret_cast<uint32_t(*)[16]>(r)); }
template <>
__device__ __forceinline__ void tmem_ld_n<8>(uint32_t addr, uint32_t* r) { tmem_ld8(addr, *reinterpret_cast<uint32_t(*)[8]>(r)); }

// acc[0..HD) = acc * alpha + O_tile(row)  reading HD fp32 columns in pieces of 32 / 16 / 8
template <int HD, int OFF = 0>
__device__ __forceinline__ void accumulate_o(uint32_t taddr, float (&acc)[HD], float alpha) {
  if constexpr (OFF < HD) {
    constexpr int W = (HD - OFF >= 32) ? 32 : ((HD - OFF >= 16) ? 16 : 8);
    uint32_t r[W];
    tmem_ld_n<W>(taddr + OFF, r);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < W; ++e) acc[OFF + e] = fmaf(acc[OFF + e], alpha, __uint_as_float(r[e]));
    accumulate_o<HD, OFF + W>(taddr, acc, alpha);
  }
}

// Pass 1 of a KV tile: row maximum over BKV logits (thread <-> query row).  kMask only on the ragged last tile of a
// segment: the unmasked variant is one FMNMX per element.  TMEM loads are software-pipelined (the next 16 columns are
// in flight while the current 16 are reduced) and the reduction uses four independent chains.
template <bool kMask>
__device__ __forceinline__ void max16(const uint32_t (&r)[16], int c0, int valid, float (&mx)[4]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    if (kMask) {
      if (c0 + e < valid) mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(r[e]));
    } else {
      mx[e & 3] = fmaxf(mx[e & 3], __uint_as_float(r[e]));
    }
  }
}

template <int BKV, bool kMask>
__device__ __forceinline__ float row_max_pass(uint32_t tS, float m, int valid) {
  float mx[4] = {m, m, m, m};
  uint32_t ra[16], rb[16];
  tmem_ld16(tS, ra);
#pragma unroll
  for (int c0 = 0; c0 < BKV; c0 += 32) {
    tmem_ld_wait();
    tmem_ld16(tS + c0 + 16, rb);
    max16<kMask>(ra, c0, valid, mx);
    tmem_ld_wait();
    if (c0 + 32 < BKV) tmem_ld16(tS + c0 + 32, ra);
    max16<kMask>(rb, c0 + 16, valid, mx);
  }
  return fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
}

// Pass 2: p = exp2(s - max), row sum, P written as 16-bit into the 128B-swizzled K-major smem tile.
template <bool kMask, bool kBf16>
__device__ __forceinline__ void exp16(const uint32_t (&r)[16], int c0, int valid, float mx, float (&sum)[2], uint8_t* sP,
                                      int row) {
  uint32_t pk[8];
#pragma unroll
  for (int e = 0; e < 16; e += 2) {
    float p0 = ex2_approx(__uint_as_float(r[e]) - mx);
    float p1 = ex2_approx(__uint_as_float(r[e + 1]) - mx);
    if (kMask) {
      p0 = (c0 + e < valid) ? p0 : 0.f;
      p1 = (c0 + e + 1 < valid) ? p1 : 0.f;
    }
    sum[0] += p0;
    sum[1] += p1;
    pk[e >> 1] = kBf16 ? pack_bf16(p0, p1) : pack_f16(p0, p1);
  }
  // 16 keys = 32 B = two 16-byte units of this row inside the 64-key chunk (128 B per row)
  uint8_t* chunk = sP + (c0 >> 6) * (128 * 128) + row * 128;
  const int u0 = (c0 & 63) >> 3;  // first 16-byte unit
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int pos = (u0 + u) ^ (row & 7);
    *reinterpret_cast<uint4*>(chunk + pos * 16) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
  }
}

template <int BKV, bool kMask, bool kBf16>
__device__ __forceinline__ float exp_store_pass(uint32_t tS, float mx, int valid, uint8_t* sP, int row) {
  float sum[2] = {0.f, 0.f};
  uint32_t ra[16], rb[16];
  tmem_ld16(tS, ra);
#pragma unroll
  for (int c0 = 0; c0 < BKV; c0 += 32) {
    tmem_ld_wait();
    tmem_ld16(tS + c0 + 16, rb);
    exp16<kMask, kBf16>(ra, c0, valid, mx, sum, sP, row);
    tmem_ld_wait();
    if (c0 + 32 < BKV) tmem_ld16(tS + c0 + 32, ra);
    exp16<kMask, kBf16>(rb, c0 + 16, valid, mx, sum, sP, row);
  }
  return sum[0] + sum[1];
}

template <int HDP, int HD, int BKV>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_kernel(const __grid_constant__ AttnParams p) {
  constexpr int HDC = HDP / 64;              // 64-element chunks along head dim
  constexpr int KVC = BKV / 64;              // 64-key chunks per KV tile
  constexpr uint32_t kQBytes = 128 * HDP * 2;
  constexpr uint32_t kKBytes = BKV * HDP * 2;
  constexpr uint32_t kStageBytes = 2 * kKBytes;  // K tile + V^T tile
  constexpr uint32_t kPBytes = 128 * BKV * 2;
  constexpr uint32_t kTmemCols = 256;        // S0: [0, BKV)  S1: [BKV, 2 BKV)  O: [2 BKV, 2 BKV + HDP)
  static_assert(2 * BKV + HDP <= 256, "TMEM budget");

  // 1024-byte alignment (128B swizzle atoms) comes from the declaration; no slack is added so that two CTAs
  // (2 x ~113 KB) fit one SM
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = smem_raw;
  uint8_t* sQ = sm;
  uint8_t* sKV = sQ + kQBytes;
  uint8_t* sP = sKV + 2 * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2   K tile landed
  uint64_t* k_empty = bars + 3;     // 2   S MMA that read it has completed
  uint64_t* v_full = bars + 5;      // 2
  uint64_t* v_empty = bars + 7;     // 2   PV MMA that read it has completed
  uint64_t* s_full = bars + 9;      // 2   S tile (double-buffered in TMEM) ready
  uint64_t* p_full = bars + 11;
  uint64_t* o_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, ns = blockIdx.z;
  const int bh = ns * p.H + h;
  const int n_self = (p.T + BKV - 1) / BKV, n_cross = (p.L + BKV - 1) / BKV;
  const int n_tiles = n_self + n_cross;

  if (warp == 4) {
    if (lane == 0) {
      prefetch_tmap(&p.tmQ); prefetch_tmap(&p.tmK); prefetch_tmap(&p.tmVt);
      prefetch_tmap(&p.tmKy); prefetch_tmap(&p.tmVyt);
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        mbar_init(&s_full[i], 1);
      }
      mbar_init(p_full, 128);
      mbar_init(o_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 2 * BKV;

  if (warp == 4) {
    if (elect_one()) {
      auto load_k = [&](int i) {
        const int st = i & 1;
        const bool cross = i >= n_self;
        const int kv0 = (cross ? i - n_self : i) * BKV;
        uint8_t* dK = sKV + st * kStageBytes;
        mbar_arrive_expect_tx(&k_full[st], kKBytes);
#pragma unroll
        for (int c = 0; c < HDC; ++c)
          tma_load_3d(dK + c * (BKV * 128), cross ? &p.tmKy : &p.tmK, &k_full[st], c * 64, kv0, bh);
      };
      auto load_v = [&](int i) {
        const int st = i & 1;
        const bool cross = i >= n_self;
        const int kv0 = (cross ? i - n_self : i) * BKV;
        uint8_t* dV = sKV + st * kStageBytes + kKBytes;
        mbar_arrive_expect_tx(&v_full[st], kKBytes);
#pragma unroll
        for (int c = 0; c < KVC; ++c)
          tma_load_3d(dV + c * (HDP * 128), cross ? &p.tmVyt : &p.tmVt, &v_full[st], kv0 + c * 64, 0, bh);
      };
      const uint32_t idesc_s = umma_idesc(128, BKV, p.dtype == MA3_BF16 ? 1 : 0);
      const uint32_t idesc_o = umma_idesc(128, HDP, p.dtype == MA3_BF16 ? 1 : 0);
      // S(i) = Q K_i^T into TMEM buffer i & 1; frees the K stage when done
      auto issue_s = [&](int i) {
        const int st = i & 1;
        mbar_wait(&k_full[st], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sKV + st * kStageBytes);
#pragma unroll
        for (int k = 0; k < HDP / 16; ++k) {
          const uint64_t da = umma_desc_kmajor(qa + (k / 4) * (128 * 128) + (k % 4) * 32, 128);
          const uint64_t db = umma_desc_kmajor(ka + (k / 4) * (BKV * 128) + (k % 4) * 32, 128);
          umma_f16(tmem_S + st * BKV, da, db, idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
      };

      mbar_arrive_expect_tx(q_full, kQBytes);
#pragma unroll
      for (int c = 0; c < HDC; ++c) tma_load_3d(sQ + c * (128 * 128), &p.tmQ, q_full, c * 64, q0, bh);
      load_k(0);
      load_v(0);
      if (n_tiles > 1) { load_k(1); load_v(1); }
      mbar_wait(q_full, 0);
      issue_s(0);
      if (n_tiles > 1) issue_s(1);   // the softmax of tile i overlaps the tensor core computing S(i+1)
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i & 1;
        if (i + 2 < n_tiles) {        // K stage st was released by S(i): refill it for tile i+2
          mbar_wait(&k_empty[st], (i >> 1) & 1);
          load_k(i + 2);
        }
        attn_trace(p, i, 8);
        mbar_wait(p_full, i & 1);     // P(i) in smem, S buffer st drained, O tile of PV(i-1) consumed
        attn_trace(p, i, 9);
        mbar_wait(&v_full[st], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP), va = smem_u32(sKV + st * kStageBytes + kKBytes);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t da = umma_desc_kmajor(pa + (k / 4) * (128 * 128) + (k % 4) * 32, 128);
          const uint64_t db = umma_desc_kmajor(va + (k / 4) * (HDP * 128) + (k % 4) * 32, 128);
          umma_f16(tmem_O, da, db, idesc_o, k != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&v_empty[st]);
        attn_trace(p, i, 10);
        if (i + 2 < n_tiles) {
          issue_s(i + 2);
          mbar_wait(&v_empty[st], (i >> 1) & 1);
          load_v(i + 2);
        }
      }
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // normalised self-attention result parked as packed bf16 pairs while the cross segment accumulates
    // (halves its register footprint so two CTAs fit per SM; it is rounded to 16 bits on output anyway)
    uint32_t stash[HD / 2];
#pragma unroll
    for (int e = 0; e < HD / 2; ++e) stash[e] = 0u;
    int it = 0;
    float acc[HD];
    for (int seg = 0; seg < 2; ++seg) {
      const int ntl = seg ? n_cross : n_self;
      const int kvlen = seg ? p.L : p.T;
      if (ntl == 0) continue;
#pragma unroll
      for (int e = 0; e < HD; ++e) acc[e] = 0.f;
      float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
      for (int j = 0; j < ntl; ++j, ++it) {
        const bool tr = warp == 0 && lane == 0;
        if (tr) attn_trace(p, it, 0);
        mbar_wait(&s_full[it & 1], (it >> 1) & 1);
        if (tr) attn_trace(p, it, 1);
        tc_fence_after();
        const uint32_t tS = tmem_S + (it & 1) * BKV + lane_base;
        const int valid = kvlen - j * BKV;        // >= BKV on every tile but the ragged last one of a segment
        const bool full = valid >= BKV;
        const float mx = full ? row_max_pass<BKV, false>(tS, m, valid) : row_max_pass<BKV, true>(tS, m, valid);
        const float alpha = ex2_approx(m - mx);
        if (tr) attn_trace(p, it, 2);
        if (j > 0) {
          mbar_wait(o_full, (it - 1) & 1);
          if (tr) attn_trace(p, it, 3);
          tc_fence_after();
          accumulate_o<HD>(tmem_O + lane_base, acc, alpha_prev);
        }
        if (tr) attn_trace(p, it, 4);
        float rowsum;
        if (p.dtype == MA3_BF16)
          rowsum = full ? exp_store_pass<BKV, false, true>(tS, mx, valid, sP, row)
                        : exp_store_pass<BKV, true, true>(tS, mx, valid, sP, row);
        else
          rowsum = full ? exp_store_pass<BKV, false, false>(tS, mx, valid, sP, row)
                        : exp_store_pass<BKV, true, false>(tS, mx, valid, sP, row);
        l = l * alpha + rowsum;
        m = mx;
        alpha_prev = alpha;
        if (tr) attn_trace(p, it, 5);
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(p_full);
        if (tr) attn_trace(p, it, 6);
      }
      mbar_wait(o_full, (it - 1) & 1);
      tc_fence_after();
      accumulate_o<HD>(tmem_O + lane_base, acc, alpha_prev);
      const float f = (seg ? tanhf(p.gate[h]) : 1.f) / l;
      if (seg == 0 && n_cross > 0) {
#pragma unroll
        for (int e = 0; e < HD; e += 2) stash[e >> 1] = pack_bf16(acc[e] * f, acc[e + 1] * f);
      } else {
#pragma unroll
        for (int e = 0; e < HD; e += 2) {
          const float2 sv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&stash[e >> 1]));
          acc[e] = fmaf(acc[e], f, sv.x);
          acc[e + 1] = fmaf(acc[e + 1], f, sv.y);
        }
      }
    }
    if (q0 + row < p.T) {
      const long long o = ((long long)ns * p.T + q0 + row) * p.D + (long long)h * HD;
#pragma unroll
      for (int e = 0; e < HD; e += 8) {
        uint4 u;
        if (p.dtype == MA3_BF16)
          u = make_uint4(pack_bf16(acc[e], acc[e + 1]), pack_bf16(acc[e + 2], acc[e + 3]),
                         pack_bf16(acc[e + 4], acc[e + 5]), pack_bf16(acc[e + 6], acc[e + 7]));
        else
          u = make_uint4(pack_f16(acc[e], acc[e + 1]), pack_f16(acc[e + 2], acc[e + 3]),
                         pack_f16(acc[e + 4], acc[e + 5]), pack_f16(acc[e + 6], acc[e + 7]));
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + o + e) = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------ v2 kernel
// Same contract and CTA shape as attn_kernel above, restructured to cut the softmax warps' instruction count ~3x
// (they, not the tensor core, bound the kernel at these head dims):
//  * S is read from TMEM once instead of once for the max and once for the exponentials, and P is written back into
//    tensor memory over the consumed S columns (tcgen05.st) and fed to the PV MMA as a TMEM A operand: no shared-memory
//    P tile, no async-proxy fence.  The tensor pipe does NOT order those A reads against a later MMA that overwrites
//    the same columns, so S(i+2) is issued only after PV(i) has completed;
//  * O stays in TMEM and the PV MMAs accumulate into it.  The running maximum is only raised when a tile exceeds it
//    by more than 2^8 (lazy rescaling: p <= 256 is harmless for bf16 P and the fp32 accumulator, and the result is
//    exact because numerator and denominator share the stale reference); the rare correction multiplies O in place;
//  * the softmax denominator is accumulated by the tensor core: V^T carries a row of ones at index HD, so column HD of
//    O is sum_j p_ij of exactly the rounded probabilities the MMA consumed;
//  * S = Q K^T only issues the ceil(HD/16) k-steps that carry data, and PV only N = HDO columns;
//  * warps whose 32 query rows all lie beyond T (ragged last Q tile) do no work.
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { return __fadd2_rn(a, b); }
// 2^x for a pair of logits on the FMA pipe (x <= ~8; clamped below at -120): round-to-nearest split x = i + f with the
// 1.5 * 2^23 trick, degree-3 minimax of 2^f on [-0.5, 0.5] (max relative error 7.5e-5, far below the 16-bit rounding of
// P), and i added into the exponent field.  Takes a quarter of a tile's exponentials off the MUFU pipe, which three
// softmax warps per scheduler otherwise saturate.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float kM = 12582912.f;
  x.x = fmaxf(x.x, -120.f);
  x.y = fmaxf(x.y, -120.f);
  const float2 t = fadd2(x, make_float2(kM, kM));
  const float2 xi = fadd2(t, make_float2(-kM, -kM));
  const float2 f = __ffma2_rn(xi, make_float2(-1.f, -1.f), x);
  float2 q = __ffma2_rn(f, make_float2(0.05517165f, 0.05517165f), make_float2(0.24261112f, 0.24261112f));
  q = __ffma2_rn(q, f, make_float2(0.69326099f, 0.69326099f));
  q = __ffma2_rn(q, f, make_float2(0.99992807f, 0.99992807f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}

// O[:, 0..HDO) *= alpha, in place in TMEM (thread <-> row)
template <int HDO>
__device__ __forceinline__ void rescale_o(uint32_t tO, float alpha) {
#pragma unroll
  for (int c = 0; c < HDO; c += 16) {
    uint32_t r[16];
    tmem_ld16(tO + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
    tmem_st16(tO + c, r);
  }
  tmem_st_wait();
}

// Segment epilogue over O columns [OFF, HD) in pieces of 16 / 8: kFinal = false parks O * f as packed bf16 pairs,
// kFinal = true adds the parked self-attention result and writes the 16-bit output row.
template <int HD, bool kFinal, int OFF = 0>
__device__ __forceinline__ void drain_o(uint32_t tO, float f, uint32_t (&stash)[HD / 2], uint16_t* orow, bool bf16,
                                        bool store) {
  if constexpr (OFF < HD) {
    constexpr int W = (HD - OFF >= 16) ? 16 : 8;
    uint32_t r[W];
    tmem_ld_n<W>(tO + OFF, r);
    tmem_ld_wait();
    if constexpr (!kFinal) {
#pragma unroll
      for (int e = 0; e < W; e += 2)
        stash[(OFF + e) >> 1] = pack_bf16(__uint_as_float(r[e]) * f, __uint_as_float(r[e + 1]) * f);
    } else {
#pragma unroll
      for (int e = 0; e < W; e += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float2 sv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&stash[(OFF + e + i) >> 1]));
          v[i] = fmaf(__uint_as_float(r[e + i]), f, sv.x);
          v[i + 1] = fmaf(__uint_as_float(r[e + i + 1]), f, sv.y);
        }
        uint4 u;
        if (bf16)
          u = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        else
          u = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
        if (store) *reinterpret_cast<uint4*>(orow + OFF + e) = u;
      }
    }
    drain_o<HD, kFinal, OFF + W>(tO, f, stash, orow, bf16, store);
  }
}

constexpr int kAttn2Threads = 288;   // warps 0-7: softmax, two threads per query row (32 keys each); warp 8: TMA + MMA

// Named barrier of one softmax warp pair.  The id must be an immediate: with a register operand ptxas cannot see which
// barriers the kernel uses, reports "used 1 barriers", the hardware then reserves a single barrier per CTA and ids
// 1-4 alias the barriers of the co-resident CTA (observed as intermittent wrong rows in the second CTA of an SM).
__device__ __forceinline__ void pair_bar_sync(int qw) {
  switch (qw) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

// Segment epilogue over this thread's O columns [C0, C1) in pieces of 8: kFinal = false parks O * f as packed bf16
// pairs, kFinal = true adds the parked self-attention result and writes the 16-bit output row.
template <int C0, int C1, bool kFinal, int NS>
__device__ __forceinline__ void drain_cols(uint32_t tO, float f, uint32_t (&stash)[NS], uint16_t* orow, bool bf16,
                                           bool store) {
  if constexpr (C0 < C1) {
    uint32_t r[8];
    tmem_ld8(tO + C0, r);
    tmem_ld_wait();
    constexpr int SI = (C0 % (2 * NS)) / 2;   // stash index of this group's first pair (thread-local)
    if constexpr (!kFinal) {
#pragma unroll
      for (int e = 0; e < 8; e += 2)
        stash[SI + (e >> 1)] = pack_bf16(__uint_as_float(r[e]) * f, __uint_as_float(r[e + 1]) * f);
    } else {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        const float2 sv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&stash[SI + (e >> 1)]));
        v[e] = fmaf(__uint_as_float(r[e]), f, sv.x);
        v[e + 1] = fmaf(__uint_as_float(r[e + 1]), f, sv.y);
      }
      uint4 u;
      if (bf16)
        u = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      else
        u = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
      if (store) *reinterpret_cast<uint4*>(orow + C0) = u;
    }
    drain_cols<C0 + 8, C1, kFinal, NS>(tO, f, stash, orow, bf16, store);
  }
}

// O[:, C0..C1) *= alpha in place (8-column pieces)
template <int C0, int C1>
__device__ __forceinline__ void rescale_cols(uint32_t tO, float alpha) {
  if constexpr (C0 < C1) {
    uint32_t r[8];
    tmem_ld8(tO + C0, r);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
    tmem_st8(tO + C0, r);
    rescale_cols<C0 + 8, C1>(tO, alpha);
  }
}

template <int HDP, int HD, int BKV>
__global__ void __launch_bounds__(kAttn2Threads, 2) attn2_kernel(const __grid_constant__ AttnParams p) {
  // BKV = 64: one 64-key chunk per KV tile.  BKV = 80: 312 latent tokens are 4 tiles instead of 5 and 154 context tokens 2
  // instead of 3 (6 KV tiles per CTA instead of 8: the per-tile synchronisation chain is the cost, not the MMAs); the 16
  // extra keys of V^T live in a second, 32B-swizzled chunk and each softmax thread owns 40 keys instead of 32.
  static_assert(BKV == 64 || BKV == 80, "KV tile of 64 or 80 keys");
  constexpr int CPT = BKV / 2;                   // S columns (keys) per softmax thread
  constexpr int HDC = HDP / 64;
  constexpr int HDO = (HD + 1 + 15) / 16 * 16;   // O columns: HD values, the row-sum column, zero padding
  constexpr int KS = (HD + 15) / 16;             // k-steps of S = Q K^T that carry data (pad columns are zero)
  static_assert(HDO <= HDP, "needs a spare V^T row for the row sums");
  // column split of O between the two threads of a row (multiples of 8; thread 0 takes the larger part)
  constexpr int HSPLIT = ((HD / 8 + 1) / 2) * 8;
  constexpr int NSTASH = HSPLIT / 2;
  constexpr uint32_t kQBytes = 128 * HDP * 2;
  constexpr uint32_t kKBytes = BKV * HDP * 2;
  constexpr uint32_t kV1Bytes = HDO * 128;       // only the HDO rows of V^T the PV MMA reads are staged (keys 0..63)
  constexpr uint32_t kV2Bytes = BKV > 64 ? HDO * 32 : 0;   // keys 64..79
  constexpr uint32_t kVBytes = kV1Bytes + kV2Bytes;
  constexpr uint32_t kStageBytes = kKBytes + kVBytes;
  constexpr uint32_t kTmemCols = 256;            // S0: [0, BKV)  S1: [BKV, 2 BKV)  O: [2 BKV, 2 BKV + HDO)
  static_assert(2 * BKV + HDO <= 256, "TMEM budget");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sQ = smem_raw;
  uint8_t* sKV = sQ + kQBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + 2 * kStageBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 11;     // 2, by tile parity
  uint64_t* o_full = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* xch = reinterpret_cast<float*>(bars + 16);   // [tile parity][column half][row]: per-tile row maxima

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, ns = blockIdx.z;
  const int bh = ns * p.H + h;
  const int n_self = (p.T + BKV - 1) / BKV, n_cross = (p.L + BKV - 1) / BKV;
  const int n_tiles = n_self + n_cross;
  const int n_active = min(4, (p.T - q0 + 31) / 32);   // 32-row groups owning at least one query row < T
  // diagnostics: per-CTA (SM id, start, end) in nanoseconds of the global timer (tools/probe_trace_attn.py)
  long long* cta_rec = nullptr;
  if (p.trace && threadIdx.x == 0) {
    cta_rec = p.trace + 256 + 4 * ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
    unsigned smid;
    unsigned long long t0;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    cta_rec[0] = smid;
    cta_rec[1] = (long long)t0;
    cta_rec[3] = clock64();
  }

  if (warp == 8) {
    if (lane == 0) {
      prefetch_tmap(&p.tmQ); prefetch_tmap(&p.tmK); prefetch_tmap(&p.tmVt);
      prefetch_tmap(&p.tmKy); prefetch_tmap(&p.tmVyt);
      mbar_init(q_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
        mbar_init(&s_full[i], 1);
      }
      mbar_init(&p_full[0], 2 * n_active);   // one arrival per active softmax warp
      mbar_init(&p_full[1], 2 * n_active);
      mbar_init(o_full, 1);
      fence_barrier_init();
      // the first loads (Q, K/V tiles 0 and 1) go out before the TMEM allocation and the CTA-wide sync, so their
      // L2 / HBM latency overlaps the rest of the prologue
      pdl_wait();
      mbar_arrive_expect_tx(q_full, kQBytes);
#pragma unroll
      for (int c = 0; c < HDC; ++c) tma_load_3d(sQ + c * (128 * 128), &p.tmQ, q_full, c * 64, q0, bh);
      for (int i = 0; i < 2 && i < n_tiles; ++i) {
        const bool cross = i >= n_self;
        const int kv0 = (cross ? i - n_self : i) * BKV;
        uint8_t* dK = sKV + i * kStageBytes;
        mbar_arrive_expect_tx(&k_full[i], kKBytes);
#pragma unroll
        for (int c = 0; c < HDC; ++c)
          tma_load_3d(dK + c * (BKV * 128), cross ? &p.tmKy : &p.tmK, &k_full[i], c * 64, kv0, bh);
        mbar_arrive_expect_tx(&v_full[i], kVBytes);
        tma_load_3d(dK + kKBytes, cross ? &p.tmVyt : &p.tmVt, &v_full[i], kv0, 0, bh);
        if constexpr (BKV > 64)
          tma_load_3d(dK + kKBytes + kV1Bytes, cross ? &p.tmVyt2 : &p.tmVt2, &v_full[i], kv0 + 64, 0, bh);
      }
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 2 * BKV;

  if (warp == 8) {
    if (elect_one()) {
      auto load_k = [&](int i) {
        const int st = i & 1;
        const bool cross = i >= n_self;
        const int kv0 = (cross ? i - n_self : i) * BKV;
        uint8_t* dK = sKV + st * kStageBytes;
        mbar_arrive_expect_tx(&k_full[st], kKBytes);
#pragma unroll
        for (int c = 0; c < HDC; ++c)
          tma_load_3d(dK + c * (BKV * 128), cross ? &p.tmKy : &p.tmK, &k_full[st], c * 64, kv0, bh);
      };
      auto load_v = [&](int i) {
        const int st = i & 1;
        const bool cross = i >= n_self;
        const int kv0 = (cross ? i - n_self : i) * BKV;
        uint8_t* dV = sKV + st * kStageBytes + kKBytes;
        mbar_arrive_expect_tx(&v_full[st], kVBytes);
        tma_load_3d(dV, cross ? &p.tmVyt : &p.tmVt, &v_full[st], kv0, 0, bh);
        if constexpr (BKV > 64) tma_load_3d(dV + kV1Bytes, cross ? &p.tmVyt2 : &p.tmVt2, &v_full[st], kv0 + 64, 0, bh);
      };
      const uint32_t idesc_s = umma_idesc(128, BKV, p.dtype == MA3_BF16 ? 1 : 0);
      const uint32_t idesc_o = umma_idesc(128, HDO, p.dtype == MA3_BF16 ? 1 : 0);
      // descriptors as (lo, shared hi) 32-bit halves: only the start-address field of lo changes between MMAs, so the
      // single issuing thread spends one integer add per operand per MMA (see the tap-GEMM issue loop)
      const uint64_t dq0 = umma_desc_kmajor(smem_u32(sQ), 128);
      const uint32_t dhi = (uint32_t)(dq0 >> 32), q_lo = (uint32_t)dq0;
      const uint32_t k_lo = (uint32_t)umma_desc_kmajor(smem_u32(sKV), 128);
      const uint32_t v_lo = (uint32_t)umma_desc_kmajor(smem_u32(sKV + kKBytes), 128);
      const uint64_t dv2 = umma_desc_kmajor(smem_u32(sKV + kKBytes + kV1Bytes), 32);   // 32B-swizzled chunk of keys 64..79
      const uint32_t v2_lo = (uint32_t)dv2, dhi32 = (uint32_t)(dv2 >> 32);
      auto issue_s = [&](int i) {
        const int st = i & 1;
        mbar_wait(&k_full[st], (i >> 1) & 1);
        const uint32_t b_lo = k_lo + (uint32_t)st * (kStageBytes >> 4);
#pragma unroll
        for (int k = 0; k < KS; ++k)
          umma_f16_lohi<1>(tmem_S + st * BKV, q_lo + (uint32_t)((k / 4) * (128 * 128 / 16) + (k % 4) * 2),
                           b_lo + (uint32_t)((k / 4) * (BKV * 128 / 16) + (k % 4) * 2), dhi, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
      };

      mbar_wait(q_full, 0);   // Q and the first two K / V tiles were requested in the prologue
      issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int i = 0; i < n_tiles; ++i) {
        const int st = i & 1;
        if (i + 2 < n_tiles) {
          mbar_wait(&k_empty[st], (i >> 1) & 1);
          load_k(i + 2);
        }
        attn_trace(p, i, 8);
        mbar_wait(&p_full[st], (i >> 1) & 1);   // P(i) in TMEM, S buffer st drained, O rescaled / consumed as needed
        attn_trace(p, i, 9);
        mbar_wait(&v_full[st], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t dv = v_lo + (uint32_t)st * (kStageBytes >> 4);
        const uint32_t fresh = (i == 0 || i == n_self) ? 0u : 1u;   // first tile of a segment overwrites O
        // P(i) lives in tensor memory, packed two keys per column over the first BKV / 2 columns of S buffer st
        const uint32_t tP = tmem_S + st * BKV;
        umma_f16_ts(tmem_O, tP, dv, dhi, idesc_o, fresh);
        umma_f16_ts(tmem_O, tP + 8, dv + 2, dhi, idesc_o, 1u);
        umma_f16_ts(tmem_O, tP + 16, dv + 4, dhi, idesc_o, 1u);
        umma_f16_ts(tmem_O, tP + 24, dv + 6, dhi, idesc_o, 1u);
        if constexpr (BKV > 64)
          umma_f16_ts(tmem_O, tP + 32, v2_lo + (uint32_t)st * (kStageBytes >> 4), dhi32, idesc_o, 1u);
        umma_commit(o_full);
        umma_commit(&v_empty[st]);
        attn_trace(p, i, 10);
        if (i + 2 < n_tiles) {
          // S(i+2) overwrites the TMEM buffer PV(i) reads P from: issue it only once PV(i) has completed
          mbar_wait(&v_empty[st], (i >> 1) & 1);
          load_v(i + 2);
          issue_s(i + 2);
        }
      }
    }
  } else if ((warp & 3) < n_active) {
    const int qw = warp & 3, half = warp >> 2;      // TMEM lane quarter; which 32 keys of a tile / which O columns
    const int row = qw * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qw * 32) << 16;
    const uint32_t tO = tmem_O + lane_base;
    const bool bf16 = p.dtype == MA3_BF16;
    uint32_t stash[NSTASH];
#pragma unroll
    for (int e = 0; e < NSTASH; ++e) stash[e] = 0u;
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + ((long long)ns * p.T + q0 + row) * p.D + (long long)h * HD;
    const bool row_ok = q0 + row < p.T;
    int it = 0;
    for (int seg = 0; seg < 2; ++seg) {
      const int ntl = seg ? n_cross : n_self;
      const int kvlen = seg ? p.L : p.T;
      if (ntl == 0) continue;
      float m = 0.f;
      for (int j = 0; j < ntl; ++j, ++it) {
        const bool tr = warp == 0 && lane == 0;
        if (tr) attn_trace(p, it, 0);
        mbar_wait(&s_full[it & 1], (it >> 1) & 1);
        if (tr) attn_trace(p, it, 1);
        tc_fence_after();
        uint32_t s[CPT];
        tmem_ld_n<32>(tmem_S + (it & 1) * BKV + half * CPT + lane_base, s);
        if constexpr (CPT > 32) tmem_ld_n<8>(tmem_S + (it & 1) * BKV + half * CPT + 32 + lane_base, s + 32);
        tmem_ld_wait();
        const int valid = kvlen - j * BKV - half * CPT;   // keys of this thread's half that exist
        if (valid < CPT) {   // ragged last tile of a segment: keys beyond the sequence get -inf
#pragma unroll
          for (int e = 0; e < CPT; ++e)
            if (e >= valid) s[e] = 0xff800000u;
        }
        float mx4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) mx4[c] = fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 4]));
#pragma unroll
        for (int e = 8; e < CPT; e += 8) {
#pragma unroll
          for (int c = 0; c < 4; ++c) mx4[c] = fmax3(mx4[c], __uint_as_float(s[e + c]), __uint_as_float(s[e + c + 4]));
        }
        float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);
        // row maximum over both halves: exchange with the partner thread (warp ^ 4) through shared memory
        float* xs = xch + (it & 1) * 256;
        xs[half * 128 + row] = mx;
        pair_bar_sync(qw);
        mx = fmaxf(mx, xs[(half ^ 1) * 128 + row]);
        if (tr) attn_trace(p, it, 2);
        if (j == 0) {
          m = mx;
        } else if (__any_sync(0xffffffffu, mx > m + 8.f)) {   // identical decision in both warps of the pair
          mbar_wait(o_full, (it - 1) & 1);   // PV(it-1) complete: O may be modified
          tc_fence_after();
          const float m_new = fmaxf(m, mx);
          const float alpha = ex2_approx(m - m_new);
          if (half == 0) rescale_cols<0, HSPLIT>(tO, alpha);
          else rescale_cols<HSPLIT, HDO>(tO, alpha);
          tmem_st_wait();
          m = m_new;
        }
        if (tr) attn_trace(p, it, 3);
        uint32_t pk[CPT / 2];
        const float2 nm = make_float2(-m, -m);
        if (bf16) {
#pragma unroll
          for (int e = 0; e < CPT; e += 2) {
            const float2 d = fadd2(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), nm);
            pk[e >> 1] = pack_bf16(ex2_approx(d.x), ex2_approx(d.y));
          }
        } else {
#pragma unroll
          for (int e = 0; e < CPT; e += 2) {
            const float2 d = fadd2(make_float2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), nm);
            pk[e >> 1] = pack_f16(ex2_approx(d.x), ex2_approx(d.y));
          }
        }
        if (tr) attn_trace(p, it, 4);
        // P goes back into tensor memory over the S columns both threads of the row have already consumed (the named
        // barrier above orders the partner's S load before this store): 16 packed columns per thread, read by the PV
        // MMA as its A operand.  No shared-memory tile, no async-proxy fence; S(it+2), which overwrites this buffer,
        // is issued after PV(it) and the tensor pipe executes in issue order.
        tmem_st16(tmem_S + (it & 1) * BKV + half * (CPT / 2) + lane_base, *reinterpret_cast<uint32_t(*)[16]>(pk));
        if constexpr (CPT > 32)
          tmem_st4(tmem_S + (it & 1) * BKV + half * (CPT / 2) + 16 + lane_base, *reinterpret_cast<uint32_t(*)[4]>(pk + 16));
        tmem_st_wait();
        // Every warp observes every phase of o_full, in order, before it arrives for the next tile: a parity wait is
        // only meaningful when the waiter is at most one phase behind, and the waits of the correction path and of the
        // segment end would otherwise see a barrier two phases back and fall through (measured: intermittent wrong
        // rows).  PV(it-1) was issued a whole tile ago, so this wait is normally already satisfied.
        if (it > 0) mbar_wait(o_full, (it - 1) & 1);
        if (tr) attn_trace(p, it, 5);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[it & 1]);
        if (tr) attn_trace(p, it, 6);
      }
      mbar_wait(o_full, (it - 1) & 1);
      tc_fence_after();
      uint32_t rl[8];
      tmem_ld8(tO + (HD / 8) * 8, rl);
      tmem_ld_wait();
      const float l = __uint_as_float(rl[HD % 8]);
      const float f = __fdividef(seg ? tanhf(p.gate[h]) : 1.f, l);
      const bool park = seg == 0 && n_cross > 0;
      if (half == 0) {
        if (park) drain_cols<0, HSPLIT, false, NSTASH>(tO, f, stash, orow, bf16, row_ok);
        else drain_cols<0, HSPLIT, true, NSTASH>(tO, f, stash, orow, bf16, row_ok);
      } else {
        if (park) drain_cols<HSPLIT, HD, false, NSTASH>(tO, f, stash, orow, bf16, row_ok);
        else drain_cols<HSPLIT, HD, true, NSTASH>(tO, f, stash, orow, bf16, row_ok);
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cta_rec) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    cta_rec[2] = (long long)t1;
    cta_rec[3] = clock64() - cta_rec[3];
  }
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ v3 kernel
// One CTA per (sample, head, group of NCTX consecutive 128-query tiles), one CTA per SM, all of tensor memory.
// The v2 kernel runs at a quarter of its MUFU / issue floor: a CTA is one latency chain per KV tile (S -> max ->
// partner exchange -> exp -> P -> PV -> next S), TMEM caps an SM at two such CTAs, every CTA pays its own prologue and
// re-loads K / V^T, and a ragged last query tile costs a whole CTA.  Here the NCTX query tiles of a head are NCTX
// independent softmax streams inside one CTA that share every K / V^T tile (loaded once per head instead of once per
// query tile); a softmax thread owns a whole query row of an 80-key tile (no partner exchange, no named barriers, 80
// exponentials per ~800 clocks of per-tile synchronisation instead of 40); every context has its own MMA-issuing warp;
// and P goes to the PV MMA through SHARED memory (128B / 32B-swizzled K-major tile written by the softmax threads), not
// over the S columns: S (single-buffered: 3 x (80 S + 80 O) = 480 TMEM columns) is handed back to the tensor core as
// soon as the softmax warps hold it in registers, so S(c, i+1) is computed under the exponentials of tile i and the
// PV -> next-S turnaround of the tensor pipe (~1400 clocks) leaves the critical path.
//   warps [0, 4 NCTX):        softmax, warp w <-> context w / 4, TMEM lane quarter w % 4
//   warp 4 NCTX:              TMA producer (Q tiles, K ring, V^T ring)
//   warps 4 NCTX + 1 + c:     MMA issue for context c (one elected lane each; the first one owns the TMEM allocation)
// Lazy rescaling, row sums from the ones row of V^T, keys 64..79 of V^T (and of P) as a 32B-swizzled chunk: as in v2.  The normalised self-attention result is parked in the
// output row (16-bit) across the cross segment and read back by the same thread at the end.
template <int HDP, int HD, int NCTX>
__global__ void __launch_bounds__(NCTX * 128 + 32 + NCTX * 32, 1) attn3_kernel(const __grid_constant__ AttnParams p) {
  constexpr int BKV = 80;
  constexpr int NST = 2;                         // K / V^T stages (K and V^T halves are released separately)
  constexpr int HDC = HDP / 64;
  constexpr int HDO = (HD + 1 + 15) / 16 * 16;
  constexpr int KS = (HD + 15) / 16;
  constexpr int CW = BKV + HDO;                  // TMEM columns of one context: S [0, 80)  O [80, 80 + HDO)
  constexpr uint32_t kP1Bytes = 128 * 128;       // P, keys 0..63: 128 rows x 128 B, 128B swizzle
  constexpr uint32_t kPBytes = kP1Bytes + 128 * 32;   // + keys 64..79: 128 rows x 32 B, 32B swizzle
  static_assert(NCTX * CW <= 512, "TMEM budget");
  static_assert(HDO <= HDP, "needs a spare V^T row for the row sums");
  constexpr uint32_t kQBytes = 128 * HDP * 2;
  constexpr uint32_t kKBytes = BKV * HDP * 2;
  constexpr uint32_t kV1Bytes = HDO * 128;       // keys 0..63 of a V^T tile (128B swizzle)
  constexpr uint32_t kV2Bytes = HDO * 32;        // keys 64..79 (32B swizzle)
  constexpr uint32_t kVBytes = kV1Bytes + kV2Bytes;
  constexpr uint32_t kStageBytes = (kKBytes + kVBytes + 1023) / 1024 * 1024;
  constexpr int kSoftWarps = 4 * NCTX;
  constexpr int kPolyPairs = MA3_ATTN_POLY_PAIRS;   // of the 8 key pairs of a 16-key group

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sQ = smem_raw;
  uint8_t* sKV = sQ + NCTX * kQBytes;
  uint8_t* sP = sKV + NST * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NCTX * kPBytes);
  uint64_t* q_full = bars;                  // NCTX (one per query tile, so that context 0 starts on its own tile)
  uint64_t* k_full = q_full + NCTX;         // NST
  uint64_t* k_empty = k_full + NST;         // NST, one commit per active context
  uint64_t* v_full = k_empty + NST;         // NST
  uint64_t* v_empty = v_full + NST;         // NST, one commit per active context
  uint64_t* s_full = v_empty + NST;         // NCTX: phase i completes with S(c, i)
  uint64_t* s_free = s_full + NCTX;         // NCTX: S(c, i) is in registers (one arrival per active softmax warp)
  uint64_t* p_full = s_free + NCTX;         // NCTX: P(c, i) is in shared memory (one arrival per active softmax warp)
  uint64_t* pv_done = p_full + NCTX;        // NCTX: phase i completes with PV(c, i)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + NCTX);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt0 = blockIdx.x * NCTX, h = blockIdx.y, ns = blockIdx.z;
  const int bh = ns * p.H + h;
  const int n_self = (p.T + BKV - 1) / BKV, n_cross = (p.L + BKV - 1) / BKV;   // KV tiles per segment
  const int n_tiles = n_self + n_cross;
  const int n_ctx = min(NCTX, (p.T + 127) / 128 - qt0);   // contexts owning at least one query row
  long long* cta_rec = nullptr;
  if (p.trace && threadIdx.x == 0) {
    cta_rec = p.trace + 256 + 4 * ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x);
    unsigned smid;
    unsigned long long t0;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    cta_rec[0] = smid;
    cta_rec[1] = (long long)t0;
    cta_rec[3] = clock64();
  }

  if (warp == kSoftWarps && lane == 0) {
    prefetch_tmap(&p.tmQ); prefetch_tmap(&p.tmK); prefetch_tmap(&p.tmVt); prefetch_tmap(&p.tmVt2);
    prefetch_tmap(&p.tmKy); prefetch_tmap(&p.tmVyt); prefetch_tmap(&p.tmVyt2);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], n_ctx);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], n_ctx);
    }
    for (int c = 0; c < NCTX; ++c) {
      const int nw = max(1, min(4, (p.T - (qt0 + c) * 128 + 31) / 32));
      mbar_init(&q_full[c], 1);
      mbar_init(&s_full[c], 1);
      mbar_init(&s_free[c], nw);
      mbar_init(&p_full[c], nw);
      mbar_init(&pv_done[c], 1);
    }
    fence_barrier_init();
  }
  if (warp == kSoftWarps + 1) tmem_alloc(tmem_slot, 512);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == kSoftWarps) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      auto load_q = [&](int c) {
        mbar_arrive_expect_tx(&q_full[c], kQBytes);
#pragma unroll
        for (int ch = 0; ch < HDC; ++ch)
          tma_load_3d(sQ + c * kQBytes + ch * (128 * 128), &p.tmQ, &q_full[c], ch * 64, (qt0 + c) * 128, bh);
      };
      load_q(0);
      int st = 0, ph = 0;   // ring position; ph = parity of the stage's current use
      for (int g = 0; g < n_tiles; ++g) {
        const bool cross = g >= n_self;
        const int kv0 = (cross ? g - n_self : g) * BKV;
        uint8_t* dK = sKV + st * kStageBytes;
        if (g >= NST) mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kKBytes);
#pragma unroll
        for (int ch = 0; ch < HDC; ++ch)
          tma_load_3d(dK + ch * (BKV * 128), cross ? &p.tmKy : &p.tmK, &k_full[st], ch * 64, kv0, bh);
        if (g == 0)
          for (int c = 1; c < n_ctx; ++c) load_q(c);
        if (g >= NST) mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kVBytes);
        tma_load_3d(dK + kKBytes, cross ? &p.tmVyt : &p.tmVt, &v_full[st], kv0, 0, bh);
        tma_load_3d(dK + kKBytes + kV1Bytes, cross ? &p.tmVyt2 : &p.tmVt2, &v_full[st], kv0 + 64, 0, bh);
        if (++st == NST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp > kSoftWarps) {
    // ---------------------------------------------------------------- MMA issue, one warp (one elected lane) per context
    // (A single thread polling all contexts reacted ~1000 clocks late: a lone warp runs its dependent instruction stream at
    // a fraction of an instruction per clock.  Per context the events come in a fixed order, so each issuing thread
    // simply blocks on the next one.)
    const int c = warp - kSoftWarps - 1;
    if (c < n_ctx && elect_one()) {
      const uint32_t idesc_s = umma_idesc(128, BKV, p.dtype == MA3_BF16 ? 1 : 0);
      const uint32_t idesc_o = umma_idesc(128, HDO, p.dtype == MA3_BF16 ? 1 : 0);
      const uint64_t dq0 = umma_desc_kmajor(smem_u32(sQ + c * kQBytes), 128);
      const uint32_t dhi = (uint32_t)(dq0 >> 32), a_lo = (uint32_t)dq0;
      const uint32_t k_lo = (uint32_t)umma_desc_kmajor(smem_u32(sKV), 128);
      const uint32_t v_lo = (uint32_t)umma_desc_kmajor(smem_u32(sKV + kKBytes), 128);
      const uint64_t dv2 = umma_desc_kmajor(smem_u32(sKV + kKBytes + kV1Bytes), 32);
      const uint32_t v2_lo = (uint32_t)dv2, dhi32 = (uint32_t)(dv2 >> 32);
      const uint32_t tS = tmem_base + c * CW, tO = tS + BKV;
      const uint32_t p_lo = (uint32_t)umma_desc_kmajor(smem_u32(sP + c * kPBytes), 128);
      const uint32_t p2_lo = (uint32_t)umma_desc_kmajor(smem_u32(sP + c * kPBytes + kP1Bytes), 32);
      const bool trm = p.trace && n_tiles <= 16 && (c == 0 || c == 2) && blockIdx.x + blockIdx.y + blockIdx.z == 0;
      auto issue_s = [&](int it, int st, int ph) {
        mbar_wait(&k_full[st], ph);
        tc_fence_after();
        const uint32_t b_lo = k_lo + (uint32_t)st * (kStageBytes >> 4);
#pragma unroll
        for (int k = 0; k < KS; ++k)
          umma_f16_lohi<1>(tS, a_lo + (uint32_t)((k / 4) * (128 * 128 / 16) + (k % 4) * 2),
                           b_lo + (uint32_t)((k / 4) * (BKV * 128 / 16) + (k % 4) * 2), dhi, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[c]);
        umma_commit(&k_empty[st]);
        if (trm) p.trace[it * 16 + 8 + c] = clock64();
      };
      mbar_wait(&q_full[c], 0);
      issue_s(0, 0, 0);
      int st = 0, ph = 0;          // ring stage of tile it and the parity of that use
      for (int it = 0; it < n_tiles; ++it) {
        int st1 = st + 1, ph1 = ph;
        if (st1 == NST) { st1 = 0; ph1 ^= 1; }
        if (it + 1 < n_tiles) {
          mbar_wait(&s_free[c], it & 1);   // S(c, it) is in registers: its columns may be overwritten
          issue_s(it + 1, st1, ph1);
        }
        mbar_wait(&p_full[c], it & 1);     // P(c, it) is in shared memory (its writers waited for PV(c, it - 1))
        mbar_wait(&v_full[st], ph);
        tc_fence_after();
        const uint32_t dv = v_lo + (uint32_t)st * (kStageBytes >> 4);
        const uint32_t fresh = (it == 0 || it == n_self) ? 0u : 1u;   // first tile of a segment overwrites O
        umma_f16_lohi<1>(tO, p_lo, dv, dhi, idesc_o, fresh);
        umma_f16_lohi<1>(tO, p_lo + 2, dv + 2, dhi, idesc_o, 1u);
        umma_f16_lohi<1>(tO, p_lo + 4, dv + 4, dhi, idesc_o, 1u);
        umma_f16_lohi<1>(tO, p_lo + 6, dv + 6, dhi, idesc_o, 1u);
        umma_f16_lohi<1>(tO, p2_lo, v2_lo + (uint32_t)st * (kStageBytes >> 4), dhi32, idesc_o, 1u);
        umma_commit(&pv_done[c]);
        umma_commit(&v_empty[st]);
        if (trm) p.trace[it * 16 + 9 + c] = clock64();
        st = st1; ph = ph1;
      }
    }
  } else if (warp < 4 * n_ctx && (warp & 3) * 32 < p.T - (qt0 + (warp >> 2)) * 128) {
    // ---------------------------------------------------------------- softmax: thread <-> query row
    const int c = warp >> 2, qw = warp & 3;
    const int q0 = (qt0 + c) * 128;
    const int row = qw * 32 + lane;
    const uint32_t tS = tmem_base + c * CW + ((uint32_t)(qw * 32) << 16);
    const uint32_t tO = tS + BKV;
    const bool bf16 = p.dtype == MA3_BF16;
    const bool tr = p.trace != nullptr && n_tiles <= 16 && warp == 0 && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 &&
                    blockIdx.z == 0;
    uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + ((long long)ns * p.T + q0 + row) * p.D + (long long)h * HD;
    const bool row_ok = q0 + row < p.T;
    float m = 0.f;
    for (int it = 0; it < n_tiles; ++it) {
      const bool cross = it >= n_self;
      const int j = cross ? it - n_self : it;
      const int valid = (cross ? p.L : p.T) - j * BKV;   // keys of this tile that exist
      if (tr) p.trace[it * 16 + 0] = clock64();
      mbar_wait(&s_full[c], it & 1);
      if (it == 0 && c > 0 && p.stagger > 0) {
        const long long t_go = clock64() + (long long)c * p.stagger;
        while (clock64() < t_go) {
        }
      }
      tc_fence_after();
      if (tr) p.trace[it * 16 + 1] = clock64();
      uint32_t s[BKV];
      tmem_ld_n<32>(tS, s);
      tmem_ld_n<32>(tS + 32, s + 32);
      tmem_ld_n<16>(tS + 64, s + 64);
      tmem_ld_wait();
      // S(c, it) is in registers: hand the columns back so that S(c, it + 1) runs under this tile's exponentials
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&s_free[c]);
      if (tr) p.trace[it * 16 + 7] = clock64();
      if (valid < BKV) {
#pragma unroll
        for (int e = 0; e < BKV; ++e)
          if (e >= valid) s[e] = 0xff800000u;
      }
      float mx4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) mx4[q] = fmaxf(__uint_as_float(s[q]), __uint_as_float(s[q + 4]));
#pragma unroll
      for (int e = 8; e < BKV; e += 8) {
#pragma unroll
        for (int q = 0; q < 4; ++q) mx4[q] = fmax3(mx4[q], __uint_as_float(s[e + q]), __uint_as_float(s[e + q + 4]));
      }
      const float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);
      if (tr) p.trace[it * 16 + 2] = clock64();
      if (j == 0) {
        m = mx;
      } else if (__any_sync(0xffffffffu, mx > m + 8.f)) {
        mbar_wait(&pv_done[c], (it - 1) & 1);   // PV(c, it - 1) complete: O may be modified in place
        tc_fence_after();
        const float m_new = fmaxf(m, mx);
        rescale_cols<0, HDO>(tO, ex2_approx(m - m_new));
        tmem_st_wait();
        tc_fence_before();
        m = m_new;
      }
      if (tr) p.trace[it * 16 + 3] = clock64();
      // p = 2^(s - m) for the 80 keys, packed to 16-bit pairs; 16-key groups beyond the sequence cost no exponentials
      uint32_t pk[BKV / 2];
      const float2 nm = make_float2(-m, -m);
#pragma unroll
      for (int g = 0; g < BKV; g += 16) {
        if (g + 16 <= valid && bf16) {
          // full group: the first kPolyPairs key pairs through the FMA-pipe polynomial, the rest through MUFU
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float2 d = fadd2(make_float2(__uint_as_float(s[g + e]), __uint_as_float(s[g + e + 1])), nm);
            if (e < 2 * kPolyPairs) {
              const float2 r = ex2_poly2(d);
              pk[(g + e) >> 1] = pack_bf16(r.x, r.y);
            } else {
              pk[(g + e) >> 1] = pack_bf16(ex2_approx(d.x), ex2_approx(d.y));
            }
          }
        } else if (g < valid) {
          if (bf16) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const float2 d = fadd2(make_float2(__uint_as_float(s[g + e]), __uint_as_float(s[g + e + 1])), nm);
              pk[(g + e) >> 1] = pack_bf16(ex2_approx(d.x), ex2_approx(d.y));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const float2 d = fadd2(make_float2(__uint_as_float(s[g + e]), __uint_as_float(s[g + e + 1])), nm);
              pk[(g + e) >> 1] = pack_f16(ex2_approx(d.x), ex2_approx(d.y));
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) pk[(g >> 1) + e] = 0u;
        }
      }
      if (tr) p.trace[it * 16 + 4] = clock64();
      // PV(c, it - 1) reads the shared-memory P tile and (first cross tile) leaves the final self-attention O: it must
      // have completed.  Every phase of pv_done is observed in order (a parity wait is only meaningful one phase back);
      // that PV was issued a whole tile ago, so this is normally already satisfied.
      if (it > 0) {
        mbar_wait(&pv_done[c], (it - 1) & 1);
        tc_fence_after();
      }
      if (cross && j == 0 && n_self > 0) {
        // segment switch: the self-attention O is final.  Park it, normalised, in the output row (16-bit, as it is
        // rounded on output anyway) before PV(c, it) overwrites O; the same thread adds the cross part at the end.
        uint32_t rl[8];
        tmem_ld8(tO + (HD / 8) * 8, rl);
        tmem_ld_wait();
        const float f = __fdividef(1.f, __uint_as_float(rl[HD % 8]));
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 8) {
          uint32_t r[8];
          tmem_ld8(tO + c0, r);
          tmem_ld_wait();
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[e]) * f;
          const uint4 u = bf16 ? make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]))
                               : make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
          if (row_ok) *reinterpret_cast<uint4*>(orow + c0) = u;
        }
      }
      // P row -> the K-major operand tile of PV(c, it): 16-byte units, XOR-swizzled like a TMA write would
      // (128B swizzle: unit ^= row & 7 in 128-byte rows; 32B swizzle: unit ^= (row >> 2) & 1 in 32-byte rows)
      {
        uint8_t* prow = sP + c * kPBytes + row * 128;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<uint4*>(prow + ((u ^ (row & 7)) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        uint8_t* prow2 = sP + c * kPBytes + kP1Bytes + row * 32;
#pragma unroll
        for (int u = 0; u < 2; ++u)
          *reinterpret_cast<uint4*>(prow2 + ((u ^ ((row >> 2) & 1)) << 4)) =
              make_uint4(pk[32 + 4 * u], pk[33 + 4 * u], pk[34 + 4 * u], pk[35 + 4 * u]);
      }
      fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's shared-memory reads
      if (tr) p.trace[it * 16 + 5] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[c]);
      if (tr) p.trace[it * 16 + 6] = clock64();
    }
    // epilogue: out = parked self part + tanh(gate) * cross part / its row sum
    uint4 prev[HD / 8];
    const bool two = n_self > 0 && n_cross > 0;
#pragma unroll
    for (int e = 0; e < HD / 8; ++e) prev[e] = (two && row_ok) ? *reinterpret_cast<const uint4*>(orow + 8 * e) : make_uint4(0u, 0u, 0u, 0u);
    mbar_wait(&pv_done[c], (n_tiles - 1) & 1);
    tc_fence_after();
    uint32_t rl[8];
    tmem_ld8(tO + (HD / 8) * 8, rl);
    tmem_ld_wait();
    const float f = __fdividef(n_cross > 0 ? tanhf(p.gate[h]) : 1.f, __uint_as_float(rl[HD % 8]));
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 8) {
      uint32_t r[8];
      tmem_ld8(tO + c0, r);
      tmem_ld_wait();
      const uint32_t pw[4] = {prev[c0 / 8].x, prev[c0 / 8].y, prev[c0 / 8].z, prev[c0 / 8].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 sv;
        if (bf16) sv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pw[e]));
        else sv = __half22float2(*reinterpret_cast<const __half2*>(&pw[e]));
        const float v0 = fmaf(__uint_as_float(r[2 * e]), f, sv.x), v1 = fmaf(__uint_as_float(r[2 * e + 1]), f, sv.y);
        o[e] = bf16 ? pack_bf16(v0, v1) : pack_f16(v0, v1);
      }
      if (row_ok) *reinterpret_cast<uint4*>(orow + c0) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cta_rec) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    cta_rec[2] = (long long)t1;
    cta_rec[3] = clock64() - cta_rec[3];
  }
  if (warp == kSoftWarps + 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int HDP, int HD, int NCTX>
static int launch_attn3(const AttnParams& p, int NS, cudaStream_t st) {
  constexpr int HDO = (HD + 1 + 15) / 16 * 16;
  constexpr size_t stage = ((size_t)80 * HDP * 2 + HDO * 160 + 1023) / 1024 * 1024;
  constexpr size_t smem = (size_t)NCTX * 128 * HDP * 2 + 2 * stage + (size_t)NCTX * (128 * 160) + (8 + 6 * NCTX) * 8 + 16;
  static DeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn3_kernel<HDP, HD, NCTX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) MA3_FAIL((int)e, "cudaFuncSetAttribute(attn3): %s", cudaGetErrorString(e));
    configured.mark();
  }
  const int nq = (p.T + 127) / 128;
  dim3 grid((unsigned)((nq + NCTX - 1) / NCTX), (unsigned)p.H, (unsigned)NS);
  cudaError_t le = launch_pdl(attn3_kernel<HDP, HD, NCTX>, grid, dim3(NCTX * 128 + 32 + NCTX * 32), smem, st, 1, p);
  if (le != cudaSuccess) MA3_FAIL((int)le, "attention launch: %s", cudaGetErrorString(le));
  MA3_LAUNCH_CHECK("attention");
  return 0;
}

template <int HDP, int HD, int BKV>
static int launch_attn2(const AttnParams& p, int NS, cudaStream_t st) {
  constexpr int HDO = (HD + 1 + 15) / 16 * 16;
  constexpr size_t smem = 128 * HDP * 2 + 2 * (BKV * HDP * 2 + HDO * 128 + (BKV > 64 ? HDO * 32 : 0)) + 128 + 2048;
  static DeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn2_kernel<HDP, HD, BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) MA3_FAIL((int)e, "cudaFuncSetAttribute(attn2): %s", cudaGetErrorString(e));
    configured.mark();
  }
  dim3 grid((unsigned)((p.T + 127) / 128), (unsigned)p.H, (unsigned)NS);
  cudaError_t le = launch_pdl(attn2_kernel<HDP, HD, BKV>, grid, dim3(kAttn2Threads), smem, st, 1, p);
  if (le != cudaSuccess) MA3_FAIL((int)le, "attention launch: %s", cudaGetErrorString(le));
  MA3_LAUNCH_CHECK("attention");
  return 0;
}

template <int HDP, int HD, int BKV>
static int launch_attn(const AttnParams& p, int NS, cudaStream_t st) {
  constexpr size_t smem = 128 * HDP * 2 + 2 * (2 * BKV * HDP * 2) + 128 * BKV * 2 + 256;
  static DeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel<HDP, HD, BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) MA3_FAIL((int)e, "cudaFuncSetAttribute(attn): %s", cudaGetErrorString(e));
    configured.mark();
  }
  dim3 grid((unsigned)((p.T + 127) / 128), (unsigned)p.H, (unsigned)NS);
  cudaError_t le = launch_pdl(attn_kernel<HDP, HD, BKV>, grid, dim3(kAttnThreads), smem, st, 1, p);
  if (le != cudaSuccess) MA3_FAIL((int)le, "attention launch: %s", cudaGetErrorString(le));
  MA3_LAUNCH_CHECK("attention");
  return 0;
}

}  // namespace ma3

using namespace ma3;

static int g_attn_version = 0;   // 0 = default (v2 where eligible), 2 / 3 = forced (diagnostics, tests)
extern "C" int ma3_debug_set_attn_version(int v) {
  g_attn_version = v;
  return 0;
}

// q, k: [NS*H, T, hdp]; vt: [NS*H, hdp, Tp]; ky: [NS*H, L, hdp]; vyt: [NS*H, hdp, Lp]; gate: [H] fp32 (raw, tanh is
// applied here); out: [NS, T, H*hd] (16-bit, same dtype as the operands).
extern "C" int ma3_attention(const void* q, const void* k, const void* vt, const void* ky, const void* vyt,
                             const float* gate, void* out, int dtype, int NS, int H, int T, int Tp, int L, int Lp,
                             int hd, int hdp, void* stream) {
  MA3_REQUIRE(q && k && vt && out, "attention: null pointer");
  MA3_REQUIRE(dtype == MA3_BF16 || dtype == MA3_F16, "attention: 16-bit operands only");
  MA3_REQUIRE(NS > 0 && H > 0 && T > 0 && L >= 0, "attention: empty problem");
  MA3_REQUIRE(L == 0 || (ky && vyt && gate), "attention: cross operands required when L > 0");
  MA3_REQUIRE(Tp % 8 == 0 && Tp >= T && (L == 0 || (Lp % 8 == 0 && Lp >= L)), "attention: padded lengths must be multiples of 8");
  MA3_REQUIRE(hd % 8 == 0 && (hdp == 64 || hdp == 128) && hdp >= hd, "attention: hd %% 8 == 0 and hd_pad in {64,128}");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.L = L; p.H = H; p.D = H * hd; p.out = out; p.gate = gate; p.dtype = dtype;
  p.trace = g_trace;
  static const int stagger = getenv("MA3_ATTN_STAGGER") ? atoi(getenv("MA3_ATTN_STAGGER")) : 0;
  p.stagger = stagger;
  const uint64_t nbh = (uint64_t)NS * H;
  int rc;
  // v2 (O resident in TMEM, row sums through the ones row of V^T) whenever the padded head has a spare row;
  // MA3_ATTN_V1=1 forces the first-generation kernel (diagnostics)
  static const bool force_v1 = getenv("MA3_ATTN_V1") != nullptr && getenv("MA3_ATTN_V1")[0] == '1';
  const bool v2 = !force_v1 && ((hdp == 64 && (hd == 16 || hd == 24 || hd == 32 || hd == 48)) ||
                                (hdp == 128 && (hd == 72 || hd == 96)));
  const uint32_t vrows = v2 ? (uint32_t)((hd + 1 + 15) / 16 * 16) : (uint32_t)hdp;   // V^T rows staged per tile
  // KV tile of 80 keys when that saves tiles (the per-tile synchronisation chain is the cost): 312 + 154 tokens are
  // 4 + 2 tiles of 80 against 5 + 3 of 64.  A tile of 80 costs ~1.1x a tile of 64.  MA3_ATTN_BKV=64|80 forces one.
  static const int force_bkv = getenv("MA3_ATTN_BKV") ? atoi(getenv("MA3_ATTN_BKV")) : 0;
  const int t64 = (T + 63) / 64 + (L + 63) / 64, t80 = (T + 79) / 80 + (L + 79) / 80;
  int BKV = (v2 && 11 * t80 < 10 * t64 && 2 * 80 + (int)vrows <= 256) ? 80 : 64;
  if (v2 && force_bkv == 64) BKV = 64;
  if (v2 && force_bkv == 80 && 2 * 80 + (int)vrows <= 256) BKV = 80;
  // v3 (several query tiles of a head per CTA, one softmax thread per row, P through shared memory) is opt-in:
  // MA3_ATTN_VER=3 or ma3_debug_set_attn_version(3).  Measured equal to v2 at the XL shape (41-42 us against 40 us;
  // DESIGN section 7b has the timeline): per-CTA fixed costs and the work granularity bound both the same way.
  static const int env_ver = getenv("MA3_ATTN_VER") ? atoi(getenv("MA3_ATTN_VER")) : 0;
  const int force_ver = g_attn_version ? g_attn_version : env_ver;
  const bool v3 = v2 && force_ver == 3;
  if (v3) BKV = 80;
  {
    uint64_t dims[3] = {(uint64_t)hdp, (uint64_t)T, nbh};
    uint64_t str[2] = {(uint64_t)hdp * 2, (uint64_t)T * hdp * 2};
    uint32_t boxq[3] = {64, 128, 1}, boxk[3] = {64, (uint32_t)BKV, 1};
    if ((rc = encode_tmap(&p.tmQ, q, 2, 3, dims, str, boxq, 128))) return rc;
    if ((rc = encode_tmap(&p.tmK, k, 2, 3, dims, str, boxk, 128))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)T, (uint64_t)hdp, nbh};
    uint64_t str[2] = {(uint64_t)Tp * 2, (uint64_t)hdp * Tp * 2};
    uint32_t box[3] = {64, vrows, 1};
    if ((rc = encode_tmap(&p.tmVt, vt, 2, 3, dims, str, box, 128))) return rc;
    uint32_t box2[3] = {16, vrows, 1};
    if (BKV > 64 && (rc = encode_tmap(&p.tmVt2, vt, 2, 3, dims, str, box2, 32))) return rc;
  }
  if (L > 0) {
    uint64_t dims[3] = {(uint64_t)hdp, (uint64_t)L, nbh};
    uint64_t str[2] = {(uint64_t)hdp * 2, (uint64_t)L * hdp * 2};
    uint32_t boxk[3] = {64, (uint32_t)BKV, 1};
    if ((rc = encode_tmap(&p.tmKy, ky, 2, 3, dims, str, boxk, 128))) return rc;
    uint64_t dimv[3] = {(uint64_t)L, (uint64_t)hdp, nbh};
    uint64_t strv[2] = {(uint64_t)Lp * 2, (uint64_t)hdp * Lp * 2};
    uint32_t boxv[3] = {64, vrows, 1};
    if ((rc = encode_tmap(&p.tmVyt, vyt, 2, 3, dimv, strv, boxv, 128))) return rc;
    uint32_t boxv2[3] = {16, vrows, 1};
    if (BKV > 64 && (rc = encode_tmap(&p.tmVyt2, vyt, 2, 3, dimv, strv, boxv2, 32))) return rc;
  } else {
    p.tmKy = p.tmK;
    p.tmVyt = p.tmVt;
    p.tmVyt2 = p.tmVt2;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (v3) {
    if (hdp == 64 && hd == 16) return launch_attn3<64, 16, 3>(p, NS, st);
    if (hdp == 64 && hd == 24) return launch_attn3<64, 24, 3>(p, NS, st);
    if (hdp == 64 && hd == 32) return launch_attn3<64, 32, 3>(p, NS, st);
    if (hdp == 64 && hd == 48) return launch_attn3<64, 48, 3>(p, NS, st);
    if (hdp == 128 && hd == 72) return launch_attn3<128, 72, 3>(p, NS, st);
    if (hdp == 128 && hd == 96) return launch_attn3<128, 96, 2>(p, NS, st);
  }
  if (v2 && BKV == 80) {
    if (hdp == 64 && hd == 16) return launch_attn2<64, 16, 80>(p, NS, st);
    if (hdp == 64 && hd == 24) return launch_attn2<64, 24, 80>(p, NS, st);
    if (hdp == 64 && hd == 32) return launch_attn2<64, 32, 80>(p, NS, st);
    if (hdp == 64 && hd == 48) return launch_attn2<64, 48, 80>(p, NS, st);
    if (hdp == 128 && hd == 72) return launch_attn2<128, 72, 80>(p, NS, st);
  }
  if (v2) {
    if (hdp == 64 && hd == 16) return launch_attn2<64, 16, 64>(p, NS, st);
    if (hdp == 64 && hd == 24) return launch_attn2<64, 24, 64>(p, NS, st);
    if (hdp == 64 && hd == 32) return launch_attn2<64, 32, 64>(p, NS, st);
    if (hdp == 64 && hd == 48) return launch_attn2<64, 48, 64>(p, NS, st);
    if (hdp == 128 && hd == 72) return launch_attn2<128, 72, 64>(p, NS, st);
    if (hdp == 128 && hd == 96) return launch_attn2<128, 96, 64>(p, NS, st);
  }
  if (hdp == 64 && hd == 24) return launch_attn<64, 24, 64>(p, NS, st);
  if (hdp == 64 && hd == 64) return launch_attn<64, 64, 64>(p, NS, st);
  if (hdp == 128 && hd == 72) return launch_attn<128, 72, 64>(p, NS, st);
  if (hdp == 128 && hd == 128) return launch_attn<128, 128, 64>(p, NS, st);
  MA3_FAIL(MA3_EINVAL, "attention: head_dim %d (pad %d) not instantiated", hd, hdp);
}
