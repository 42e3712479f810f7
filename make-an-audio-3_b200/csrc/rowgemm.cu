// Row-owning gated-residual GEMM with the following RMSNorm + adaLN modulate fused in (Next-DiT block,
// flag_large_dit.py:79-91, flag_large_dit_moe.py:63-81):
//
//     h[m, :]  <-  h[m, :] + gate_s * (A[m, :] . W^T)                       (fp32 residual stream, updated in place)
//     u[m, :]  <-  16-bit( h_new[m, :] * rsqrt(mean_D(h_new[m, :]^2) + eps) * wn_s + shift_s )      wn_s = w (1 + scale_s)
//
// for the wo and w2 projections of a block (N = D = hidden size), s = m / rows_per_sample.
//
// Why a kernel of its own: the tap-GEMM tiles N = D into 128 x 192 pieces, which (a) cannot know a row's mean square,
// so the normalisation was a stand-alone pass over h between two GEMMs (1344 launches, ~10 % of the sampling step),
// (b) leaves 234 tiles on 148 SMs (1.58 waves), and (c) had to add into h with L2 reductions.  Here a CLUSTER of
// D / 384 CTAs owns 128 complete rows: every CTA accumulates a 128 x 384 slice in 384 TMEM columns (one wave:
// ceil(M / 128) * D / 384 CTAs, 117 for the XL step), updates its slice of h with plain loads and stores, the
// per-row sums of squares are exchanged through distributed shared memory, and each CTA writes its slice of the
// normalised, modulated operand of the next GEMM.  No atomics anywhere: results are bit-reproducible.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue.
// Shared memory: 3 operand stages of 64 KB (A 128 x 64, B 384 x 64, 128B swizzle) that are re-used, once the last MMA
// has completed, as the fp32 tile [128][388] through which the accumulator is transposed (TMEM hands out thread <->
// row; global memory wants lanes along columns) and in which h_new waits for the cluster-wide row sums.
#include "host_common.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace ma3 {

constexpr int kRgThreads = 320;
constexpr int kRgBM = 128, kRgCN = 384, kRgBK = 64, kRgStages = 3;
constexpr int kRgPitch = 388;                                   // floats; = 4 mod 32: conflict-free 128-bit accesses both ways
constexpr uint32_t kRgABytes = kRgBM * kRgBK * 2;               // 16 KB
constexpr uint32_t kRgBBytes = kRgCN * kRgBK * 2;               // 48 KB
constexpr uint32_t kRgStageBytes = kRgABytes + kRgBBytes;       // 64 KB
constexpr uint32_t kRgTileBytes = kRgBM * kRgPitch * 4;         // 198 656 B >= 3 stages (196 608 B)
constexpr int kRgMaxCluster = 4;
constexpr size_t kRgSmem = 1024 + kRgTileBytes + 128 + kRgMaxCluster * kRgBM * sizeof(float);

struct RowGemmParams {
  CUtensorMap tmA, tmB;
  int M, K, D;
  float* h;                 // [M][D] fp32, in place
  const float* gate;        // [samples][mod_ld]
  const float* wn;          // [samples][mod_ld]   w * (1 + scale)
  const float* shift;       // [samples][mod_ld]
  long long mod_ld;
  int rows_per_sample;
  float inv_rps, inv_D, eps;
  void* u_out;              // [M][D] 16-bit, or nullptr: residual update only
  int bf16;
  uint32_t idesc256, idesc128;
  int dbg;                  // diagnostics (MA3_RG_DBG): 1 no h stores, 2 no h_old loads, 4 no row sums; results are garbage
  long long* trace;         // diagnostics (ma3_debug_set_gemm_trace): CTA 0 records clock64() at its phase boundaries
};

__device__ __forceinline__ void rg_trace(const RowGemmParams& p, int slot) {
  if (p.trace && blockIdx.x == 0) p.trace[slot] = clock64();
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ int rg_div(int a, float inv) { return __float2int_rd(((float)a + 0.5f) * inv); }
__device__ __forceinline__ float4 rg_lds4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void rg_sts4(float* p, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float rg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// per-sample vectors of this lane's 12 columns (3 float4)
struct RgVec {
  float4 v[3];
};
__device__ __forceinline__ void rg_load_vec(const float* base, long long row_off, int col0, int lane, RgVec& out) {
#pragma unroll
  for (int i = 0; i < 3; ++i) out.v[i] = __ldg(reinterpret_cast<const float4*>(base + row_off + col0 + 4 * (lane + 32 * i)));
}

__global__ void __launch_bounds__(kRgThreads, 1) rowgemm_norm_kernel(const __grid_constant__ RowGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* tile = reinterpret_cast<float*>(base);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kRgTileBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kRgStages;
  uint64_t* tfull = bars + 2 * kRgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRgStages + 1);
  float* rs_all = reinterpret_cast<float*>(bars + 16);   // [cluster rank][128]: per-row sums of squares of every CTA's slice

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t csize = (uint32_t)(p.D / kRgCN);
  const bool no_cluster = (p.dbg & 16) != 0;   // diagnostics: launched without a cluster (row sums stay CTA-local: wrong)
  const uint32_t rank = no_cluster ? blockIdx.x % csize : cluster_ctarank();
  const int m0 = (int)(blockIdx.x / csize) * kRgBM;
  const int n0 = (int)rank * kRgCN;
  const int kchunks = (p.K + kRgBK - 1) / kRgBK;
  const bool norm = p.u_out != nullptr;

  if (threadIdx.x == 0) rg_trace(p, 0);
  if (p.trace && threadIdx.x == 0) {   // launch span over all CTAs (ns): first entry [32], last exit [33]; CTA 0: [34], [35]
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMin(reinterpret_cast<unsigned long long*>(p.trace) + 32, t);
    if (blockIdx.x == 0) p.trace[34] = (long long)t;
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kRgStages; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncwarp();
  if (no_cluster) __syncthreads();
  else cluster_sync_all();   // every CTA of the cluster is running (its shared memory may be written remotely from here on)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef RG_CHECK_TMEM
  if (tmem_base != 0) {
    if (lane == 0) printf("rowgemm: block %d warp %d tmem_base %08x\n", blockIdx.x, warp, tmem_base);
    __trap();
  }
#endif
  pdl_wait();
  if (threadIdx.x == 0) rg_trace(p, 1);

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tiles_u32 = smem_u32(base), full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
      int s = 0;
      uint32_t ph = 1;
      int kx = 0;
      for (int i = 0; i < kchunks; ++i) {
        while (!mbar_try_wait(empty_u32 + 8 * s, ph)) {
        }
        const uint32_t dst = tiles_u32 + (uint32_t)s * kRgStageBytes, bar = full_u32 + 8 * s;
        mbar_arrive_expect_tx_u32(bar, kRgStageBytes);
        tma_load_3d_u32(dst, &p.tmA, bar, kx, m0, 0);
        tma_load_3d_u32(dst + kRgABytes, &p.tmB, bar, kx, n0, 0);                       // rows n0 .. n0+191
        tma_load_3d_u32(dst + kRgABytes + 192 * 128, &p.tmB, bar, kx, n0 + 192, 0);     // rows n0+192 .. n0+383
        if (++s == kRgStages) { s = 0; ph ^= 1; }
        kx += kRgBK;
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint64_t desc0 = umma_desc_kmajor(smem_u32(base), 128);
      const uint32_t desc_hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
      const uint32_t full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
      const uint32_t stage16 = kRgStageBytes >> 4, a16 = kRgABytes >> 4, b1_16 = (192 * 128) >> 4;
      const uint32_t id192 = p.idesc256;   // two symmetric N = 192 MMAs per k-step: one per TMA box of B
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < kchunks; ++i) {
        while (!mbar_try_wait(full_u32 + 8 * s, ph)) {
        }
        if (i == 0) rg_trace(p, 2);
        if (i == kchunks / 2) rg_trace(p, 3);
        tc_fence_after();
        const uint32_t alo = lo0 + (uint32_t)s * stage16, blo = alo + a16;
        // first the four K = 16 steps of slice columns 0..191, then the four of columns 192..383 (TMEM 256..447)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!(p.dbg & 128)) umma_f16_lohi<1>(tmem_base, alo + 2 * k, blo + 2 * k, desc_hi, id192, (i | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!(p.dbg & 32))
            umma_f16_lohi<1>(tmem_base + 256, alo + 2 * k, blo + b1_16 + 2 * k, desc_hi, id192, (i | k) != 0 ? 1u : 0u);
        umma_commit_u32<1>(empty_u32 + 8 * s);
        if (++s == kRgStages) { s = 0; ph ^= 1; }
      }
      umma_commit(tfull);
    }
  } else {
    const int ew = warp - 2, r0 = ew * 16;
    // Everything this warp needs from global memory is requested as early as the registers allow: the first four rows
    // of h_old and the gate vector while the mainloop still runs, then row batches double-buffered against the
    // processing of the previous batch -- the phase is bound by how many bytes are in flight, not by instructions.
    auto load_rows = [&](int rb, float4(&ho)[4][3]) {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int m = m0 + r0 + rb + rr;
        if (m < p.M && !(p.dbg & 2)) {
          const float* hp = p.h + (long long)m * p.D + n0;
#pragma unroll
          for (int i = 0; i < 3; ++i) ho[rr][i] = *reinterpret_cast<const float4*>(hp + 4 * (lane + 32 * i));
        }
      }
    };
    float4 hoA[4][3], hoB[4][3];
    RgVec gate;
    const int m_first = m0 + r0 < p.M ? m0 + r0 : p.M - 1;
    int cur_s = rg_div(m_first, p.inv_rps);
    load_rows(0, hoA);
    rg_load_vec(p.gate, (long long)cur_s * p.mod_ld, n0, lane, gate);

    // ---- phase A: accumulator (thread <-> row) -> fp32 tile in shared memory (the operand stages are dead by now)
    {
      const int q = warp & 3, hf = ew >> 2;
      mbar_wait(tfull, 0);
      if (ew == 0 && lane == 0) rg_trace(p, 4);
      tc_fence_after();
      if (p.dbg & 64) goto done;   // diagnostics: mainloop only
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      float* trow = tile + (q * 32 + lane) * kRgPitch;
      for (int c = hf; c < kRgCN / 32; c += 2) {
        uint32_t r[32];
        tmem_ld32(taddr + (c < 6 ? c * 32 : 256 + (c - 6) * 32), r);   // slice columns 192.. live at TMEM column 256..
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          rg_sts4(trow + c * 32 + e, make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]),
                                                  __uint_as_float(r[e + 3])));
      }
      tc_fence_before();
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps: the tile is complete
    if (ew == 0 && lane == 0) rg_trace(p, 5);

    // ---- phase B: rows r0 .. r0+15, lanes along columns (3 float4 per lane): h_new = h + gate * acc
    float my_ss = 0.f;
    auto process = [&](int rb, float4(&ho)[4][3]) {
      float ssr[4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = r0 + rb + rr, m = m0 + r;
        ssr[rr] = 0.f;
        if (m < p.M) {   // warp-uniform
          const int s = rg_div(m, p.inv_rps);
          if (s != cur_s) {
            rg_load_vec(p.gate, (long long)s * p.mod_ld, n0, lane, gate);
            cur_s = s;
          }
          float* tp = tile + r * kRgPitch;
          float* hp = p.h + (long long)m * p.D + n0;
          float part[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int c = 4 * (lane + 32 * i);
            const float4 a = rg_lds4(tp + c);
            float4 hn;
            hn.x = fmaf(gate.v[i].x, a.x, ho[rr][i].x); hn.y = fmaf(gate.v[i].y, a.y, ho[rr][i].y);
            hn.z = fmaf(gate.v[i].z, a.z, ho[rr][i].z); hn.w = fmaf(gate.v[i].w, a.w, ho[rr][i].w);
            if (!(p.dbg & 1)) *reinterpret_cast<float4*>(hp + c) = hn;
            if (norm) rg_sts4(tp + c, hn);
            part[i] = fmaf(hn.x, hn.x, hn.y * hn.y) + fmaf(hn.z, hn.z, hn.w * hn.w);
          }
          ssr[rr] = part[0] + part[1] + part[2];
        }
      }
      // four row sums across the warp in 6 shuffles (transposing butterfly) instead of 4 x 5: after the xor-16 and
      // xor-8 steps every lane carries ONE row (row = 2 * bit4 + bit3 of the lane), three more steps finish it
#ifdef RG_OLD_SUM
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const float t = rg_warp_sum(ssr[rr]);
        if (lane == rb + rr) my_ss = t;
      }
      return;
#endif
      const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
      float x0 = hi16 ? ssr[2] : ssr[0], y0 = hi16 ? ssr[0] : ssr[2];
      float x1 = hi16 ? ssr[3] : ssr[1], y1 = hi16 ? ssr[1] : ssr[3];
      x0 += __shfl_xor_sync(0xffffffffu, y0, 16);
      x1 += __shfl_xor_sync(0xffffffffu, y1, 16);
      float x = hi8 ? x1 : x0;
      const float y = hi8 ? x0 : x1;
      x += __shfl_xor_sync(0xffffffffu, y, 8);
      x += __shfl_xor_sync(0xffffffffu, x, 4);
      x += __shfl_xor_sync(0xffffffffu, x, 2);
      x += __shfl_xor_sync(0xffffffffu, x, 1);
      // row rr of the batch now sits in lanes with (bit4, bit3) = (rr >> 1, rr & 1); lane rb + rr keeps it
      const float t = __shfl_sync(0xffffffffu, x, ((lane & 2) << 3) | ((lane & 1) << 3));
      if ((lane >> 2) == (rb >> 2) && lane < 16) my_ss = t;
    };
    load_rows(4, hoB);
    process(0, hoA);
    load_rows(8, hoA);
    process(4, hoB);
    load_rows(12, hoB);
    process(8, hoA);
    process(12, hoB);
    if (ew == 0 && lane == 0) rg_trace(p, 6);

    if (norm) {
      // the modulation vectors of phase C are requested before the cluster-wide wait
      RgVec wn, sh;
      int cs = rg_div(m_first, p.inv_rps);
      rg_load_vec(p.wn, (long long)cs * p.mod_ld, n0, lane, wn);
      rg_load_vec(p.shift, (long long)cs * p.mod_ld, n0, lane, sh);
      if (lane < 16) {
        // this CTA's partial row sums go to slot [rank] of EVERY CTA of the cluster (distributed shared memory)
        const uint32_t slot = smem_u32(rs_all + rank * kRgBM + r0 + lane);
        if (!(p.dbg & 24))
          for (uint32_t k = 0; k < csize; ++k) st_cluster_f32(mapa_u32(slot, k), my_ss);
      }
      __syncwarp();
      if (!(p.dbg & 24)) cluster_sync_all();   // release / acquire: every CTA's partial sums are visible in every CTA's rs_all
      if (ew == 0 && lane == 0) rg_trace(p, 7);
      // ---- phase C: u = h_new * rstd * wn_s + shift_s for the same rows (h_new is still in this warp's tile rows)
      float rstd = 0.f;
      if (lane < 16) {
        float tot = 0.f;
        for (uint32_t k = 0; k < csize; ++k) tot += rs_all[k * kRgBM + r0 + lane];   // rank order: same sum in every CTA
        rstd = rsqrtf(tot * p.inv_D + p.eps);
      }
      uint16_t* ub = reinterpret_cast<uint16_t*>(p.u_out);
      for (int rb = 0; rb < 16; rb += 4) {
        float4 hn[4][3];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const float* tp = tile + (r0 + rb + rr) * kRgPitch;
#pragma unroll
          for (int i = 0; i < 3; ++i) hn[rr][i] = rg_lds4(tp + 4 * (lane + 32 * i));
        }
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int m = m0 + r0 + rb + rr;
          const float rs = __shfl_sync(0xffffffffu, rstd, rb + rr);
          if (m < p.M) {   // warp-uniform
            const int s = rg_div(m, p.inv_rps);
            if (s != cs) {
              rg_load_vec(p.wn, (long long)s * p.mod_ld, n0, lane, wn);
              rg_load_vec(p.shift, (long long)s * p.mod_ld, n0, lane, sh);
              cs = s;
            }
            uint16_t* up = ub + (long long)m * p.D + n0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const float4 v = hn[rr][i];
              const float v0 = fmaf(v.x * rs, wn.v[i].x, sh.v[i].x), v1 = fmaf(v.y * rs, wn.v[i].y, sh.v[i].y);
              const float v2 = fmaf(v.z * rs, wn.v[i].z, sh.v[i].z), v3 = fmaf(v.w * rs, wn.v[i].w, sh.v[i].w);
              uint2 u;
              if (p.bf16) { u.x = pack_bf16(v0, v1); u.y = pack_bf16(v2, v3); }
              else { u.x = pack_f16(v0, v1); u.y = pack_f16(v2, v3); }
              *reinterpret_cast<uint2*>(up + 4 * (lane + 32 * i)) = u;
            }
          }
        }
      }
    }
  }
done:
  if (warp < 2 && norm && !(p.dbg & 24)) {   // the producer / MMA warps take part in the cluster-wide barrier
    __syncwarp();                            // (.aligned: the elected lane has to have rejoined its warp)
    cluster_sync_all();
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) rg_trace(p, 8);
  if (p.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(reinterpret_cast<unsigned long long*>(p.trace) + 33, t);
    if (blockIdx.x == 0) p.trace[35] = (long long)t;
  }
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ma3

using namespace ma3;

// a [M][K] (16-bit, row pitch a_ld), w [D][K] (row pitch w_ld); h fp32 [M][D] in place; gate / wn / shift fp32
// [samples][mod_ld]; u_out 16-bit [M][D] or NULL.  D must be 768, 1152 or 1536 (cluster of D / 384 CTAs).
extern "C" int ma3_gemm_rownorm(const void* a, int64_t a_ld, const void* w, int64_t w_ld, int dtype, int M, int K, int D,
                                float* h, const float* gate, const float* wn, const float* shift, int64_t mod_ld,
                                int rows_per_sample, void* u_out, float eps, void* stream) {
  MA3_REQUIRE(a && w && h && gate, "gemm_rownorm: null pointer");
  MA3_REQUIRE(dtype == MA3_BF16 || dtype == MA3_F16, "gemm_rownorm: operand dtype must be bf16 or f16");
  MA3_REQUIRE(D % kRgCN == 0 && D / kRgCN >= 1 && D / kRgCN <= kRgMaxCluster, "gemm_rownorm: D=%d must be 384 * {1..4}", D);
  MA3_REQUIRE(M > 0 && K > 0 && K % 16 == 0 && a_ld % 8 == 0 && w_ld % 8 == 0 && a_ld >= K && w_ld >= K,
              "gemm_rownorm: K %% 16 == 0, leading dimensions multiples of 8 and >= K");
  MA3_REQUIRE(rows_per_sample > 0 && mod_ld % 4 == 0 && aligned16(h) && aligned16(gate) && aligned16(a) && aligned16(w),
              "gemm_rownorm: rows_per_sample > 0, mod_ld %% 4 == 0, 16-byte aligned pointers");
  MA3_REQUIRE(u_out == nullptr || (wn && shift && aligned16(u_out) && aligned16(wn) && aligned16(shift)),
              "gemm_rownorm: wn and shift required with u_out");
  RowGemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.K = K; p.D = D; p.h = h; p.gate = gate; p.wn = wn; p.shift = shift; p.mod_ld = mod_ld;
  p.rows_per_sample = rows_per_sample; p.inv_rps = 1.0f / (float)rows_per_sample; p.inv_D = 1.0f / (float)D; p.eps = eps;
  p.u_out = u_out; p.bf16 = dtype == MA3_BF16 ? 1 : 0;
  p.idesc256 = umma_idesc(kRgBM, 192, p.bf16);
  p.idesc128 = 0;
  p.trace = g_trace;
  {
    static const int dbg = getenv("MA3_RG_DBG") ? atoi(getenv("MA3_RG_DBG")) : 0;
    p.dbg = dbg;
  }
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, 1};
    uint64_t str[2] = {(uint64_t)a_ld * 2, (uint64_t)M * a_ld * 2};
    uint32_t box[3] = {(uint32_t)kRgBK, (uint32_t)kRgBM, 1};
    if ((rc = encode_tmap(&p.tmA, a, 2, 3, dims, str, box, 128))) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)D, 1};
    uint64_t str[2] = {(uint64_t)w_ld * 2, (uint64_t)D * w_ld * 2};
    uint32_t box[3] = {(uint32_t)kRgBK, 192, 1};
    if ((rc = encode_tmap(&p.tmB, w, 2, 3, dims, str, box, 128))) return rc;
  }
  static DeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(rowgemm_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRgSmem);
    if (e != cudaSuccess) MA3_FAIL((int)e, "cudaFuncSetAttribute(rowgemm_norm): %s", cudaGetErrorString(e));
    configured.mark();
  }
  const int csize = D / kRgCN;
  const int grid = ((M + kRgBM - 1) / kRgBM) * csize;
  cudaError_t e = launch_pdl(rowgemm_norm_kernel, dim3((unsigned)grid), dim3(kRgThreads), kRgSmem,
                             reinterpret_cast<cudaStream_t>(stream), (p.dbg & 16) ? 1 : csize, p);
  if (e != cudaSuccess) MA3_FAIL((int)e, "gemm_rownorm launch: %s", cudaGetErrorString(e));
  MA3_LAUNCH_CHECK("gemm_rownorm");
  return 0;
}
