// Tap-GEMM for sm_100a: persistent, warp-specialised, TMA -> smem ring -> tcgen05.mma -> TMEM (double-buffered
// accumulator) -> fused epilogue.  One kernel serves nn.Linear, Conv1d (any k / dilation, channels-last) and the
// phases of ConvTranspose1d; see include/ma3_b200.h for the contract and the reference code it replaces.
//
// Roles (320 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2..9 = epilogue (TMEM lane quarter = warp % 4, two warps per quarter on alternating 32-column chunks).
#include "host_common.h"
#include "ptx.cuh"

#include <math.h>
#include <stdlib.h>

namespace ma3 {

constexpr int kBM = 128;
constexpr int kEpiWarps = 8;                      // two warps per TMEM lane quarter, alternating column chunks
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;

struct GemmKParams {
  CUtensorMap tmA, tmB;
  int M, N, K, taps;
  int a_shift[MA3_MAX_TAPS], b_row[MA3_MAX_TAPS];
  int a_batched, b_batched;
  int BN, BK, stages;
  int tiles_m, tiles_n, batch;
  uint32_t idesc, tmem_cols, tmem_stage_cols;
  // epilogue
  void* out;
  int out_dtype;
  long long out_ld, out_batch_stride;
  int out_row_mul, out_row_off;
  const float* bias;
  int bias_per_row;
  const void* res;
  int res_dtype;
  long long res_ld, res_batch_stride;
  float alpha;
  int act;
  int accumulate;
  int vec_ok;
  const float* gate;
  long long gate_ld, gate_batch_stride;
  int rows_per_sample;
  void *q_out, *k_out, *vt_out;
  const float* rope;
  int model_dim, head_dim, head_dim_pad, heads, tokens, tokens_pad;
  float q_scale;
  int first_section;
  int qkv_fast;      // QKV_ROPE: lean chunk code for full, single-section chunks (MA3_QKV_FAST=0 switches it off)
  int op_dtype;
  float inv_rows_per_sample, inv_tokens, inv_head_dim;  // exact-division helpers (see fast_div)
  int debug_mode;    // diagnostics: 1 = no TMA loads, 2 = no MMAs
  // narrow-conv kernel (conv_narrow_kernel): rows of the staged A tile, smallest tap shift, padded N, A stages
  int cn_rows_a, cn_min_shift, cn_bnp, cn_stages;
  int cn_lin, cn_tap_step;   // taps equally spaced (a_shift[t] = a_shift[0] + t * cn_tap_step): descriptor increments
  // 65..128 channels: two 64-wide k-chunks per staged tile (cn_kchunks), and the output columns split over cn_nsplit
  // CTAs of cn_ncols columns each, so that a CTA keeps only its share of the weights resident
  int cn_kchunks, cn_nsplit, cn_ncols, cn_rows_a8;
  // fused RMSNorm (see ma3_gemm_t): producer outputs of GATE_RES, consumer pre-op of any epilogue
  void* norm_out;
  const float* norm_w;
  float* ss_out;
  const float* row_ss;
  int ss_cols;
  float ss_inv_dim, ss_eps;
  const float* col_bias2;
  long long col_bias2_ld;
  int stream_k;      // GATE_RES only: workers take equal contiguous ranges of (tile, k-iteration) instead of whole tiles
  long long* trace;  // diagnostics: when non-null, CTA 0 records clock64() at pipeline events (tools/probe_trace.py)
};

// Work distribution.  Tile mode: worker w owns tiles w, w + n_workers, ... with the full reduction each.  Stream-K mode
// (epilogues that add into the output with fire-and-forget reductions): the tiles x k-iterations space is cut into
// n_workers equal contiguous ranges, so every SM gets the same number of MMA k-steps whatever the tile count; a tile cut
// by a range boundary is finished by two workers, each adding its partial product.
struct WorkState {
  long long cur, end;
  int step;
};
__device__ __forceinline__ WorkState work_begin(const GemmKParams& p, int worker, int n_workers, int total_tiles, int iters) {
  WorkState ws;
  if (p.stream_k) {
    const long long total = (long long)total_tiles * iters;
    ws.cur = total * worker / n_workers;
    ws.end = total * (worker + 1) / n_workers;
    ws.step = 0;
  } else {
    ws.cur = worker;
    ws.end = total_tiles;
    ws.step = n_workers;
  }
  return ws;
}
// next item: tile index, first k-iteration and number of k-iterations; false when the worker is done
__device__ __forceinline__ bool work_next(const GemmKParams& p, WorkState& ws, int iters, int& tile, int& k0, int& kn) {
  if (ws.cur >= ws.end) return false;
  if (p.stream_k) {
    tile = (int)(ws.cur / iters);
    k0 = (int)(ws.cur - (long long)tile * iters);
    const long long left = ws.end - ws.cur;
    kn = left < (long long)(iters - k0) ? (int)left : iters - k0;
    ws.cur += kn;
  } else {
    tile = (int)ws.cur;
    k0 = 0;
    kn = iters;
    ws.cur += ws.step;
  }
  return true;
}

__device__ __forceinline__ void trace_evt(const GemmKParams& p, int tile_local, int slot) {
  if (p.trace && blockIdx.x == 0 && tile_local < 16) p.trace[tile_local * 16 + slot] = clock64();
}

// floor(a / b) for 0 <= a < 2^22 via a float reciprocal (inv = 1/b): exact because the rounding error of
// (a + 0.5) * inv is far below the 0.5 / b margin for the sizes used here (rows < 4M, b < 64K).
__device__ __forceinline__ int fast_div(int a, float inv) { return __float2int_rd(((float)a + 0.5f) * inv); }

__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Explicit shared-space accesses for the epilogue staging patch: the patch pointer is derived by pointer arithmetic
// from the dynamic shared-memory base and ptxas otherwise falls back to generic LD / ST (slower address path, and the
// loads then sit on the long scoreboard next to the global ones).
__device__ __forceinline__ float4 lds_f4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ float lds_f1(const float* p) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void sts_u4(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------------------------------------ epilogues
__device__ __forceinline__ void load8(const void* base, int dtype, long long idx, bool vec, int n, float (&v)[8]) {
  if (dtype == MA3_F32) {
    const float* p = reinterpret_cast<const float*>(base) + idx;
    if (vec && n == 8) {
      float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = e < n ? p[e] : 0.f;
    }
  } else if (dtype == MA3_BF16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + idx;
    if (vec && n == 8) {
      uint4 u = *reinterpret_cast<const uint4*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) { float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = e < n ? __bfloat162float(p[e]) : 0.f;
    }
  } else {
    const __half* p = reinterpret_cast<const __half*>(base) + idx;
    if (vec && n == 8) {
      uint4 u = *reinterpret_cast<const uint4*>(p);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) { float2 f = __half22float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = e < n ? __half2float(p[e]) : 0.f;
    }
  }
}

__device__ __forceinline__ void store8(void* base, int dtype, long long idx, bool vec, int n, const float (&v)[8]) {
  if (dtype == MA3_F32) {
    float* p = reinterpret_cast<float*>(base) + idx;
    if (vec && n == 8) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) if (e < n) p[e] = v[e];
    }
  } else if (dtype == MA3_BF16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + idx;
    if (vec && n == 8) {
      uint4 u = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      *reinterpret_cast<uint4*>(p) = u;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) if (e < n) p[e] = __float2bfloat16_rn(v[e]);
    }
  } else {
    __half* p = reinterpret_cast<__half*>(base) + idx;
    if (vec && n == 8) {
      uint4 u = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
      *reinterpret_cast<uint4*>(p) = u;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) if (e < n) p[e] = __float2half_rn(v[e]);
    }
  }
}

// Epilogue of one 32-row x w-column (w = 16 or 32) accumulator chunk owned by one warp.  tcgen05.ld delivers the chunk
// with thread <-> row; writing global memory in that mapping costs 32 wavefronts per instruction (each lane touches a
// different row).  The chunk is therefore transposed through a per-warp shared-memory patch (pitch 33 floats, conflict
// free both ways) and processed row-wise: a group of lanes covers contiguous columns of one row, so loads of the
// residual / gate / RoPE table and the stores are coalesced.
constexpr int kStagePitch = 36;  // floats; 16-byte aligned rows, conflict-free for 128-bit shared accesses both ways

// Per-tile, per-lane row bookkeeping hoisted out of the chunk loop (the epilogue warps are instruction-latency bound:
// one warp per scheduler, dependent chains).  Meaning per epilogue:
//   STORE    : 4 rows (m0 + pass*8 + lane/4): off = out offset of the row start, aux = res offset, brow = row bias
//   GATE_RES : 8 rows (m0 + pass*4 + lane/8): off = out offset of the row start, aux = gate row offset
//   SWIGLU   : 2 rows (m0 + pass*16 + lane/2): off = out offset of the row start
//   QKV_ROPE : 4 rows as STORE: off = sample*H*tokens*hdp + t*hdp (q/k row part), aux = t*head_dim (RoPE row offset)
// off < 0 marks a row outside [0, M).
struct RowCtx {
  long long off[8];
  long long aux[8];
  float brow[4];
};

template <int EPI>
__device__ __forceinline__ void make_row_ctx(const GemmKParams& p, int z, int m0, int lane, RowCtx& rc) {
  if constexpr (EPI == 4) {   // MA3_EPI_GATE_RES_NORM keeps no per-row table (see norm_pre_issue)
  } else if constexpr (EPI == MA3_EPI_GATE_RES) {
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int m = m0 + pass * 4 + (lane >> 3);
      rc.off[pass] = m < p.M ? (long long)z * p.out_batch_stride + (long long)m * p.out_ld : -1;
      rc.aux[pass] = m < p.M ? (long long)z * p.gate_batch_stride + (long long)fast_div(m, p.inv_rows_per_sample) * p.gate_ld : 0;
    }
  } else if constexpr (EPI == MA3_EPI_SWIGLU) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int m = m0 + pass * 16 + (lane >> 1);
      rc.off[pass] = m < p.M ? (long long)z * p.out_batch_stride + (long long)m * p.out_ld : -1;
    }
  } else if constexpr (EPI == MA3_EPI_STORE) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int m = m0 + pass * 8 + (lane >> 2);
      const long long orow = (long long)m * p.out_row_mul + p.out_row_off;
      rc.off[pass] = m < p.M ? (long long)z * p.out_batch_stride + orow * p.out_ld : -1;
      rc.aux[pass] = (long long)z * p.res_batch_stride + orow * p.res_ld;
      rc.brow[pass] = (p.bias && p.bias_per_row && m < p.M) ? p.bias[m] : 0.f;
    }
  } else {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int m = m0 + pass * 8 + (lane >> 2);
      const int sample = fast_div(m, p.inv_tokens), t = m - sample * p.tokens;
      rc.off[pass] = m < p.M ? ((long long)sample * p.heads * p.tokens + t) * p.head_dim_pad : -1;
      rc.aux[pass] = (long long)t * p.head_dim;
    }
    // V^T token groups (8 consecutive tokens from m0 + 8 g; never straddle a sample when tokens % 8 == 0):
    // offset of (sample, head 0, d 0, t)
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      const int m = m0 + gq * 8;
      const int sample = fast_div(m, p.inv_tokens), t = m - sample * p.tokens;
      rc.off[4 + gq] = m < p.M ? (long long)sample * p.heads * p.head_dim_pad * p.tokens_pad + t : -1;
    }
  }
}

// Epilogue of one 32-row x w-column (w = 16 or 32) accumulator chunk owned by one warp.  tcgen05.ld delivers the chunk
// with thread <-> row; writing global memory in that mapping costs 32 wavefronts per instruction (each lane touches a
// different row).  The chunk is therefore transposed through a per-warp shared-memory patch (pitch 33 floats, conflict
// free both ways) and processed row-wise: a group of lanes covers contiguous columns of one row, so loads of the
// residual / RoPE table and the stores are coalesced.  All global loads of a chunk are issued before its first store.
__device__ __forceinline__ void unpack16(uint4 u, int dtype, float (&v)[8]);

// 16-bit residual pieces of one chunk in the STORE epilogue's row-wise layout (4 passes x 8 columns per lane), requested
// before the accumulator chunk is loaded and staged so that their global-memory latency overlaps that work.
struct ResPre {
  uint4 v[4];
  bool on;
};
__device__ __forceinline__ void res_pre_issue(const GemmKParams& p, int n0, int w, int lane, const RowCtx& rc, ResPre& rp) {
  const int cg = (lane & 3) * 8, col = n0 + cg;
  rp.on = p.res != nullptr && p.vec_ok && p.res_dtype != MA3_F32 && cg < w && col + 8 <= p.N;
  if (!rp.on) return;
  const uint16_t* res = reinterpret_cast<const uint16_t*>(p.res);
#pragma unroll
  for (int pass = 0; pass < 4; ++pass)
    if (rc.off[pass] >= 0) rp.v[pass] = *reinterpret_cast<const uint4*>(res + rc.aux[pass] + col);
}
// Fused-RMSNorm producer (GATE_RES with norm_out; instantiated as its own epilogue kind so that the row bookkeeping of
// the reduction path does not occupy registers here).  Row-wise mapping as GATE_RES: 8 lanes x float4 cover the 32
// columns of a row, 4 rows per pass, 8 passes; lane rows are rbase + 4 * pass.  NormPre holds everything a lane needs
// from global memory for one chunk; it is requested BEFORE the accumulator chunk is loaded and staged so that the L2
// latency overlaps that work.  The 32 rows of a warp lie in at most two samples (rows_per_sample >= 32, host check),
// so the per-sample gate and wn vectors are two float4 each instead of one per row.
constexpr int MA3_EPI_GATE_RES_NORM = 4;   // internal: MA3_EPI_GATE_RES with norm_out != NULL

struct NormPre {
  float4 h[8];
  float4 g0, w0, g1, w1;
  int nvalid;   // passes [0, nvalid) are rows inside M
  int split;    // passes >= split belong to the lane's second sample
};
__device__ __forceinline__ void norm_pre_issue(const GemmKParams& p, int m0, int n0, int lane, NormPre& pf) {
  const int col = n0 + (lane & 7) * 4;
  const int rbase = m0 + (lane >> 3);
  const float* hp = reinterpret_cast<const float*>(p.out) + (long long)rbase * p.out_ld + col;
  const int left = p.M - rbase;
  pf.nvalid = left <= 0 ? 0 : (left >= 29 ? 8 : (left + 3) >> 2);
#pragma unroll
  for (int pass = 0; pass < 8; ++pass)
    if (pass < pf.nvalid) pf.h[pass] = *reinterpret_cast<const float4*>(hp + (long long)pass * 4 * p.out_ld);
  const int rb = rbase < p.M ? rbase : p.M - 1;
  const int s0 = fast_div(rb, p.inv_rows_per_sample);
  const int to_next = (s0 + 1) * p.rows_per_sample - rbase;      // rows until the next sample starts (>= 1)
  pf.split = to_next >= 29 ? 8 : (to_next + 3) >> 2;
  const bool two = __any_sync(0xffffffffu, pf.split < pf.nvalid);
  const float* gp = p.gate + (long long)s0 * p.gate_ld + col;
  const float* wp = p.norm_w + (long long)s0 * p.gate_ld + col;
  pf.g0 = *reinterpret_cast<const float4*>(gp);
  pf.w0 = *reinterpret_cast<const float4*>(wp);
  if (two) {
    const long long o1 = pf.split < pf.nvalid ? p.gate_ld : 0;   // lanes without a second sample re-read the first
    pf.g1 = *reinterpret_cast<const float4*>(gp + o1);
    pf.w1 = *reinterpret_cast<const float4*>(wp + o1);
  } else {
    pf.g1 = pf.g0;
    pf.w1 = pf.w0;
  }
}

// This tile owns its elements of h (no split-K): the update is a plain load-add-store, and the same pass emits the next
// GEMM's 16-bit operand h_new * wn_s and the chunk's per-row sum of squares (the 8 lanes of a row reduce by shuffles;
// one store per row and chunk, so the partial sums are deterministic).
__device__ __forceinline__ void gate_res_norm_chunk(const GemmKParams& p, int m0, int n0, const uint32_t* r, float* stg,
                                                    int lane, const NormPre& pf) {
#pragma unroll
  for (int e = 0; e < 32; e += 4) sts_u4(stg + lane * kStagePitch + e, r[e], r[e + 1], r[e + 2], r[e + 3]);
  __syncwarp();
  const int cg = (lane & 7) * 4, col = n0 + cg;
  const int rbase = m0 + (lane >> 3);
  float* hp = reinterpret_cast<float*>(p.out) + (long long)rbase * p.out_ld + col;
  uint16_t* gp = reinterpret_cast<uint16_t*>(p.norm_out) + (long long)rbase * p.out_ld + col;
  float* ssp = p.ss_out + (long long)rbase * p.ss_cols + (n0 >> 5);
  const long long step = 4 * p.out_ld, ss_step = 4LL * p.ss_cols;
  const float* sp = stg + (lane >> 3) * kStagePitch + cg;
  const bool bf16 = p.op_dtype == MA3_BF16;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    float ss = 0.f;
    const bool ok = pass < pf.nvalid;
    if (ok) {
      const bool first = pass < pf.split;
      const float4 gv = first ? pf.g0 : pf.g1;
      const float4 wv = first ? pf.w0 : pf.w1;
      const float4 a = lds_f4(sp);
      float4 hn;
      hn.x = fmaf(gv.x, a.x, pf.h[pass].x); hn.y = fmaf(gv.y, a.y, pf.h[pass].y);
      hn.z = fmaf(gv.z, a.z, pf.h[pass].z); hn.w = fmaf(gv.w, a.w, pf.h[pass].w);
      *reinterpret_cast<float4*>(hp) = hn;
      uint2 u;
      if (bf16) {
        u.x = pack_bf16(hn.x * wv.x, hn.y * wv.y); u.y = pack_bf16(hn.z * wv.z, hn.w * wv.w);
      } else {
        u.x = pack_f16(hn.x * wv.x, hn.y * wv.y); u.y = pack_f16(hn.z * wv.z, hn.w * wv.w);
      }
      *reinterpret_cast<uint2*>(gp) = u;
      ss = fmaf(hn.x, hn.x, fmaf(hn.y, hn.y, fmaf(hn.z, hn.z, hn.w * hn.w)));
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    if ((lane & 7) == 0 && ok) *ssp = ss;
    hp += step; gp += step; ssp += ss_step;
    sp += 4 * kStagePitch;
  }
  __syncwarp();
}

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, int m0, int n0, int w, const uint32_t* r,
                                               float* stg, int lane, const RowCtx& rc, const ResPre* rp = nullptr) {
  // V^T: with tokens % 8 == 0 the chunk is transposed through the staging patch below (16-byte stores along the token
  // axis); otherwise it is scattered straight from registers (lanes = consecutive tokens, 2-byte stores per column).
  const bool vt_fast = EPI == MA3_EPI_QKV_ROPE && (p.tokens & 7) == 0;
  if (EPI == MA3_EPI_QKV_ROPE && !vt_fast) {
    const int m = m0 + lane;
    const int sec0 = (n0 >= p.model_dim) + (n0 >= 2 * p.model_dim);
    const int sec1 = (n0 + w - 1 >= p.model_dim) + (n0 + w - 1 >= 2 * p.model_dim);
    if (m < p.M && (sec0 + p.first_section == 2 || sec1 + p.first_section == 2)) {
      const int sample = fast_div(m, p.inv_tokens), t = m - sample * p.tokens;
      for (int g = 0; g < w; g += 8) {
        const int col = n0 + g;
        if (col >= p.N) break;
        const int sec = (col >= p.model_dim) + (col >= 2 * p.model_dim);
        if (sec + p.first_section != 2) continue;
        const int within = col - sec * p.model_dim;
        const int head = fast_div(within, p.inv_head_dim);
        const int d = within - head * p.head_dim;
        const long long base = (((long long)sample * p.heads + head) * p.head_dim_pad + d) * p.tokens_pad + t;
        if (p.op_dtype == MA3_BF16) {
          __nv_bfloat16* vt = reinterpret_cast<__nv_bfloat16*>(p.vt_out);
#pragma unroll
          for (int e = 0; e < 8; ++e) vt[base + (long long)e * p.tokens_pad] = __float2bfloat16_rn(__uint_as_float(r[g + e]));
        } else {
          __half* vt = reinterpret_cast<__half*>(p.vt_out);
#pragma unroll
          for (int e = 0; e < 8; ++e) vt[base + (long long)e * p.tokens_pad] = __float2half_rn(__uint_as_float(r[g + e]));
        }
      }
    }
  }
  // registers (thread = row) -> staging patch, 128-bit stores
#pragma unroll
  for (int e = 0; e < 32; e += 4)
    if (e < w)
      sts_u4(stg + lane * kStagePitch + e, r[e], r[e + 1], r[e + 2], r[e + 3]);
  __syncwarp();

  if constexpr (EPI == MA3_EPI_GATE_RES) {
    // 8 lanes x float4 cover 32 columns of one row; 4 rows per pass.
    // h += gate * acc as vectorised fire-and-forget reductions (REDG.ADD.F32x4): the fp32 residual stream is never
    // loaded by the epilogue, so no HBM round trip sits on its critical path.  One reduction per element per GEMM
    // (no split-K) keeps the result deterministic.
    float* out = reinterpret_cast<float*>(p.out);
    const int cg = (lane & 7) * 4;
    const int col = n0 + cg;
    if (cg < w && col < p.N) {  // N % 4 == 0 enforced on the host
      float4 gv[8];
#pragma unroll
      for (int pass = 0; pass < 8; ++pass)
        if (rc.off[pass] >= 0) gv[pass] = *reinterpret_cast<const float4*>(p.gate + rc.aux[pass] + col);
      const float* sp = stg + (lane >> 3) * kStagePitch + cg;
#pragma unroll
      for (int pass = 0; pass < 8; ++pass) {
        if (rc.off[pass] >= 0) {
          const float4 a = lds_f4(sp);
          red_add_f32x4(out + rc.off[pass] + col, gv[pass].x * a.x, gv[pass].y * a.y, gv[pass].z * a.z, gv[pass].w * a.w);
        }
        sp += 4 * kStagePitch;
      }
    }
  } else if constexpr (EPI == MA3_EPI_SWIGLU) {
    // B rows interleave w1 (even) and w3 (odd): out[m, n/2] = silu(acc[n]) * acc[n+1]; 2 lanes x 16 columns per row
    const int cg = (lane & 1) * 16;
    const int col = n0 + cg;
    if (cg < w && col < p.N) {  // N % 16 == 0 enforced on the host
      const float* sp = stg + (lane >> 1) * kStagePitch + cg;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        if (rc.off[pass] >= 0) {
          float v[8];
          if (p.act == 4) {   // gated tanh-GELU (T5 v1.1 DenseGatedActDense: gelu_new(wi_0 x) * wi_1 x), warp-uniform
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 a = lds_f4(sp + 4 * e);
              v[2 * e] = gelu_tanh_f(a.x) * a.y;
              v[2 * e + 1] = gelu_tanh_f(a.z) * a.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 a = lds_f4(sp + 4 * e);
              v[2 * e] = silu_f(a.x) * a.y;
              v[2 * e + 1] = silu_f(a.z) * a.w;
            }
          }
          store8(p.out, p.out_dtype, rc.off[pass] + (col >> 1), true, 8, v);
        }
        sp += 16 * kStagePitch;
      }
    }
  } else if constexpr (EPI == MA3_EPI_STORE) {
    // 4 lanes x 8 columns cover 32 columns of one row; 8 rows per pass
    const int cg = (lane & 3) * 8;
    const int col = n0 + cg;
    if (cg < w && col < p.N) {
      const int n = min(8, p.N - col);
      const bool vec = p.vec_ok != 0;
      float rv[4][8], ov[4][8], bcol[8];
      if (p.bias && !p.bias_per_row) {
#pragma unroll
        for (int e = 0; e < 8; ++e) bcol[e] = e < n ? p.bias[col + e] : 0.f;
      }
      if (p.res) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          if (rc.off[pass] < 0) continue;
          if (rp != nullptr && rp->on) unpack16(rp->v[pass], p.res_dtype, rv[pass]);
          else load8(p.res, p.res_dtype, rc.aux[pass] + col, vec, n, rv[pass]);
        }
      }
      if (p.accumulate) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass)
          if (rc.off[pass] >= 0) load8(p.out, p.out_dtype, rc.off[pass] + col, vec, n, ov[pass]);
      }
      const float* sp = stg + (lane >> 2) * kStagePitch + cg;
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        if (rc.off[pass] >= 0) {
          float v[8];
          {
            const float4 a = lds_f4(sp), b = lds_f4(sp + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          }
          if (p.bias) {
            if (p.bias_per_row) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] += rc.brow[pass];
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] += bcol[e];
            }
          }
          if (p.res) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += rv[pass][e];
          }
          if (p.alpha != 1.f) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] *= p.alpha;
          }
          if (p.act == 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = silu_f(v[e]);
          } else if (p.act == 2) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.5f * v[e] * (1.f + erff(v[e] * 0.70710678118654752f));
          } else if (p.act == 3) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = tanhf(v[e]);
          }
          if (p.accumulate) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += ov[pass][e];
          }
          store8(p.out, p.out_dtype, rc.off[pass] + col, vec, n, v);
        }
        sp += 8 * kStagePitch;
      }
    }
  } else {  // MA3_EPI_QKV_ROPE
    if (vt_fast) {
      // lane <-> column (one head dimension), 4 groups of 8 consecutive tokens: conflict-free column reads of the
      // patch, one 16-byte store per group into vt[(sample, head, d), t .. t+8)
      const int vcol = n0 + lane;
      const int vsec = (vcol >= p.model_dim) + (vcol >= 2 * p.model_dim);
      if (lane < w && vcol < p.N && vsec + p.first_section == 2) {
        const int within = vcol - vsec * p.model_dim;
        const int head = fast_div(within, p.inv_head_dim);
        const int d = within - head * p.head_dim;
        const long long cpart = ((long long)head * p.head_dim_pad + d) * p.tokens_pad;
        uint16_t* vt = reinterpret_cast<uint16_t*>(p.vt_out);
        const float* sp = stg + lane;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          if (rc.off[4 + gq] >= 0) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = lds_f1(sp + (gq * 8 + e) * kStagePitch);
            uint4 u;
            if (p.op_dtype == MA3_BF16)
              u = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            else
              u = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
            *reinterpret_cast<uint4*>(vt + rc.off[4 + gq] + cpart) = u;
          }
        }
      }
    }
    // q / k columns: 4 lanes x 8 columns per row
    const int cg = (lane & 3) * 8;
    const int col = n0 + cg;
    const int sec = (col >= p.model_dim) + (col >= 2 * p.model_dim);
    const int which = sec + p.first_section;  // 0 q, 1 k, 2 v
    if (cg < w && col < p.N && which != 2) {
      const int within = col - sec * p.model_dim;
      const int head = fast_div(within, p.inv_head_dim);
      const int d = within - head * p.head_dim;  // multiple of 8 (head_dim % 8 == 0 enforced on the host)
      const long long colpart = (long long)head * p.tokens * p.head_dim_pad + d;
      const float sc = which == 0 ? p.q_scale : 1.0f;
      void* dst = which == 0 ? p.q_out : p.k_out;
      float4 cs[4][2];
      if (p.rope) {
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          if (rc.off[pass] >= 0) {
            const float4* c4 = reinterpret_cast<const float4*>(p.rope + rc.aux[pass] + d);
            cs[pass][0] = c4[0];
            cs[pass][1] = c4[1];
          }
        }
      }
      const float* sp = stg + (lane >> 2) * kStagePitch + cg;
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        if (rc.off[pass] >= 0) {
          float v[8];
          if (p.rope) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float4 f = cs[pass][e];  // (cos, sin) of two consecutive pairs
              const float4 xv = lds_f4(sp + 4 * e);
              const float x0 = xv.x, x1 = xv.y, x2 = xv.z, x3 = xv.w;
              v[4 * e] = (x0 * f.x - x1 * f.y) * sc;
              v[4 * e + 1] = (x0 * f.y + x1 * f.x) * sc;
              v[4 * e + 2] = (x2 * f.z - x3 * f.w) * sc;
              v[4 * e + 3] = (x2 * f.w + x3 * f.z) * sc;
            }
          } else {
            const float4 a = lds_f4(sp), b = lds_f4(sp + 4);
            v[0] = a.x * sc; v[1] = a.y * sc; v[2] = a.z * sc; v[3] = a.w * sc;
            v[4] = b.x * sc; v[5] = b.y * sc; v[6] = b.z * sc; v[7] = b.w * sc;
          }
          store8(dst, p.op_dtype, rc.off[pass] + colpart, true, 8, v);
        }
        sp += 8 * kStagePitch;
      }
    }
  }
  __syncwarp();
}

// QKV_ROPE fast path for a full 32 x 32 chunk that lies in one section (q, k or v) with all 32 rows inside M and
// tokens % 8 == 0: the same arithmetic, memory accesses and layouts as epilogue_chunk<QKV_ROPE>, with every row / dtype
// / section test hoisted out of the pass loops.  The epilogue warps (two per scheduler, dependent chains) are
// instruction-latency bound -- ncu: ~350 warp-instructions per chunk of which 140 are loads, math and stores -- so the
// instruction count is the epilogue time, and this epilogue is longer than its tile's mainloop.
template <bool kBf16>
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  if constexpr (kBf16) return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  else return make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
}

template <bool kBf16, bool kScale, bool kRope>
__device__ __forceinline__ void qk_rows_fast(uint16_t* dst, const RowCtx& rc, const float* sp, const float* rope_d, float sc) {
  float4 cs[4][2];
  if constexpr (kRope) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const float4* c4 = reinterpret_cast<const float4*>(rope_d + rc.aux[pass]);
      cs[pass][0] = __ldg(c4);
      cs[pass][1] = __ldg(c4 + 1);
    }
  }
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const float4 xa = lds_f4(sp), xb = lds_f4(sp + 4);
    float v[8];
    if constexpr (kRope) {
      const float4 ca = cs[pass][0], cb = cs[pass][1];
      v[0] = xa.x * ca.x - xa.y * ca.y; v[1] = xa.x * ca.y + xa.y * ca.x;
      v[2] = xa.z * ca.z - xa.w * ca.w; v[3] = xa.z * ca.w + xa.w * ca.z;
      v[4] = xb.x * cb.x - xb.y * cb.y; v[5] = xb.x * cb.y + xb.y * cb.x;
      v[6] = xb.z * cb.z - xb.w * cb.w; v[7] = xb.z * cb.w + xb.w * cb.z;
    } else {
      v[0] = xa.x; v[1] = xa.y; v[2] = xa.z; v[3] = xa.w; v[4] = xb.x; v[5] = xb.y; v[6] = xb.z; v[7] = xb.w;
    }
    if constexpr (kScale) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= sc;
    }
    *reinterpret_cast<uint4*>(dst + rc.off[pass]) = pack8<kBf16>(v);
    sp += 8 * kStagePitch;
  }
}

template <bool kBf16>
__device__ __forceinline__ void qkv_chunk_fast(const GemmKParams& p, int n0, int which, int sec, const uint32_t* r,
                                               float* stg, int lane, const RowCtx& rc) {
#pragma unroll
  for (int e = 0; e < 32; e += 4) sts_u4(stg + lane * kStagePitch + e, r[e], r[e + 1], r[e + 2], r[e + 3]);
  __syncwarp();
  const int hd = p.head_dim;
  if (which == 2) {
    // lane <-> column (one head dimension); 4 groups of 8 consecutive tokens, one 16-byte store each
    const int within = n0 + lane - sec * p.model_dim;
    const int head = fast_div(within, p.inv_head_dim);
    const int d = within - head * hd;
    uint16_t* vt = reinterpret_cast<uint16_t*>(p.vt_out) + ((long long)head * p.head_dim_pad + d) * p.tokens_pad;
    const float* sp = stg + lane;
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = lds_f1(sp + (gq * 8 + e) * kStagePitch);
      *reinterpret_cast<uint4*>(vt + rc.off[4 + gq]) = pack8<kBf16>(v);
    }
  } else {
    // 4 lanes x 8 columns per row, 8 rows per pass; the 8 columns never straddle a head (head_dim % 8 == 0)
    const int cg = (lane & 3) * 8;
    const int within = n0 + cg - sec * p.model_dim;
    const int head = fast_div(within, p.inv_head_dim);
    const int d = within - head * hd;
    uint16_t* dst = reinterpret_cast<uint16_t*>(which == 0 ? p.q_out : p.k_out) + (long long)head * p.tokens * p.head_dim_pad + d;
    const float* sp = stg + (lane >> 2) * kStagePitch + cg;
    const bool scale = which == 0 && p.q_scale != 1.0f;
    if (p.rope) {
      if (scale) qk_rows_fast<kBf16, true, true>(dst, rc, sp, p.rope + d, p.q_scale);
      else qk_rows_fast<kBf16, false, true>(dst, rc, sp, p.rope + d, 1.0f);
    } else {
      if (scale) qk_rows_fast<kBf16, true, false>(dst, rc, sp, nullptr, p.q_scale);
      else qk_rows_fast<kBf16, false, false>(dst, rc, sp, nullptr, 1.0f);
    }
  }
  __syncwarp();
}

// STORE epilogue for narrow tiles (BN <= 64): thread <-> output row, all BN columns of the row straight from the
// TMEM registers to global memory in 8-column (16-byte) pieces.  For the channels-last conv outputs this path serves
// (rows of 2 * BN contiguous bytes) the warp's stores cover a contiguous span, no shared-memory transpose or warp
// synchronisation sits in the chain, and the residual pieces are requested before the accumulator is even ready.
template <int kMaxGroups>
struct RowRes {
  uint4 v[kMaxGroups];
};

__device__ __forceinline__ void unpack16(uint4 u, int dtype, float (&v)[8]) {
  if (dtype == MA3_BF16) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  }
}

// whether the 16-bit residual of this launch can be prefetched as one uint4 per 8-column group
__device__ __forceinline__ bool rowdirect_res16(const GemmKParams& p) {
  return p.res != nullptr && p.vec_ok && p.res_dtype != MA3_F32;
}

__device__ __forceinline__ void rowdirect_prefetch(const GemmKParams& p, int z, int m, int n0, int bn, RowRes<8>& rr) {
  if (!rowdirect_res16(p) || m >= p.M) return;
  const long long orow = (long long)m * p.out_row_mul + p.out_row_off;
  const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.res) + (long long)z * p.res_batch_stride + orow * p.res_ld + n0;
#pragma unroll
  for (int gq = 0; gq < 8; ++gq)
    if (gq * 8 < bn && n0 + gq * 8 + 8 <= p.N) rr.v[gq] = *reinterpret_cast<const uint4*>(rp + gq * 8);
}

// previous output values of a row (16-bit out, accumulate), requested together with the residual
__device__ __forceinline__ bool rowdirect_acc16(const GemmKParams& p) {
  return p.accumulate && p.vec_ok && p.out_dtype != MA3_F32;
}
__device__ __forceinline__ void rowdirect_prefetch_acc(const GemmKParams& p, int z, int m, int n0, int bn, RowRes<8>& ro) {
  if (!rowdirect_acc16(p) || m >= p.M) return;
  const long long orow = (long long)m * p.out_row_mul + p.out_row_off;
  const uint16_t* op = reinterpret_cast<const uint16_t*>(p.out) + (long long)z * p.out_batch_stride + orow * p.out_ld + n0;
#pragma unroll
  for (int gq = 0; gq < 8; ++gq)
    if (gq * 8 < bn && n0 + gq * 8 + 8 <= p.N) ro.v[gq] = *reinterpret_cast<const uint4*>(op + gq * 8);
}

__device__ __forceinline__ void rowdirect_group(const GemmKParams& p, int z, int m, int col, const uint32_t* r,
                                                bool has_pre, uint4 pre, float brow, const float* sbias = nullptr,
                                                bool has_acc = false, uint4 acc = make_uint4(0u, 0u, 0u, 0u)) {
  // 8 columns [col, col + 8) of row m
  const int n = min(8, p.N - col);
  if (n <= 0 || m >= p.M) return;
  const bool vec = p.vec_ok != 0;
  const long long orow = (long long)m * p.out_row_mul + p.out_row_off;
  const long long ooff = (long long)z * p.out_batch_stride + orow * p.out_ld + col;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[e]);
  if (p.bias) {
    if (p.bias_per_row) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += brow;
    } else if (sbias != nullptr) {   // per-column bias staged in shared memory (zero beyond N): two broadcast loads
      const float4 b0 = lds_f4(sbias + col), b1 = lds_f4(sbias + col + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) if (e < n) v[e] += p.bias[col + e];
    }
  }
  if (p.res) {
    float rv[8];
    if (has_pre && n == 8) unpack16(pre, p.res_dtype, rv);
    else load8(p.res, p.res_dtype, (long long)z * p.res_batch_stride + orow * p.res_ld + col, vec, n, rv);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += rv[e];
  }
  if (p.alpha != 1.f) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= p.alpha;
  }
  if (p.act == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = silu_f(v[e]);
  } else if (p.act == 2) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.5f * v[e] * (1.f + erff(v[e] * 0.70710678118654752f));
  } else if (p.act == 3) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tanhf(v[e]);
  }
  if (p.accumulate) {
    float ov[8];
    if (has_acc && n == 8) unpack16(acc, p.out_dtype, ov);
    else load8(p.out, p.out_dtype, ooff, vec, n, ov);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += ov[e];
  }
  store8(p.out, p.out_dtype, ooff, vec, n, v);
}

// ------------------------------------------------------------------------------------------------ kernel
// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile: each
// CTA stages its own 128 rows of A and HALF of the B tile, the leader issues one M=256 MMA that reads both CTAs' shared
// memory, and each CTA drains its own 128 accumulator rows from its own TMEM.  Halving the B bytes every SM has to pull
// through L2 is what lifts the L2-feed bound of the large GEMMs.
template <int EPI, int CG, bool kNarrow = false>
__global__ void __launch_bounds__(kGemmThreads, 1) tap_gemm_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = kBM * p.BK * 2, b_bytes = (p.BN / CG) * p.BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;   // per CTA
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* staging = reinterpret_cast<float*>(bars + 2 * kMaxStages + 8);  // kEpiWarps x 32 x kStagePitch floats
  float* bias_stage = staging + kEpiWarps * 32 * kStagePitch;            // kEpiWarps x 256 floats (fused-norm consumer)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  if (threadIdx.x == 0) trace_evt(p, 15, 0);   // kernel entry
  if (p.trace && threadIdx.x == 0) {            // launch span over all CTAs (ns): first entry, last exit
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMin(reinterpret_cast<unsigned long long*>(p.trace) + 250, t);
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        // CG = 2: the leader's copy collects both CTAs' epilogue warps; row-direct STORE: one warp set per stage
        mbar_init(&tempty[i], (kNarrow ? kEpiWarps / 2 : kEpiWarps) * CG);
      }
      fence_barrier_init();
    }
    __syncwarp();
    if constexpr (CG == 2) tmem_alloc2(tmem_slot, p.tmem_cols);
    else tmem_alloc(tmem_slot, p.tmem_cols);
  }
  pdl_launch_dependents();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) trace_evt(p, 15, 1);   // setup done

  // work items: CG = 1 -> 128-row tiles over all CTAs; CG = 2 -> 256-row tiles over CTA pairs
  const int total_tiles = p.tiles_m * p.tiles_n * p.batch;
  const int worker = blockIdx.x / CG, n_workers = gridDim.x / CG;
  const int kchunks = (p.K + p.BK - 1) / p.BK;   // a ragged last chunk is zero-filled by TMA (box beyond the tensor)
  const int iters = p.taps * kchunks;

  // The producer and the MMA issuer are single threads running dependent instruction chains (~5 clk per instruction):
  // their loops are kept to a handful of instructions per k-iteration -- parameters hoisted out of constant memory,
  // stage index / phase advanced incrementally (no division), descriptors rebuilt from one 32-bit add per stage and the
  // four K = 16 steps of a 64-wide stage unrolled.  (A generic loop measured ~155 clk per MMA, above the 128-clk
  // tensor-pipe floor of a 128 x 256 x 16 MMA, which made every tile shape issue-bound.)
  if (warp == 0) {
    if (elect_one()) {
      const int stages = p.stages, BK = p.BK, BN = p.BN, tiles_n = p.tiles_n, tiles_m = p.tiles_m;
      const bool a_b = p.a_batched != 0, b_b = p.b_batched != 0;
      const int dbg = p.debug_mode;
      const uint32_t tiles_u32 = smem_u32(tiles), full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
      int s = 0;
      uint32_t ph = 1;   // parity to wait for on empty[s]: a fresh barrier passes a wait on parity 1
      WorkState ws = work_begin(p, worker, n_workers, total_tiles, iters);
      int tile, k0, kn;
      while (work_next(p, ws, iters, tile, k0, kn)) {
        const int n_t = tile % tiles_n;
        const int rest = tile / tiles_n;
        const int m_t = rest % tiles_m;
        const int z = rest / tiles_m;
        const int za = a_b ? z : 0, zb = b_b ? z : 0;
        const int arow0 = (m_t * CG + (int)rank) * kBM, brow0 = n_t * BN + (int)rank * (BN / CG);
        int tap = k0 / kchunks, kc = k0 - tap * kchunks;
        int arow = arow0 + p.a_shift[tap], brow = brow0 + p.b_row[tap], kx = kc * BK;
        for (int i = 0; i < kn; ++i) {
          while (!mbar_try_wait(empty_u32 + 8 * s, ph)) {
          }
          const uint32_t dst = tiles_u32 + (uint32_t)s * stage_bytes, bar = full_u32 + 8 * s;
          if (dbg == 1) {
            if (leader) mbar_arrive_u32(bar);
          } else if constexpr (CG == 2) {
            if (leader) mbar_arrive_expect_tx_u32(bar, 2 * stage_bytes);   // bytes of both CTAs land on the leader
            tma_load_3d_2cta_u32(dst, &p.tmA, bar, kx, arow, za);
            tma_load_3d_2cta_u32(dst + a_bytes, &p.tmB, bar, kx, brow, zb);
          } else {
            mbar_arrive_expect_tx_u32(bar, stage_bytes);
            tma_load_3d_u32(dst, &p.tmA, bar, kx, arow, za);
            tma_load_3d_u32(dst + a_bytes, &p.tmB, bar, kx, brow, zb);
          }
          if (++s == stages) { s = 0; ph ^= 1; }
          kx += BK;
          if (++kc == kchunks) {
            kc = 0; kx = 0; ++tap;
            if (i + 1 < kn) { arow = arow0 + p.a_shift[tap]; brow = brow0 + p.b_row[tap]; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      int lt = 0;
      const int stages = p.stages, dbg = p.debug_mode;
      const uint32_t idesc = p.idesc, acc_cols = p.tmem_stage_cols;
      const int sw = p.BK * 2;
      const int ksteps = p.BK / 16;
      // descriptor = lo | hi << 32 with only the 14-bit start-address field of lo changing from stage to stage
      const uint64_t desc0 = umma_desc_kmajor(smem_u32(tiles), sw);
      const uint32_t desc_hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
      const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4;
      const uint32_t full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
      int s = 0;
      uint32_t ph = 0;
      WorkState ws = work_begin(p, worker, n_workers, total_tiles, iters);
      int tile, k0, kn;
      for (; work_next(p, ws, iters, tile, k0, kn); ++lt) {
        const int as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        trace_evt(p, lt, 0);
        mbar_wait(&tempty[as], aph ^ 1);
        trace_evt(p, lt, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * acc_cols;
        for (int i = 0; i < kn; ++i) {
          while (!mbar_try_wait(full_u32 + 8 * s, ph)) {
          }
          const uint32_t alo = lo0 + (uint32_t)s * stage16, blo = alo + a16;
          if (dbg != 2) {
            if (ksteps == 4) {
              umma_f16_lohi<CG>(d_tmem, alo, blo, desc_hi, idesc, i != 0 ? 1u : 0u);
              umma_f16_lohi<CG>(d_tmem, alo + 2, blo + 2, desc_hi, idesc, 1u);
              umma_f16_lohi<CG>(d_tmem, alo + 4, blo + 4, desc_hi, idesc, 1u);
              umma_f16_lohi<CG>(d_tmem, alo + 6, blo + 6, desc_hi, idesc, 1u);
            } else {
              for (int k = 0; k < ksteps; ++k)
                umma_f16_lohi<CG>(d_tmem, alo + 2 * k, blo + 2 * k, desc_hi, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_u32<CG>(empty_u32 + 8 * s);   // frees the stage (in both CTAs of a pair)
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        if constexpr (CG == 2) umma_commit_2cta(&tfull[as]);
        else umma_commit(&tfull[as]);
        trace_evt(p, lt, 2);
      }
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int ew = warp - 2;                // epilogue warp index 0..7
    const int half = ew >> 2;               // which alternating set of 32-column chunks it owns
    int lt = 0;
    WorkState ws = work_begin(p, worker, n_workers, total_tiles, iters);
    int tile, k0, kn;
    for (; work_next(p, ws, iters, tile, k0, kn); ++lt) {
      const int n_t = tile % p.tiles_n;
      const int rest = tile / p.tiles_n;
      const int m_t = rest % p.tiles_m;
      const int z = rest / p.tiles_m;
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      const int m0 = (m_t * CG + (int)rank) * kBM + q * 32;
      if constexpr (kNarrow) {
        // narrow tiles: the two warp sets take alternate tiles (set = accumulator stage), thread <-> row
        if (as != half) continue;
        const int m = m0 + lane, n0 = n_t * p.BN;
        RowRes<8> rr;
        rowdirect_prefetch(p, z, m, n0, p.BN, rr);
        const float brow = (p.bias && p.bias_per_row && m < p.M) ? p.bias[m] : 0.f;
        const bool pre = rowdirect_res16(p);
        mbar_wait(&tfull[as], aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * p.tmem_stage_cols;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          if (c0 < p.BN) {
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
            rowdirect_group(p, z, m, n0 + c0, r, pre, rr.v[c0 >> 3], brow);
            if (c0 + 8 < p.BN) rowdirect_group(p, z, m, n0 + c0 + 8, r + 8, pre, rr.v[(c0 >> 3) + 1], brow);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_leader_relaxed(&tempty[as]);
          else mbar_arrive_relaxed(&tempty[as]);
        }
        continue;
      }
      float* stg = staging + ew * (32 * kStagePitch);
      // fused-RMSNorm consumer: this thread's accumulator row (TMEM lane) is m0 + lane; its rstd comes from the
      // producer's per-chunk sums of squares, its bias row from the row's sample (computed while the mainloop runs)
      // (done while the mainloop of the tile runs).  The bias values of the warp's own column chunks -- for the at most
      // two samples its 32 rows lie in -- are staged in a private 1 KB patch of shared memory, so the chunk loop reads
      // them with broadcast LDS instead of L2-latency loads on its critical path.
      float pre_rstd = 0.f;
      const float* pre_bias = nullptr;   // this lane's sample row of the staged patch
      constexpr bool kPreOp = EPI == MA3_EPI_QKV_ROPE || EPI == MA3_EPI_SWIGLU;   // the two consumers of a norm in the DiT
      if (kPreOp && p.row_ss != nullptr) {
        const int m = m0 + lane;
        const int mc = m < p.M ? m : p.M - 1;
        // every load of this prologue is issued before the first use (fully unrolled, predicated): a rolled loop would
        // pay one L2 round trip per iteration and delay the warp's start on the tile by more than a mainloop
        const float4* ssr = reinterpret_cast<const float4*>(p.row_ss + (long long)mc * p.ss_cols);
        const int n4 = p.ss_cols >> 2;   // <= 16 (host check)
        float4 sv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) sv[j] = j < n4 ? __ldg(ssr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int s_first = fast_div(m0 < p.M ? m0 : p.M - 1, p.inv_rows_per_sample);
        const int s_mine = fast_div(mc, p.inv_rows_per_sample);
        const int s_last = __shfl_sync(0xffffffffu, s_mine, 31);
        float* bs = bias_stage + ew * 256;               // [2 samples][4 chunks][32 columns]
        // lane l stages float4 #(l & 7) of chunk (l >> 3) for both samples
        const int bcol = n_t * p.BN + half * 32 + (lane >> 3) * 64 + (lane & 7) * 4;
        float4 bv0 = make_float4(0.f, 0.f, 0.f, 0.f), bv1 = bv0;
        if (bcol < p.N && (lane >> 3) * 64 + half * 32 < p.BN) {
          bv0 = __ldg(reinterpret_cast<const float4*>(p.col_bias2 + (long long)s_first * p.col_bias2_ld + bcol));
          bv1 = __ldg(reinterpret_cast<const float4*>(p.col_bias2 + (long long)s_last * p.col_bias2_ld + bcol));
        }
        float s4 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) s4 += (sv[j].x + sv[j].y) + (sv[j].z + sv[j].w);
        pre_rstd = rsqrtf(s4 * p.ss_inv_dim + p.ss_eps);
        *reinterpret_cast<float4*>(bs + lane * 4) = bv0;
        *reinterpret_cast<float4*>(bs + 128 + lane * 4) = bv1;
        __syncwarp();
        pre_bias = bs + (s_mine == s_first ? 0 : 128);
      }
      RowCtx rc;
      make_row_ctx<EPI>(p, z, m0, lane, rc);
      if (ew == 0 && lane == 0) trace_evt(p, lt, 4);
      mbar_wait(&tfull[as], aph);
      if (ew == 0 && lane == 0) trace_evt(p, lt, 5);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * p.tmem_stage_cols;
      // QKV_ROPE: all 32 rows of this warp inside M and 16-byte V^T stores possible -> lean chunk code
      const bool qkv_fast = EPI == MA3_EPI_QKV_ROPE && p.qkv_fast && (p.tokens & 7) == 0 && m0 + 32 <= p.M;
      for (int c0 = half * 32; c0 < p.BN; c0 += 32 * (kEpiWarps / 4)) {
        uint32_t r[32];
        const int w = min(32, p.BN - c0);
        const bool tr0 = ew == 0 && lane == 0 && c0 == 0;
        if (tr0) trace_evt(p, lt, 8);
        ResPre rp;
        // (the same trick for the accumulate operand was measured slower: 16 more live registers spill)
        if constexpr (EPI == MA3_EPI_STORE) res_pre_issue(p, n_t * p.BN + c0, w, lane, rc, rp);
        NormPre pf;
        if constexpr (EPI == MA3_EPI_GATE_RES_NORM) {
          if (n_t * p.BN + c0 >= p.N) continue;   // chunk of a ragged last column tile wholly beyond N
          norm_pre_issue(p, m0, n_t * p.BN + c0, lane, pf);
        }
        if (w == 32) {
          tmem_ld32(taddr + c0, r);
        } else {
          uint32_t r16[16];
          tmem_ld16(taddr + c0, r16);
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = r16[e];
        }
        tmem_ld_wait();
        if (tr0) trace_evt(p, lt, 9);
        if (kPreOp && pre_bias != nullptr) {   // acc <- acc * rstd[row] + (shift_s W^T)[n]  (staged: zero beyond N)
          const float* b4 = pre_bias + ((c0 - half * 32) >> 6) * 32;
          const float2 rs2 = make_float2(pre_rstd, pre_rstd);
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b = lds_f4(b4 + e);   // same address in every lane of a sample: broadcast
            const float2 lo = ffma2(make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), rs2, make_float2(b.x, b.y));
            const float2 hi = ffma2(make_float2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), rs2, make_float2(b.z, b.w));
            r[e] = __float_as_uint(lo.x); r[e + 1] = __float_as_uint(lo.y);
            r[e + 2] = __float_as_uint(hi.x); r[e + 3] = __float_as_uint(hi.y);
          }
        }
        if constexpr (EPI == MA3_EPI_STORE) {
          epilogue_chunk<EPI>(p, m0, n_t * p.BN + c0, w, r, stg, lane, rc, &rp);
        } else if constexpr (EPI == MA3_EPI_QKV_ROPE) {
          const int n0 = n_t * p.BN + c0;
          const int sec = (n0 >= p.model_dim) + (n0 >= 2 * p.model_dim);
          const int sec_last = (n0 + 31 >= p.model_dim) + (n0 + 31 >= 2 * p.model_dim);
          if (qkv_fast && w == 32 && sec == sec_last && n0 + 32 <= p.N) {
            if (p.op_dtype == MA3_BF16) qkv_chunk_fast<true>(p, n0, sec + p.first_section, sec, r, stg, lane, rc);
            else qkv_chunk_fast<false>(p, n0, sec + p.first_section, sec, r, stg, lane, rc);
          } else {
            epilogue_chunk<EPI>(p, m0, n0, w, r, stg, lane, rc);
          }
        } else if constexpr (EPI == MA3_EPI_GATE_RES_NORM) {
          gate_res_norm_chunk(p, m0, n_t * p.BN + c0, r, stg, lane, pf);
        } else {
          epilogue_chunk<EPI>(p, m0, n_t * p.BN + c0, w, r, stg, lane, rc);
        }
        if (tr0) trace_evt(p, lt, 10);
      }
      if (ew == 0 && lane == 0) trace_evt(p, lt, 6);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_leader_relaxed(&tempty[as]);
        else mbar_arrive_relaxed(&tempty[as]);
      }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all();
  else __syncthreads();
  if (threadIdx.x == 0) trace_evt(p, 15, 2);   // all roles done
  if (p.trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(reinterpret_cast<unsigned long long*>(p.trace) + 251, t);
  }
  if (warp == 1) {
    __syncwarp();
    if constexpr (CG == 2) tmem_dealloc2(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------ narrow conv kernel
// Conv1d / ConvTranspose1d phases with K <= 64 input channels and N <= 64 output channels (the 48- and 24-channel
// stages of BigVGAN and conv_post).  The tap-GEMM above re-fetches the 128-row input tile once per tap and pays a
// pipeline round trip per (tap, k-chunk); with so few channels that fixed cost is the whole run time.  Here
//   * the weights of all taps stay resident in shared memory for the life of the persistent CTA,
//   * the input tile is staged ONCE per output tile as (128 + halo) rows, and a tap is nothing but a row offset of the
//     UMMA A descriptor inside that tile: the 128B swizzle is a function of the absolute shared-memory address, so a
//     K-major operand may start at any row of a swizzled tile with base-offset 0 (measured exact for every shift,
//     tools/probe note in DESIGN.md section 7),
//   * four TMEM accumulator stages and two alternating sets of row-direct epilogue warps keep two tile epilogues in
//     flight.
constexpr int kCnAcc = 4;   // accumulator stages of 64 columns

// Lean epilogue of the narrow-conv kernel for the layers that make up the vocoder (16-bit channels-last output, whole
// 8-column groups, per-column bias or none, residual and previous output of the output's type, no activation):
// out = (acc + bias + res) * alpha (+ out).  The generic row-direct code serves every epilogue option at run time and
// is ~8000 instructions; at two epilogue warps per scheduler the kernel then waits on instruction fetch and branches
// (profiles/r02_ncu_full_convn_summary.txt: stall reasons no_instruction 2.1, branch_resolving 1.3 per issue).
template <bool kBf16>
__device__ __forceinline__ uint4 cn_lean_group(const uint32_t* r, const float* sb, bool has_res, uint4 rv, float alpha,
                                               bool has_acc, uint4 av) {
  const float4 b0 = lds_f4(sb), b1 = lds_f4(sb + 4);
  float v[8] = {__uint_as_float(r[0]) + b0.x, __uint_as_float(r[1]) + b0.y, __uint_as_float(r[2]) + b0.z,
                __uint_as_float(r[3]) + b0.w, __uint_as_float(r[4]) + b1.x, __uint_as_float(r[5]) + b1.y,
                __uint_as_float(r[6]) + b1.z, __uint_as_float(r[7]) + b1.w};
  if (has_res) {
    float x[8];
    unpack16(rv, kBf16 ? MA3_BF16 : MA3_F16, x);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += x[e];
  }
  if (has_acc) {
    float x[8];
    unpack16(av, kBf16 ? MA3_BF16 : MA3_F16, x);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], alpha, x[e]);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= alpha;
  }
  if (kBf16) return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  return make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
}

template <bool kBf16>
__device__ __forceinline__ void cn_lean_epilogue(const GemmKParams& p, int z, int m, int n0, uint32_t taddr, uint64_t* t_full,
                                                 uint32_t t_parity, uint64_t* t_empty, const float* sbias, int lane) {
  const int ngr = p.cn_ncols >> 3;          // whole 8-column groups of this CTA (host: multiple of 8, <= 64), from column n0
  const bool valid = m < p.M, has_res = p.res != nullptr, has_acc = p.accumulate != 0;
  const long long orow = (long long)m * p.out_row_mul + p.out_row_off;
  const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.res) + (long long)z * p.res_batch_stride + orow * p.res_ld + n0;
  uint16_t* op = reinterpret_cast<uint16_t*>(p.out) + (long long)z * p.out_batch_stride + orow * p.out_ld + n0;
  uint4 rr[8], ro[8];
#pragma unroll
  for (int gq = 0; gq < 8; ++gq) {
    rr[gq] = make_uint4(0u, 0u, 0u, 0u);
    ro[gq] = make_uint4(0u, 0u, 0u, 0u);
    if (valid && gq < ngr) {
      if (has_res) rr[gq] = *reinterpret_cast<const uint4*>(rp + gq * 8);
      if (has_acc) ro[gq] = *reinterpret_cast<const uint4*>(op + gq * 8);
    }
  }
  mbar_wait(t_full, t_parity);
  tc_fence_after();
  uint32_t r[4][16];
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c * 16 < p.cn_bnp) tmem_ld16(taddr + c * 16, r[c]);
  tmem_ld_wait();
  // the accumulator stage is free as soon as it is in registers: the MMA warp gets it back before the stores go out
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_relaxed(t_empty);
  if (!valid) return;
#pragma unroll
  for (int gq = 0; gq < 8; ++gq) {
    if (gq < ngr) {
      const uint4 o = cn_lean_group<kBf16>(&r[gq >> 1][(gq & 1) * 8], sbias + gq * 8, has_res, rr[gq], p.alpha, has_acc, ro[gq]);
      *reinterpret_cast<uint4*>(op + gq * 8) = o;
    }
  }
}

// All MMAs of one output tile when the taps are equally spaced: the operand descriptors advance by constant increments,
// so the single issuing thread spends two integer adds per MMA.  (Looking every tap's shift up in the parameter block
// cost ~75 dependent uniform-datapath instructions per tap: the issuing thread, not the tensor pipe (12 % busy), the
// loads or the epilogue, set the tile rate of the 32- and 48-channel layers at ~2000 clocks per tile.)
template <int KS, int KS1>   // k-steps of the first and (channels 64..127) second k-chunk
__device__ __forceinline__ void cn_issue_tile(uint32_t d_tmem, uint32_t at, uint32_t wt, uint32_t a_step, uint32_t w_step,
                                              int taps, uint32_t dhi, uint32_t idesc, uint32_t a_c1, uint32_t w_c1) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) umma_f16_lohi<1>(d_tmem, at + 2 * ks, wt + 2 * ks, dhi, idesc, ks != 0 ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < KS1; ++ks) umma_f16_lohi<1>(d_tmem, at + a_c1 + 2 * ks, wt + w_c1 + 2 * ks, dhi, idesc, 1u);
#pragma unroll 1
  for (int tap = 1; tap < taps; ++tap) {
    at += a_step;
    wt += w_step;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) umma_f16_lohi<1>(d_tmem, at + 2 * ks, wt + 2 * ks, dhi, idesc, 1u);
#pragma unroll
    for (int ks = 0; ks < KS1; ++ks) umma_f16_lohi<1>(d_tmem, at + a_c1 + 2 * ks, wt + w_c1 + 2 * ks, dhi, idesc, 1u);
  }
}

// kLean: 0 = generic row-direct epilogue, 1 = lean fp16, 2 = lean bf16
template <int kLean>
__global__ void __launch_bounds__(kGemmThreads, 1) conv_narrow_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t w_chunk = (uint32_t)p.cn_bnp * 128u, a_chunk = (uint32_t)p.cn_rows_a8 * 128u;   // one 64-wide k-chunk
  const uint32_t w_tap_bytes = (uint32_t)p.cn_kchunks * w_chunk;
  const uint32_t w_bytes = (uint32_t)p.taps * w_tap_bytes;
  const uint32_t a_bytes = (uint32_t)(p.cn_kchunks * p.cn_rows_a) * 128u;    // bytes the loads of one tile deliver
  const uint32_t a_stage = ((uint32_t)p.cn_kchunks * a_chunk + 1023u) & ~1023u;
  // CTA -> (column slice, worker): the CTAs of a slice share the tiles of that slice
  const int ns = (int)blockIdx.x % p.cn_nsplit, worker = (int)blockIdx.x / p.cn_nsplit;
  const int n_workers = (int)gridDim.x / p.cn_nsplit, n0 = ns * p.cn_ncols;
  uint8_t* sW = base;
  uint8_t* sA = base + ((w_bytes + 1023u) & ~1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)p.cn_stages * a_stage);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;
  uint64_t* a_empty = a_full + kMaxStages;
  uint64_t* t_full = a_empty + kMaxStages;
  uint64_t* t_empty = t_full + kCnAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + kCnAcc);
  float* sbias = reinterpret_cast<float*>(bars + 32);       // 64 floats at byte 256 of the barrier block

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
  }
  const bool col_bias = p.bias != nullptr && !p.bias_per_row;
  if (warp >= 2 && threadIdx.x - 64 < 64) {
    const int c = threadIdx.x - 64;
    sbias[c] = (col_bias && c < p.cn_ncols && n0 + c < p.N) ? p.bias[n0 + c] : 0.f;
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int i = 0; i < p.cn_stages; ++i) {
        mbar_init(&a_full[i], 1);
        mbar_init(&a_empty[i], 1);
      }
      for (int i = 0; i < kCnAcc; ++i) {
        mbar_init(&t_full[i], 1);
        mbar_init(&t_empty[i], kEpiWarps / 2);   // one warp set (4 warps) drains a stage
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kCnAcc * 64);
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int total_tiles = p.tiles_m * p.batch;
  const int ksteps = (p.K + 15) >> 4;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, w_bytes);
      for (int tap = 0; tap < p.taps; ++tap)
        for (int c = 0; c < p.cn_kchunks; ++c)
          tma_load_3d(sW + (size_t)tap * w_tap_bytes + (size_t)c * w_chunk, &p.tmB, w_full, c * 64, p.b_row[tap] + n0, 0);
      int s = 0;
      uint32_t ph = 1;
      for (int tile = worker; tile < total_tiles; tile += n_workers) {
        const int m_t = tile % p.tiles_m, z = tile / p.tiles_m;
        mbar_wait(&a_empty[s], ph);
        mbar_arrive_expect_tx(&a_full[s], a_bytes);
        for (int c = 0; c < p.cn_kchunks; ++c)
          tma_load_3d(sA + (size_t)s * a_stage + (size_t)c * a_chunk, &p.tmA, &a_full[s], c * 64, m_t * kBM + p.cn_min_shift,
                      p.a_batched ? z : 0);
        if (++s == p.cn_stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc(kBM, p.cn_bnp, p.op_dtype == MA3_BF16 ? 1 : 0);
      const uint64_t d0 = umma_desc_kmajor(smem_u32(sA), 128);
      const uint32_t dhi = (uint32_t)(d0 >> 32), a_lo0 = (uint32_t)d0;
      const uint32_t w_lo0 = (uint32_t)umma_desc_kmajor(smem_u32(sW), 128);
      const uint32_t a_stage16 = a_stage >> 4, w_tap16 = w_tap_bytes >> 4;
      const uint32_t a_first = (uint32_t)(p.a_shift[0] - p.cn_min_shift) * 8u;
      const uint32_t a_c1 = a_chunk >> 4, w_c1 = w_chunk >> 4;
      const int ks0 = p.cn_kchunks > 1 ? 4 : ksteps, ks1 = p.cn_kchunks > 1 ? ksteps - 4 : 0;
      mbar_wait(w_full, 0);
      int s = 0, lt = 0;
      uint32_t ph = 0;
      for (int tile = worker; tile < total_tiles; tile += n_workers, ++lt) {
        const int as = lt & (kCnAcc - 1);
        mbar_wait(&t_empty[as], ((lt / kCnAcc) & 1) ^ 1);
        mbar_wait(&a_full[s], ph);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 64;
        const uint32_t a_lo = a_lo0 + (uint32_t)s * a_stage16;
        if (p.cn_lin) {
          const uint32_t at0 = a_lo + a_first, a_step = (uint32_t)(p.cn_tap_step * 8);
#define CN_ISSUE(A, B) cn_issue_tile<A, B>(d_tmem, at0, w_lo0, a_step, w_tap16, p.taps, dhi, idesc, a_c1, w_c1)
          switch (ksteps) {
            case 1: CN_ISSUE(1, 0); break;
            case 2: CN_ISSUE(2, 0); break;
            case 3: CN_ISSUE(3, 0); break;
            case 4: CN_ISSUE(4, 0); break;
            case 5: CN_ISSUE(4, 1); break;
            case 6: CN_ISSUE(4, 2); break;
            case 7: CN_ISSUE(4, 3); break;
            default: CN_ISSUE(4, 4); break;
          }
#undef CN_ISSUE
        } else {
          for (int tap = 0; tap < p.taps; ++tap) {
            const uint32_t at = a_lo + (uint32_t)(p.a_shift[tap] - p.cn_min_shift) * 8u;   // 128-byte rows = 8 x 16 B
            const uint32_t wt = w_lo0 + (uint32_t)tap * w_tap16;
            for (int ks = 0; ks < ks0; ++ks)
              umma_f16_lohi<1>(d_tmem, at + 2 * ks, wt + 2 * ks, dhi, idesc, (tap | ks) != 0 ? 1u : 0u);
            for (int ks = 0; ks < ks1; ++ks) umma_f16_lohi<1>(d_tmem, at + a_c1 + 2 * ks, wt + w_c1 + 2 * ks, dhi, idesc, 1u);
          }
        }
        umma_commit(&a_empty[s]);
        umma_commit(&t_full[as]);
        if (++s == p.cn_stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3, ew = warp - 2, set = ew >> 2;
    int lt = 0;
    for (int tile = worker; tile < total_tiles; tile += n_workers, ++lt) {
      if ((lt & 1) != set) continue;
      const int m_t = tile % p.tiles_m, z = tile / p.tiles_m;
      const int as = lt & (kCnAcc - 1);
      const int m = m_t * kBM + q * 32 + lane;
      if constexpr (kLean != 0) {
        cn_lean_epilogue<kLean == 2>(p, z, m, n0, tmem_base + ((uint32_t)(q * 32) << 16) + as * 64, &t_full[as],
                                     (uint32_t)((lt / kCnAcc) & 1), &t_empty[as], sbias, lane);
        continue;
      }
      RowRes<8> rr;
      rowdirect_prefetch(p, z, m, 0, p.cn_bnp, rr);
      RowRes<8> ro;
      rowdirect_prefetch_acc(p, z, m, 0, p.cn_bnp, ro);
      const float brow = (p.bias && p.bias_per_row && m < p.M) ? p.bias[m] : 0.f;
      const bool pre = rowdirect_res16(p), pacc = rowdirect_acc16(p);
      mbar_wait(&t_full[as], (lt / kCnAcc) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 64;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        if (c0 < p.cn_bnp) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          rowdirect_group(p, z, m, c0, r, pre, rr.v[c0 >> 3], brow, col_bias ? sbias : nullptr, pacc, ro.v[c0 >> 3]);
          rowdirect_group(p, z, m, c0 + 8, r + 8, pre, rr.v[(c0 >> 3) + 1], brow, col_bias ? sbias : nullptr, pacc,
                          ro.v[(c0 >> 3) + 1]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(&t_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kCnAcc * 64);
  }
}

// host side of the narrow-conv path; returns 1 when the problem is not eligible (caller falls back to the tap-GEMM)
static int try_conv_narrow(const ma3_gemm_t* g, GemmKParams& kp, cudaStream_t st) {
  static const bool off = getenv("MA3_CONV_NARROW") != nullptr && getenv("MA3_CONV_NARROW")[0] == '0';
  static const bool wide_off = getenv("MA3_CONV_NARROW_WIDE") != nullptr && getenv("MA3_CONV_NARROW_WIDE")[0] == '0';
  if (off || g->epi != MA3_EPI_STORE || g->K > 128 || g->N > 128 || g->taps < 2 || g->cta_group == 2 || g->tile_n > 0 ||
      g_gemm_debug_mode != 0 || g_trace != nullptr || g->b_batch_stride != 0)
    return 1;
  // lean epilogue: 16-bit output in whole 8-column groups, residual / previous output of the same type, column bias or
  // none, no activation (MA3_CONV_LEAN=0 keeps the generic epilogue)
  static const bool lean_off = getenv("MA3_CONV_LEAN") != nullptr && getenv("MA3_CONV_LEAN")[0] == '0';
  const bool lean = !lean_off && kp.vec_ok && g->out_dtype != MA3_F32 && g->N % 8 == 0 && g->act == 0 &&
                    !(g->bias && g->bias_per_row) && (!g->res || g->res_dtype == g->out_dtype);
  // 65..128 channels (the 96-channel stage of BigVGAN): two k-chunks per staged tile, output columns split over two
  // CTAs so that each keeps only its half of the weights resident; lean epilogue only
  const int kchunks = g->K > 64 ? 2 : 1, nsplit = g->N > 64 ? 2 : 1;
  if ((kchunks > 1 || nsplit > 1) && (wide_off || !lean || g->N % (16 * nsplit) != 0)) return 1;
  const int ncols = g->N / nsplit;
  int lo = g->a_shift[0], hi = g->a_shift[0];
  for (int i = 1; i < g->taps; ++i) { lo = g->a_shift[i] < lo ? g->a_shift[i] : lo; hi = g->a_shift[i] > hi ? g->a_shift[i] : hi; }
  const int rows_a = kBM + (hi - lo), rows_a8 = (rows_a + 7) / 8 * 8;
  const int bnp = nsplit > 1 ? ncols : (g->N + 15) / 16 * 16;
  if (rows_a > 256) return 1;
  const size_t w_bytes = ((size_t)g->taps * kchunks * bnp * 128 + 1023) & ~(size_t)1023;
  const size_t a_stage = ((size_t)kchunks * rows_a8 * 128 + 1023) & ~(size_t)1023;
  const size_t tail = 512;
  // single-chunk layers keep the 200 KB budget they were tuned with; the two-chunk ones may use the whole SM
  const size_t budget = (kchunks > 1 ? 225 : 200) * 1024;
  if (w_bytes + tail + 2 * a_stage > budget) return 1;
  int stages = (int)((budget - w_bytes - tail) / a_stage);
  if (stages > 4) stages = 4;
  kp.cn_rows_a = rows_a; kp.cn_min_shift = lo; kp.cn_bnp = bnp; kp.cn_stages = stages;
  kp.cn_kchunks = kchunks; kp.cn_nsplit = nsplit; kp.cn_ncols = ncols; kp.cn_rows_a8 = rows_a8;
  kp.cn_tap_step = g->a_shift[1] - g->a_shift[0];
  kp.cn_lin = 1;
  for (int i = 2; i < g->taps; ++i)
    if (g->a_shift[i] - g->a_shift[i - 1] != kp.cn_tap_step) kp.cn_lin = 0;
  kp.BN = bnp; kp.BK = 64;
  kp.tiles_m = (g->M + kBM - 1) / kBM; kp.tiles_n = 1;
  const bool a_batched = g->a_batch_stride != 0;
  {
    uint64_t dims[3] = {(uint64_t)g->K, (uint64_t)g->a_rows, (uint64_t)(a_batched ? g->batch : 1)};
    uint64_t str[2] = {(uint64_t)g->a_ld * 2, (uint64_t)(a_batched ? g->a_batch_stride : g->a_rows * g->a_ld) * 2};
    uint32_t box[3] = {64, (uint32_t)rows_a, 1};
    int rc = encode_tmap(&kp.tmA, g->a, 2, 3, dims, str, box, 128);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)g->K, (uint64_t)g->b_rows, 1};
    uint64_t str[2] = {(uint64_t)g->b_ld * 2, (uint64_t)g->b_rows * g->b_ld * 2};
    uint32_t box[3] = {64, (uint32_t)bnp, 1};
    int rc = encode_tmap(&kp.tmB, g->b, 2, 3, dims, str, box, 128);
    if (rc) return rc;
  }
  const size_t smem = 1024 + w_bytes + stages * a_stage + tail;
  static DeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e = cudaFuncSetAttribute(conv_narrow_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_narrow_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_narrow_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) MA3_FAIL((int)e, "cudaFuncSetAttribute(conv_narrow): %s", cudaGetErrorString(e));
    configured.mark();
  }
  const int total_tiles = kp.tiles_m * g->batch;      // per column slice
  int grid = total_tiles * nsplit < num_sms() ? total_tiles * nsplit : num_sms();
  grid -= grid % nsplit;
  const size_t smem_l = smem < 120 * 1024 ? 120 * 1024 : smem;
  // at least half of the SM's shared memory so that one CTA (and its 256 TMEM columns) lives per SM
  cudaError_t e;
  if (!lean) e = launch_pdl(conv_narrow_kernel<0>, dim3((unsigned)grid), dim3(kGemmThreads), smem_l, st, 1, kp);
  else if (g->out_dtype == MA3_F16) e = launch_pdl(conv_narrow_kernel<1>, dim3((unsigned)grid), dim3(kGemmThreads), smem_l, st, 1, kp);
  else e = launch_pdl(conv_narrow_kernel<2>, dim3((unsigned)grid), dim3(kGemmThreads), smem_l, st, 1, kp);
  if (e != cudaSuccess) MA3_FAIL((int)e, "conv_narrow launch: %s", cudaGetErrorString(e));
  MA3_LAUNCH_CHECK("conv_narrow");
  return 0;
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

template <int EPI, int CG, bool kNarrow>
static int launch_cg(const GemmKParams& kp, size_t smem, int grid, cudaStream_t st) {
  static DeviceOnce configured;
  if (configured.pending()) {
    cudaError_t e =
        cudaFuncSetAttribute(tap_gemm_kernel<EPI, CG, kNarrow>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) MA3_FAIL((int)e, "cudaFuncSetAttribute(tap_gemm): %s", cudaGetErrorString(e));
    configured.mark();
  }
  cudaError_t e = launch_pdl(tap_gemm_kernel<EPI, CG, kNarrow>, dim3((unsigned)grid), dim3(kGemmThreads), smem, st, CG, kp);
  if (e != cudaSuccess) MA3_FAIL((int)e, "tap_gemm launch: %s", cudaGetErrorString(e));
  MA3_LAUNCH_CHECK("tap_gemm");
  return 0;
}

template <int EPI>
static int launch(const GemmKParams& kp, size_t smem, int grid, int cta_group, cudaStream_t st) {
  if constexpr (EPI == MA3_EPI_STORE) {
    if (kp.BN <= 64)   // narrow tiles: row-direct epilogue
      return cta_group == 2 ? launch_cg<EPI, 2, true>(kp, smem, grid, st) : launch_cg<EPI, 1, true>(kp, smem, grid, st);
  }
  return cta_group == 2 ? launch_cg<EPI, 2, false>(kp, smem, grid, st) : launch_cg<EPI, 1, false>(kp, smem, grid, st);
}

// Modelled critical path (SM clocks) of one launch for a tile shape.  Per 64-wide k-iteration the tensor pipe needs
// 2 * BN clocks (128 x BN x 64 at 8192 flop/clk), the single issuing thread ~95 clocks per MMA, and the TMA feed
// 120 + 1.5 BN (one CTA) or 200 + 0.6 BN (CTA pair: half of B per SM); a tile's epilogue (clocks per accumulator
// column, per epilogue kind) overlaps the next tile's mainloop, so a tile costs max(mainloop, epilogue) and the last
// epilogue is exposed.  Numbers from tools/probe_trace.py on B200.
static double tile_cost(const ma3_gemm_t* g, int BN, int CG, int BK) {
  const int workers = num_sms() / CG;
  const long tiles_m = (g->M + kBM * CG - 1) / (kBM * CG), tiles_n = (g->N + BN - 1) / BN;
  const long tiles = tiles_m * tiles_n * g->batch;
  const long iters = (long)g->taps * ((g->K + BK - 1) / BK);
  const double ks = BK / 64.0;
  const double mma = fmax(2.0 * BN * ks, 95.0 * (BK / 16));
  const double feed = CG == 1 ? 120.0 + 1.5 * BN * ks : 200.0 + 0.6 * BN * ks;
  const double iter = fmax(mma, feed);
  double per_col;
  switch (g->epi) {
    case MA3_EPI_GATE_RES: per_col = 27.0; break;
    case MA3_EPI_SWIGLU: per_col = 21.0; break;
    case MA3_EPI_QKV_ROPE: per_col = 33.0; break;
    default: per_col = g->out_dtype == MA3_F32 ? 45.0 : 37.0; break;
  }
  const double epi = 400.0 + per_col * BN;
  const double tile = fmax((double)iters * iter, epi);
  const long waves = (tiles + workers - 1) / workers;
  return (double)waves * tile + epi + (CG == 2 ? 900.0 : 0.0);
}

}  // namespace ma3

extern "C" int ma3_gemm(const ma3_gemm_t* g, void* stream) {
  using namespace ma3;
  MA3_REQUIRE(g != nullptr, "gemm: null descriptor");
  MA3_REQUIRE(g->a && g->b, "gemm: null operand");
  MA3_REQUIRE(g->dtype == MA3_BF16 || g->dtype == MA3_F16, "gemm: operand dtype must be bf16 or f16");

  MA3_REQUIRE(g->M > 0 && g->N > 0 && g->batch > 0, "gemm: empty problem M=%d N=%d batch=%d", g->M, g->N, g->batch);
  MA3_REQUIRE(g->K > 0 && g->K % 16 == 0, "gemm: K=%d must be a positive multiple of 16", g->K);
  MA3_REQUIRE(g->taps >= 1 && g->taps <= MA3_MAX_TAPS, "gemm: taps=%d out of range", g->taps);
  MA3_REQUIRE(aligned16(g->a) && aligned16(g->b), "gemm: operands must be 16-byte aligned");
  MA3_REQUIRE(g->a_ld % 8 == 0 && g->b_ld % 8 == 0, "gemm: leading dimensions must be multiples of 8 elements");
  MA3_REQUIRE(g->a_batch_stride % 8 == 0 && g->b_batch_stride % 8 == 0, "gemm: batch strides must be multiples of 8");
  MA3_REQUIRE(g->a_ld >= g->K && g->b_ld >= g->K, "gemm: leading dimension smaller than K");

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  // 64-wide k-chunks whenever K > 32: a ragged last chunk (K = 48, 96, ...) is zero-filled by TMA, which costs idle
  // tensor-pipe work but a third of the k-iterations (and their fixed pipeline cost) of 16- or 32-wide chunks
  const int BK = g->K > 32 ? 64 : (g->K > 16 ? 32 : 16);
  int BN = g->tile_n;
  int CG = g->cta_group;
  {
    // experiment hook: MA3_TILE_<epilogue id>="tile_n,cta_group" overrides the automatic choice for large problems
    // (in-step A/B runs of bench.py; the isolated probes see L2-warm operands and a different clock)
    static int ov[5][2] = {{-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}};   // [4]: GATE_RES with K > 2048
    static bool parsed = false;
    if (!parsed) {
      for (int e = 0; e < 5; ++e) {
        char name[32];
        snprintf(name, sizeof(name), "MA3_TILE_%d", e);
        const char* v = getenv(name);
        if (v) sscanf(v, "%d,%d", &ov[e][0], &ov[e][1]);
      }
      parsed = true;
    }
    const int oi = (g->epi == MA3_EPI_GATE_RES && g->K > 2048) ? 4 : g->epi;
    if (oi >= 0 && oi < 5 && ov[oi][0] > 0 && BN <= 0 && CG == 0 && g->M >= 2048 && g->N >= 512) {
      BN = ov[oi][0];
      CG = ov[oi][1];
    }
  }
  MA3_REQUIRE(CG >= 0 && CG <= 2, "gemm: cta_group must be 0 (auto), 1 or 2");
  if (BN <= 0 || CG == 0) {
    // pick (tile_n, cta_group) by the modelled critical path in SM clocks (constants measured with tools/probe_trace.py)
    int cands[4] = {256, 192, 128, 0}, ncand = 3;
    if (g->N < 256) { cands[0] = (g->N + 15) / 16 * 16; ncand = 1; }
    double best = -1.0;
    int bBN = cands[0], bCG = 1;
    for (int ci = 0; ci < ncand; ++ci) {
      if (g->tile_n > 0 && ci > 0) break;
      const int bn = g->tile_n > 0 ? g->tile_n : cands[ci];
      for (int cg = 1; cg <= 2; ++cg) {
        if (g->cta_group != 0 && cg != g->cta_group) continue;
        if (cg == 2 && (bn % 32 != 0 || bn < 64 || g->M < 256)) continue;
        // (a 3 % bonus for the CTA pair on 3-tap convolutions, suggested by the isolated probe
        // profiles/r02z_probe_conv_tiles.log, measured WORSE inside the step: 84 -> 110 us at 384 channels, 78 -> 103 us
        // at 192, 62 -> 76 us at 768; in-step A/B runs decide, not the isolated probe)
        const double c = tile_cost(g, bn, cg, BK);
        if (best < 0 || c < best) { best = c; bBN = bn; bCG = cg; }
      }
    }
    if (best < 0) { bBN = g->tile_n > 0 ? g->tile_n : cands[0]; bCG = g->cta_group ? g->cta_group : 1; }
    BN = bBN;
    CG = bCG;
  }
  MA3_REQUIRE(BN >= 16 && BN <= 256 && BN % 16 == 0, "gemm: tile_n=%d must be a multiple of 16 in [16,256]", BN);
  if (CG == 2) MA3_REQUIRE(BN % 32 == 0, "gemm: cta_group 2 needs tile_n %% 32 == 0 (got %d)", BN);
  const size_t stage_bytes = (size_t)(kBM + BN / CG) * BK * 2;   // per CTA
  // barriers + epilogue staging + per-warp bias patches of the fused-norm consumer
  const size_t kTail = 256 + kEpiWarps * 32 * kStagePitch * sizeof(float) + kEpiWarps * 256 * sizeof(float);
  const size_t budget = 232448 - 1024 - kTail;
  int stages = (int)(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  MA3_REQUIRE(stages >= 2, "gemm: tile does not fit shared memory");
  // keep one CTA per SM (TMEM is allocated per CTA): request more than half of the SM's shared memory
  size_t smem = 1024 + stages * stage_bytes + kTail;
  if (smem < 120 * 1024) smem = 120 * 1024;

  kp.M = g->M; kp.N = g->N; kp.K = g->K; kp.taps = g->taps;
  for (int i = 0; i < g->taps; ++i) { kp.a_shift[i] = g->a_shift[i]; kp.b_row[i] = g->b_row[i]; }
  kp.a_batched = g->a_batch_stride != 0; kp.b_batched = g->b_batch_stride != 0;
  kp.BN = BN; kp.BK = BK; kp.stages = stages;
  kp.tiles_m = (g->M + kBM * CG - 1) / (kBM * CG); kp.tiles_n = (g->N + BN - 1) / BN; kp.batch = g->batch;
  kp.idesc = umma_idesc(kBM * CG, BN, g->dtype == MA3_BF16 ? 1 : 0);
  kp.tmem_stage_cols = pow2_cols(BN);
  kp.tmem_cols = 2 * kp.tmem_stage_cols;
  kp.op_dtype = g->dtype;

  {
    uint64_t dims[3] = {(uint64_t)g->K, (uint64_t)g->a_rows, (uint64_t)(kp.a_batched ? g->batch : 1)};
    uint64_t str[2] = {(uint64_t)g->a_ld * 2, (uint64_t)(kp.a_batched ? g->a_batch_stride : g->a_rows * g->a_ld) * 2};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)kBM, 1};
    int rc = encode_tmap(&kp.tmA, g->a, 2, 3, dims, str, box, BK * 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)g->K, (uint64_t)g->b_rows, (uint64_t)(kp.b_batched ? g->batch : 1)};
    uint64_t str[2] = {(uint64_t)g->b_ld * 2, (uint64_t)(kp.b_batched ? g->b_batch_stride : g->b_rows * g->b_ld) * 2};
    uint32_t box[3] = {(uint32_t)BK, (uint32_t)(BN / CG), 1};
    int rc = encode_tmap(&kp.tmB, g->b, 2, 3, dims, str, box, BK * 2);
    if (rc) return rc;
  }

  kp.out = g->out; kp.out_dtype = g->out_dtype; kp.out_ld = g->out_ld; kp.out_batch_stride = g->out_batch_stride;
  kp.out_row_mul = g->out_row_mul == 0 ? 1 : g->out_row_mul; kp.out_row_off = g->out_row_off;
  kp.bias = g->bias; kp.bias_per_row = g->bias_per_row;
  kp.res = g->res; kp.res_dtype = g->res_dtype; kp.res_ld = g->res_ld; kp.res_batch_stride = g->res_batch_stride;
  kp.alpha = g->alpha; kp.accumulate = g->accumulate; kp.act = g->act; kp.first_section = g->first_section;
  kp.gate = g->gate; kp.gate_ld = g->gate_ld; kp.gate_batch_stride = g->gate_batch_stride; kp.rows_per_sample = g->rows_per_sample;
  kp.q_out = g->q_out; kp.k_out = g->k_out; kp.vt_out = g->vt_out; kp.rope = g->rope;
  kp.model_dim = g->model_dim; kp.head_dim = g->head_dim; kp.head_dim_pad = g->head_dim_pad;
  kp.heads = g->head_dim > 0 ? g->model_dim / g->head_dim : 0;
  kp.tokens = g->tokens; kp.tokens_pad = g->tokens_pad; kp.q_scale = g->q_scale;
  kp.inv_rows_per_sample = g->rows_per_sample > 0 ? 1.0f / (float)g->rows_per_sample : 0.f;
  kp.inv_tokens = g->tokens > 0 ? 1.0f / (float)g->tokens : 0.f;
  kp.inv_head_dim = g->head_dim > 0 ? 1.0f / (float)g->head_dim : 0.f;
  {
    static const bool qf = !(getenv("MA3_QKV_FAST") && getenv("MA3_QKV_FAST")[0] == '0');
    kp.qkv_fast = qf ? 1 : 0;
  }
  kp.trace = g_trace;
  kp.debug_mode = g_gemm_debug_mode;
  if (g->row_ss != nullptr) {
    MA3_REQUIRE(g->col_bias2 && g->rows_per_sample >= 32 && g->ss_dim > 0 && g->ss_cols > 0 && g->ss_cols % 4 == 0 &&
                    g->ss_cols <= 64,
                "gemm/fused-norm consumer: col_bias2, rows_per_sample >= 32, ss_dim, ss_cols %% 4 == 0 (<= 64) required");
    MA3_REQUIRE(g->N % 32 == 0 && BN % 32 == 0 && g->batch == 1 && aligned16(g->row_ss) && aligned16(g->col_bias2) &&
                    g->col_bias2_ld % 4 == 0,
                "gemm/fused-norm consumer: N and tile_n must be multiples of 32, batch 1, 16-byte aligned tables");
    MA3_REQUIRE(g->epi == MA3_EPI_QKV_ROPE || g->epi == MA3_EPI_SWIGLU,
                "gemm/fused-norm consumer: only the QKV_ROPE and SWIGLU epilogues take a normalised input in the DiT");
    kp.row_ss = g->row_ss; kp.ss_cols = g->ss_cols; kp.ss_inv_dim = 1.0f / (float)g->ss_dim; kp.ss_eps = g->ss_eps;
    kp.col_bias2 = g->col_bias2; kp.col_bias2_ld = g->col_bias2_ld;
  }

  const int total_tiles = kp.tiles_m * kp.tiles_n * kp.batch;
  const int workers = num_sms() / CG;   // CTAs (CG = 1) or CTA pairs (CG = 2)
  const int grid = (total_tiles < workers ? total_tiles : workers) * CG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  switch (g->epi) {
    case MA3_EPI_STORE: {
      MA3_REQUIRE(g->out, "gemm/store: null out");
      MA3_REQUIRE(g->out_dtype >= MA3_F32 && g->out_dtype <= MA3_F16, "gemm/store: bad out dtype");
      const size_t ob = dtype_bytes(g->out_dtype);
      bool vec = aligned16(g->out) && (g->out_ld * ob) % 16 == 0 && (g->out_batch_stride * ob) % 16 == 0 &&
                 (g->out_dtype != MA3_F32 || (g->out_ld % 8 == 0 && g->out_batch_stride % 8 == 0));
      if (g->out_dtype == MA3_F32)
        vec = aligned16(g->out) && g->out_ld % 4 == 0 && g->out_batch_stride % 4 == 0;
      else
        vec = aligned16(g->out) && g->out_ld % 8 == 0 && g->out_batch_stride % 8 == 0;
      if (g->res) {
        if (g->res_dtype == MA3_F32)
          vec = vec && aligned16(g->res) && g->res_ld % 4 == 0 && g->res_batch_stride % 4 == 0;
        else
          vec = vec && aligned16(g->res) && g->res_ld % 8 == 0 && g->res_batch_stride % 8 == 0;
      }
      kp.vec_ok = vec ? 1 : 0;
      if (kp.alpha == 0.f) kp.alpha = 1.f;
      {
        const int rc = try_conv_narrow(g, kp, st);   // narrow multi-tap layers: input tile staged once (see above)
        if (rc != 1) return rc;
      }
      return launch<MA3_EPI_STORE>(kp, smem, grid, CG, st);
    }
    case MA3_EPI_GATE_RES:
      MA3_REQUIRE(g->out && g->gate && g->rows_per_sample > 0, "gemm/gate_res: out, gate, rows_per_sample required");
      // batch > 1: z selects a column slice of the same residual stream (out_batch_stride / gate_batch_stride columns)
      MA3_REQUIRE(g->batch >= 1 && g->out_batch_stride % 4 == 0 && g->gate_batch_stride % 4 == 0,
                  "gemm/gate_res: batch strides of out and gate must be multiples of 4");
      MA3_REQUIRE(g->batch == 1 || g->stream_k != 1, "gemm/gate_res: stream-K is not available with batch > 1");
      MA3_REQUIRE(g->N % 4 == 0 && g->out_ld % 4 == 0 && g->gate_ld % 4 == 0 && aligned16(g->out) && aligned16(g->gate),
                  "gemm/gate_res: N, out_ld, gate_ld must be multiples of 4 and pointers 16-byte aligned");
      MA3_REQUIRE(g->stream_k >= -1 && g->stream_k <= 1, "gemm/gate_res: stream_k must be -1, 0 or 1");
      if (g->norm_out != nullptr) {
        MA3_REQUIRE(g->norm_w && g->ss_out && g->N % 32 == 0 && BN % 32 == 0 && g->stream_k != 1,
                    "gemm/gate_res fused norm: norm_w, ss_out, N %% 32 == 0, tile_n %% 32 == 0, no stream-K");
        MA3_REQUIRE(g->ss_cols >= g->N / 32 && g->ss_cols % 4 == 0, "gemm/gate_res fused norm: ss_cols >= N/32, %% 4 == 0");
        MA3_REQUIRE(g->rows_per_sample >= 32, "gemm/gate_res fused norm: rows_per_sample must be >= 32");
        kp.ss_cols = g->ss_cols;
        MA3_REQUIRE(aligned16(g->norm_out) && aligned16(g->norm_w) && aligned16(g->ss_out) && g->out_ld % 8 == 0,
                    "gemm/gate_res fused norm: 16-byte aligned pointers, out_ld %% 8 == 0");
        kp.norm_out = g->norm_out; kp.norm_w = g->norm_w; kp.ss_out = g->ss_out;
        return launch<MA3_EPI_GATE_RES_NORM>(kp, smem, grid, CG, st);
      }
      {
        // stream-K when whole tiles would leave a large part of the last wave idle and every worker still gets a
        // reasonable run of k-iterations
        const long long total_iters = (long long)total_tiles * kp.taps * (kp.K / kp.BK);
        const int waves = (total_tiles + workers - 1) / workers;
        const bool unbalanced = (long long)total_tiles * 100 < (long long)waves * workers * 92;
        // measured neutral-to-slower on the DiT shapes (the partial tiles add RED-bound epilogues), so only on request
        const bool on = g->stream_k == 1 && unbalanced;
        if (on && total_iters >= workers) {
          kp.stream_k = 1;
          return launch<MA3_EPI_GATE_RES>(kp, smem, workers * CG, CG, st);
        }
      }
      return launch<MA3_EPI_GATE_RES>(kp, smem, grid, CG, st);
    case MA3_EPI_SWIGLU:
      MA3_REQUIRE(g->out && g->batch >= 1, "gemm/swiglu: out required");
      MA3_REQUIRE(g->N % 16 == 0 && g->out_ld % 8 == 0 && g->out_batch_stride % 8 == 0 && aligned16(g->out) &&
                      g->out_dtype != MA3_F32,
                  "gemm/swiglu: N %% 16, out_ld %% 8, out_batch_stride %% 8, 16-bit out required");
      return launch<MA3_EPI_SWIGLU>(kp, smem, grid, CG, st);
    case MA3_EPI_QKV_ROPE:
      MA3_REQUIRE(g->q_out && g->k_out && g->vt_out, "gemm/qkv_rope: q_out, k_out, vt_out required");
      MA3_REQUIRE(g->first_section == 0 || g->first_section == 1, "gemm/qkv_rope: first_section must be 0 or 1");
      MA3_REQUIRE(g->batch == 1 && g->N == (3 - g->first_section) * g->model_dim,
                  "gemm/qkv_rope: N must be (3 - first_section)*model_dim, batch 1");
      MA3_REQUIRE(g->head_dim % 8 == 0 && g->head_dim_pad % 8 == 0 && g->model_dim % g->head_dim == 0 &&
                      g->head_dim_pad >= g->head_dim,
                  "gemm/qkv_rope: head_dim must be a multiple of 8 dividing model_dim");
      MA3_REQUIRE(g->tokens > 0 && g->M % g->tokens == 0 && g->tokens_pad >= g->tokens,
                  "gemm/qkv_rope: M must be samples*tokens");
      MA3_REQUIRE(aligned16(g->q_out) && aligned16(g->k_out), "gemm/qkv_rope: outputs must be 16-byte aligned");
      return launch<MA3_EPI_QKV_ROPE>(kp, smem, grid, CG, st);
    default:
      MA3_FAIL(MA3_EINVAL, "gemm: unknown epilogue %d", g->epi);
  }
}

/* diagnostics: device buffer of >= 256 int64 that CTA 0 of subsequent ma3_gemm launches fills with clock64()
 * timestamps (per tile: MMA thread slots 0-2, epilogue warp slots 4-6); pass NULL to switch tracing off. */
extern "C" int ma3_debug_set_gemm_mode(int mode) {
  ma3::g_gemm_debug_mode = mode;
  return 0;
}
extern "C" int ma3_debug_set_gemm_trace(void* buf) {
  ma3::g_trace = reinterpret_cast<long long*>(buf);
  return 0;
}
