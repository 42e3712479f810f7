/*
 * ma3_b200.h -- C ABI of libma3b200.so: the sm_100a kernels behind the Make-An-Audio-3 sampling path.
 *
 * The reference (GiovanniPriore/Make-An-Audio-3) has no FFI: its plugin boundary is the YAML `target:` string
 * resolved by ldm/util.py:110-125 (instantiate_from_config).  The Python classes in ma3_b200/ mirror the reference
 * classes behind that boundary and call the entry points below through ctypes; each entry point cites the reference
 * code whose arithmetic it replaces.
 *
 * Conventions (all entry points):
 *   - plain C: raw device pointers, sizes, strides; no torch / C++ types in the signatures;
 *   - caller allocates every output and workspace; nothing is allocated or freed inside the library;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), never synchronised -> CUDA-graph capturable;
 *   - returns 0 on success, a negative MA3_E* code on bad arguments, or a positive cudaError_t from the launch;
 *   - sm_100a only: there is no other code path.
 */
#ifndef MA3_B200_H_
#define MA3_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MA3_OK 0
#define MA3_EINVAL (-22)
#define MA3_ENOSYS (-38)

/* element types */
#define MA3_F32 0
#define MA3_BF16 1
#define MA3_F16 2

/* library / device introspection */
int ma3_version(void);
/* 0 when the current device is sm_100 and the driver exposes cuTensorMapEncodeTiled; negative otherwise. */
int ma3_check_device(void);
/* number of kernel launches enqueued by this library in this process so far (bench.py reports the delta). */
int64_t ma3_launch_count(void);
/* Programmatic dependent launch for the launches that follow (kernel attribute, baked into captured graph nodes): 1 on,
 * 0 off, -1 back to the default (off).  The environment variable MA3_PDL=0|1 overrides every request.  On = each kernel's
 * prologue overlaps its predecessor's tail: worth 3-5 % on single-clip workloads, -1..2 % on power-capped large batches. */
int ma3_set_pdl(int on);
/* L2 access-policy window for launches on `stream` (persisting hits on [ptr, ptr+bytes), streaming elsewhere); NULL
 * removes the stream's window.  Used for the DiT's fp32 residual stream, which every block reads twice and reduces
 * into twice.  NOTE: this is the one entry point with a device-wide side effect: it sizes the device's persisting-L2
 * set-aside (cudaLimitPersistingL2CacheSize) to the window, and kernels captured while the window was set keep using
 * it on replay.  ma3_l2_persist_release() gives the set-aside back (limit 0 + cudaCtxResetPersistingL2Cache). */
int ma3_l2_persist(const void* ptr, size_t bytes, void* stream);
int ma3_l2_persist_release(void);
const char* ma3_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Tap-GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators):
 *
 *     acc[z, m, n] = sum_{j < taps} sum_{k < K}  A[z*a_batched, m + a_shift[j], k] * B[z*b_batched, n + b_row[j], k]
 *
 * Both operands are "K-major" (k contiguous), 16-bit (bf16 or fp16), fp32 accumulation.  Rows of A outside
 * [0, a_rows) read as zero (TMA out-of-bounds fill), which is what gives conv1d its zero padding.
 * One kernel covers: nn.Linear (taps = 1), Conv1d of any kernel size / dilation on channels-last activations
 * (tap j: a_shift = j*dilation - padding, b_row = j*C_out), and each phase of ConvTranspose1d.
 * Replaces: every nn.Linear in flag_large_dit.py:56-124,177-210 / flag_large_dit_moe.py:325-408,488-489;
 * torch.nn.Conv1d in autoencoder1d.py:176-295,415-517; Conv1d/ConvTranspose1d in vocoder/bigvgan/models.py:32-205.
 *
 * Epilogues (`epi`):
 *   MA3_EPI_STORE      v = acc (+ bias[n] or bias[m]) (+ res[z,m,n]) ; v = act(alpha*v) ; (+ out_old if accumulate) -> out
 *                      out row = m*out_row_mul + out_row_off (strided rows: transposed-conv phases)
 *   MA3_EPI_GATE_RES   out(f32)[z,m,n] += gate[(m / rows_per_sample), z*gate_batch_stride + n] * acc   (flag_large_dit.py:83-91)
 *                      (+ the fused-RMSNorm producer outputs described at norm_out below)
 *   MA3_EPI_SWIGLU     out[m, n/2] = silu(acc[m, n]) * acc[m, n+1], n even         (flag_large_dit_moe.py:484-489;
 *                      w1 rows interleaved with w3 rows in B); act == 4: tanh-GELU instead of SiLU (the gated-GELU
 *                      feed-forward of the T5 v1.1 text encoder, ldm/modules/encoders/modules.py:178-191)
 *   MA3_EPI_QKV_ROPE   columns [0,D) q, [D,2D) k, [2D,3D) v of one fused projection; rotary embedding on q,k
 *                      (flag_large_dit_moe.py:240-271; rope == NULL: no rotation), q pre-multiplied by q_scale; scatter to
 *                      q,k: [sample, head, t, hd_pad]   v: [sample, head, hd_pad, t_pad] (transposed)
 * ------------------------------------------------------------------------------------------------------------------ */
#define MA3_EPI_STORE 0
#define MA3_EPI_GATE_RES 1
#define MA3_EPI_SWIGLU 2
#define MA3_EPI_QKV_ROPE 3

#define MA3_MAX_TAPS 16

typedef struct ma3_gemm {
  /* operands */
  const void* a;          /* [a_batch][a_rows][K] with row pitch a_ld elements */
  int64_t a_rows, a_ld, a_batch_stride; /* a_batch_stride in elements; 0 = shared by all z */
  const void* b;          /* [b_batch][b_rows][K] with row pitch b_ld elements */
  int64_t b_rows, b_ld, b_batch_stride;
  int32_t dtype;          /* MA3_BF16 or MA3_F16 (both operands: a kind::f16 MMA whose instruction descriptor names
                           * different A and B formats raises an illegal-instruction fault on B200 -- measured) */
  int32_t batch;          /* grid z */
  int32_t M, N, K;        /* output rows per z, output columns, reduction length per tap (multiple of 16) */
  int32_t taps;
  int32_t a_shift[MA3_MAX_TAPS];
  int32_t b_row[MA3_MAX_TAPS];
  /* epilogue */
  int32_t epi;
  void* out;
  int32_t out_dtype;      /* MA3_F32 / MA3_BF16 / MA3_F16 */
  int64_t out_ld, out_batch_stride;
  int32_t out_row_mul, out_row_off;
  const float* bias;      /* nullable, fp32 */
  int32_t bias_per_row;
  const void* res;        /* nullable residual, same indexing as out (incl. row_mul/off) */
  int32_t res_dtype;
  int64_t res_ld, res_batch_stride;
  float alpha;
  int32_t act;            /* STORE: activation after alpha: 0 none, 1 SiLU, 2 GELU(erf), 3 tanh */
  int32_t accumulate;     /* 1: add previous contents of out */
  /* GATE_RES */
  const float* gate;      /* [samples][gate_ld] fp32 */
  int64_t gate_ld;
  int64_t gate_batch_stride; /* elements added to the gate COLUMN per z (batched column slices of one residual stream:
                              * the frequency experts of flag_large_dit_moe.py:516-538 as one launch) */
  int32_t rows_per_sample;
  /* QKV_ROPE */
  void* q_out; void* k_out; void* vt_out;   /* operand dtype */
  const float* rope;      /* [T_max][hd/2][2] (cos, sin) fp32 */
  int32_t model_dim, head_dim, head_dim_pad, tokens, tokens_pad;
  float q_scale;
  int32_t first_section;  /* 0: columns are q|k|v (N = 3*model_dim); 1: k|v only (cross K/V, N = 2*model_dim) */
  /* tiling overrides: 0 = library heuristic */
  int32_t tile_n;
  int32_t cta_group;      /* 1: one CTA per 128-row tile; 2: CTA pair (tcgen05 cta_group::2) per 256-row tile */
  int32_t stream_k;       /* MA3_EPI_GATE_RES only.  0: library heuristic; 1: split the tiles x k-iterations space evenly
                           * over the SMs (partial products are added by separate reductions: fp32 sums may differ in the
                           * last bit from run to run); -1: whole tiles only (bit-reproducible) */
  /* Fused RMSNorm + adaLN modulate (flag_large_dit.py:79-91, flag_large_dit_moe.py:63-81) -- the stand-alone
   * normalisation pass between a gated-residual GEMM and the GEMM that consumes the normalised rows is removed by
   * splitting  u = rms(h) * w * (1 + scale_s) + shift_s  into a row scalar and a per-sample bias:
   *     u W^T = rstd[m] * ((h * wn_s) W^T) + (shift_s W^T),      wn_s = w * (1 + scale_s),  rstd = rsqrt(mean(h^2) + eps)
   * Producer (MA3_EPI_GATE_RES with norm_out != NULL; whole tiles, N % 32 == 0): the epilogue owns its elements
   * (no atomics): h_new = h_old + gate * acc -> out (fp32);  norm_out[m, n] = 16-bit(h_new * norm_w[sample, n]);
   * ss_out[m, n / 32] = sum over the 32-column chunk of h_new^2 (deterministic per-chunk partial sums). */
  void* norm_out;         /* [M][out_ld], operand dtype */
  const float* norm_w;    /* [samples][gate_ld] fp32 (same row pitch as gate) */
  float* ss_out;          /* [M][ss_cols] fp32, ss_cols >= N / 32 and a multiple of 4 (pad columns must hold zeros) */
  /* Consumer (MA3_EPI_QKV_ROPE / MA3_EPI_SWIGLU; N % 32 == 0, rows_per_sample > 0): before the epilogue proper
   *     acc[m, n] <- acc[m, n] * rsqrt(sum_j row_ss[m, j] / ss_dim + ss_eps) + col_bias2[m / rows_per_sample, n] */
  const float* row_ss;    /* [M][ss_cols] fp32 partial sums of squares (a producer's ss_out); NULL = off */
  int32_t ss_cols, ss_dim;
  float ss_eps;
  const float* col_bias2; /* [samples][col_bias2_ld] fp32 = shift_s W^T */
  int64_t col_bias2_ld;
} ma3_gemm_t;

int ma3_gemm(const ma3_gemm_t* g, void* stream);

/* Row-owning gated-residual GEMM with the following RMSNorm + adaLN modulate fused in -- the wo and w2 projections of a
 * Next-DiT block together with the normalisation that feeds the next projection (flag_large_dit.py:79-91,
 * flag_large_dit_moe.py:63-81):
 *     h[m, :] <- h[m, :] + gate[s, :] * (a[m, :] . w^T)                                   (fp32, in place)
 *     u_out[m, :] <- 16-bit( h_new[m, :] * rsqrt(mean_D(h_new[m, :]^2) + eps) * wn[s, :] + shift[s, :] )
 * with s = m / rows_per_sample and wn = norm_weight * (1 + scale) (ma3_norm_weights).  a [M][K] and w [D][K] are 16-bit
 * (`dtype`), row pitches a_ld / w_ld in elements; gate / wn / shift are fp32 [samples][mod_ld].  u_out == NULL: only the
 * residual update (last block).  D must be 384 * {1, 2, 3, 4}: a thread-block cluster of D / 384 CTAs owns 128 complete
 * rows (384 TMEM columns per CTA), the row sums of squares cross the cluster through distributed shared memory.  No
 * atomics: bit-reproducible. */
int ma3_gemm_rownorm(const void* a, int64_t a_ld, const void* w, int64_t w_ld, int dtype, int M, int K, int D, float* h,
                     const float* gate, const float* wn, const float* shift, int64_t mod_ld, int rows_per_sample,
                     void* u_out, float eps, void* stream);
/* diagnostics only: device buffer (>= 256 int64) that CTA 0 of later ma3_gemm launches fills with clock64() stamps of
 * its pipeline events; NULL switches tracing off (tools/probe_trace.py). */
int ma3_debug_set_gemm_trace(void* buf);
/* diagnostics only: 0 = normal; 1 = the producer signals stages full without issuing TMA loads; 2 = the MMA thread
 * releases stages without issuing MMAs.  Isolates the feed side from the tensor side of the mainloop; the output of
 * later ma3_gemm calls is garbage until mode 0 is restored (tools/probe_trace.py). */
int ma3_debug_set_gemm_mode(int mode);
/* Attention kernel generation: 0 = default, 2 / 3 = force v2 / v3 (v3: several query tiles of a head per CTA; kept as a
 * measured alternative).  Diagnostics and tests only. */
int ma3_debug_set_attn_version(int v);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused flash attention of one Next-DiT block: self-attention over the T latent tokens plus tanh-gated
 * cross-attention over the L context tokens, sharing the rotary-embedded Q tile.
 *   out[ns, t, h*hd + d] = softmax(q k^T) v + tanh(gate[h]) * softmax(q ky^T) vy
 * Replaces Attention.forward's two F.scaled_dot_product_attention calls and the gate combine
 * (flag_large_dit_moe.py:382-406).  q must already carry RoPE and the factor log2(e)/sqrt(hd) (MA3_EPI_QKV_ROPE).
 * Layouts (16-bit, `dtype`): q,k [NS*H, T, hdp]; vt [NS*H, hdp, Tp]; ky [NS*H, L, hdp]; vyt [NS*H, hdp, Lp];
 * gate [H] fp32 (raw parameter); out [NS, T, H*hd].  hdp in {64,128}; Tp, Lp multiples of 8.
 * Pad columns [hd, hdp) of q, k, ky must be zero.  When hd < hdp, row hd of vt and of vyt must hold 1.0 for every
 * token and rows (hd, hdp) zero: the P.V tensor-core product then also yields the softmax row sums (column hd of the
 * accumulator), so the denominator is formed from exactly the rounded probabilities the MMA consumed.  The host side
 * allocates such buffers once (ops.alloc_vt); MA3_EPI_QKV_ROPE only ever writes rows [0, hd).
 * ------------------------------------------------------------------------------------------------------------------ */
int ma3_attention(const void* q, const void* k, const void* vt, const void* ky, const void* vyt, const float* gate,
                  void* out, int dtype, int NS, int H, int T, int Tp, int L, int Lp, int hd, int hdp, void* stream);

/* out[m,:] = x[m,:] * rsqrt(mean(x^2)+eps) * w * (1 + scale[s,:]) + shift[s,:],  s = m / rows_per_sample, with
 * shift = mod[s, shift_off : shift_off+D], scale = mod[s, scale_off : ...] (fp32).  w == NULL: no weight;
 * mod == NULL: plain RMSNorm.  x fp32 [M, D].  Replaces RMSNorm + modulate (flag_large_dit_moe.py:34-81) as used at
 * flag_large_dit.py:83-91. */
int ma3_rmsnorm_modulate(const float* x, const float* w, const float* mod, int64_t mod_ld, int shift_off,
                         int scale_off, int rows_per_sample, void* out, int out_dtype, int M, int D, float eps,
                         void* stream);

/* FinalLayer (flag_large_dit.py:101-124): LayerNorm(no affine, eps) -> modulate -> Linear(D -> Cout), written
 * transposed as v_out[N, Cout, T] fp32 (the 'b t c -> b c t' rearrange of flag_large_dit.py:209). */
int ma3_final_layer(const float* h, const float* mod, int64_t mod_ld, int shift_off, int scale_off, const float* W,
                    const float* bias, int N, int T, int D, int Cout, float eps, float* v_out, void* stream);

/* FinalLayer fused with the classifier-free-guidance combine and the Euler update: rows [0, N/2) are the
 * unconditional half, [N/2, N) the conditional half (cfm1_audio.py:154-161);
 *   v = v_u + guidance (v_c - v_u);  x_out = x_in + dt v;   v_out (nullable) receives v.   x, v: [N/2, Cout, T]. */
int ma3_final_layer_cfg_euler(const float* h, const float* mod, int64_t mod_ld, int shift_off, int scale_off,
                              const float* W, const float* bias, int N, int T, int D, int Cout, float eps,
                              float guidance, float dt, const float* x_in, float* x_out, float* v_out, void* stream);

/* Stand-alone guidance combine + Euler update on velocities v [2*B or B, ...] (cfg = 1 / 0); elems = numel(x). */
int ma3_cfg_euler_update(const float* v, const float* x, float* out, int64_t elems, float dt, float guidance, int cfg,
                         void* stream);

/* h[n*T + t, :] = x[n % x_batch, :, t] Wt + b   (proj_in on the 'b c t -> b t c' view, flag_large_dit.py:186-187);
 * Wt is the TRANSPOSED nn.Linear weight, [C, D] row-major. */
int ma3_proj_in(const float* x, const float* Wt, const float* b, float* h, int N, int x_batch, int C, int T, int D,
                void* stream);

/* Sinusoidal timestep embedding [cos | sin] (flag_large_dit_moe.py:110-127); t int64 [M] -> out [M, dim]. */
int ma3_timestep_embed(const int64_t* t, void* out, int out_dtype, int M, int dim, void* stream);

/* out[n,:] = LayerNorm_affine(mean over L of ctx[n,:,:])   (flag_large_dit.py:193-198, first half of cap_embedder). */
int ma3_pool_layernorm(const void* ctx, int in_dtype, const float* w, const float* b, void* out, int out_dtype, int N,
                       int L, int Cd, float eps, void* stream);

/* Row-wise affine LayerNorm (trailing nn.LayerNorm of ConditionEmbedder, flag_large_dit_moe.py:151-162). */
int ma3_layernorm_rows(const void* x, int in_dtype, const float* w, const float* b, void* out, int out_dtype, int M,
                       int D, float eps, void* stream);

/* out[s*N + n, :] = silu(temb[s*ts_s + n*ts_n, :] + cap[n, :])  -- adaln_input = t_emb + cap_emb followed by the SiLU
 * at the head of every adaLN_modulation (flag_large_dit.py:50-51,200). */
int ma3_adaln_input(const float* temb, const float* cap, void* out, int out_dtype, int S, int N, int D, int ts_s,
                    int ts_n, void* stream);

/* hi / lo split of fp32 operands: for slice b = columns [col0 + b*col_step, +cols) of x [rows][ld],
 * out[b] = [2*rows][cols] bf16 with rows [0, rows) = bf16(x) and rows [rows, 2*rows) = bf16(x - bf16(x)).  A tap-GEMM
 * over (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo) then carries ~16 mantissa bits through the bf16 tensor cores.  Used for
 * the step-invariant conditioning path: TimestepEmbedder (flag_large_dit_moe.py:96-133), cap_embedder
 * (flag_large_dit.py:171-174,198), every adaLN_modulation Linear (flag_large_dit.py:50-51,120-124) and the shift_s W^T
 * bias tables of the fused RMSNorm (ma3_gemm_t.col_bias2). */
int ma3_split_bf16(const float* x, int64_t ld, int col0, int col_step, int nb, int rows, int cols, void* out,
                   void* stream);

/* qk_norm=True variant of the QKV path (flag_large_dit_moe.py:199-207,345-352): x fp32 [M][ld] holds the raw projections
 * q | k | v (first_section 0) or k | v (first_section 1, cross-attention with ky_norm); per row: LayerNorm over the full
 * model dim of q (qw, qb) and k (kw, kb) (NULL weights = no norm), RoPE (rope NULL = none), q * q_scale, and the scatter
 * into q, k [sample, head, t, hd_pad] and V^T [sample, head, hd_pad, t_pad] (same layouts as MA3_EPI_QKV_ROPE). */
int ma3_qknorm_rope(const float* x, int64_t ld, int first_section, const float* qw, const float* qb, const float* kw,
                    const float* kb, float eps, const float* rope, void* q_out, void* k_out, void* vt_out, int dtype, int M,
                    int tokens, int tokens_pad, int D, int hd, int hdp, float q_scale, void* stream);

/* Mel front-end (preprocess/NAT_mel.py:65-85, MelNet.forward, center=False): the elementwise glue around two tap-GEMMs
 * (STFT as a 4-tap GEMM over hop-sized rows with the windowed DFT basis, then the mel filterbank), all 16-bit operands as
 * (hi, lo) bf16 splits.
 *   ma3_melnet_prep: y [B][n] fp32 -> hops [B][2][nh][hop] bf16 of clamp(y, -1, 1) reflect-padded by `pad` on both sides
 *   ma3_melnet_mag:  S [B*F][ld] fp32 (re, im interleaved) -> mag [B][2][F][bins_pad] bf16 split of sqrt(re^2+im^2+1e-9)
 *   ma3_melnet_log:  mel [B*F][n_mels] fp32 -> out [B][n_mels][F] = log10(max(mel, 1e-5)) */
int ma3_melnet_prep(const float* y, void* hops, int B, int n, int pad, int nh, int hop, void* stream);
int ma3_melnet_mag(const float* S, int64_t ld, void* mag, int B, int F, int bins, int bins_pad, void* stream);
int ma3_melnet_log(const float* mel, float* out, int B, int F, int n_mels, void* stream);

/* mod[r, tail_off + (2i+j)*D + d] = norm_w[i][j][d] * (1 + mod[r, 6*D*i + (j ? 4*D : D) + d]) for r < rows, i < depth,
 * j in {0: attention_norm, 1: ffn_norm}: wn_s = w * (1 + scale_s) of flag_large_dit.py:83-91, written behind the
 * modulation columns of the same row (ma3_gemm_t.norm_w). */
int ma3_norm_weights(float* mod, int64_t ld, const float* norm_w, int rows, int depth, int D, int tail_off, void* stream);

/* GroupNorm(groups, eps, affine) optionally followed by swish on channels-last x [B, T, C]
 * (Normalize + nonlinearity, autoencoder1d.py:169-175). */
int ma3_groupnorm_swish(const void* x, int in_dtype, const float* w, const float* b, void* out, int out_dtype, int B,
                        int T, int C, int groups, float eps, int swish, void* stream);

/* P[r, 0:n] = softmax(scale * S[r, 0:n]), P[r, n:ld_out] = 0   (AttnBlock1D, autoencoder1d.py:265-270). */
int ma3_softmax_rows(const float* S, void* P, int out_dtype, int rows, int n, int64_t ld_in, int64_t ld_out,
                     float scale, void* stream);

/* Layout changes between the reference's [B, C, T] fp32 tensors and channels-last 16-bit activations. */
int ma3_nct_to_ntc(const float* x, void* out, int out_dtype, int B, int C, int T, int Cp, float scale, void* stream);
int ma3_ntc_to_nct(const void* x, int in_dtype, float* out, int B, int C, int T, int64_t ld, void* stream);
/* nearest-neighbour x2 along T (Upsample1D, autoencoder1d.py:291-292) on channels-last 16-bit rows. */
int ma3_upsample_nearest2(const void* x, void* out, int64_t rows, int C, void* stream);
int ma3_cast(const void* x, int in_dtype, void* out, int out_dtype, int64_t n, void* stream);
/* Token embedding lookup of the text encoders (the nn.Embedding inside the CLAP-BERT and T5 encoders that
 * ldm/modules/encoders/modules.py:178-191 calls): out[m, :] = table[ids[m], :] (+ pos[m % T, :]) (+ type0[:]), fp32.
 * ids are int64 (torch.long); the caller validates them (an id outside [0, vocab) reads row 0); pos / type0 may be NULL. */
int ma3_embed_rows(const float* table, int64_t vocab, const int64_t* ids, const float* pos, const float* type0, float* out,
                   int M, int T, int D, void* stream);

/* Fused anti-aliased periodic activation (Activation1d.forward, vocoder/bigvgan/alias_free_torch/act.py:23-28):
 * replicate-pad -> x2 up-sampling with the 12-tap Kaiser-sinc filter (resample.py:25-33) -> SnakeBeta / Snake
 * (activations.py:48-59,107-119; beta == NULL selects Snake) -> replicate-pad -> low-pass, stride 2
 * (filter.py:86-95), one pass over HBM on channels-last x [B, T, C], C % 16 == 0.
 * ma3_act1d_set_filter uploads the 12 taps (host pointer) once per process. */
int ma3_act1d_set_filter(const float* taps12, void* stream);
int ma3_act1d(const void* x, int in_dtype, void* out, int out_dtype, const float* alpha, const float* beta, int B,
              int T, int C, int logscale, void* stream);
/* Activation1d kernel generation for fp16 -> fp16, T % 8 == 0: 0 = default (tiles staged by TMA loads / stores, a warp
 * per 16 channels x 128 outputs), 1 = the first-generation kernel (cp.async staging).  Diagnostics and tests only. */
int ma3_debug_set_act_version(int v);

#ifdef __cplusplus
}
#endif
#endif /* MA3_B200_H_ */
