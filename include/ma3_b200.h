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
 *   MA3_EPI_GATE_RES   out(f32)[m,n] += gate[(m / rows_per_sample), n] * acc        (flag_large_dit.py:83-91)
 *   MA3_EPI_SWIGLU     out[m, n/2] = silu(acc[m, n]) * acc[m, n+1], n even         (flag_large_dit_moe.py:484-489;
 *                      w1 rows interleaved with w3 rows in B)
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
  int32_t dtype;          /* MA3_BF16 or MA3_F16 (both operands) */
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
  int32_t rows_per_sample;
  /* QKV_ROPE */
  void* q_out; void* k_out; void* vt_out;   /* operand dtype */
  const float* rope;      /* [T_max][hd/2][2] (cos, sin) fp32 */
  int32_t model_dim, head_dim, head_dim_pad, tokens, tokens_pad;
  float q_scale;
  int32_t first_section;  /* 0: columns are q|k|v (N = 3*model_dim); 1: k|v only (cross K/V, N = 2*model_dim) */
  /* tiling override: 0 = library heuristic */
  int32_t tile_n;
} ma3_gemm_t;

int ma3_gemm(const ma3_gemm_t* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MA3_B200_H_ */
