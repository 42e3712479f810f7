// Mel front-end on the GPU (preprocess/NAT_mel.py:42-85, MelNet.forward with center=False, complex=False):
//   clamp to [-1, 1] -> reflect-pad (n_fft - hop) / 2 -> STFT (Hann window) -> sqrt(re^2 + im^2 + 1e-9) -> mel filterbank
//   -> log10(clamp(., 1e-5)).
// The STFT is a tap-GEMM on the tensor cores: with n_fft = 4 * hop a frame is four consecutive hop-sized rows of the
// padded signal, so  STFT[r, :] = sum_j hops[r + j, :] . Basis_j^T  with Basis_j the windowed DFT rows of samples
// [j*hop, (j+1)*hop): the frames are never materialised.  All 16-bit operands are (hi, lo) bf16 splits (three products
// per term, fp32 accumulation), which keeps ~16 mantissa bits -- a log-mel of quiet bins does not survive 8.
// The three kernels here are the elementwise glue around the two GEMMs (ma3_gemm): HBM-bound, trivially small.
#include "host_common.h"
#include "ptx.cuh"

namespace ma3 {

__device__ __forceinline__ void split_store(float v, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  *hi = h;
  *lo = __float2bfloat16_rn(v - __bfloat162float(h));
}

// y [B][n] fp32 -> hops [B][2][nh][hop] bf16 (plane 0 = hi, plane 1 = lo) of the clamped, reflect-padded signal;
// positions beyond the padded length are zero.
__global__ void melnet_prep_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ hops, int B, int n, int pad,
                                   int nh, int hop) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (long long)nh * hop;
  if (i >= (long long)B * per) return;
  const int b = (int)(i / per);
  const long long p = i - (long long)b * per;          // position in the padded signal
  float v = 0.f;
  if (p < (long long)n + 2 * pad) {
    long long s = p - pad;                             // reflect (no edge repeat), as F.pad(mode='reflect')
    if (s < 0) s = -s;
    if (s >= n) s = 2LL * (n - 1) - s;
    v = fminf(fmaxf(y[(long long)b * n + s], -1.f), 1.f);
  }
  __nv_bfloat16* base = hops + (long long)b * 2 * per;
  split_store(v, base + p, base + per + p);
}

// S [B*F][ld] fp32 with (re, im) interleaved per bin -> mag [B][2][F][bins_pad] bf16 split of sqrt(re^2 + im^2 + 1e-9)
__global__ void melnet_mag_kernel(const float* __restrict__ S, long long ld, __nv_bfloat16* __restrict__ mag, int B, int F,
                                  int bins, int bins_pad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * F * bins_pad) return;
  const int k = (int)(i % bins_pad);
  const long long bf = i / bins_pad;
  const int f = (int)(bf % F), b = (int)(bf / F);
  float v = 0.f;
  if (k < bins) {
    const float2 c = *reinterpret_cast<const float2*>(S + bf * ld + 2 * k);
    v = sqrtf(c.x * c.x + c.y * c.y + 1e-9f);
  }
  __nv_bfloat16* base = mag + (long long)b * 2 * F * bins_pad;
  split_store(v, base + (long long)f * bins_pad + k, base + (long long)(F + f) * bins_pad + k);
}

// mel [B*F][n_mels] fp32 -> out [B][n_mels][F] = log10(max(mel, 1e-5))
__global__ void melnet_log_kernel(const float* __restrict__ mel, float* __restrict__ out, int B, int F, int n_mels) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * F * n_mels) return;
  const int f = (int)(i % F);
  const long long bm = i / F;
  const int m = (int)(bm % n_mels), b = (int)(bm / n_mels);
  out[i] = log10f(fmaxf(mel[((long long)b * F + f) * n_mels + m], 1e-5f));
}

}  // namespace ma3

using namespace ma3;
static inline unsigned mel_nblk(long long n) { return (unsigned)((n + 255) / 256); }

extern "C" {

int ma3_melnet_prep(const float* y, void* hops, int B, int n, int pad, int nh, int hop, void* stream) {
  MA3_REQUIRE(y && hops && B > 0 && n > pad && pad >= 0 && nh > 0 && hop > 0, "melnet_prep: bad arguments (n must exceed the pad)");
  MA3_REQUIRE((long long)nh * hop >= (long long)n + 2 * pad, "melnet_prep: nh * hop must cover the padded signal");
  melnet_prep_kernel<<<mel_nblk((long long)B * nh * hop), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      y, (__nv_bfloat16*)hops, B, n, pad, nh, hop);
  MA3_LAUNCH_CHECK("melnet_prep");
  return 0;
}

int ma3_melnet_mag(const float* S, int64_t ld, void* mag, int B, int F, int bins, int bins_pad, void* stream) {
  MA3_REQUIRE(S && mag && B > 0 && F > 0 && bins > 0 && bins_pad >= bins && ld >= 2 * bins && ld % 2 == 0, "melnet_mag: bad arguments");
  melnet_mag_kernel<<<mel_nblk((long long)B * F * bins_pad), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      S, ld, (__nv_bfloat16*)mag, B, F, bins, bins_pad);
  MA3_LAUNCH_CHECK("melnet_mag");
  return 0;
}

int ma3_melnet_log(const float* mel, float* out, int B, int F, int n_mels, void* stream) {
  MA3_REQUIRE(mel && out && B > 0 && F > 0 && n_mels > 0, "melnet_log: bad arguments");
  melnet_log_kernel<<<mel_nblk((long long)B * F * n_mels), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mel, out, B, F, n_mels);
  MA3_LAUNCH_CHECK("melnet_log");
  return 0;
}

}  // extern "C"
