// Library-wide state and the TMA descriptor encoder (driver entry point resolved at run time so the .so loads on a
// machine without libcuda, e.g. the CPU-only CI container).
#include "host_common.h"

#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>

namespace ma3 {

std::atomic<int64_t> g_launches{0};
thread_local char g_err[512] = {0};
long long* g_trace = nullptr;
int g_gemm_debug_mode = 0;

// Programmatic dependent launch.  Measured (round 2, same box A/B): +2.7 % / +5 % on the single-clip configurations (M,
// video/MoE: not power-capped, ~3000 launches of ~8 us each), -1.3 % / -1.9 % on the power-capped 8- and 16-prompt
// batches (overlapped prologues only lower the clock the controller grants).  So the host layer switches it per
// workload (ma3_set_pdl, called by the sampler / pipeline before a plan is captured); MA3_PDL=0|1 pins it.
static int g_pdl_request = -1;   // -1: no request (off); 0 / 1: set by ma3_set_pdl
bool pdl_enabled() {
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("MA3_PDL");
    env = e ? (e[0] == '1' ? 1 : 0) : -1;
  }
  if (env >= 0) return env != 0;
  return g_pdl_request == 1;
}

int num_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 63];
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
  }
  return v;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  auto fn = get_encode();
  if (!fn) MA3_FAIL(MA3_ENOSYS, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  // 16-bit operands are bf16 or fp16 depending on the caller: the maps only ever copy (no OOB NaN fill, no
  // arithmetic), so they are typed as plain 16-bit words rather than mislabelled as one of the two.
  CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    MA3_FAIL(MA3_EINVAL,
             "cuTensorMapEncodeTiled failed (%d): base=%p rank=%d dims=[%llu,%llu,%llu] strides=[%llu,%llu] "
             "box=[%u,%u,%u] sw=%d",
             (int)r, base, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 1 ? gstr[0] : 0),
             (unsigned long long)(rank > 2 ? gstr[1] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
             swizzle_bytes);
  return 0;
}

}  // namespace ma3

extern "C" {

int ma3_version(void) { return 100; }

int64_t ma3_launch_count(void) { return ma3::g_launches.load(); }

int ma3_set_pdl(int on) {
  ma3::g_pdl_request = on > 0 ? 1 : (on == 0 ? 0 : -1);
  return 0;
}

const char* ma3_last_error(void) { return ma3::g_err; }

int ma3_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) MA3_FAIL(MA3_ENOSYS, "no CUDA device: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) MA3_FAIL(MA3_ENOSYS, "device is sm_%d%d; this library is sm_100a only", major, minor);
  if (!ma3::get_encode()) MA3_FAIL(MA3_ENOSYS, "driver lacks cuTensorMapEncodeTiled");
  return 0;
}

/* Mark [ptr, ptr + bytes) as L2-persisting for kernels launched (or captured) on `stream` from now on: hits keep the
 * lines resident, other addresses are treated as streaming.  ptr == NULL removes the window.  The set-aside is sized
 * to the window (capped by the device limit). */
int ma3_l2_persist(const void* ptr, size_t bytes, void* stream) {
  int dev = 0;
  cudaGetDevice(&dev);
  int max_persist = 0, max_window = 0;
  cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
  cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
  cudaStreamAttrValue v;
  memset(&v, 0, sizeof(v));
  if (ptr != nullptr && bytes > 0) {
    size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
    size_t keep = win < (size_t)max_persist ? win : (size_t)max_persist;
    cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, keep);
    if (e != cudaSuccess) MA3_FAIL((int)e, "cudaDeviceSetLimit(persisting L2): %s", cudaGetErrorString(e));
    v.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
    v.accessPolicyWindow.num_bytes = win;
    v.accessPolicyWindow.hitRatio = win > 0 ? (float)keep / (float)win : 0.f;
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  }
  cudaError_t e = cudaStreamSetAttribute(reinterpret_cast<cudaStream_t>(stream), cudaStreamAttributeAccessPolicyWindow, &v);
  if (e != cudaSuccess) MA3_FAIL((int)e, "cudaStreamSetAttribute(access policy window): %s", cudaGetErrorString(e));
  return 0;
}

int ma3_l2_persist_release(void) {
  cudaError_t e = cudaCtxResetPersistingL2Cache();
  if (e == cudaSuccess) e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  if (e != cudaSuccess) MA3_FAIL((int)e, "ma3_l2_persist_release: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
