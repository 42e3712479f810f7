// Host-side helpers shared by the C-ABI entry points: error reporting, launch accounting, TMA descriptor encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "ma3_b200.h"

namespace ma3 {

extern std::atomic<int64_t> g_launches;
extern thread_local char g_err[512];
extern int g_gemm_debug_mode;  // diagnostics (ma3_debug_set_gemm_mode): 1 = skip TMA loads, 2 = skip MMAs; results are garbage
extern long long* g_trace;  // diagnostics buffer (ma3_debug_set_gemm_trace); nullptr = off

#define MA3_FAIL(code, ...)                       \
  do {                                            \
    snprintf(::ma3::g_err, sizeof(::ma3::g_err), __VA_ARGS__); \
    return (code);                                \
  } while (0)

#define MA3_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) MA3_FAIL(MA3_EINVAL, __VA_ARGS__); \
  } while (0)

// Call after every kernel launch: counts it and converts launch errors to a return code.
#define MA3_LAUNCH_CHECK(name)                                                         \
  do {                                                                                 \
    ::ma3::g_launches.fetch_add(1, std::memory_order_relaxed);                         \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) MA3_FAIL((int)e__, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

int num_sms();

// "Done once" flag for state that lives on the device (function attributes, __constant__ uploads): one bit per device
// ordinal, so a second GPU used from the same process configures its own copy instead of inheriting the first one's.
struct DeviceOnce {
  std::atomic<uint64_t> done{0};
  static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool pending() const { return !((done.load(std::memory_order_relaxed) >> dev()) & 1ull); }
  void mark() { done.fetch_or(1ull << dev(), std::memory_order_relaxed); }
};

// elem_bytes = 2 (bf16/f16) or 4 (f32).  dims/strides innermost first; strides[i] = byte pitch of dim i+1.
// swizzle_bytes in {0, 32, 64, 128}.  Returns 0 or an error code (message in g_err).
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// Programmatic dependent launch: every hot kernel is launched with programmaticStreamSerialization so that its
// launch latency and its prologue (barrier init, TMEM allocation, descriptor prefetch) overlap the tail of the kernel
// before it; the kernel calls pdl_wait() (griddepcontrol.wait) before touching global memory.  Opt-in with MA3_PDL=1.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

inline size_t dtype_bytes(int dt) { return dt == MA3_F32 ? 4 : 2; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace ma3
