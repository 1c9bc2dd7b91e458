// Host-side helpers shared by all translation units of liblsvs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <utility>
#include "../../include/lsvs_b200.h"

namespace lsvs {

extern thread_local char g_err[512];
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- programmatic dependent launch ------------------------------------------------------------------------
// Kernels launched through launch_pdl may become resident while their predecessor in the stream is still draining:
// their prologue (barrier init, tensor-memory allocation, tensor-map prefetch) and the grid launch latency overlap the
// predecessor's tail.  Such a kernel executes pdl_wait() (griddepcontrol.wait) before its first global-memory access
// and pdl_launch_dependents() at its top.  LSVS_PDL=0 in the environment turns the attribute off (A/B measurements).
inline bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("LSVS_PDL"); return !(v && v[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---- optional per-kernel-class CUDA-event profiler (bench.py's roofline numbers; off by default) ----------
enum ProfCat { PROF_GEMM = 0, PROF_ATTENTION = 1, PROF_ELEMENTWISE = 2, PROF_SMALL_F32 = 3, PROF_SIM3 = 4, PROF_ATTENTION_GLOBAL = 5, PROF_NCAT = 6 };
extern bool g_prof_on;
void prof_begin(int cat, cudaStream_t st, double flops, double bytes);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st; bool on;
  ProfScope(int cat, cudaStream_t s, double flops = 0, double bytes = 0) : st(s), on(g_prof_on) { if (on) prof_begin(cat, s, flops, bytes); }
  ~ProfScope() { if (on) prof_end(st); }
};

#define LSVS_CHECK_ARG(cond, ...) \
  do { if (!(cond)) return ::lsvs::fail(LSVS_EINVAL, __VA_ARGS__); } while (0)

#define LSVS_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return ::lsvs::fail(LSVS_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
  } while (0)

#define LSVS_LAUNCH_CHECK()                                                                   \
  do {                                                                                        \
    ::lsvs::count_launch();                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess)                                                                    \
      return ::lsvs::fail(LSVS_ECUDA, "%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

}  // namespace lsvs
