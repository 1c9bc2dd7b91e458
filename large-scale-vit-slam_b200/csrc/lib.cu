// Library-level state and trivial entry points.
#include "host_common.h"

namespace lsvs {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace lsvs

extern "C" const char* lsvs_last_error(void) { return lsvs::g_err; }
extern "C" int lsvs_version(void) { return 1; }
extern "C" unsigned long long lsvs_launch_count(void) { return lsvs::g_launches.load(); }

// ---------------------------------------------------------------------------------------------------
// CUDA-event profiler: one (start, stop) event pair per launch, accumulated per kernel class on read.
#include <vector>
namespace lsvs {
bool g_prof_on = false;
namespace {
struct Rec { int cat; cudaEvent_t a, b; double flops, bytes; };
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
void prof_begin(int cat, cudaStream_t st, double flops, double bytes) {
  Rec r{cat, get_event(), get_event(), flops, bytes};
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
}
void prof_end(cudaStream_t st) { if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, st); }
}  // namespace lsvs

extern "C" int lsvs_profile_enable(int on) {
  lsvs::g_prof_on = on != 0;
  return LSVS_OK;
}

// Synchronises the device, sums the recorded launches per class and clears the log.
// ms / flops / bytes / launches: arrays of LSVS_PROF_NCAT entries (any may be NULL).
extern "C" int lsvs_profile_read(double* ms, double* flops, double* bytes, long long* launches) {
  LSVS_CUDA(cudaDeviceSynchronize());
  for (int c = 0; c < lsvs::PROF_NCAT; ++c) { if (ms) ms[c] = 0; if (flops) flops[c] = 0; if (bytes) bytes[c] = 0; if (launches) launches[c] = 0; }
  for (auto& r : lsvs::g_recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    if (ms) ms[r.cat] += t;
    if (flops) flops[r.cat] += r.flops;
    if (bytes) bytes[r.cat] += r.bytes;
    if (launches) launches[r.cat] += 1;
    lsvs::g_pool.push_back(r.a); lsvs::g_pool.push_back(r.b);
  }
  lsvs::g_recs.clear();
  return LSVS_OK;
}
