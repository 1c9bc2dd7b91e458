// Library-level state and trivial entry points.
#include "host_common.h"

namespace lsvs {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace lsvs

extern "C" const char* lsvs_last_error(void) { return lsvs::g_err; }
extern "C" int lsvs_version(void) { return 1; }
extern "C" unsigned long long lsvs_launch_count(void) { return lsvs::g_launches.load(); }
