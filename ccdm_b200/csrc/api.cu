// Library-level entry points: version, thread-local error string, launch counter.
#include <cstring>

#include "common.cuh"

namespace ccdm {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace ccdm

extern "C" int ccdm_version(void) { return 100; }
extern "C" const char* ccdm_last_error(void) { return ccdm::g_err; }
extern "C" int64_t ccdm_launch_count(void) { return ccdm::g_launches.load(); }
extern "C" int ccdm_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(ccdm_tapgemm_args);
    case 1: return (int)sizeof(ccdm_view);
    case 2: return (int)sizeof(ccdm_step_args);
    case 3: return (int)sizeof(ccdm_qsample_args);
    case 4: return (int)sizeof(ccdm_loss_args);
    case 5: return (int)sizeof(ccdm_wgrad_args);
    default: return -1;
  }
}
