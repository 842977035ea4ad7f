// Error reporting, version and launch accounting of the C ABI (include/esf.h).
#include <atomic>
#include <string.h>

#include "esf_host.h"

namespace esf {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace esf

extern "C" const char* esf_last_error(void) { return esf::g_err; }
extern "C" int esf_version(void) { return 100; }
extern "C" int64_t esf_launch_count(void) { return esf::g_launches.load(std::memory_order_relaxed); }
