// Library-wide state: last-error string, launch counter, version.
#include "common.cuh"

namespace osfm {
static thread_local char g_last_error[1024] = "";
std::atomic<int64_t> g_kernel_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_last_error; }
}  // namespace osfm

extern "C" {
const char* osfm_last_error(void) { return osfm::get_last_error(); }
int osfm_version(void) { return 100; }
int64_t osfm_kernel_launch_count(void) { return osfm::g_kernel_launches.load(); }
}
