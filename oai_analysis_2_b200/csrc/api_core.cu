// Library-wide state of the C ABI: per-thread error message, launch counter, version.
#include "../../include/oai_b200.h"
#include "api_common.h"

namespace oai {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
std::atomic<long long> g_launches{0};
}  // namespace oai

extern "C" const char* oai_last_error(void) { return oai::last_error_buf(); }
extern "C" int oai_version(void) { return 100; }
extern "C" long long oai_launch_count(void) { return oai::g_launches.load(); }
