// Version + thread-local error string of the C ABI (include/edgeconv_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace ecb200 {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace ecb200

extern "C" int ecb200_version(void) { return ECB200_VERSION; }
extern "C" const char* ecb200_last_error(void) { return ecb200::g_err; }
