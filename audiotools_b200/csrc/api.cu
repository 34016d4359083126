// api.cu -- version + thread-local error message of libb2a.so (include/b2a.h).
#include "b2a_common.h"

namespace b2a {
char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace b2a

extern "C" int b2a_version(void) { return B2A_VERSION; }
extern "C" const char* b2a_last_error(void) { return b2a::err_buf(); }
