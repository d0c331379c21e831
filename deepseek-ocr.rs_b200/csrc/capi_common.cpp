// Error plumbing shared by every extern "C" entry point.
#include "dsocr.h"
#include "util.h"

namespace dsocr {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
long long& launch_counter() {
  static long long n = 0;
  return n;
}
}  // namespace dsocr

extern "C" const char* dsocr_last_error(void) { return dsocr::g_last_error.c_str(); }
extern "C" const char* dsocr_version(void) { return "dsocr-b200 0.1 (sm_100a)"; }
