// Error plumbing shared by every C-ABI entry point: errors are values, never exceptions
// (the reference returns ResponsePart::Error instead of panicking, dispatcher.rs:323-325).
#include "../../include/ollamamq_b200.h"
#include <cstdarg>
#include <cstdio>

namespace mq {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mq

extern "C" {
const char* mq_last_error(void) { return mq::g_err; }
const char* mq_version(void) { return "ollamamq_b200 0.1.0 (sm_100a)"; }
}
