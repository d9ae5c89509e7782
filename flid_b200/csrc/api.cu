// Library-wide state: thread-local error string, ABI version, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace flid {
static thread_local char g_error[512] = "";
int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace flid

extern "C" {
const char* flid_last_error(void) { return flid::g_error; }
int flid_abi_version(void) { return 1; }
int64_t flid_launch_count(void) { return flid::g_launches; }
}
