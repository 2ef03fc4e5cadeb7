// Library-level state of libmbseg: thread-local error string and launch accounting.
#include <cstdarg>
#include <cstdio>

#include "../../include/mbseg.h"
#include "common.cuh"

namespace {
thread_local char g_err[1024] = "";
thread_local long long g_launches = 0;
}  // namespace

namespace mbs {
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }
}  // namespace mbs

extern "C" const char *mbs_last_error(void) { return g_err; }
extern "C" int mbs_version(void) { return 1; }
extern "C" int64_t mbs_launch_count(int reset) {
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
