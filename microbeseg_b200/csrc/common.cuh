// Shared helpers for libmbseg (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

namespace mbs {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define MBS_CHECK_CUDA(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            mbs::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                           cudaGetErrorString(_e));                                           \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

#define MBS_CHECK_LAUNCH()                                                                    \
    do {                                                                                      \
        mbs::count_launch();                                                                  \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            mbs::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,             \
                           cudaGetErrorString(_e));                                           \
            return 3;                                                                         \
        }                                                                                     \
    } while (0)

#define MBS_REQUIRE(cond, ...)                                                                \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            mbs::set_error(__VA_ARGS__);                                                      \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Per-device caches: function attributes (opt-in shared memory), SM counts and occupancy are properties of ONE
// device, so anything cached about them is indexed by the current device ordinal (a process may use cuda:0 and
// then cuda:1).  Racing threads write the same value, so plain ints are enough.
constexpr int kMaxDevices = 64;
static inline int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
    return d;
}

}  // namespace mbs
