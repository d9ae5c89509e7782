// Shared host/device helpers for the flid_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "flid_b200.h"

namespace flid {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern int64_t g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

#define FLID_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            flid::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FLID_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define FLID_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        flid::count_launch();                                                             \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            flid::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FLID_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define FLID_REQUIRE(cond, ...)                                                           \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            flid::set_error(__VA_ARGS__);                                                 \
            return FLID_ERR_INVALID;                                                      \
        }                                                                                 \
    } while (0)

#define FLID_TRY(expr)                                                                    \
    do {                                                                                  \
        int _s = (expr);                                                                  \
        if (_s != FLID_OK) return _s;                                                     \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grow-only device buffer (cudaMalloc is synchronous; growth only happens on the
// first calls of a given shape).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return FLID_OK;
        if (p) {
            FLID_CUDA(cudaDeviceSynchronize());
            FLID_CUDA(cudaFree(p));
            p = nullptr;
            cap = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        FLID_CUDA(cudaMalloc(&p, want));
        cap = want;
        return FLID_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// ---------------------------------------------------------------- device math
#ifdef __CUDACC__
constexpr int WARP = 32;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// cos(x) for any float32 x with |x| < ~3e9, accurate to ~1 ulp.
// The time-encoder argument fma(dt, w, b) reaches 1e6..1e8 rad, where
// __cosf is useless and cosf() takes the slow Payne-Hanek path.  The argument is
// reduced in float64 with a two-term pi/2 (error < 1e-16 * |n|), then the Cephes
// single-precision minimax polynomials on [-pi/4, pi/4] are evaluated in float32.
__device__ __forceinline__ float cos_accurate(float x) {
    const double xd = (double)x;
    const double q = rint(xd * 0.63661977236758134308);
    double r = fma(q, -1.57079632679489655800, xd);
    r = fma(q, -6.12323399573676603587e-17, r);
    const int n = (int)q;
    const float rf = (float)r;
    const float r2 = rf * rf;
    float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    sp = fmaf(sp, r2, -1.6666654611e-1f);
    const float s = fmaf(rf * r2, sp, rf);
    float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    cp = fmaf(cp, r2, 4.166664568298827e-2f);
    const float c = fmaf(r2 * r2, cp, fmaf(r2, -0.5f, 1.0f));
    float v = (n & 1) ? s : c;
    // n mod 4: 0 -> c, 1 -> -s, 2 -> -c, 3 -> s
    return (((n + 1) & 2) != 0) ? -v : v;
}

// TimeEncoder (models/modules.py:35-38): cos of the single-rounded fma(dt, w, b).
__device__ __forceinline__ float time_channel(float dt, float w, float b) { return cos_accurate(fmaf(dt, w, b)); }
#endif  // __CUDACC__

}  // namespace flid
