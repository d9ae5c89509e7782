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

// cos(x) for any float32 x with |x| < ~3e9, absolute error <= ~2e-7.
// The time-encoder argument fma(dt, w, b) reaches 1e6..1e8 rad, where __cosf is
// useless and cosf() takes the slow Payne-Hanek path.  cos(x) = (-1)^k cos(x - k*pi):
//   |x| < 4e6 : float32 only.  k = rint(x / pi) by the 1.5*2^23 magic-number trick (one FMA, no
//               conversion instructions), r = x - k*pi by a two-term Cody-Waite with FMAs.
//               fl(1/pi) is only good to 2^-25, so k can be off by one next to a half-way point
//               and |r| can reach ~1.75; the even degree-8 polynomial below is good to 2.4e-7 there.
//   otherwise : the reduction is done in float64.
// cos(r) on |r| <= 1.75 as a degree-8 even polynomial (least-squares fit on Chebyshev nodes in r^2; max abs
// error 2.4e-7 in fp32 Horner, the same as the degree-12 Taylor form it replaced, with two FMAs less)
#define FLID_COS_C0 0.9999998807907104f
#define FLID_COS_C1 -0.4999977648258209f
#define FLID_COS_C2 0.04166082665324211f
#define FLID_COS_C3 -0.0013835163554176688f
#define FLID_COS_C4 2.277065323141869e-05f
__device__ __forceinline__ float cos_poly_signed(float rf, int n) {
    const float r2 = rf * rf;
    float p = fmaf(r2, FLID_COS_C4, FLID_COS_C3);
    p = fmaf(p, r2, FLID_COS_C2);
    p = fmaf(p, r2, FLID_COS_C1);
    p = fmaf(p, r2, FLID_COS_C0);
    return __int_as_float(__float_as_int(p) ^ (n << 31));  // (-1)^k
}
constexpr float COS_FAST_LIMIT = 4.0e6f;
__device__ __forceinline__ float cos_fast(float x) {  // |x| < COS_FAST_LIMIT
    const float t = fmaf(x, 0.31830987334251404f, 12582912.0f);
    const float kf = t - 12582912.0f;
    float rf = fmaf(kf, -3.1415927410125732f, x);
    rf = fmaf(kf, 8.742277657347586e-08f, rf);
    return cos_poly_signed(rf, __float_as_int(t));
}
__device__ __forceinline__ float cos_slow(float x) {
    const double xd = (double)x;
    const double q = rint(xd * 0.31830988618379067154);
    double r = fma(q, -3.14159265358979311600, xd);
    r = fma(q, -1.2246467991473532072e-16, r);
    return cos_poly_signed((float)r, (int)q);
}
__device__ __forceinline__ float cos_accurate(float x) { return fabsf(x) < COS_FAST_LIMIT ? cos_fast(x) : cos_slow(x); }

// sin(x) with the same range reduction (the time-encoder gradient of the training path, attn_train.cu):
// sin(x) = (-1)^k sin(x - k*pi), sin(r) = r * P(r^2) on |r| <= 1.75, degree-9 least-squares fit, abs. error 1.8e-7
#define FLID_SIN_C1 -0.16666646301746368f
#define FLID_SIN_C2 0.008332798257470131f
#define FLID_SIN_C3 -0.0001979204243980348f
#define FLID_SIN_C4 2.570016476965975e-06f
__device__ __forceinline__ float sin_poly_signed(float rf, int n) {
    const float r2 = rf * rf;
    float p = fmaf(r2, FLID_SIN_C4, FLID_SIN_C3);
    p = fmaf(p, r2, FLID_SIN_C2);
    p = fmaf(p, r2, FLID_SIN_C1);
    p = fmaf(p * r2, rf, rf);
    return __int_as_float(__float_as_int(p) ^ (n << 31));
}
__device__ __forceinline__ float sin_accurate(float x) {
    if (fabsf(x) < COS_FAST_LIMIT) {
        const float t = fmaf(x, 0.31830987334251404f, 12582912.0f);
        const float kf = t - 12582912.0f;
        float rf = fmaf(kf, -3.1415927410125732f, x);
        rf = fmaf(kf, 8.742277657347586e-08f, rf);
        return sin_poly_signed(rf, __float_as_int(t));
    }
    const double xd = (double)x;
    const double q = rint(xd * 0.31830988618379067154);
    double r = fma(q, -3.14159265358979311600, xd);
    r = fma(q, -1.2246467991473532072e-16, r);
    return sin_poly_signed((float)r, (int)q);
}

// TimeEncoder (models/modules.py:35-38): cos of the single-rounded fma(dt, w, b).
__device__ __forceinline__ float time_channel(float dt, float w, float b) { return cos_accurate(fmaf(dt, w, b)); }
#endif  // __CUDACC__

}  // namespace flid
