// Library-wide state: thread-local error string, ABI version, launch counter.
#include <stdarg.h>

#include "common.cuh"
#include "gemm.cuh"
#include "gemm_tc.cuh"

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

extern "C" int flid_debug_gemm(int backend, const float* a0, int64_t lda0, const int32_t* idx0, int w0, const float* a1,
                               int64_t lda1, int w1, const float* w, int64_t ldw, const float* bias, float* c,
                               int64_t ldc, int64_t m, int n, int relu, flid_stream stream) {
    using namespace flid;
    cudaStream_t st = (cudaStream_t)stream;
    if (backend == 0) {
        GemmArgs g0{a0, lda0, idx0, w, ldw, c, ldc, w1 > 0 ? nullptr : bias, m, n, w0, 0, w1 > 0 ? 0 : relu};
        FLID_TRY(launch_gemm(g0, st));
        if (w1 > 0) {
            GemmArgs g1{a1, lda1, nullptr, w + w0, ldw, c, ldc, bias, m, n, w1, 1, relu};
            FLID_TRY(launch_gemm(g1, st));
        }
        return FLID_OK;
    }
    TcWeight tw;
    FLID_TRY(tc_prepare_weight(w, ldw, n, w0 + w1, &tw, st));
    TcGemmArgs g;
    g.A0 = a0, g.lda0 = lda0, g.idx0 = idx0, g.w0 = w0, g.A1 = a1, g.lda1 = lda1, g.w1 = w1;
    g.C = c, g.ldc = ldc, g.bias = bias, g.M = m, g.relu = relu;
    int status = tc_gemm(g, tw, st);
    cudaStreamSynchronize(st);
    tc_free_weight(&tw);
    return status;
}

// Timing hook for tools/gemm_probe.py: tile the weight once, launch the tcgen05 GEMM `reps`
// times on `stream` and return the mean device time per launch (CUDA events on that stream).
extern "C" int flid_debug_gemm_time(const float* a0, int64_t lda0, const int32_t* idx0, int w0, const float* a1,
                                    int64_t lda1, int w1, const float* w, int64_t ldw, const float* bias, float* c,
                                    int64_t ldc, int64_t m, int n, int relu, int reps, float* ms_per_launch,
                                    flid_stream stream) {
    using namespace flid;
    FLID_REQUIRE(reps > 0 && ms_per_launch, "flid_debug_gemm_time: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    TcWeight tw;
    FLID_TRY(tc_prepare_weight(w, ldw, n, w0 + w1, &tw, st));
    TcGemmArgs g;
    g.A0 = a0, g.lda0 = lda0, g.idx0 = idx0, g.w0 = w0, g.A1 = a1, g.lda1 = lda1, g.w1 = w1;
    g.C = c, g.ldc = ldc, g.bias = bias, g.M = m, g.relu = relu;
    int status = tc_gemm(g, tw, st);  // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps && status == FLID_OK; ++r) status = tc_gemm(g, tw, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(ms_per_launch, e0, e1);
    *ms_per_launch /= (float)reps;
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    tc_free_weight(&tw);
    return status;
}

