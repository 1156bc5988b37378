// common.cuh — shared host/device plumbing for libsquidpy_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/squidpy_b200.h"

// ---------------------------------------------------------------------------------------------
// error handling: every C-ABI entry point returns 0 on success or a negative sqb_status; the message is
// kept in a thread-local buffer readable through sqb_last_error().
// ---------------------------------------------------------------------------------------------
void sqb_set_error(const char* fmt, ...);

#define SQB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            sqb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,  \
                          cudaGetErrorString(_e));                                                  \
            return SQB_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define SQB_CHECK(cond, code, ...)   \
    do {                             \
        if (!(cond)) {               \
            sqb_set_error(__VA_ARGS__); \
            return (code);           \
        }                            \
    } while (0)

#define SQB_TRY(expr)           \
    do {                        \
        int _rc = (expr);       \
        if (_rc != SQB_OK) return _rc; \
    } while (0)

// kernel classes for the per-class launch accounting / event timing (sqb_ctx_profile_get)
enum {
    SQB_K_NHOOD_FILL = 0,
    SQB_K_NHOOD_SHUFFLE = 1,
    SQB_K_NHOOD_TRANSPOSE = 2,
    SQB_K_NHOOD_COUNT = 3,
    SQB_K_AUTOCORR_PREP = 4,
    SQB_K_AUTOCORR_MAIN = 5,
    SQB_K_AUTOCORR_FINAL = 6,
    SQB_K_PAIRS = 7,
    SQB_K_MISC = 8,
    SQB_K_NCLASS = 16
};

struct sqb_ctx;

// RAII-less device buffer helper (explicit free).  Unbound buffers use synchronous cudaMalloc / cudaFree; a buffer bound
// to a stream uses the stream-ordered allocator (cudaMallocAsync / cudaFreeAsync on the device's default memory pool,
// whose release threshold sqb_ctx_create raises so freed blocks are reused): creating and destroying a handle per API
// call then costs microseconds instead of ~1 ms of device synchronisation per buffer.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t stream = nullptr;
    bool async = false;
    void bind(cudaStream_t s) {
        stream = s;
        async = true;
    }
    int alloc(size_t count) {
        if (count <= n && p) return SQB_OK;
        release();
        if (count == 0) count = 1;
        cudaError_t e = async ? cudaMallocAsync((void**)&p, count * sizeof(T), stream) : cudaMalloc((void**)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            n = 0;
            sqb_set_error("device allocation of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
            return SQB_ERR_OOM;
        }
        n = count;
        return SQB_OK;
    }
    void release() {
        if (p) {
            if (async)
                cudaFreeAsync(p, stream);
            else
                cudaFree(p);
        }
        p = nullptr;
        n = 0;
    }
};

struct sqb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    size_t smem_optin = 0;   // max opt-in dynamic shared memory per block
    int64_t launches = 0;
    bool profile = false;
    double k_ms[SQB_K_NCLASS] = {0};
    int64_t k_n[SQB_K_NCLASS] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // Large scratch buffers that live as long as the context (GB-sized cudaMalloc/cudaFree per API call would
    // otherwise dominate the host-facing latency).  Work on one ctx is stream ordered, so handles can share them.
    DevBuf<uint8_t> scratch[3];
    // pinned staging ring for large host-to-device copies of pageable caller memory (sqb_h2d)
    static constexpr int kStage = 16;
    void* stage[kStage] = {};
    cudaEvent_t stage_ev[kStage] = {};
    bool stage_used[kStage] = {};
};

// Host-to-device copy of caller-owned (pageable) memory on the ctx stream.  Large copies go through a ring of pinned
// staging buffers filled by several host threads while the previous chunk is on the bus: a plain cudaMemcpyAsync from
// pageable memory ran at ~3-7 GB/s on the B200 hosts (3.2 GB of expression matrix = 1 s of a 1.08 s Moran call).
// The source has been read completely when the call returns; the copy itself completes in stream order.
int sqb_h2d(sqb_ctx* c, void* dst, const void* src, size_t bytes);
int sqb_h2d_gather(sqb_ctx* c, void* dst, const void* src, size_t elem, const int64_t* start, const int64_t* cnt,
                   const int64_t* out_ptr, int64_t rows);

// Launch accounting.  In profile mode each launch is bracketed by events on the ctx stream and the elapsed
// time accumulated per kernel class (the launch is synchronised; never used inside a timed bench region
// except for the dedicated roofline pass).
struct SqbLaunchScope {
    sqb_ctx* c;
    int kclass;
    SqbLaunchScope(sqb_ctx* ctx, int k) : c(ctx), kclass(k) {
        if (c->profile) cudaEventRecord(c->ev0, c->stream);
    }
    ~SqbLaunchScope() {
        c->launches += 1;
        c->k_n[kclass] += 1;
        if (c->profile) {
            cudaEventRecord(c->ev1, c->stream);
            cudaEventSynchronize(c->ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->ev0, c->ev1);
            c->k_ms[kclass] += ms;
        }
    }
};

#define SQB_POST_LAUNCH()                                                                     \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            sqb_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,            \
                          cudaGetErrorString(_e));                                            \
            return SQB_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
