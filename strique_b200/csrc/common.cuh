// Shared host/device plumbing of libstrique_b200: context, device arena, error handling.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/strique_b200.h"

// SeqAn's "minus infinity" for float scores is FLT_MIN/2, a POSITIVE denormal
// (seqan/align/dp_cell.h:137-145).  Parity needs this exact constant and non-flushed denormals:
// never build this library with --use_fast_math / -ftz=true.
#define STRIQUE_SEQAN_INF 5.87747175411143754e-39f

struct strique_ctx;
namespace strique {
struct HmmModel;
struct Target;
}

// A growable device buffer owned by the context (looked up by name, reused across calls).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(strique_ctx *ctx, size_t bytes);
    template <typename T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

struct strique_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // host -> device upload of raw reads, overlapped with conditioning
    cudaEvent_t copy_ev[8] = {nullptr};
    int num_sms = 0;
    std::string error;
    int64_t launches = 0;
    // statistics of the last alignment call
    int64_t last_align_cells = 0;
    float last_scan_ms = 0.f;
    int64_t last_viterbi_edges = 0;
    // profile models: sequences decoded by the fixed-point kernel / declined by it and decoded in float64 (last call)
    int64_t last_viterbi_fixed = 0, last_viterbi_declined = 0;
    bool viterbi_exact = false;           // float64 kernel only (strique_set_viterbi_exact)
    std::vector<strique_viterbi_result> vit_res_scratch;
    float stage_ms[8] = {0};              // last detect call: see STRIQUE_STAGE_* in the public header
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t stage_ev[16] = {nullptr};
    std::vector<void *> owned;            // device allocations living as long as the context (models)
    std::vector<strique::HmmModel *> models;
    std::vector<strique::Target *> targets;
    std::map<std::string, DevBuf> bufs;   // persistent device scratch
    int64_t last_mod_bytes = 0;           // bytes of methylation patterns the last detect call produced / needs
    // cudaMemGetInfo is a driver round trip that was measured at 16 ... 65 ms per call on some (virtualised) GPU
    // boxes: the free-memory figure behind the chunking budgets is cached and refreshed only after an allocation
    size_t free_cached = 0, total_cached = 0;
    DevBuf &buf(const char *name) { return bufs[name]; }
    ~strique_ctx();
};

extern std::string g_strique_create_error;

#define CUDA_TRY(ctx, expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            (ctx)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                   \
            return STRIQUE_ECUDA;                                                                 \
        }                                                                                         \
    } while (0)

#define FAIL(ctx, code, msg)                                                                      \
    do {                                                                                          \
        (ctx)->error = (msg);                                                                     \
        return (code);                                                                            \
    } while (0)

inline int DevBuf::ensure(strique_ctx *ctx, size_t bytes) {
    if (bytes <= cap && p) return STRIQUE_OK;
    ctx->free_cached = 0;   // the cached free-memory figure is stale after this
    if (bytes == 0) bytes = 16;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        p = nullptr;
        ctx->error = "cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e);
        return STRIQUE_ENOMEM;
    }
    cap = want;
    return STRIQUE_OK;
}

// free / total device memory for the chunking budgets (cached, see strique_ctx::free_cached)
static inline cudaError_t ctx_mem_info(strique_ctx *ctx, size_t *free_b, size_t *total_b) {
    if (ctx->free_cached == 0) {
        const cudaError_t e = cudaMemGetInfo(&ctx->free_cached, &ctx->total_cached);
        if (e != cudaSuccess) { ctx->free_cached = 0; return e; }
        if (ctx->free_cached == 0) ctx->free_cached = 1;
    }
    *free_b = ctx->free_cached;
    *total_b = ctx->total_cached;
    return cudaSuccess;
}

#define TRY(expr)                                                                                 \
    do {                                                                                          \
        int _rc = (expr);                                                                         \
        if (_rc != STRIQUE_OK) return _rc;                                                        \
    } while (0)

// STRIQUE_HOST_TIMING=1: report host-side sections that take longer than 3 ms (GPU-idle gaps between stages)
#include <chrono>
#include <cstdio>
#include <cstdlib>
struct HostTimer {
    const char *what;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit HostTimer(const char *w) : what(w), t0(std::chrono::steady_clock::now()), on(getenv("STRIQUE_HOST_TIMING") != nullptr) {}
    ~HostTimer() {
        if (!on) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 3.0) fprintf(stderr, "[host] %-28s %.1f ms\n", what, ms);
    }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// per-stage device timing: events 2*stage (begin) and 2*stage+1 (end) on the context's stream
static inline void stage_mark(strique_ctx *ctx, int idx) {
    if (!ctx->stage_ev[idx]) cudaEventCreate(&ctx->stage_ev[idx]);
    cudaEventRecord(ctx->stage_ev[idx], ctx->stream);
}
static inline void stage_collect(strique_ctx *ctx, int stage) {   // call after a stream synchronize
    float ms = 0.f;
    if (ctx->stage_ev[2 * stage] && ctx->stage_ev[2 * stage + 1] &&
        cudaEventElapsedTime(&ms, ctx->stage_ev[2 * stage], ctx->stage_ev[2 * stage + 1]) == cudaSuccess)
        ctx->stage_ms[stage] += ms;
    else
        cudaGetLastError();
}
static inline void stage_reset(strique_ctx *ctx) {
    for (int i = 0; i < 8; ++i) ctx->stage_ms[i] = 0.f;
}
