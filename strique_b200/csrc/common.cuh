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
    int num_sms = 0;
    std::string error;
    int64_t launches = 0;
    // statistics of the last alignment call
    int64_t last_align_cells = 0;
    float last_scan_ms = 0.f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::map<std::string, DevBuf> bufs;   // persistent device scratch
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

#define TRY(expr)                                                                                 \
    do {                                                                                          \
        int _rc = (expr);                                                                         \
        if (_rc != STRIQUE_OK) return _rc;                                                        \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
