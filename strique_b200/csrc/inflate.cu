// fast5 Signal chunks -> raw samples on the device.  The reference reads every read's Signal dataset
// through h5py (STRique_lib/fast5Index.py:76-84), i.e. zlib's inflate on one CPU core per worker: ≈ 1 ms per read,
// which caps `STRique.py count` at ≈ 1 k reads/s per core while the kernels behind it consume 44 k.  Here the host
// only locates the chunks; the compressed bytes cross PCIe (fewer than the samples would) and `inflate_kernel`
// decodes them: ONE LANE PER CHUNK (a chunk is an independent zlib stream of ≤ 16 KB of samples; a batch of 8192
// reads holds ≈ 44 000 of them), Huffman tables of a warp's 32 lanes interleaved in shared memory, output straight
// into the batch's raw-sample buffer.  8192 reads (654 MB of samples from 434 MB stored): 18.9 ms = 34.6 GB/s of
// samples, against ≈ 0.085 GB/s per CPU core for zlib.  inflate_core.h is the decoder (also compiled for the host and held against
// zlib by tests/test_inflate_emul.py).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "inflate_core.h"

namespace strique {
namespace {

#ifndef STRIQUE_INF_THREADS
#define STRIQUE_INF_THREADS 384
#endif
constexpr int INF_THREADS = STRIQUE_INF_THREADS;   // 12 warps x 32 lanes x 576 B of tables = 216 KB of shared memory: one CTA per SM
constexpr int INF_TAB_ENTRIES = (1 << inf::LIT_BITS) + (1 << inf::DIST_BITS);

struct ChunkDev {
    int64_t src_off, dst_off, spill_off;
    int32_t src_len, keep, full, pad;
};

// The warp alternates between two phases.  Service (divergent, rare): lanes whose stream needs a block header, its
// Huffman tables, the trailer, or a new chunk from the queue get it while the others wait.  Run (convergent): every
// lane with a block in progress executes inf::lane_step, one symbol or eight match bytes per iteration, until some
// lane asks for service again.
__global__ void __launch_bounds__(INF_THREADS) inflate_kernel(const uint8_t *__restrict__ comp, const ChunkDev *__restrict__ chunks,
                                                              int n_chunks, uint8_t *out, uint8_t *spill, int32_t *status,
                                                              int *queue) {
    extern __shared__ uint16_t tables[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t *lit = tables + (size_t)warp * 32 * INF_TAB_ENTRIES + lane;
    uint16_t *dist = lit + 32 * (1 << inf::LIT_BITS);
    constexpr unsigned FULL = 0xffffffffu;
    inf::Scratch scratch;
    inf::Lane L;
    L.need = inf::DONE;
    L.status = inf::INF_OK;
    L.o = 0; L.keep = 0;
    int cur = -1;
    bool idle = false;
    for (;;) {
        while (!idle && L.need != inf::RUN) {
            if (L.need == inf::DONE) {
                if (cur >= 0) status[cur] = L.status == inf::INF_OK && L.o < L.keep ? (int)inf::INF_SHORT_OUTPUT : L.status;
                cur = atomicAdd(queue, 1);       // chunks differ in size (the last one of a read, compressibility)
                if (cur >= n_chunks) {
                    idle = true;
                    break;
                }
                const ChunkDev ch = chunks[cur];
                inf::lane_begin(L, comp + ch.src_off, (int64_t)ch.src_len, out + ch.dst_off, spill + ch.spill_off,
                                (uint32_t)ch.keep, (uint32_t)ch.full);
            } else {
                inf::lane_service(L, lit, dist, 32, scratch);
            }
        }
        __syncwarp();
        if (__all_sync(FULL, idle)) break;
        for (;;) {
            const bool running = !idle && L.need == inf::RUN;
            if (running) inf::lane_step(L, lit, dist, 32, scratch);
            __syncwarp();
            const bool wants = !idle && L.need != inf::RUN;
            if (__any_sync(FULL, wants) || !__any_sync(FULL, running)) break;
        }
    }
}

}  // namespace
}  // namespace strique

using namespace strique;

extern "C" int strique_inflate_batch(strique_ctx *ctx, const void *comp, int64_t comp_bytes, int comp_memspace,
                                     const strique_inflate_chunk *chunks, int n_chunks, int64_t out_bytes,
                                     int32_t *status_host, void **out_dev) {
    if (!ctx) return STRIQUE_EINVAL;
    if (n_chunks < 0 || comp_bytes < 0 || out_bytes < 0 || (n_chunks > 0 && (!comp || !chunks || !status_host)) || !out_dev)
        FAIL(ctx, STRIQUE_EINVAL, "strique_inflate_batch: bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevBuf &d_out = ctx->buf("inf.out"), &d_comp = ctx->buf("inf.comp"), &d_chunks = ctx->buf("inf.chunks"),
           &d_spill = ctx->buf("inf.spill"), &d_status = ctx->buf("inf.status"), &d_queue = ctx->buf("inf.queue");
    TRY(d_out.ensure(ctx, (size_t)out_bytes + 16));
    *out_dev = d_out.p;
    if (n_chunks == 0) return STRIQUE_OK;
    std::vector<ChunkDev> cd((size_t)n_chunks);
    int64_t spill = 0;
    for (int i = 0; i < n_chunks; ++i) {
        const strique_inflate_chunk &c = chunks[i];
        if (c.src_off < 0 || c.src_len < 0 || c.src_off + c.src_len > comp_bytes || c.keep < 0 || c.full < c.keep ||
            c.dst_off < 0 || c.dst_off + c.keep > out_bytes)
            FAIL(ctx, STRIQUE_EINVAL, "strique_inflate_batch: chunk " + std::to_string(i) + " lies outside its buffers");
        cd[i] = ChunkDev{c.src_off, c.dst_off, spill, c.src_len, c.keep, c.full, 0};
        spill += c.full - c.keep;
    }
    const uint8_t *comp_dev = static_cast<const uint8_t *>(comp);
    if (comp_memspace != STRIQUE_DEVICE) {
        TRY(d_comp.ensure(ctx, (size_t)comp_bytes + 16));       // the bit reader loads whole aligned words
        CUDA_TRY(ctx, cudaMemcpyAsync(d_comp.p, comp, (size_t)comp_bytes, cudaMemcpyHostToDevice, ctx->stream));
        comp_dev = d_comp.as<uint8_t>();
    }
    TRY(d_chunks.ensure(ctx, cd.size() * sizeof(ChunkDev)));
    TRY(d_spill.ensure(ctx, (size_t)spill + 16));
    TRY(d_status.ensure(ctx, (size_t)n_chunks * 4));
    TRY(d_queue.ensure(ctx, 16));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_chunks.p, cd.data(), cd.size() * sizeof(ChunkDev), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(d_queue.p, 0, 4, ctx->stream));
    const size_t smem = (size_t)(INF_THREADS / 32) * 32 * INF_TAB_ENTRIES * sizeof(uint16_t);
    CUDA_TRY(ctx, cudaFuncSetAttribute(inflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::max(1, std::min(ctx->num_sms, (n_chunks + INF_THREADS - 1) / INF_THREADS));
    inflate_kernel<<<grid, INF_THREADS, smem, ctx->stream>>>(comp_dev, d_chunks.as<ChunkDev>(), n_chunks, d_out.as<uint8_t>(),
                                                           d_spill.as<uint8_t>(), d_status.as<int32_t>(), d_queue.as<int>());
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(status_host, d_status.p, (size_t)n_chunks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));          // cd and the caller's buffers are read by the copies
    return STRIQUE_OK;
}
