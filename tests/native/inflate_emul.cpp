// TEST INFRASTRUCTURE: the inflate kernel's per-thread decoder (strique_b200/csrc/inflate_core.h) compiled for the
// host, so that tests can hold it against zlib without a GPU.  Nothing here is used by the product.
#include <string.h>

#include <vector>

#include "../../strique_b200/csrc/inflate_core.h"

// misalign: the stream is placed at this byte offset (0..15) of a 16-byte aligned buffer, as chunks are inside a batch
extern "C" int strique_test_inflate(const uint8_t *src, long long n, int misalign, uint8_t *out, unsigned keep, uint8_t *spill,
                                    unsigned full, unsigned *produced) {
    using namespace strique::inf;
    std::vector<uint32_t> words((size_t)(n + misalign) / 4 + 16, 0xA5A5A5A5u);    // garbage around the stream
    uint8_t *base = reinterpret_cast<uint8_t *>(words.data());
    base += (16 - (reinterpret_cast<uintptr_t>(base) & 15)) & 15;
    memcpy(base + misalign, src, (size_t)n);
    std::vector<uint16_t> lit(1 << LIT_BITS), dist(1 << DIST_BITS);
    Scratch s;
    Lane L;
    lane_begin(L, base + misalign, n, out, spill, keep, full);
    while (L.need != DONE) {
        if (L.need == RUN) lane_step(L, lit.data(), dist.data(), 1, s);
        else lane_service(L, lit.data(), dist.data(), 1, s);
    }
    *produced = L.o;
    return L.status;
}
