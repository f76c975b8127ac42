// zlib / DEFLATE (RFC 1950 / 1951) decoder of ONE stream by ONE thread -- host and device.  HDF5's deflate filter
// (the `gzip` compression of fast5 Signal datasets the reference reads through h5py, STRique_lib/fast5Index.py:76-84)
// stores every dataset chunk as an independent zlib stream, so a batch of reads is tens of thousands of independent
// 16 KB streams: `inflate_kernel` (inflate.cu) gives each to one lane of a warp, and the same code compiled by g++
// (tests/native/inflate_emul.cpp) is checked against zlib itself on the CPU.
//
// Written for SIMT: the 32 lanes of a warp decode 32 different streams, so the decoder is a STATE MACHINE whose
// `step` does one small unit of work -- decode one literal/length symbol, decode one distance symbol, or copy up to
// COPY_STEP bytes of a pending match -- with no loops over data-dependent counts and no early exits.  Lanes in different
// states execute the same loop iteration (each predicated section once), instead of each lane's match copies and
// table look-ups being serialised against the others'.  The rare, long operations (block headers, building the
// Huffman tables, stored blocks, the trailer) happen in `service`, which a lane asks for through `need`.
//
// Decoding: a primary look-up table per Huffman code (2^LIT_BITS / 2^DIST_BITS uint16 entries = length << 12 | symbol,
// indexed by the low bits of the bit buffer; on the device the tables of the 32 lanes of a warp are interleaved in
// shared memory, entry i of lane l at [i * 32 + l]: at most 2 lanes per bank); codes longer than that are found by
// comparing the next 15 bits against the left-justified code range of every length (canonical codes are ordered),
// with those ranges in registers.  The tables are deliberately SMALL (8 / 5 bits: 576 B per lane, 12 warps per SM):
// the kernel is latency bound per warp, and resident warps bought more than table hits (measured on 8192 reads:
// 10 / 7 bits at 3 warps per SM 55.7 ms, 9 / 6 at 6 warps 31.3 ms, 8 / 5 at 12 warps 18.9 ms, 7 / 5 at 16 warps 28.3 ms).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define INF_HD __host__ __device__ __forceinline__
#else
#define INF_HD inline
#endif

#ifndef STRIQUE_INF_LIT_BITS
#define STRIQUE_INF_LIT_BITS 8
#endif
#ifndef STRIQUE_INF_DIST_BITS
#define STRIQUE_INF_DIST_BITS 5
#endif
#ifndef STRIQUE_INF_COPY_STEP
#define STRIQUE_INF_COPY_STEP 4
#endif

namespace strique {
namespace inf {

constexpr int LIT_BITS = STRIQUE_INF_LIT_BITS, DIST_BITS = STRIQUE_INF_DIST_BITS;
constexpr int MAX_LIT = 288, MAX_DIST = 32;
constexpr uint32_t ADLER_BASE = 65521u, ADLER_NMAX = 5552u;
constexpr int COPY_STEP = STRIQUE_INF_COPY_STEP;      // match bytes per step

enum Status {
    INF_OK = 0,
    INF_BAD_HEADER = 1,       // not a zlib stream with method 8, or a preset dictionary
    INF_BAD_BLOCK = 2,        // block type 3, stored length check, code-length header
    INF_BAD_CODE = 3,         // over-subscribed code or a bit pattern no code matches
    INF_OVERFLOW = 4,         // more output than the chunk holds
    INF_BAD_DISTANCE = 5,     // match reaching before the start of the output
    INF_INPUT_OVERRUN = 6,    // stream longer than its stored size
    INF_BAD_CHECKSUM = 7,     // Adler-32 mismatch
    INF_SHORT_OUTPUT = 8      // fewer bytes than the caller keeps
};

#define INF_LBASE {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258}
#define INF_LEXT {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0}
#define INF_DBASE {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577}
#define INF_DEXT {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13}
#define INF_CLORDER {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15}
#ifdef __CUDACC__
__constant__ uint16_t d_lbase[29] = INF_LBASE;
__constant__ uint8_t d_lext[29] = INF_LEXT;
__constant__ uint16_t d_dbase[30] = INF_DBASE;
__constant__ uint8_t d_dext[30] = INF_DEXT;
__constant__ uint8_t d_clorder[19] = INF_CLORDER;
#endif
static const uint16_t h_lbase[29] = INF_LBASE;
static const uint8_t h_lext[29] = INF_LEXT;
static const uint16_t h_dbase[30] = INF_DBASE;
static const uint8_t h_dext[30] = INF_DEXT;
static const uint8_t h_clorder[19] = INF_CLORDER;
#ifdef __CUDA_ARCH__
#define INF_TAB(name) d_##name
#else
#define INF_TAB(name) h_##name
#endif

// LSB-first bit reader over aligned 32-bit words (the stream itself may start at any byte).  Never reads past the
// stream's last word and masks the bytes behind its end, so what a damaged stream decodes to depends on the stream
// alone; past the end it reads zeros until the output bound stops it.
struct Reader {
    const uint32_t *w;
    int64_t pos, limit;      // word after `nxt` / first word not to be read
    uint64_t bb;
    uint32_t nxt, nmask;     // the word after the ones in bb, loaded one refill ahead so that its latency hides behind a
                             // step -- and its mask, applied only when the word is used (masking at once would wait for it)
    uint32_t tail_mask;      // of the last word
    int bc, skip;

    INF_HD uint32_t raw(int64_t i) const {
        if (i >= limit) return 0u;
#ifdef __CUDA_ARCH__
        return __ldg(w + i);
#else
        return w[i];
#endif
    }
    INF_HD uint32_t mask(int64_t i) const { return i >= limit ? 0u : (i == limit - 1 ? tail_mask : 0xffffffffu); }
    INF_HD void init(const uint8_t *src, int64_t n) {
        const uintptr_t a = (uintptr_t)src;
        skip = (int)(a & 3);
        w = (const uint32_t *)(a - skip);
        limit = (skip + n + 3) / 4;
        const int tail = (int)((skip + n) & 3);
        tail_mask = tail ? (1u << (8 * tail)) - 1u : 0xffffffffu;
        bb = (uint64_t)((raw(0) & mask(0)) >> (8 * skip));
        bc = 32 - 8 * skip;
        nxt = raw(1);
        nmask = mask(1);
        pos = 2;
    }
    INF_HD void refill() {            // afterwards at least 33 bits are in the buffer
        if (bc <= 32) {
            bb |= (uint64_t)(nxt & nmask) << bc;
            bc += 32;
            nxt = raw(pos);
            nmask = mask(pos);
            ++pos;
        }
    }
    INF_HD uint32_t peek(int k) const { return (uint32_t)bb & ((1u << k) - 1u); }
    INF_HD void drop(int k) { bb >>= k; bc -= k; }
    INF_HD uint32_t take(int k) { const uint32_t v = peek(k); drop(k); return v; }
    INF_HD int64_t consumed_bits() const { return (pos - 1) * 32 - skip * 8 - bc; }
};

// One Huffman code: counts per length and symbols in code order (canonical), plus, for the codes longer than the
// primary table, the left-justified 15-bit code ranges per length.
struct Code {
    uint16_t cnt[16];
    uint16_t end[16];        // (first code of length l + cnt[l]) << (15 - l): a 15-bit window v has length l iff
    uint16_t first[16];      //   end[l-1] <= v < end[l]; first[l] = first code << (15 - l)
    uint16_t offs[16];       // index in `sorted` of the first symbol of length l
    uint16_t sorted[MAX_LIT];
};

// canonical Huffman code of n symbols with code lengths lens[]; fills `c` and the primary table (entries stride
// apart).  false: over-subscribed.
INF_HD bool build_table(const uint8_t *lens, int n, Code &c, uint16_t *tab, int stride, int bits) {
    for (int l = 0; l < 16; ++l) c.cnt[l] = 0;
    for (int i = 0; i < n; ++i) c.cnt[lens[i]]++;
    c.cnt[0] = 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= c.cnt[l];
        if (left < 0) return false;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + c.cnt[l];
    for (int l = 1; l < 16; ++l) c.offs[l] = offs[l];
    for (int i = 0; i < n; ++i)
        if (lens[i]) c.sorted[offs[lens[i]]++] = (uint16_t)i;
    for (int j = 0; j < (1 << bits); ++j) tab[j * stride] = 0;          // 0: longer than the table, or no code
    uint32_t code = 0;
    int idx = 0;
    c.end[0] = 0; c.first[0] = 0; c.offs[0] = 0;
    for (int l = 1; l < 16; ++l) {
        c.first[l] = (uint16_t)(code << (15 - l));
        if (l <= bits) {
            for (int k = 0; k < c.cnt[l]; ++k) {
                const uint32_t sym = c.sorted[idx++];
                uint32_t rev = 0;                                        // the code is sent most significant bit first
                for (int t = 0; t < l; ++t) rev |= ((code >> t) & 1u) << (l - 1 - t);
                for (uint32_t j = rev; j < (1u << bits); j += 1u << l) tab[j * stride] = (uint16_t)((l << 12) | sym);
                ++code;
            }
        } else {
            code += c.cnt[l];
        }
        // (code <= 2^l because the code is not over-subscribed; 2^15 does not fit: a complete code ends at 0x8000)
        c.end[l] = (uint16_t)((code << (15 - l)) > 0x7fffu ? 0x8000u : (code << (15 - l)));
        code <<= 1;
    }
    return true;
}

INF_HD uint32_t reverse15(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __brev(x) >> 17;
#else
    uint32_t r = 0;
    for (int t = 0; t < 15; ++t) r |= ((x >> t) & 1u) << (14 - t);
    return r;
#endif
}

INF_HD int decode_symbol(Reader &r, const uint16_t *tab, int stride, int bits, const Code &c) {
    const uint32_t e = tab[r.peek(bits) * stride];
    if (e) {
        r.drop((int)(e >> 12));
        return (int)(e & 0xfffu);
    }
    // longer than the table: the code's length is where the next 15 bits (first bit most significant) fall between
    // the left-justified code ranges
    const uint32_t v = reverse15(r.peek(15));
    int len = bits + 1;
#pragma unroll
    for (int l = DIST_BITS + 1; l < 15; ++l)
        len += (l > bits && v >= c.end[l]) ? 1 : 0;
    if (v >= c.end[15]) return -1;
    const int sym = c.sorted[c.offs[len] + ((v - c.first[len]) >> (15 - len))];
    r.drop(len);
    return sym;
}

struct Scratch {                         // per thread, local memory
    uint8_t lens[MAX_LIT + MAX_DIST];
    Code lit, dist;                      // (dist uses the first MAX_DIST entries of `sorted`)
};

enum Need { RUN = 0, NEW_BLOCK = 1, FINISH = 2, DONE = 3 };
constexpr int N_LONG = 15 - DIST_BITS;   // lengths above the smaller primary table
enum State { ST_LITLEN = 0, ST_DIST = 1, ST_COPY = 2 };

static_assert(DIST_BITS <= LIT_BITS, "decode_symbol scans the lengths above DIST_BITS");

// one stream being decoded by one lane
struct Lane {
    Reader r;
    uint8_t *out, *hi;           // byte i of the stream's output lives at (i < keep ? out : hi)[i]  (hi = spill - keep)
    uint32_t keep, full, o;
    uint32_t a, b, pending;      // Adler-32 of the output so far (reduced before ADLER_NMAX bytes have gone in)
    uint32_t mlen, mdist;        // pending match
    int64_t nbits;               // stored size of the stream
    int state, need, status;
    uint32_t last;               // the current block is the last one
    // the codes longer than the primary tables, per length DIST_BITS + 1 + k (Code::end / first / offs): in REGISTERS,
    // because the per-thread scratch does not stay in what shared memory leaves of L1
    uint32_t lit_end[N_LONG], lit_first[N_LONG], lit_offs[N_LONG];
    uint32_t dist_end[N_LONG], dist_first[N_LONG], dist_offs[N_LONG];

    INF_HD uint8_t *at(uint32_t i) const { return (i < keep ? out : hi) + i; }
    INF_HD void put(uint32_t c) {
        *at(o) = (uint8_t)c;
        ++o;
        a += c;
        b += a;
    }
    INF_HD void settle(uint32_t n) {            // after at most COPY_STEP puts
        pending += n;
        if (pending >= ADLER_NMAX - COPY_STEP) { a %= ADLER_BASE; b %= ADLER_BASE; pending = 0; }
    }
    INF_HD void fail(int st) { status = st; need = DONE; }
};

// zlib header; afterwards the lane asks for its first block
INF_HD void lane_begin(Lane &L, const uint8_t *src, int64_t n, uint8_t *out, uint8_t *spill, uint32_t keep, uint32_t full) {
    L.r.init(src, n);
    L.out = out; L.hi = spill - keep; L.keep = keep; L.full = full;
    L.o = 0; L.a = 1; L.b = 0; L.pending = 0; L.mlen = 0; L.mdist = 1;
    L.nbits = n * 8;
    L.state = ST_LITLEN; L.need = NEW_BLOCK; L.status = INF_OK; L.last = 0;
    L.r.refill();
    const uint32_t cmf = L.r.take(8), flg = L.r.take(8);
    if (n < 6 || (cmf & 15u) != 8u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) L.fail(INF_BAD_HEADER);
}

// One unit of work of a lane with need == RUN: the bytes of a pending match are LOADED first and STORED last, so
// that their latency (they were written moments ago: an L2 round trip) hides behind the symbol decode of the lanes
// that are not copying; the decode is one section for both codes.
INF_HD void lane_step(Lane &L, const uint16_t *lit_tab, const uint16_t *dist_tab, int stride, const Scratch &s) {
    const bool copying = L.state == ST_COPY;
    // byte i of a match equals byte (i mod dist) of the dist bytes before it: independent loads
    const uint32_t n = copying ? (L.mlen < (uint32_t)COPY_STEP ? L.mlen : (uint32_t)COPY_STEP) : 0u;
    const uint32_t base = L.o - L.mdist;
    uint8_t c[COPY_STEP];
    uint32_t idx = 0;
#pragma unroll
    for (int i = 0; i < COPY_STEP; ++i) {
        c[i] = (uint32_t)i < n ? *L.at(base + idx) : (uint8_t)0;
        ++idx;
        if (idx == L.mdist) idx = 0;
    }
    uint32_t emitted = n;
    if (!copying) {
        L.r.refill();
        const bool want_dist = L.state == ST_DIST;
        const int bits = want_dist ? DIST_BITS : LIT_BITS;
        const uint32_t e = (want_dist ? dist_tab : lit_tab)[L.r.peek(bits) * stride];
        int sym;
        if (e) {
            L.r.drop((int)(e >> 12));
            sym = (int)(e & 0xfffu);
        } else {
            // longer than the table: the code's length is where the next 15 bits (first bit most significant) fall
            // between the left-justified code ranges of the lengths
            const uint32_t v = reverse15(L.r.peek(15));
            int len = bits + 1;
            uint32_t last_end = 0;
#pragma unroll
            for (int k = 0; k < N_LONG; ++k) {
                const uint32_t end = want_dist ? L.dist_end[k] : L.lit_end[k];
                if (k < N_LONG - 1) len += (DIST_BITS + 1 + k > bits && v >= end) ? 1 : 0;
                else last_end = end;
            }
            uint32_t first = 0, offs = 0;
#pragma unroll
            for (int k = 0; k < N_LONG; ++k)
                if (len == DIST_BITS + 1 + k) {
                    first = want_dist ? L.dist_first[k] : L.lit_first[k];
                    offs = want_dist ? L.dist_offs[k] : L.lit_offs[k];
                }
            if (v >= last_end) {
                sym = -1;
            } else {
                sym = (want_dist ? s.dist : s.lit).sorted[offs + ((v - first) >> (15 - len))];
                L.r.drop(len);
            }
        }
        if (sym < 0) {
            L.fail(INF_BAD_CODE);
        } else if (want_dist) {
            if (sym >= 30) {
                L.fail(INF_BAD_CODE);
            } else {
                // distance code d >= 4: (d >> 1) - 1 extra bits on the base ((2 + (d & 1)) << extra) + 1   (RFC 1951 3.2.5)
                const int extra = sym < 4 ? 0 : (sym >> 1) - 1;
                L.mdist = (sym < 4 ? (uint32_t)sym : ((2u + (uint32_t)(sym & 1)) << extra)) + 1u + L.r.take(extra);
                if (L.mdist > L.o) L.fail(INF_BAD_DISTANCE);
                else if (L.mlen > L.full - L.o) L.fail(INF_OVERFLOW);
                else L.state = ST_COPY;
            }
        } else if (sym < 256) {
            if (L.o >= L.full) {
                L.fail(INF_OVERFLOW);
            } else {
                L.put((uint32_t)sym);
                emitted = 1;
            }
        } else if (sym == 256) {
            L.need = L.last ? FINISH : NEW_BLOCK;
        } else {
            sym -= 257;
            if (sym >= 29) {
                L.fail(INF_BAD_CODE);
            } else {
                // length code c = sym - 257 >= 8: (c - 4) >> 2 extra bits on ((4 + (c & 3)) << extra) + 3; c = 28: 258
                const int extra = (sym < 8 || sym == 28) ? 0 : (sym - 4) >> 2;
                L.mlen = (sym < 8 ? (uint32_t)sym : (sym == 28 ? 255u : (4u + (uint32_t)(sym & 3)) << extra)) + 3u + L.r.take(extra);
                L.state = ST_DIST;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < COPY_STEP; ++i)
        if ((uint32_t)i < n) L.put(c[i]);
    L.mlen -= n;
    if (copying && L.mlen == 0) L.state = ST_LITLEN;
    L.settle(emitted);
}

// need == NEW_BLOCK: block header (stored blocks are copied here), Huffman tables.  need == FINISH: trailer.
INF_HD void lane_service(Lane &L, uint16_t *lit_tab, uint16_t *dist_tab, int stride, Scratch &s) {
    Reader &r = L.r;
    while (L.need == NEW_BLOCK) {
        r.refill();
        if (r.consumed_bits() > L.nbits) { L.fail(INF_INPUT_OVERRUN); return; }
        L.last = r.take(1);
        const uint32_t type = r.take(2);
        if (type == 0) {
            r.drop(r.bc & 7);
            r.refill();
            const uint32_t len = r.take(16), nlen = r.take(16);
            if ((len ^ 0xffffu) != nlen) { L.fail(INF_BAD_BLOCK); return; }
            if (len > L.full - L.o) { L.fail(INF_OVERFLOW); return; }
            for (uint32_t i = 0; i < len; ++i) {
                r.refill();
                L.put(r.take(8));
                L.settle(1);
            }
            if (L.last) L.need = FINISH;
            continue;
        }
        if (type == 3) { L.fail(INF_BAD_BLOCK); return; }
        int nlit, ndist;
        if (type == 1) {
            nlit = 288; ndist = 30;
            for (int i = 0; i < 144; ++i) s.lens[i] = 8;
            for (int i = 144; i < 256; ++i) s.lens[i] = 9;
            for (int i = 256; i < 280; ++i) s.lens[i] = 7;
            for (int i = 280; i < 288; ++i) s.lens[i] = 8;
            for (int i = 0; i < 30; ++i) s.lens[288 + i] = 5;
        } else {
            nlit = (int)r.take(5) + 257;
            ndist = (int)r.take(5) + 1;
            const int ncl = (int)r.take(4) + 4;
            if (nlit > 286 || ndist > 30) { L.fail(INF_BAD_BLOCK); return; }
            uint8_t cl[19];
            for (int i = 0; i < 19; ++i) cl[i] = 0;
            for (int i = 0; i < ncl; ++i) {
                r.refill();
                cl[INF_TAB(clorder)[i]] = (uint8_t)r.take(3);
            }
            // the code-length code borrows the distance table and the literal code's scratch
            if (!build_table(cl, 19, s.lit, dist_tab, stride, DIST_BITS)) { L.fail(INF_BAD_CODE); return; }
            int i = 0;
            while (i < nlit + ndist) {
                r.refill();
                const int sym = decode_symbol(r, dist_tab, stride, DIST_BITS, s.lit);
                if (sym < 0) { L.fail(INF_BAD_CODE); return; }
                if (sym < 16) {
                    s.lens[i++] = (uint8_t)sym;
                    continue;
                }
                uint8_t v = 0;
                int rep;
                if (sym == 16) {
                    if (i == 0) { L.fail(INF_BAD_BLOCK); return; }
                    v = s.lens[i - 1];
                    rep = 3 + (int)r.take(2);
                } else if (sym == 17) {
                    rep = 3 + (int)r.take(3);
                } else {
                    rep = 11 + (int)r.take(7);
                }
                if (i + rep > nlit + ndist) { L.fail(INF_BAD_BLOCK); return; }
                while (rep--) s.lens[i++] = v;
            }
            if (s.lens[256] == 0) { L.fail(INF_BAD_BLOCK); return; }
        }
        // (the distance lengths follow the literal lengths in s.lens)
        if (!build_table(s.lens + nlit, ndist, s.dist, dist_tab, stride, DIST_BITS) ||
            !build_table(s.lens, nlit, s.lit, lit_tab, stride, LIT_BITS)) {
            L.fail(INF_BAD_CODE);
            return;
        }
#pragma unroll
        for (int k = 0; k < N_LONG; ++k) {
            const int l = DIST_BITS + 1 + k;
            L.lit_end[k] = s.lit.end[l]; L.lit_first[k] = s.lit.first[l]; L.lit_offs[k] = s.lit.offs[l];
            L.dist_end[k] = s.dist.end[l]; L.dist_first[k] = s.dist.first[l]; L.dist_offs[k] = s.dist.offs[l];
        }
        L.state = ST_LITLEN;
        L.need = RUN;
        return;
    }
    if (L.need == FINISH) {
        r.drop(r.bc & 7);
        r.refill();
        uint32_t want = 0;
        for (int i = 0; i < 4; ++i) want = (want << 8) | r.take(8);
        L.need = DONE;
        if (r.consumed_bits() > L.nbits) { L.status = INF_INPUT_OVERRUN; return; }
        L.a %= ADLER_BASE;
        L.b %= ADLER_BASE;
        if (((L.b << 16) | L.a) != want) L.status = INF_BAD_CHECKSUM;
    }
}

}  // namespace inf
}  // namespace strique
