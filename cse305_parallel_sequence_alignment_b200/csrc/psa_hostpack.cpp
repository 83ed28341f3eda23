// psa_hostpack.cpp -- host-only (compiled by the host compiler, no device code): ASCII reads -> the fixed-stride 2-bit layout psa_align_batch_packed takes
// ("multi-threaded parse/pack").
//
// psa_pack_bases (psa_capi.cu) is the one-sequence, byte-at-a-time statement of the layout; psa_pack_reads below is
// the batch form a caller with 10^6 ASCII reads per step uses: 8 bases per 64-bit word operation, split over host
// threads.  Both produce the same words and the same count of unrepresentable bytes.
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/psa.h"

namespace {

constexpr uint64_t ONES = 0x0101010101010101ull;

// 8 ASCII bases (little-endian: the first base in the low byte) -> 16 bits, base r in bits 2r: code = (c >> 1) & 3.
inline uint32_t pack8(uint64_t w) {
    uint64_t x = (w >> 1) & (3 * ONES);
    x = (x | (x >> 6)) & 0x000F000F000F000Full;
    x = (x | (x >> 12)) & 0x000000FF000000FFull;
    x = (x | (x >> 24)) & 0xFFFFull;
    return (uint32_t)x;
}

// Bytes of w that are not one of 'A' 'C' 'G' 'T', as a mask (non-zero byte = unrepresentable).  The letter a 2-bit code
// stands for is 0x41 + {0, 2, 0x13, 6}[code] = 0x41 + 2 b0 + 0x13 b1 - 0x0F (b0 & b1): rebuild it per byte and compare.
inline uint64_t mismatch8(uint64_t w) {
    const uint64_t b0 = (w >> 1) & ONES, b1 = (w >> 2) & ONES;
    const uint64_t expect = 0x41 * ONES + (b0 << 1) + b1 * 0x13 - (b0 & b1) * 0x0F;      // every byte stays below 0x100: no carries
    return expect ^ w;
}
inline unsigned nonzero_bytes(uint64_t t) {
    const uint64_t nz = (((t & (0x7F * ONES)) + 0x7F * ONES) | t) & (0x80 * ONES);       // bit 7 of every non-zero byte
    return (unsigned)(((nz >> 7) * ONES) >> 56);                                         // sum of the eight 0/1 bytes
}

size_t pack_range(const uint8_t* bases, size_t first, size_t last, size_t len, size_t src_stride, uint32_t* packed, size_t words) {
    size_t bad = 0;
    const size_t full = len / 16, tail = len % 16;
    for (size_t k = first; k < last; ++k) {
        const uint8_t* s = bases + k * src_stride;
        uint32_t* out = packed + k * words;
        for (size_t w = 0; w < full; ++w) {
            uint64_t lo, hi;
            std::memcpy(&lo, s + w * 16, 8);
            std::memcpy(&hi, s + w * 16 + 8, 8);
            const uint64_t tl = mismatch8(lo), th = mismatch8(hi);
            if (__builtin_expect((tl | th) != 0, 0)) bad += nonzero_bytes(tl) + nonzero_bytes(th);
            out[w] = pack8(lo) | (pack8(hi) << 16);
        }
        if (tail) {                                             // the last, partial word: padded with 'A' (code 0, representable)
            uint8_t buf[16];
            std::memset(buf, 'A', 16);
            std::memcpy(buf, s + full * 16, tail);
            uint64_t lo, hi;
            std::memcpy(&lo, buf, 8);
            std::memcpy(&hi, buf + 8, 8);
            const uint64_t tl = mismatch8(lo), th = mismatch8(hi);
            if (__builtin_expect((tl | th) != 0, 0)) bad += nonzero_bytes(tl) + nonzero_bytes(th);
            out[full] = pack8(lo) | (pack8(hi) << 16);
        }
    }
    return bad;
}

}  // namespace

extern "C" size_t psa_pack_reads(const uint8_t* bases, size_t n_reads, size_t len, size_t src_stride, uint32_t* packed,
                                 int n_threads) {
    if (!bases || !packed || n_reads == 0 || len == 0) return 0;
    const size_t words = (len + 15) / 16;
    size_t threads = n_threads > 0 ? (size_t)n_threads : (size_t)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    const size_t min_reads_per_thread = 4096;                  // a thread start costs more than packing a few hundred reads
    if (threads > (n_reads + min_reads_per_thread - 1) / min_reads_per_thread) threads = (n_reads + min_reads_per_thread - 1) / min_reads_per_thread;
    if (threads <= 1) return pack_range(bases, 0, n_reads, len, src_stride, packed, words);
    std::vector<size_t> bad(threads, 0);
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (size_t t = 0; t < threads; ++t) {
        const size_t first = n_reads * t / threads, last = n_reads * (t + 1) / threads;
        pool.emplace_back([=, &bad] { bad[t] = pack_range(bases, first, last, len, src_stride, packed, words); });
    }
    size_t total = 0;
    for (size_t t = 0; t < threads; ++t) { pool[t].join(); total += bad[t]; }
    return total;
}
