// BGZF block writer for the BAM stream of `groot align` (the reference gets this from biogo/hts: bam.NewWriter(w, h, 0),
// src/pipeline/boss.go:86-99,225-241 — BAM bytes are not pinned by the reference, any valid deflate stream decodes to
// the same records).
//
// Why it is not just zlib: theBoss writes one sam.Record per (read, path) — alignment.go:296-315 emits the SAME read
// once for every path through the start node, 17 records per aligned read against arg-annot.90 — so consecutive
// records of one (read, graph) pair are byte-identical except for refID, pos, bin and the secondary flag. zlib finds
// those repeats again by hashing every byte; the writer below is TOLD where the previous record lies (a hint: "this
// record probably equals the bytes `dist` back") and turns equal stretches into LZ77 matches with a 16-byte-wide
// compare. The tokens of a block (literals, matches, runs of one byte as matches at distance 1) are then written as ONE
// deflate block, with the fixed code of RFC 1951 3.2.6 or a Huffman code built for the block (3.2.7), whichever is
// shorter. It is correct by construction: a match is only emitted for bytes that WERE compared equal, whatever the
// hints say. A block the hints do not help (single-path databases: every record is a different read) is handed to zlib
// at the configured level instead, so the default output is never much larger than the reference's.
#pragma once
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace groot_host {

constexpr size_t kBgzfBlock = 0xff00;       // uncompressed bytes per block (htslib's choice: a stored block always fits in 64 KiB)
constexpr size_t kBgzfMaxData = 65536 - 26; // deflate bytes that fit behind the 18-byte header and in front of crc32 + isize

namespace bgzf_detail {
inline uint32_t bit_reverse(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i); return r; }

constexpr int kLitLen = 286, kDist = 30, kCodeLen = 19;

// RFC 1951 3.2.5 length symbols and the fixed codes of 3.2.6, ready for an LSB-first bit stream (Huffman codes are
// stored bit-reversed, extra bits behind them)
struct Tables {
    uint8_t len_sym[259], len_extra_n[29]; uint16_t len_base[29];     // match length -> symbol - 257; extra bits of a symbol
    uint16_t fix_lit_bits[256]; uint8_t fix_lit_n[256];
    uint32_t fix_len_bits[259]; uint8_t fix_len_n[259];               // code + extra bits of a match length
    uint8_t fix_sym_n[kLitLen];
    Tables() {
        static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        for (int c = 0; c < 29; c++) { len_base[c] = base[c]; len_extra_n[c] = extra[c]; }
        for (int s = 0; s < kLitLen; s++) fix_sym_n[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
        for (int b = 0; b < 256; b++) {
            fix_lit_bits[b] = static_cast<uint16_t>(b < 144 ? bit_reverse(0x30u + b, 8) : bit_reverse(0x190u + (b - 144), 9));
            fix_lit_n[b] = fix_sym_n[b];
        }
        memset(len_sym, 0, sizeof len_sym);
        for (int len = 3; len <= 258; len++) {
            int c = 28;
            if (len < 258) { c = 0; while (c + 1 < 28 && base[c + 1] <= len) c++; }
            len_sym[len] = static_cast<uint8_t>(c);
            const int sym = 257 + c, n = fix_sym_n[sym];
            const uint32_t code = sym < 280 ? bit_reverse(static_cast<uint32_t>(sym - 256), 7) : bit_reverse(0xC0u + (sym - 280), 8);
            fix_len_bits[len] = code | static_cast<uint32_t>(len - base[c]) << n;
            fix_len_n[len] = static_cast<uint8_t>(n + extra[c]);
        }
    }
};
inline const Tables& tables() { static const Tables t; return t; }

// CRC-32 (the gzip polynomial) of a block. Every byte of the BAM passes through it, at ~1.3 cycles per byte in zlib's
// table-driven code — a fifth of a worker's time once deflate itself is cheap. On x86-64 with carry-less multiply the
// block is folded 64 bytes at a time (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ
// Instruction", Intel 2009; constants for the reflected polynomial 0x1DB710641), ~10x faster; the tail of fewer than 16
// bytes and machines without the instruction go through zlib. tests/cpp/bgzf_fuzz.cpp inflates every block with zlib,
// which checks this CRC against zlib's own.
#if defined(__x86_64__)
__attribute__((target("pclmul,sse4.1")))
inline __m128i crc32_fold16(__m128i x, __m128i k, __m128i data) {   // x * x^N mod P (both halves) + the next 16 bytes
    return _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x, k, 0x00), _mm_clmulepi64_si128(x, k, 0x11)), data);
}
__attribute__((target("pclmul,sse4.1")))
inline __m128i crc32_load16(const uint8_t* p) { return _mm_loadu_si128(reinterpret_cast<const __m128i*>(p)); }
__attribute__((target("pclmul,sse4.1")))
inline uint32_t crc32_fold(const uint8_t* buf, size_t len, uint32_t crc) {   // len >= 64 and a multiple of 16; inverted state in and out
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ULL, 0x01c6e41596ULL};   // x^(4*128+32) mod P, x^(4*128-32) mod P
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ULL, 0x00ccaa009eULL};   // x^(128+32) mod P, x^(128-32) mod P
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ULL, 0x0000000000ULL};   // x^64 mod P
    alignas(16) static const uint64_t poly[2] = {0x01db710641ULL, 0x01f7011641ULL};   // P, floor(x^64 / P)
    __m128i x1 = crc32_load16(buf), x2 = crc32_load16(buf + 16), x3 = crc32_load16(buf + 32), x4 = crc32_load16(buf + 48);
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128(static_cast<int>(crc)));
    __m128i k = _mm_load_si128(reinterpret_cast<const __m128i*>(k1k2));
    buf += 64; len -= 64;
    while (len >= 64) {
        x1 = crc32_fold16(x1, k, crc32_load16(buf)); x2 = crc32_fold16(x2, k, crc32_load16(buf + 16));
        x3 = crc32_fold16(x3, k, crc32_load16(buf + 32)); x4 = crc32_fold16(x4, k, crc32_load16(buf + 48));
        buf += 64; len -= 64;
    }
    k = _mm_load_si128(reinterpret_cast<const __m128i*>(k3k4));
    x1 = crc32_fold16(x1, k, x2); x1 = crc32_fold16(x1, k, x3); x1 = crc32_fold16(x1, k, x4);        // four lanes into one
    while (len >= 16) { x1 = crc32_fold16(x1, k, crc32_load16(buf)); buf += 16; len -= 16; }
    // 128 -> 64 -> 32 bits (Barrett reduction)
    const __m128i mask = _mm_setr_epi32(~0, 0, ~0, 0);
    __m128i t = _mm_clmulepi64_si128(x1, k, 0x10);
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), t);
    k = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(k5k0));
    t = _mm_srli_si128(x1, 4);
    x1 = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(x1, mask), k, 0x00), t);
    k = _mm_load_si128(reinterpret_cast<const __m128i*>(poly));
    t = _mm_and_si128(_mm_clmulepi64_si128(_mm_and_si128(x1, mask), k, 0x10), mask);
    x1 = _mm_xor_si128(x1, _mm_clmulepi64_si128(t, k, 0x00));
    return static_cast<uint32_t>(_mm_extract_epi32(x1, 1));
}
#endif
inline uint32_t crc32_of(const uint8_t* p, size_t n) {
    uint32_t crc = 0;
#if defined(__x86_64__)
    static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (have && n >= 64) {
        const size_t chunk = n & ~static_cast<size_t>(15);
        crc = ~crc32_fold(p, chunk, ~crc);
        p += chunk; n -= chunk;
    }
#endif
    return n ? static_cast<uint32_t>(crc32(crc, p, static_cast<uInt>(n))) : crc;
}

struct BitWriter {
    uint8_t* p; uint64_t acc = 0; int nb = 0;
    explicit BitWriter(uint8_t* out) : p(out) {}
    inline void put(uint32_t bits, int n) {        // n <= 32
        acc |= static_cast<uint64_t>(bits) << nb; nb += n;
        if (nb >= 32) { const uint32_t w = static_cast<uint32_t>(acc); memcpy(p, &w, 4); p += 4; acc >>= 32; nb -= 32; }
    }
    inline uint8_t* finish() { while (nb > 0) { *p++ = static_cast<uint8_t>(acc); acc >>= 8; nb -= 8; } nb = 0; return p; }
};

// distance 1..32768 -> symbol 0..29, number of extra bits, their value (RFC 1951 3.2.5)
inline void dist_symbol(uint32_t d, uint32_t* sym, int* extra_n, uint32_t* extra) {
    if (d <= 4) { *sym = d - 1; *extra_n = 0; *extra = 0; return; }
    const uint32_t x = d - 1;
    const int hb = 31 - __builtin_clz(x);
    *sym = 2u * hb + ((x >> (hb - 1)) & 1u);
    *extra_n = hb - 1;
    *extra = x & ((1u << (hb - 1)) - 1u);
}

// Huffman code lengths of at most max_bits for the symbols with freq != 0 (at least two symbols get a code, so the code
// is always complete: what zlib's build_tree does for inflaters that insist on it). Plain two-queue Huffman; if the tree
// comes out too deep the frequencies are halved (rounding up) and it is built again — that flattens it until it fits.
inline void huffman_lengths(const uint32_t* freq, int n, int max_bits, uint8_t* len) {
    struct Leaf { uint32_t f; int s; };
    Leaf leaf[288];
    int m = 0;
    for (int s = 0; s < n; s++) if (freq[s]) leaf[m++] = {freq[s], s};
    for (int s = 0; m < 2 && s < n; s++) if (!freq[s]) leaf[m++] = {1u, s};
    // ascending by (frequency, symbol): the leaves are in symbol order, so a stable radix sort on the frequency does it
    // (two or four 8-bit passes; a block has at most 65 280 symbols, so two passes almost always)
    {
        Leaf tmp[288];
        uint32_t all = 0;
        for (int i = 0; i < m; i++) all |= leaf[i].f;
        for (int shift = 0; shift < 32 && (all >> shift); shift += 8) {
            int count[257] = {0};
            for (int i = 0; i < m; i++) count[((leaf[i].f >> shift) & 0xffu) + 1]++;
            for (int b = 0; b < 256; b++) count[b + 1] += count[b];
            for (int i = 0; i < m; i++) tmp[count[(leaf[i].f >> shift) & 0xffu]++] = leaf[i];
            memcpy(leaf, tmp, static_cast<size_t>(m) * sizeof(Leaf));
        }
    }
    memset(len, 0, static_cast<size_t>(n));
    uint64_t f[576]; int parent[576], depth[576];
    while (true) {
        for (int i = 0; i < m; i++) f[i] = leaf[i].f;
        int a = 0, b = m, k = m;                       // next unmerged leaf / internal node, next free internal slot
        while (k < 2 * m - 1) {
            int pick[2];
            for (int j = 0; j < 2; j++) pick[j] = (a < m && (b >= k || f[a] <= f[b])) ? a++ : b++;
            f[k] = f[pick[0]] + f[pick[1]];
            parent[pick[0]] = parent[pick[1]] = k;
            k++;
        }
        depth[2 * m - 2] = 0;
        int deepest = 0;
        for (int i = 2 * m - 3; i >= 0; i--) { depth[i] = depth[parent[i]] + 1; if (i < m) deepest = std::max(deepest, depth[i]); }
        if (deepest <= max_bits) break;
        for (int i = 0; i < m; i++) leaf[i].f = (leaf[i].f + 1u) >> 1;     // monotone: the order stays sorted
    }
    for (int i = 0; i < m; i++) len[leaf[i].s] = static_cast<uint8_t>(depth[i]);
}
// canonical codes of RFC 1951 3.2.2, bit-reversed for the LSB-first stream
inline void canonical_codes(const uint8_t* len, int n, uint16_t* code) {
    uint32_t count[16] = {0}, next[16] = {0};
    for (int s = 0; s < n; s++) count[len[s]]++;
    count[0] = 0;
    uint32_t c = 0;
    for (int bits = 1; bits < 16; bits++) { c = (c + count[bits - 1]) << 1; next[bits] = c; }
    for (int s = 0; s < n; s++) code[s] = len[s] ? static_cast<uint16_t>(bit_reverse(next[len[s]]++, len[s])) : 0;
}
}  // namespace bgzf_detail

// One worker's output stream: records are appended to a pending buffer (reserve / commit), whole blocks are deflated
// by drain(). Blocks are cut every kBgzfBlock bytes wherever that falls — a BAM record may straddle two blocks.
class BgzfDeflater {
  public:
    // level: zlib's (-1 default, 0 stored .. 9); delta: use the hint-driven encoder where it pays (never at level 0)
    BgzfDeflater(int level, bool delta) : level_(level), delta_(delta && level != 0), scratch_(kBgzfBlock * 4 / 3 + 1024) {}   // worst case: every 3 bytes a 31-bit match

    // room for one record behind the pending bytes; the pointer is valid until the next reserve / drain, and the bytes
    // in front of it are the previous records (a caller may copy from `ptr - len_of_previous`)
    uint8_t* reserve(size_t len) {
        if (raw_.size() < fill_ + len) raw_.resize(std::max(raw_.size() * 2, fill_ + len + (1u << 16)));
        return raw_.data() + fill_;
    }
    // the record is in place; dist = how far back bytes that probably equal it start (0: nothing known)
    void commit(size_t len, uint32_t dist) {
        if (delta_) hints_.push_back({static_cast<int64_t>(fill_), static_cast<uint32_t>(len), dist <= fill_ && dist <= 32768u ? dist : 0u});
        fill_ += len;
    }
    size_t pending() const { return fill_; }

    // deflates every whole block (all pending bytes when final) and appends the BGZF blocks to out
    void drain(bool final, std::vector<uint8_t>& out) {
        const size_t upto = final ? fill_ : fill_ / kBgzfBlock * kBgzfBlock;
        size_t h = 0;
        for (size_t at = 0; at < upto; at += kBgzfBlock) {
            const size_t m = std::min(upto - at, kBgzfBlock);
            size_t clen = 0;
            bool done = false;
            if (delta_) {
                while (h < hints_.size() && hints_[h].off + hints_[h].len <= static_cast<int64_t>(at)) h++;
                clen = encode_delta(at, m, h);
                done = clen <= kBgzfMaxData && clen * 3 <= m + 64;          // worth it: at most a third of the input
                delta_blocks_ += done ? 1 : 0;
                dynamic_blocks_ += done && last_dynamic_ ? 1 : 0;
            }
            if (!done) clen = encode_zlib(raw_.data() + at, m);
            zlib_blocks_ += done ? 0 : 1;
            wrap(raw_.data() + at, m, clen, out);
        }
        // keep the tail; hints that still reach into it are rebased (a straddling record gets a negative start)
        size_t keep = 0;
        for (size_t i = h; i < hints_.size(); i++) {
            if (hints_[i].off + hints_[i].len <= static_cast<int64_t>(upto)) continue;
            hints_[keep] = hints_[i]; hints_[keep].off -= static_cast<int64_t>(upto); keep++;
        }
        hints_.resize(keep);
        if (upto && fill_ > upto) memmove(raw_.data(), raw_.data() + upto, fill_ - upto);
        fill_ -= upto;
    }
    uint64_t delta_blocks() const { return delta_blocks_; }
    uint64_t zlib_blocks() const { return zlib_blocks_; }
    uint64_t dynamic_blocks() const { return dynamic_blocks_; }    // of the delta blocks: written with a code of their own

    // deflate + BGZF framing of data[0, n) with zlib only (what the header and single records use)
    static void compress_plain(const uint8_t* data, size_t n, int level, std::vector<uint8_t>& out) {
        BgzfDeflater z(level, false);
        for (size_t at = 0; at < n; at += kBgzfBlock) {
            const size_t m = std::min(n - at, kBgzfBlock);
            z.wrap(data + at, m, z.encode_zlib(data + at, m), out);
        }
    }

  private:
    struct Hint { int64_t off; uint32_t len, dist; };
    static constexpr uint32_t kMatch = 0x80000000u;        // token: literal byte, or kMatch | length << 16 | (distance - 1)

    size_t encode_zlib(const uint8_t* data, size_t m) {
        z_stream zs{};
        if (deflateInit2(&zs, level_ < 0 ? Z_DEFAULT_COMPRESSION : level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("deflateInit2 failed");
        zs.next_in = const_cast<Bytef*>(data); zs.avail_in = static_cast<uInt>(m);
        zs.next_out = scratch_.data(); zs.avail_out = static_cast<uInt>(std::min(scratch_.size(), kBgzfMaxData));
        const int rc = deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END) throw std::runtime_error("deflate failed");
        return clen;
    }

    void wrap(const uint8_t* data, size_t m, size_t clen, std::vector<uint8_t>& out) const {
        uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
        const uint16_t bsize = static_cast<uint16_t>(clen + 25);
        hdr[16] = static_cast<uint8_t>(bsize); hdr[17] = static_cast<uint8_t>(bsize >> 8);
        const size_t o = out.size();
        out.resize(o + 18 + clen + 8);
        memcpy(out.data() + o, hdr, 18);
        memcpy(out.data() + o + 18, scratch_.data(), clen);
        const uint32_t tail[2] = {bgzf_detail::crc32_of(data, m), static_cast<uint32_t>(m)};
        memcpy(out.data() + o + 18 + clen, tail, 8);
    }

    // ---- tokens of one block -------------------------------------------------------------------------------------
    // a match of n >= 3 bytes at distance dist, cut into pieces of at most 258 bytes none of which is shorter than 3
    inline void tok_match(size_t n, uint32_t dist) {
        const bgzf_detail::Tables& T = bgzf_detail::tables();
        if (dist != ds_dist_) { ds_dist_ = dist; int en; uint32_t ev; bgzf_detail::dist_symbol(dist, &ds_sym_, &en, &ev); }    // a block sees few distinct distances
        while (n > 0) {
            size_t take = std::min<size_t>(n, 258);
            if (n - take > 0 && n - take < 3) take = n - 3;
            *tok_end_++ = kMatch | static_cast<uint32_t>(take) << 16 | (dist - 1u);
            lfreq_[257 + T.len_sym[take]]++; dfreq_[ds_sym_]++;
            n -= take;
        }
    }
    // literals of raw_[p, q), runs of one byte as matches at distance 1 (the run's first byte stays a literal)
    inline void tok_literals(size_t p, size_t q) {
        const uint8_t* d = raw_.data();
        while (p < q) {
            const uint8_t b = d[p];
            *tok_end_++ = b; lfreq_[b]++;
            p++;
            if (p + 3 <= q && d[p] == b && d[p + 1] == b && d[p + 2] == b) {
                size_t r = 3;
                while (p + r < q && d[p + r] == b) r++;
                tok_match(r, 1);
                p += r;
            }
        }
    }
    // length of the common prefix of a[0, n) and b[0, n)
    static inline size_t equal_prefix(const uint8_t* a, const uint8_t* b, size_t n) {
        size_t e = 0;
#if defined(__SSE2__)
        while (e + 16 <= n) {
            const unsigned neq = ~static_cast<unsigned>(_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(a + e)),
                                                                                         _mm_loadu_si128(reinterpret_cast<const __m128i*>(b + e))))) & 0xffffu;
            if (neq) return e + static_cast<size_t>(__builtin_ctz(neq));
            e += 16;
        }
#endif
        while (e + 8 <= n) {
            uint64_t x, y;
            memcpy(&x, a + e, 8); memcpy(&y, b + e, 8);
            if (x != y) return e + (static_cast<size_t>(__builtin_ctzll(x ^ y)) >> 3);
            e += 8;
        }
        while (e < n && a[e] == b[e]) e++;
        return e;
    }
    // raw_[p, q) against the bytes dist back: equal stretches of 3+ bytes become matches, the rest literals
    inline void tok_compared(size_t p, size_t q, uint32_t dist) {
        const uint8_t* d = raw_.data();
        size_t lit0 = p;                       // start of the literal stretch not written yet
        while (p < q) {
            if (d[p] != d[p - dist]) { p++; continue; }
            const size_t e = p + equal_prefix(d + p, d + p - dist, q - p);      // [p, e) equals the bytes dist back
            if (e - p >= 3) {
                if (lit0 < p) tok_literals(lit0, p);
                tok_match(e - p, dist);
                lit0 = e;
            }
            p = e;                              // shorter: too short to pay for a match, stays in the literal stretch
        }
        if (lit0 < q) tok_literals(lit0, q);
    }

    // ---- one deflate block for raw_[at, at + m); h = first hint that reaches into it --------------------------------
    // Returns the deflate size (in scratch_), which may exceed kBgzfMaxData — the caller then falls back to zlib.
    size_t encode_delta(size_t at, size_t m, size_t h) {
        using namespace bgzf_detail;
        const Tables& T = tables();
        if (tok_.size() < m + 16) tok_.resize(m + 16);                    // at most one token per byte
        tok_end_ = tok_.data();
        memset(lfreq_, 0, sizeof lfreq_); memset(dfreq_, 0, sizeof dfreq_);
        ds_dist_ = 0;
        const int64_t b0 = static_cast<int64_t>(at), b1 = static_cast<int64_t>(at + m);
        int64_t p = b0;
        for (size_t i = h; i < hints_.size() && hints_[i].off < b1; i++) {
            const Hint& hn = hints_[i];
            const int64_t s = std::max(hn.off, b0), e = std::min(hn.off + static_cast<int64_t>(hn.len), b1);
            if (e <= s) continue;
            if (p < s) tok_literals(static_cast<size_t>(p), static_cast<size_t>(s));      // bytes no record claims
            // a reference may not reach in front of the block: literal up to b0 + dist
            const int64_t ms = hn.dist ? std::max(s, b0 + static_cast<int64_t>(hn.dist)) : e;
            if (ms < e) {
                if (s < ms) tok_literals(static_cast<size_t>(s), static_cast<size_t>(ms));
                tok_compared(static_cast<size_t>(ms), static_cast<size_t>(e), hn.dist);
            } else tok_literals(static_cast<size_t>(s), static_cast<size_t>(e));
            p = e;
        }
        if (p < b1) tok_literals(static_cast<size_t>(p), static_cast<size_t>(b1));
        lfreq_[256] = 1;                                                  // end of block

        // the block's own code (RFC 1951 3.2.7) and what it costs against the fixed one
        uint8_t llen[kLitLen], dlen[kDist];
        huffman_lengths(lfreq_, kLitLen, 15, llen);
        huffman_lengths(dfreq_, kDist, 15, dlen);
        int nlit = kLitLen, ndist = kDist;
        while (nlit > 257 && !llen[nlit - 1]) nlit--;
        while (ndist > 1 && !dlen[ndist - 1]) ndist--;
        // the code lengths themselves, run-length coded with the symbols 16 (repeat), 17 / 18 (zeros)
        uint8_t seq[kLitLen + kDist];
        memcpy(seq, llen, static_cast<size_t>(nlit)); memcpy(seq + nlit, dlen, static_cast<size_t>(ndist));
        struct ClTok { uint8_t sym, extra; };
        ClTok cl[kLitLen + kDist];
        int ncl_tok = 0;
        uint32_t clfreq[kCodeLen] = {0};
        auto emit = [&](int sym, int extra) { cl[ncl_tok++] = {static_cast<uint8_t>(sym), static_cast<uint8_t>(extra)}; clfreq[sym]++; };
        for (int i = 0, n = nlit + ndist; i < n;) {
            const int v = seq[i];
            int run = 1;
            while (i + run < n && seq[i + run] == v) run++;
            i += run;
            if (v == 0) {
                while (run >= 11) { const int t = std::min(run, 138); emit(18, t - 11); run -= t; }
                if (run >= 3) { emit(17, run - 3); run = 0; }
            } else {
                emit(v, 0); run--;
                while (run >= 3) { const int t = std::min(run, 6); emit(16, t - 3); run -= t; }
            }
            while (run-- > 0) emit(v, 0);
        }
        uint8_t cllen[kCodeLen];
        huffman_lengths(clfreq, kCodeLen, 7, cllen);
        static const uint8_t order[kCodeLen] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        int ncl = kCodeLen;
        while (ncl > 4 && !cllen[order[ncl - 1]]) ncl--;
        uint64_t extra_bits = 0, dyn_bits = 3 + 5 + 5 + 4 + 3 * static_cast<uint64_t>(ncl), fix_bits = 3;
        for (int c = 0; c < 29; c++) extra_bits += static_cast<uint64_t>(lfreq_[257 + c]) * T.len_extra_n[c];
        for (int d = 4; d < kDist; d++) extra_bits += static_cast<uint64_t>(dfreq_[d]) * static_cast<uint64_t>(d / 2 - 1);
        for (int i = 0; i < ncl_tok; i++) dyn_bits += cllen[cl[i].sym] + (cl[i].sym == 16 ? 2 : cl[i].sym == 17 ? 3 : cl[i].sym == 18 ? 7 : 0);
        for (int s = 0; s < kLitLen; s++) { dyn_bits += static_cast<uint64_t>(lfreq_[s]) * llen[s]; fix_bits += static_cast<uint64_t>(lfreq_[s]) * T.fix_sym_n[s]; }
        for (int d = 0; d < kDist; d++) { dyn_bits += static_cast<uint64_t>(dfreq_[d]) * dlen[d]; fix_bits += static_cast<uint64_t>(dfreq_[d]) * 5u; }
        dyn_bits += extra_bits; fix_bits += extra_bits;

        BitWriter bw(scratch_.data());
        last_dynamic_ = false;
        if (fix_bits <= dyn_bits) {
            bw.put(3u, 3);                                                   // BFINAL = 1, BTYPE = 01
            uint32_t last_dist = 0, dbits = 0; int dn = 0;
            for (const uint32_t* tp = tok_.data(); tp != tok_end_; tp++) {
                const uint32_t t = *tp;
                if (!(t & kMatch)) { bw.put(T.fix_lit_bits[t], T.fix_lit_n[t]); continue; }
                const uint32_t len = (t >> 16) & 0x1ffu, dist = (t & 0x7fffu) + 1u;
                if (dist != last_dist) {
                    uint32_t sym, ev; int en;
                    dist_symbol(dist, &sym, &en, &ev);
                    last_dist = dist; dbits = bit_reverse(sym, 5) | ev << 5; dn = 5 + en;
                }
                bw.put(T.fix_len_bits[len], T.fix_len_n[len]);
                bw.put(dbits, dn);
            }
            bw.put(0u, 7);                                                   // end of block (symbol 256)
            return static_cast<size_t>(bw.finish() - scratch_.data());
        }
        last_dynamic_ = true;
        uint16_t lcode[kLitLen], dcode[kDist], clcode[kCodeLen];
        canonical_codes(llen, kLitLen, lcode); canonical_codes(dlen, kDist, dcode); canonical_codes(cllen, kCodeLen, clcode);
        bw.put(5u, 3);                                                       // BFINAL = 1, BTYPE = 10
        bw.put(static_cast<uint32_t>(nlit - 257), 5); bw.put(static_cast<uint32_t>(ndist - 1), 5); bw.put(static_cast<uint32_t>(ncl - 4), 4);
        for (int i = 0; i < ncl; i++) bw.put(cllen[order[i]], 3);
        for (int i = 0; i < ncl_tok; i++) {
            bw.put(clcode[cl[i].sym], cllen[cl[i].sym]);
            if (cl[i].sym >= 16) bw.put(cl[i].extra, cl[i].sym == 16 ? 2 : cl[i].sym == 17 ? 3 : 7);
        }
        uint32_t last_dist = 0, dbits = 0; int dn = 0;
        uint32_t lit[256];                                                   // code | length << 16 of the literals
        for (int b = 0; b < 256; b++) lit[b] = lcode[b] | static_cast<uint32_t>(llen[b]) << 16;
        for (const uint32_t* tp = tok_.data(); tp != tok_end_; tp++) {
            const uint32_t t = *tp;
            if (!(t & kMatch)) { bw.put(lit[t] & 0xffffu, static_cast<int>(lit[t] >> 16)); continue; }
            const uint32_t len = (t >> 16) & 0x1ffu, dist = (t & 0x7fffu) + 1u;
            if (dist != last_dist) {
                uint32_t sym, ev; int en;
                dist_symbol(dist, &sym, &en, &ev);
                last_dist = dist; dbits = dcode[sym] | ev << dlen[sym]; dn = dlen[sym] + en;
            }
            const int c = T.len_sym[len], ls = 257 + c;
            bw.put(lcode[ls] | static_cast<uint32_t>(len - T.len_base[c]) << llen[ls], llen[ls] + T.len_extra_n[c]);
            bw.put(dbits, dn);
        }
        bw.put(lcode[256], llen[256]);
        return static_cast<size_t>(bw.finish() - scratch_.data());
    }

    int level_;
    bool delta_;
    std::vector<uint8_t> raw_, scratch_;
    size_t fill_ = 0;
    std::vector<Hint> hints_;
    std::vector<uint32_t> tok_;
    uint32_t* tok_end_ = nullptr;
    uint32_t lfreq_[bgzf_detail::kLitLen], dfreq_[bgzf_detail::kDist];
    uint32_t ds_dist_ = 0, ds_sym_ = 0;                    // last distance -> symbol
    uint64_t delta_blocks_ = 0, zlib_blocks_ = 0, dynamic_blocks_ = 0;
    bool last_dynamic_ = false;
};

}  // namespace groot_host
