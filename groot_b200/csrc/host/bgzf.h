// BGZF block writer for the BAM stream of `groot align` (the reference gets this from biogo/hts: bam.NewWriter(w, h, 0),
// src/pipeline/boss.go:86-99,225-241 — BAM bytes are not pinned by the reference, any valid deflate stream decodes to
// the same records).
//
// Why it is not just zlib: theBoss writes one sam.Record per (read, path) — alignment.go:296-315 emits the SAME read
// once for every path through the start node, 17 records per aligned read against arg-annot.90 — so consecutive
// records of one (read, graph) pair are byte-identical except for refID, pos, bin and the secondary flag. zlib finds
// those repeats again by hashing every byte (~60-150 MB/s per thread); the writer below is TOLD where the previous
// record lies (a hint: "this record probably equals the bytes `dist` back") and turns equal stretches into LZ77 matches
// with one 8-byte-wide compare — one fixed-Huffman deflate block per BGZF block. It is correct by construction: a match
// is only emitted for bytes that WERE compared equal, whatever the hints say; bytes without a usable hint are literals
// (with run-length matches at distance 1). A block the hints do not help (single-path databases: every record is a
// different read) is handed to zlib at the configured level instead, so the default output is never much larger than
// the reference's.
#pragma once
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
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

// RFC 1951 3.2.6 fixed Huffman codes, stored ready for an LSB-first bit stream (code bit-reversed, extra bits behind it)
struct FixedTables {
    uint16_t lit_bits[256]; uint8_t lit_n[256];
    uint32_t len_bits[259]; uint8_t len_n[259];
    FixedTables() {
        for (int b = 0; b < 256; b++) {
            if (b < 144) { lit_bits[b] = static_cast<uint16_t>(bit_reverse(0x30u + b, 8)); lit_n[b] = 8; }
            else { lit_bits[b] = static_cast<uint16_t>(bit_reverse(0x190u + (b - 144), 9)); lit_n[b] = 9; }
        }
        static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        for (int len = 3; len <= 258; len++) {
            int c = 28;
            if (len < 258) { c = 0; while (c + 1 < 28 && base[c + 1] <= len) c++; }
            const int sym = 257 + c;
            uint32_t code; int n;
            if (sym < 280) { code = bit_reverse(static_cast<uint32_t>(sym - 256), 7); n = 7; }
            else { code = bit_reverse(0xC0u + (sym - 280), 8); n = 8; }
            len_bits[len] = code | static_cast<uint32_t>(len - base[c]) << n;
            len_n[len] = static_cast<uint8_t>(n + extra[c]);
        }
    }
};
inline const FixedTables& tables() { static const FixedTables t; return t; }

struct BitWriter {
    uint8_t* p; uint64_t acc = 0; int nb = 0;
    explicit BitWriter(uint8_t* out) : p(out) {}
    inline void put(uint32_t bits, int n) {        // n <= 32
        acc |= static_cast<uint64_t>(bits) << nb; nb += n;
        if (nb >= 32) { const uint32_t w = static_cast<uint32_t>(acc); memcpy(p, &w, 4); p += 4; acc >>= 32; nb -= 32; }
    }
    inline uint8_t* finish() { while (nb > 0) { *p++ = static_cast<uint8_t>(acc); acc >>= 8; nb -= 8; } nb = 0; return p; }
};

// distance 1..32768 -> 5-bit code (reversed) + extra bits
inline void dist_code(uint32_t d, uint32_t* bits, int* n) {
    if (d <= 4) { *bits = bit_reverse(d - 1, 5); *n = 5; return; }
    const uint32_t x = d - 1;
    const int hb = 31 - __builtin_clz(x);
    const uint32_t code = 2u * hb + ((x >> (hb - 1)) & 1u);
    *bits = bit_reverse(code, 5) | (x & ((1u << (hb - 1)) - 1u)) << 5;
    *n = 5 + hb - 1;
}
}  // namespace bgzf_detail

// One worker's output stream: records are appended to a pending buffer (reserve / commit), whole blocks are deflated
// by drain(). Blocks are cut every kBgzfBlock bytes wherever that falls — a BAM record may straddle two blocks.
class BgzfDeflater {
  public:
    // level: zlib's (-1 default, 0 stored .. 9); delta: use the hint-driven encoder where it pays (never at level 0)
    BgzfDeflater(int level, bool delta) : level_(level), delta_(delta && level != 0), scratch_(kBgzfBlock * 4 / 3 + 256) {}   // worst case: every 3 bytes a 31-bit match

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
        const uint32_t tail[2] = {static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), data, static_cast<uInt>(m))), static_cast<uint32_t>(m)};
        memcpy(out.data() + o + 18 + clen, tail, 8);
    }

    // literals of raw_[p, q), runs of one byte as matches at distance 1 (the run's first byte stays a literal)
    inline void put_literals(bgzf_detail::BitWriter& bw, size_t p, size_t q) {
        const bgzf_detail::FixedTables& T = bgzf_detail::tables();
        const uint8_t* d = raw_.data();
        while (p < q) {
            const uint8_t b = d[p];
            bw.put(T.lit_bits[b], T.lit_n[b]);
            p++;
            if (p + 3 <= q && d[p] == b && d[p + 1] == b && d[p + 2] == b) {
                size_t r = 3;
                while (p + r < q && d[p + r] == b) r++;
                put_match(bw, r, 1);
                p += r;
            }
        }
    }
    // a match of n >= 3 bytes at distance dist, cut into pieces of at most 258 bytes none of which is shorter than 3
    inline void put_match(bgzf_detail::BitWriter& bw, size_t n, uint32_t dist) {
        const bgzf_detail::FixedTables& T = bgzf_detail::tables();
        if (dist != dc_dist_) { dc_dist_ = dist; bgzf_detail::dist_code(dist, &dc_bits_, &dc_n_); }    // a block sees few distinct distances
        while (n > 0) {
            size_t take = std::min<size_t>(n, 258);
            if (n - take > 0 && n - take < 3) take = n - 3;
            bw.put(T.len_bits[take], T.len_n[take]);
            bw.put(dc_bits_, dc_n_);
            n -= take;
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
    inline void put_compared(bgzf_detail::BitWriter& bw, size_t p, size_t q, uint32_t dist) {
        const uint8_t* d = raw_.data();
        size_t lit0 = p;                       // start of the literal stretch not written yet
        while (p < q) {
            if (d[p] != d[p - dist]) { p++; continue; }
            const size_t e = p + equal_prefix(d + p, d + p - dist, q - p);      // [p, e) equals the bytes dist back
            if (e - p >= 3) {
                if (lit0 < p) put_literals(bw, lit0, p);
                put_match(bw, e - p, dist);
                lit0 = e;
            }
            p = e;                              // shorter: too short to pay for a match, stays in the literal stretch
        }
        if (lit0 < q) put_literals(bw, lit0, q);
    }

    // one fixed-Huffman block for raw_[at, at + m); h = first hint that reaches into it. Returns the deflate size
    // (in scratch_), which may exceed kBgzfMaxData — the caller then falls back to zlib.
    size_t encode_delta(size_t at, size_t m, size_t h) {
        bgzf_detail::BitWriter bw(scratch_.data());
        bw.put(3u, 3);                                                   // BFINAL = 1, BTYPE = 01
        const int64_t b0 = static_cast<int64_t>(at), b1 = static_cast<int64_t>(at + m);
        int64_t p = b0;
        for (size_t i = h; i < hints_.size() && hints_[i].off < b1; i++) {
            const Hint& hn = hints_[i];
            const int64_t s = std::max(hn.off, b0), e = std::min(hn.off + static_cast<int64_t>(hn.len), b1);
            if (e <= s) continue;
            if (p < s) put_literals(bw, static_cast<size_t>(p), static_cast<size_t>(s));      // bytes no record claims
            // a reference may not reach in front of the block: literal up to b0 + dist
            const int64_t ms = hn.dist ? std::max(s, b0 + static_cast<int64_t>(hn.dist)) : e;
            if (ms < e) {
                if (s < ms) put_literals(bw, static_cast<size_t>(s), static_cast<size_t>(ms));
                put_compared(bw, static_cast<size_t>(ms), static_cast<size_t>(e), hn.dist);
            } else put_literals(bw, static_cast<size_t>(s), static_cast<size_t>(e));
            p = e;
        }
        if (p < b1) put_literals(bw, static_cast<size_t>(p), static_cast<size_t>(b1));
        bw.put(0u, 7);                                                   // end of block (symbol 256)
        return static_cast<size_t>(bw.finish() - scratch_.data());
    }

    int level_;
    bool delta_;
    std::vector<uint8_t> raw_, scratch_;
    size_t fill_ = 0;
    std::vector<Hint> hints_;
    uint64_t delta_blocks_ = 0, zlib_blocks_ = 0;
    uint32_t dc_dist_ = 0, dc_bits_ = 0; int dc_n_ = 0;   // last distance code
};

}  // namespace groot_host
