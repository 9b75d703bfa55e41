// Fuzz of the BGZF block writer (groot_b200/csrc/host/bgzf.h): random byte streams — copies of earlier stretches with a few
// bytes changed, runs, noise — appended as "records" with hints that are right, wrong, too far or absent, drained at
// random moments; every block is inflated again with zlib and the whole stream compared with what went in.
//   bgzf_fuzz <seed> <iterations>      prints "ok <blocks> <delta blocks> <own-code blocks>"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../groot_b200/csrc/host/bgzf.h"

using groot_host::BgzfDeflater;

static bool inflate_all(const std::vector<uint8_t>& bgzf, std::vector<uint8_t>& out, size_t* blocks) {
    size_t at = 0;
    while (at < bgzf.size()) {
        if (at + 26 > bgzf.size() || bgzf[at] != 0x1f || bgzf[at + 1] != 0x8b || bgzf[at + 12] != 'B' || bgzf[at + 13] != 'C') return false;
        const size_t bsize = (bgzf[at + 16] | bgzf[at + 17] << 8) + 1u;
        if (at + bsize > bgzf.size()) return false;
        uint32_t crc, isize;
        memcpy(&crc, &bgzf[at + bsize - 8], 4); memcpy(&isize, &bgzf[at + bsize - 4], 4);
        if (isize > groot_host::kBgzfBlock) return false;
        std::vector<uint8_t> buf(isize + 1);
        z_stream zs{};
        if (inflateInit2(&zs, -15) != Z_OK) return false;
        zs.next_in = const_cast<Bytef*>(&bgzf[at + 18]); zs.avail_in = static_cast<uInt>(bsize - 26);
        zs.next_out = buf.data(); zs.avail_out = static_cast<uInt>(buf.size());
        const int rc = inflate(&zs, Z_FINISH);
        const bool ok = rc == Z_STREAM_END && zs.total_out == isize && zs.avail_in == 0;
        inflateEnd(&zs);
        if (!ok || static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), buf.data(), isize)) != crc) return false;
        out.insert(out.end(), buf.begin(), buf.begin() + isize);
        at += bsize; (*blocks)++;
    }
    return true;
}

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 ? static_cast<unsigned>(atoi(argv[1])) : 1;
    const int iters = argc > 2 ? atoi(argv[2]) : 20;
    std::mt19937 rng(seed);
    auto rnd = [&](uint32_t n) { return static_cast<uint32_t>(rng() % n); };
    size_t blocks = 0; uint64_t delta = 0, own = 0;
    for (int it = 0; it < iters; it++) {
        const int level = (int[]){-1, 1, 9, 0}[rnd(4)];
        BgzfDeflater z(level, rnd(8) != 0);
        std::vector<uint8_t> all, bgzf;
        const uint32_t n_rec = 1 + rnd(3000);
        const uint32_t alphabet = (uint32_t[]){2, 4, 20, 256}[rnd(4)];
        size_t prev_len = 0;
        for (uint32_t r = 0; r < n_rec; r++) {
            uint32_t len = rnd(20) == 0 ? rnd(90000) : 1 + rnd(400);
            const uint32_t kind = rnd(10);
            uint8_t* dst;
            uint32_t dist = 0;
            if (kind < 6 && prev_len > 0 && prev_len <= z.pending()) {              // a copy of the previous record with a few bytes changed
                len = static_cast<uint32_t>(prev_len);
                dst = z.reserve(len);
                memcpy(dst, dst - len, len);
                for (uint32_t k = rnd(6); k > 0; k--) dst[rnd(len)] = static_cast<uint8_t>(rnd(256));
                dist = len;
            } else {
                dst = z.reserve(len);
                for (uint32_t i = 0; i < len;) {
                    if (rnd(4) == 0) { const uint32_t run = std::min(len - i, 1 + rnd(600)); memset(dst + i, static_cast<int>(rnd(alphabet)), run); i += run; }
                    else { const uint32_t run = std::min(len - i, 1 + rnd(50)); for (uint32_t k = 0; k < run; k++) dst[i + k] = static_cast<uint8_t>(rnd(alphabet)); i += run; }
                }
            }
            switch (rnd(6)) {                                                        // what the writer is told
                case 0: dist = 0; break;
                case 1: dist = 1 + rnd(70000); break;                                // anything, also beyond the window / the pending bytes
                case 2: dist = 1 + rnd(8); break;
                default: break;                                                      // the truth (or nothing for fresh bytes)
            }
            all.insert(all.end(), dst, dst + len);
            z.commit(len, dist);
            prev_len = len;
            if (rnd(50) == 0) z.drain(false, bgzf);
            if (rnd(400) == 0) { z.drain(true, bgzf); prev_len = 0; }
        }
        z.drain(true, bgzf);
        delta += z.delta_blocks(); own += z.dynamic_blocks();
        std::vector<uint8_t> back;
        if (!inflate_all(bgzf, back, &blocks) || back != all) { printf("MISMATCH seed %u iteration %d (%zu bytes in, %zu out)\n", seed, it, all.size(), back.size()); return 1; }
    }
    printf("ok %zu %llu %llu\n", blocks, static_cast<unsigned long long>(delta), static_cast<unsigned long long>(own));
    return 0;
}
