// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product path
// (groot_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// CPU restatement of github.com/will-rowe/nthash v0.2.0 (go.mod:16 of the reference), the
// rolling canonical ntHash (v1, plain 64-bit rotates) plus its "multi-hash" extension, as used
// by the reference at src/minhash/khf.go:38,44 (NewHasher / MultiHash(canonical, S)).
//
// PARITY STATUS: the module is NOT vendored under /root/reference (third-party dependency), so
// the algorithm is restated from its published form (ntHash, Mohamadi et al. 2016, alg. 3, and
// the NTMC64 multi-hash extension). External anchors checked in tests/test_oracle_kat.py:
//   ntf64("TGCAG",k=5)=0x0bafa6728fc6dabf  ntr64=0x8cf2d4072cca480e  canonical=0x0bafa6728fc6dabf
//   ntf64("ACGTC",k=5)=0xa7d01e3fb5593252  ntr64=0x480202d54e8ebecd  canonical=0x480202d54e8ebecd
// (known-answer values of the ntHash v1 test-suites: the TGCAG forward/reverse/canonical and the
// ACGTC canonical values are the published ones) and the reference's own RC-invariance test
// (src/minhash/minhash_test.go:111-157).  Hash VALUES are otherwise "parity unpinned" by the
// reference repo itself — it holds no golden hash vectors.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace oracle {

constexpr uint64_t kSeedA = 0x3c8bfbb395c60474ULL;
constexpr uint64_t kSeedC = 0x3193c18562a02b4cULL;
constexpr uint64_t kSeedG = 0x20323ed082572324ULL;
constexpr uint64_t kSeedT = 0x295549f54be24456ULL;
constexpr uint64_t kSeedN = 0x0000000000000000ULL;
constexpr uint64_t kMultiSeed = 0x90b45d39fb6da1faULL;
constexpr unsigned kMultiShift = 27;
constexpr uint8_t kCompMask = 0x07;  // seedTab[b & 7] is the seed of b's complement

inline uint64_t rol64(uint64_t v, unsigned s) { s &= 63; return s ? (v << s) | (v >> (64 - s)) : v; }
inline uint64_t ror64(uint64_t v, unsigned s) { s &= 63; return s ? (v >> s) | (v << (64 - s)) : v; }

// 256-entry seed table: forward seeds at 'A','C','G','T' (both cases); entries 1,3,4,7 hold the
// complement seeds so that tab[b & 7] is the seed of the complement of b ('A'&7=1 -> T, 'C'&7=3 ->
// G, 'T'&7=4 -> A, 'G'&7=7 -> C). Everything else (incl. 'N') is 0.
inline const uint64_t* seed_tab() {
    static uint64_t tab[256];
    static bool init = false;
    if (!init) {
        for (auto& t : tab) t = kSeedN;
        tab[1] = kSeedT; tab[3] = kSeedG; tab[4] = kSeedA; tab[7] = kSeedC;
        tab['A'] = tab['a'] = kSeedA;
        tab['C'] = tab['c'] = kSeedC;
        tab['G'] = tab['g'] = kSeedG;
        tab['T'] = tab['t'] = kSeedT;
        init = true;
    }
    return tab;
}

// forward hash of seq[0..k)
inline uint64_t ntf64(const uint8_t* s, unsigned k) {
    const uint64_t* tab = seed_tab();
    uint64_t h = 0;
    for (unsigned i = 0; i < k; i++) { h = rol64(h, 1); h ^= tab[s[i]]; }
    return h;
}
// reverse-complement hash of seq[0..k)
inline uint64_t ntr64(const uint8_t* s, unsigned k) {
    const uint64_t* tab = seed_tab();
    uint64_t h = 0;
    for (unsigned i = 0; i < k; i++) { h = rol64(h, 1); h ^= tab[s[k - 1 - i] & kCompMask]; }
    return h;
}

// Rolling hasher (nthash.NewHasher + Next). ok()==false mirrors the constructor error
// "k > len(seq)" which the reference turns into a panic at src/pipeline/boss.go:164-166.
class NtHasher {
  public:
    NtHasher(const uint8_t* seq, size_t len, unsigned k) : seq_(seq), len_(len), k_(k) {
        ok_ = (k >= 1 && len >= k);
        if (ok_) { fh_ = ntf64(seq, k); rh_ = ntr64(seq, k); }
    }
    bool ok() const { return ok_; }
    // returns false when exhausted; *out receives the (canonical) hash of the next k-mer
    bool next(bool canonical, uint64_t* out) {
        if (!ok_ || idx_ + k_ > len_) return false;
        if (idx_ != 0) {
            const uint64_t* tab = seed_tab();
            uint8_t prev = seq_[idx_ - 1], end = seq_[idx_ + k_ - 1];
            fh_ = rol64(fh_, 1) ^ rol64(tab[prev], k_) ^ tab[end];
            rh_ = ror64(rh_, 1) ^ ror64(tab[prev & kCompMask], 1) ^ rol64(tab[end & kCompMask], k_ - 1);
        }
        idx_++;
        *out = canonical ? (rh_ < fh_ ? rh_ : fh_) : fh_;
        return true;
    }

  private:
    const uint8_t* seq_; size_t len_; unsigned k_;
    size_t idx_ = 0; uint64_t fh_ = 0, rh_ = 0; bool ok_ = false;
};

// nthash MultiHash: m[0]=h; m[i] = x ^ (x >> 27) with x = h * (i ^ k*multiSeed), i = 1..S-1
inline void multi_hash(uint64_t h, unsigned k, unsigned S, uint64_t* out) {
    out[0] = h;
    for (uint64_t i = 1; i < S; i++) {
        uint64_t t = h * (i ^ (uint64_t)k * kMultiSeed);
        t ^= t >> kMultiShift;
        out[i] = t;
    }
}

}  // namespace oracle
