// K1+K2: per-read KHF MinHash sketch (ntHash v1 rolling canonical hash + S multi-hashes, per-slot min)
// fused with the LSH Ensemble probe and the full containment check.
//
// Replaces, per read:
//   Sequence.RunMinHash -> KHFsketch.AddSequence -> nthash MultiHash   (src/seqio/seqio.go:40-68, src/minhash/khf.go:35-56)
//   ContainmentIndex.Query -> lshensemble Query + Containment > t      (src/lshe/lshe.go:153-175)
//   the sketch worker loop of theBoss.mapReads                         (src/pipeline/boss.go:145-201)
//
// Mapping: ONE THREAD PER READ for the hashing. The S running minima (2*S registers) and the two rolling hashes stay
// in that thread's registers: no shuffles, no idle lanes (a warp-per-read layout would waste 26 of 96 lane-slots on
// 70 k-mers), and the rolling hash stays a 1-step roll. Read bytes come in through a TMA bulk copy (cp.async.bulk,
// global -> shared, mbarrier completion) of each warp's 32-read tile, double buffered; the 100-byte read stride is
// co-prime with the 32 banks so the per-thread byte reads are conflict free. The probe is warp-cooperative: the
// candidate buckets of the warp's 32 reads are pooled and verified 32 at a time (warp_probe). With a single band
// probed (the default) the kernel runs in two passes — a first-band prescreen over all reads, the full sketch only
// for the reads that found a bucket (SEED_PRESCREEN / SEED_QUEUED); the second pass also writes the 2-bit copies of the
// seeded reads the align kernels walk (pack_read_from_smem). Integer-issue bound (~18k integer ops per 100 bp read versus
// ~320 algorithmic bytes), see DESIGN.md §4. The sketch is tracked as raw k-mer products ordered by their top 27 bits
// (khf_update_raw), queries that need every slot equal look the whole sketch up (DevIndex::full), and sketch sizes /
// maxK outside the compiled set run the run-time-S kernels at the end of this file.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.cuh"

namespace groot {

constexpr int kSeedThreads = 128;  // threads per block (4 independent warps)
#ifndef GROOT_SEED_MIN_BLOCKS
#define GROOT_SEED_MIN_BLOCKS 4
#endif

struct SeedTabs {
    uint64_t in[256];   // seed[b]
    uint64_t out[256];  // rol(seed[b], k)
    uint64_t c0[8];     // seed[comp(b)] via b & 7
    uint64_t cin[8];    // rol(c0, k-1)
    uint64_t cout[8];   // ror(c0, 1)
};

__device__ __forceinline__ uint64_t rol64v(uint64_t v, unsigned s) {
    s &= 63u;
    return s ? (v << s) | (v >> (64u - s)) : v;
}
__device__ __forceinline__ uint64_t seed_of(unsigned b) {
    // 256-entry seedTab of nthash: ACGT (both cases) forward seeds, entries 1,3,4,7 = complement seeds
    switch (b) {
        case 'A': case 'a': case 4: return GROOT_SEED_A;
        case 'C': case 'c': case 7: return GROOT_SEED_C;
        case 'G': case 'g': case 3: return GROOT_SEED_G;
        case 'T': case 't': case 1: return GROOT_SEED_T;
        default: return 0ULL;
    }
}
__device__ inline void build_seed_tabs(SeedTabs* T, unsigned k) {
    for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) {
        uint64_t s = seed_of(i);
        T->in[i] = s;
        T->out[i] = rol64v(s, k);
        if (i < 8) { T->c0[i] = s; T->cin[i] = rol64v(s, k - 1); T->cout[i] = rol64v(s, 63); }
    }
}

// ---- the sketch: registers only -------------------------------------------------------------
#ifndef GROOT_KHF_VARIANT
#define GROOT_KHF_VARIANT 6   // 6: raw products tracked, order decided by their top 27 bits (below); 3: every value finished and compared in full;
                               // 4 / 5: c0 + low_i split of the multiplier (ptxas splits the 64-bit addend off again: no gain, kept for reference)
#endif
// m = min(m, x) with the two conditional moves issued as predicated multiply-adds (x * 1 + 0): they go to the FMA
// pipe, which has room, instead of the ALU pipe, which is the one that bounds this kernel (2 SEL per hash out of 8 ALU
// instructions; DESIGN.md "Kernels").
__device__ __forceinline__ void min_u64_fma(uint64_t& m, uint64_t x, uint32_t one) {
    uint32_t mlo = static_cast<uint32_t>(m), mhi = static_cast<uint32_t>(m >> 32);
    const uint32_t xlo = static_cast<uint32_t>(x), xhi = static_cast<uint32_t>(x >> 32);
    asm("{\n\t.reg .pred p;\n\t.reg .u64 a, b;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%0, %1};\n\tsetp.lt.u64 p, a, b;\n\t"
        "@p mad.lo.u32 %0, %2, %4, 0;\n\t@p mad.lo.u32 %1, %3, %4, 0;\n\t}"
        : "+r"(mlo), "+r"(mhi)
        : "r"(xlo), "r"(xhi), "r"(one));
    m = (static_cast<uint64_t>(mhi) << 32) | mlo;
}

// One k-mer hash h into the S running minima: m_0 = h, m_i = x ^ (x >> 27) with x = h * c_i, c_i = i ^ (k * multiSeed)
// (khf.go:44-53 / nthash MultiHash). For i < 32 the xor only touches the low 5 bits of the multiplier: c_i = c0 + low_i
// with c0 = C & ~31 and low_i = (C & 31) ^ i < 32, hence h * c_i = h * c0 + h * low_i (exact mod 2^64): ONE 64-bit
// product per k-mer (A = h * c0) and per hash a 32x32+64 multiply-add plus the high-word correction — 2 FMA-pipe
// instructions instead of the 4 of a full 64x64 product. lw[] holds the low_i in REGISTERS (21 of them: the sketch
// itself is 42): from the constant bank every use costs an extra LDC issue slot.
template <int S>
__device__ __forceinline__ void khf_update(uint64_t h, const MultTable& M, const uint32_t (&lw)[S], uint64_t (&sk)[S]) {
#if GROOT_KHF_VARIANT == 3
    min_u64_fma(sk[0], h, M.one);
#pragma unroll
    for (int i = 1; i < S; i++) {
        uint64_t x = h * M.c[i];
        x ^= x >> GROOT_MULTI_SHIFT;
        min_u64_fma(sk[i], x, M.one);
    }
#else
    min_u64_fma(sk[0], h, M.one);
    const uint32_t h_lo = static_cast<uint32_t>(h), h_hi = static_cast<uint32_t>(h >> 32);
    const uint64_t A = h * M.c0;
#pragma unroll
    for (int i = 1; i < S; i++) {
#if GROOT_KHF_VARIANT == 4
        const uint32_t li = M.low[i];
#else
        const uint32_t li = lw[i];
#endif
        uint32_t x_lo, x_hi;
        asm("{\n .reg .u64 w;\n mad.wide.u32 w, %2, %3, %4;\n mov.b64 {%0, %1}, w;\n}" : "=r"(x_lo), "=r"(x_hi) : "r"(h_lo), "r"(li), "l"(A));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x_hi) : "r"(h_hi), "r"(li));
        uint64_t x = (static_cast<uint64_t>(x_hi) << 32) | x_lo;
        x ^= x >> GROOT_MULTI_SHIFT;
        min_u64_fma(sk[i], x, M.one);
    }
#endif
}

// ---- variant 6: the sketch tracked as RAW products ------------------------------------------------------------------
// Slot i >= 1 is min over the k-mers of f(m), m = h * c_i mod 2^64, f(m) = m ^ (m >> 27). The shift only reaches bits
// 36..0, so bits 63..37 of f(m) ARE bits 63..37 of m: two products whose top 27 bits differ are ordered like their f
// values, and no k-mer has to be finished (2 shifts + 2 xors) or compared in full (2 set-predicates) to lose the race.
// Per slot the thread keeps the raw product that currently wins; per k-mer and slot it computes the product (the high
// word of c_i is the same for every i < 32, so h_lo * c_hi is shared: one wide multiply + one multiply-add + one add),
// compares the top 27 bits, overwrites on "smaller" (predicated multiply-adds: FMA pipe) and raises a flag on "equal".
// A raised flag (probability ~2^-27 per comparison) re-runs that k-mer with the exact comparison of f values for every
// slot — idempotent for the slots already updated. f is applied once per slot at the end. 4 FMA-pipe + 4-5 ALU-pipe
// instructions per hash instead of 6 + 6 and a constant load.
__device__ __forceinline__ uint64_t khf_finish(uint32_t lo, uint32_t hi) {
    const uint64_t m = (static_cast<uint64_t>(hi) << 32) | lo;
    return m ^ (m >> GROOT_MULTI_SHIFT);
}
template <int S>
__device__ __forceinline__ void khf_update_exact(uint64_t h, const MultTable& M, uint32_t (&rlo)[S], uint32_t (&rhi)[S]) {
#pragma unroll   // static indices: the raw products stay in registers (a call taking the arrays by reference would put them in local memory)
    for (int i = 1; i < S; i++) {
        const uint64_t m = h * M.c[i];
        if ((m ^ (m >> GROOT_MULTI_SHIFT)) < khf_finish(rlo[i], rhi[i])) { rlo[i] = static_cast<uint32_t>(m); rhi[i] = static_cast<uint32_t>(m >> 32); }
    }
}
template <int S, bool FIRST>
__device__ __forceinline__ void khf_update_raw(uint64_t h, const MultTable& M, uint64_t& s0, uint32_t (&rlo)[S], uint32_t (&rhi)[S]) {
    const uint32_t h_lo = static_cast<uint32_t>(h), h_hi = static_cast<uint32_t>(h >> 32);
    if (FIRST) s0 = h; else min_u64_fma(s0, h, M.one);
    const uint32_t P = h_lo * M.c_hi;
    const uint32_t kl = M.key_low;
    bool tie = false;
#pragma unroll
    for (int i = 1; i < S; i++) {
        const uint32_t clo = static_cast<uint32_t>(M.c[i]);
        uint32_t xl, xh;
        asm("{\n .reg .u64 w;\n mul.wide.u32 w, %2, %3;\n mov.b64 {%0, %1}, w;\n}" : "=r"(xl), "=r"(xh) : "r"(h_lo), "r"(clo));
        uint32_t t;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(h_hi), "r"(clo), "r"(P));
        xh += t;
        if (FIRST) { rlo[i] = xl; rhi[i] = xh; continue; }      // MaxUint64 (khf.go:18-32) loses to anything
        const uint32_t v = xh | kl, b = rhi[i] | kl;
        tie |= (v == b);
        asm("{\n .reg .pred p;\n setp.lt.u32 p, %2, %3;\n @p mad.lo.u32 %0, %4, %6, 0;\n @p mad.lo.u32 %1, %5, %6, 0;\n}"
            : "+r"(rlo[i]), "+r"(rhi[i])
            : "r"(v), "r"(b), "r"(xl), "r"(xh), "r"(M.one));
    }
    if (!FIRST && tie) khf_update_exact<S>(h, M, rlo, rhi);
}

// p: read bases (shared or global), len >= k
template <int S>
__device__ __forceinline__ void khf_sketch(const uint8_t* __restrict__ p, uint32_t len, uint32_t k, const SeedTabs& T,
                                           const MultTable& M, uint64_t (&sk)[S]) {
    uint64_t fh = 0, rh = 0;
    for (uint32_t j = 0; j < k; j++) {
        fh = ((fh << 1) | (fh >> 63)) ^ T.in[p[j]];
        rh = ((rh << 1) | (rh >> 63)) ^ T.c0[p[k - 1 - j] & 7u];
    }
    const uint32_t n = len - k + 1;
#if GROOT_KHF_VARIANT == 6
    uint32_t rlo[S], rhi[S];
    uint64_t s0;
    khf_update_raw<S, true>(rh < fh ? rh : fh, M, s0, rlo, rhi);
    for (uint32_t j = 1; j < n; j++) {
        const uint32_t bo = p[j - 1], bi = p[j + k - 1];
        fh = ((fh << 1) | (fh >> 63)) ^ T.out[bo] ^ T.in[bi];
        rh = ((rh >> 1) | (rh << 63)) ^ T.cout[bo & 7u] ^ T.cin[bi & 7u];
        khf_update_raw<S, false>(rh < fh ? rh : fh, M, s0, rlo, rhi);
    }
    sk[0] = s0;
#pragma unroll
    for (int i = 1; i < S; i++) sk[i] = khf_finish(rlo[i], rhi[i]);
#else
#pragma unroll
    for (int i = 0; i < S; i++) sk[i] = ~0ULL;
    uint32_t lw[S];
#pragma unroll
    for (int i = 0; i < S; i++) {
#if GROOT_KHF_VARIANT == 5
        asm volatile("mov.u32 %0, %1;" : "=r"(lw[i]) : "r"(M.low[i & 31]));   // volatile: stays a register, is not re-read from the constant bank per use
#else
        lw[i] = 0;
#endif
    }
    khf_update<S>(rh < fh ? rh : fh, M, lw, sk);
    for (uint32_t j = 1; j < n; j++) {
        const uint32_t bo = p[j - 1], bi = p[j + k - 1];
        fh = ((fh << 1) | (fh >> 63)) ^ T.out[bo] ^ T.in[bi];
        rh = ((rh >> 1) | (rh << 63)) ^ T.cout[bo & 7u] ^ T.cin[bi & 7u];
        khf_update<S>(rh < fh ? rh : fh, M, lw, sk);
    }
#endif
}

// ---- the probe: LSH band lookup + containment check, all per thread ----------------------------
// Calls emit(window) for every window that passes, in ascending window id per band; a window reached
// through several bands is reported once (lshensemble de-duplicates per forest query).
template <int S, int MAXK, class Emit>
__device__ __forceinline__ void lsh_probe(const DevIndex& ix, const uint64_t (&sk)[S], LenParam lp, Emit emit) {
    constexpr int NB = S / MAXK;
    if (lp.eq_min > S || lp.K == 0) return;
    const uint32_t kmask = (1u << lp.K) - 1u;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        if (b >= lp.L) break;
        uint32_t key[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < MAXK; j++) key[j] = j < lp.K ? static_cast<uint32_t>(sk[b * MAXK + j]) : 0u;
        const LshTable tab = ix.tables[(lp.K - 1) * NB + b];
        uint32_t h = band_key_hash(key) & tab.mask;
        uint32_t start = 0, count = 0;
        while (true) {
            const uint4* sp = reinterpret_cast<const uint4*>(tab.slots + h);
            uint4 kq = __ldg(sp), rest = __ldg(sp + 1);
            if (rest.y == 0) break;  // empty
            if (kq.x == key[0] && kq.y == key[1] && kq.z == key[2] && kq.w == key[3]) { start = rest.x; count = rest.y; break; }
            h = (h + 1) & tab.mask;
        }
        for (uint32_t c = 0; c < count; c++) {
            const uint32_t w = __ldg(tab.wins + start + c);
            const uint64_t* ws = ix.sketches + static_cast<size_t>(w) * S;
            uint32_t eq = 0, low = 0;
#pragma unroll
            for (int i = 0; i < S; i++) {
                uint64_t v = __ldg(ws + i);
                eq += (v == sk[i]);
                low |= static_cast<uint32_t>(static_cast<uint32_t>(v) == static_cast<uint32_t>(sk[i])) << i;
            }
            bool dup = false;
#pragma unroll
            for (int b2 = 0; b2 < NB; b2++)
                if (b2 < b && ((low >> (b2 * MAXK)) & kmask) == kmask) dup = true;
            if (eq >= lp.eq_min && !dup) emit(w);
        }
    }
}

// Warp-cooperative form of lsh_probe for 32 reads at once (lane == read). Every lane looks up its own band
// slot; the buckets found are then POOLED: an inclusive scan of the bucket sizes turns them into one list of
// (owner lane, candidate) items which the 32 lanes verify 32 at a time, fetching the owner's sketch slots by
// shuffle. (With one thread walking its own bucket, a warp ran the candidate loop with ~2 active lanes: the
// ~11 candidates of a seeded read — neighbouring windows share their first K minima — against 0 for the
// other half of the reads; 35 % of the kernel's issue slots, profiles/r01_ncu_summary.md.)
// r = this lane's read (its stage slots are stage[r * HSTAGE ..]); returns this lane's hit count. Same hits as lsh_probe, in
// the same order (band, then window id ascending).
template <int S, int MAXK>
__device__ __forceinline__ uint32_t warp_probe(const DevIndex& ix, const uint64_t (&sk)[S], LenParam lp, bool valid, uint32_t lane,
                                               uint32_t* __restrict__ stage, uint32_t r) {
    constexpr int NB = S / MAXK;
    constexpr uint32_t FULL = 0xffffffffu;
    const bool probing = valid && lp.eq_min <= S && lp.K != 0;
    const uint32_t myL = probing ? lp.L : 0u;
    const bool use_full = probing && lp.eq_min == S && lp.L == 1 && ix.full.slots != nullptr;   // every slot has to match: look the whole sketch up (DevIndex::full)
    const uint32_t maxL = __reduce_max_sync(FULL, myL);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t nh = 0;
#pragma unroll 1
    for (uint32_t b = 0; b < maxL && b < NB; b++) {
        // (1) every lane: its band slot -> bucket [start, start + count)
        uint32_t start = 0, count = 0;
        if (b < myL) {
            uint32_t key[4] = {0, 0, 0, 0};
            if (use_full) {
                sketch_digest(sk, S, key);
            } else {
#pragma unroll
                for (int bb = 0; bb < NB; bb++) {
                    if (bb == static_cast<int>(b)) {
#pragma unroll
                        for (int j = 0; j < MAXK; j++) key[j] = j < lp.K ? static_cast<uint32_t>(sk[bb * MAXK + j]) : 0u;
                    }
                }
            }
            const LshTable tab = use_full ? ix.full : ix.tables[(lp.K - 1) * NB + b];
            uint32_t h = band_key_hash(key) & tab.mask;
            while (true) {
                const uint4* sp = reinterpret_cast<const uint4*>(tab.slots + h);
                const uint4 kq = __ldg(sp), rest = __ldg(sp + 1);
                if (rest.y == 0) break;  // empty
                if (kq.x == key[0] && kq.y == key[1] && kq.z == key[2] && kq.w == key[3]) { start = rest.x; count = rest.y; break; }
                h = (h + 1) & tab.mask;
            }
        }
        __syncwarp();
        // (2) pool the buckets
        uint32_t incl = count;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); if (lane >= static_cast<uint32_t>(d)) incl += v; }
        const uint32_t excl = incl - count;
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        // (3) verify 32 (owner, candidate) items per step
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t item = base + lane;
            const bool live = item < total;
            uint32_t o = 0;  // owner = first lane whose inclusive sum exceeds item
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t v = __shfl_sync(FULL, incl, (o + step - 1) & 31u);
                if (v <= item) o += step;
            }
            o &= 31u;
            const uint32_t o_excl = __shfl_sync(FULL, excl, o), o_incl = __shfl_sync(FULL, incl, o);
            const uint32_t o_start = __shfl_sync(FULL, start, o);
            const uint32_t o_K = __shfl_sync(FULL, static_cast<uint32_t>(lp.K), o), o_eqmin = __shfl_sync(FULL, static_cast<uint32_t>(lp.eq_min), o);
            const uint32_t o_nh = __shfl_sync(FULL, nh, o), o_r = __shfl_sync(FULL, r, o);
            const bool o_full = __shfl_sync(FULL, static_cast<uint32_t>(use_full), o) != 0;
            uint32_t w = 0;
            const uint64_t* ws = ix.sketches;
            if (live) {
                const uint32_t* bucket = o_full ? ix.full.wins : ix.tables[(o_K - 1) * NB + b].wins;
                w = __ldg(bucket + o_start + (item - o_excl));
                ws += static_cast<size_t>(w) * S;
            }
            uint32_t eq = 0, low = 0;
#pragma unroll
            for (int i = 0; i < S; i++) {
                const uint64_t mine = __shfl_sync(FULL, sk[i], o);
                const uint64_t v = live ? __ldg(ws + i) : 0ULL;
                eq += (v == mine);
                low |= static_cast<uint32_t>(static_cast<uint32_t>(v) == static_cast<uint32_t>(mine)) << i;
            }
            const uint32_t kmask = (1u << o_K) - 1u;
            bool dup = false;  // already reported through an earlier band (lshensemble de-duplicates per forest query)
#pragma unroll
            for (int b2 = 0; b2 < NB; b2++)
                if (b2 < static_cast<int>(b) && ((low >> (b2 * MAXK)) & kmask) == kmask) dup = true;
            const bool pass = live && eq >= o_eqmin && !dup;
            const uint32_t pass_mask = __ballot_sync(FULL, pass);
            // lanes of this step that hold items of owner x: [max(excl_x, base), min(incl_x, base + 32)) - base
            auto range_mask = [&](uint32_t ex, uint32_t in) {
                const uint32_t lo = (ex > base ? ex : base) - base;
                const uint32_t hi = (in < base + 32 ? in : base + 32);
                if (hi <= base + lo) return 0u;
                const uint32_t n = hi - base - lo;
                return (n >= 32 ? FULL : ((1u << n) - 1u)) << lo;
            };
            if (pass) {
                const uint32_t slot = o_nh + __popc(pass_mask & range_mask(o_excl, o_incl) & lt_mask);
                if (slot < HSTAGE) stage[static_cast<size_t>(o_r) * HSTAGE + slot] = w;
            }
            if (incl > base && excl < base + 32) nh += __popc(pass_mask & range_mask(excl, incl));
        }
    }
    return nh;
}

struct SeedArgs {
    const uint8_t* seq;        // read bases, readable 64 bytes past the end
    const uint32_t* off;       // [n+1]
    uint32_t n_reads;
    uint32_t max_len;          // len_params has max_len+1 entries
    const LenParam* len_params;
    uint32_t* n_hits;          // [n]
    uint32_t* stage;           // [n*HSTAGE]
    uint64_t* sketches_out;    // [n*S] or nullptr
    uint32_t* tile_counter;    // zeroed before launch
    int* error;                // [0]=code, [1]=read index
    uint32_t tile_bytes;       // shared-memory bytes per tile buffer (0 => read straight from global)
    uint32_t* queue;           // SEED_PRESCREEN: reads whose first band found a bucket (out); SEED_QUEUED: the reads to process (in)
    uint32_t* n_queue;         // device scalar
    // SEED_QUEUED with staged reads: the 2-bit copies of the seeded reads for the align kernels are written here, from the
    // bases the warp already holds in shared memory (layout: pack_reads_kernel); nw32 == 0: not wanted / done by pack_reads_kernel
    uint32_t* reads2;
    uint8_t* read_ok2;
    uint4* read_oh;
    uint32_t nw32;
};

// reverse the order of the sixteen 2-bit groups of a word
__device__ __forceinline__ uint32_t rev_pairs16(uint32_t x) {
    const uint32_t r = __brev(x);
    return ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
}

// One thread packs one read whose bases lie in shared memory: the same output as pack_reads_kernel (2-bit forward copy,
// reverse complement shifted to the end of its words, validity flag, one-hot 9-base prefixes of both strands).
__device__ __forceinline__ void pack_read_from_smem(const uint8_t* __restrict__ p, uint32_t len, uint32_t nw32, uint32_t* __restrict__ out,
                                                    uint8_t* __restrict__ ok_out, uint4* __restrict__ oh_out) {
    bool ok = len <= nw32 * 16u;
    if (ok) {
        for (uint32_t j = 0; j < nw32; j++) {
            uint32_t acc = 0;
            const uint32_t b0 = 16u * j;
            const uint32_t nb = b0 < len ? (len - b0 < 16u ? len - b0 : 16u) : 0u;
#pragma unroll
            for (uint32_t t = 0; t < 16u; t++) {
                if (t < nb) {
                    const uint32_t b = p[b0 + t];
                    const uint32_t c = (b >> 1) & 3u;
                    ok = ok && b == ((0x47544341u >> (8u * c)) & 0xffu);     // 'A' 'C' 'T' 'G' by code: anything else (lower case, N) goes byte-wise
                    acc |= c << (2u * t);
                }
            }
            out[j] = acc;
            out[nw32 + (nw32 - 1u - j)] = rev_pairs16(acc) ^ 0xAAAAAAAAu;
        }
    }
    uint64_t f = 0, c = 0;
    if (ok) {
        for (uint32_t i = 0; i < 9u && i < len; i++) {
            f |= static_cast<uint64_t>(1u << ((p[i] >> 1) & 3u)) << (4 * i);
            c |= static_cast<uint64_t>(1u << (((p[len - 1u - i] >> 1) & 3u) ^ 2u)) << (4 * i);
        }
    }
    *ok_out = ok ? 1 : 0;
    *oh_out = make_uint4(static_cast<uint32_t>(f), static_cast<uint32_t>(f >> 4), static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 4));
}

// seed_kernel modes. SEED_FULL: every read gets its full sketch and probe in one pass. The two-pass form exploits
// that most reads of a metagenome seed nowhere: SEED_PRESCREEN computes only the first band's MAXK sketch slots (a
// fifth of the hashing at S = 21), looks the band key up and queues the reads that found a bucket (possible when
// the optimiser probes a single band, L == 1: the reference default); SEED_QUEUED then runs the full sketch + pooled
// verification for the queued reads only. Reads that are not queued have no candidates, hence n_hits = 0 — the same
// result as SEED_FULL.
enum { SEED_FULL = 0, SEED_PRESCREEN = 1, SEED_QUEUED = 2 };

__device__ __forceinline__ void set_error(int* err, int code, uint32_t read) {
    if (atomicCAS(err, 0, code) == 0) err[1] = static_cast<int>(read);
}

// TMA bulk copy helpers (cp.async.bulk + mbarrier); SASS: UBLKCP / SYNCS
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(addr),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))),
                 "l"(gmem_src), "r"(bytes), "r"(static_cast<unsigned>(__cvta_generic_to_shared(bar)))
                 : "memory");
}

// Persistent kernel: every WARP pulls 32-read tiles from an atomic ticket and owns two shared-memory tile
// buffers + two mbarriers; tile t+1 streams in (TMA bulk copy issued by lane 0) while tile t is being hashed.
// Warps never synchronise with each other: a warp waiting on the L2 for its probe does not hold back the
// other warps of the block (an earlier block-wide version lost ~30 % of its issue slots at __syncthreads).
constexpr int kTileReads = 32;

template <int S, int MAXK, int MODE>
__global__ void __launch_bounds__(kSeedThreads, MODE == SEED_PRESCREEN ? 8 : GROOT_SEED_MIN_BLOCKS) seed_kernel(DevIndex ix, SeedArgs a, MultTable M) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SeedTabs* T = reinterpret_cast<SeedTabs*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(SeedTabs));  // 2 mbarriers per warp
    uint8_t* bufs_all = smem_raw + sizeof(SeedTabs) + 64;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_units = MODE == SEED_QUEUED ? *a.n_queue : a.n_reads;   // reads, or queue entries
    const uint32_t n_tiles = (n_units + kTileReads - 1) / kTileReads;
    const bool staged = MODE != SEED_QUEUED && a.tile_bytes != 0;       // queued reads are scattered: no tile to bulk-copy
    uint8_t* bufs = bufs_all + static_cast<size_t>(warp) * 2 * a.tile_bytes;
    uint64_t* bar = bars + warp * 2;

    build_seed_tabs(T, ix.k);
    if (lane == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // make the init visible to the async (TMA) proxy
    }
    __syncthreads();  // the only block-wide barrier: seed tables + mbarrier init

    auto issue = [&](uint32_t tile, int buf) {  // lane 0 only
        const uint32_t r0 = tile * kTileReads, r1 = min(a.n_reads, r0 + kTileReads);
        const uint32_t b0 = a.off[r0] & ~15u, b1 = (a.off[r1] + 15u) & ~15u;
        const uint32_t bytes = b1 - b0;
        if (bytes <= a.tile_bytes && bytes > 0) {
            mbar_expect_tx(&bar[buf], bytes);
            bulk_g2s(bufs + static_cast<size_t>(buf) * a.tile_bytes, a.seq + b0, bytes, &bar[buf]);
        } else {
            mbar_expect_tx(&bar[buf], 0);  // oversize tile: threads read global memory directly
        }
    };
    auto next_tile = [&]() {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(a.tile_counter, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    unsigned phase[2] = {0, 0};
    int cur = 0;
    uint32_t tile = next_tile();
    if (staged && lane == 0 && tile < n_tiles) issue(tile, 0);
    while (tile < n_tiles) {
        const uint32_t tnext = next_tile();
        if (staged && lane == 0 && tnext < n_tiles) issue(tnext, cur ^ 1);  // the other buffer was released by the __syncwarp below
        const uint32_t r0 = tile * kTileReads, r1 = min(n_units, r0 + kTileReads);
        uint32_t r = r0 + lane;
        const bool live = r < n_units;
        if (MODE == SEED_QUEUED) r = live ? a.queue[r] : 0u;
        bool in_smem = false;
        uint32_t b0 = 0;
        if (staged) {
            mbar_wait(&bar[cur], phase[cur]);
            phase[cur] ^= 1u;
            b0 = a.off[r0] & ~15u;
            const uint32_t b1 = (a.off[r1] + 15u) & ~15u;
            in_smem = (b1 - b0) <= a.tile_bytes;
        }
        // ---- sketch: one thread per read, registers only ----
        uint64_t sk[S];
        LenParam lp{0, 0, 0xffff};
        bool valid = false;
        const uint32_t o = live ? a.off[r] : 0u, len = live ? a.off[r + 1] - o : 0u;
        const uint8_t* qread = nullptr;
        if (MODE == SEED_QUEUED && a.tile_bytes != 0) {
            // queued reads are scattered over the batch: the warp copies its 32 reads into its shared-memory buffer, one
            // read per step with the lanes striding over aligned 32-bit words, at a word stride that is odd (the per-thread
            // byte reads of the hash loop then fall into 32 different banks)
            const uint32_t stride = a.tile_bytes / kTileReads;
            const uint32_t max_nw = (a.max_len + 6u) >> 2;
            for (uint32_t wbase = 0; wbase < max_nw; wbase += 32) {
                const uint32_t wj = wbase + lane;
#pragma unroll
                for (uint32_t i0 = 0; i0 < kTileReads; i0 += 8) {   // eight reads' loads in flight before the first store
                    uint32_t v[8];
                    bool ok[8];
#pragma unroll
                    for (uint32_t j = 0; j < 8; j++) {
                        const uint32_t oi = __shfl_sync(0xffffffffu, o, i0 + j), li = __shfl_sync(0xffffffffu, len, i0 + j);
                        const uint32_t nw = (li == 0 || li > a.max_len) ? 0u : ((oi & 3u) + li + 3u) >> 2;
                        ok[j] = wj < nw;
                        v[j] = ok[j] ? __ldg(reinterpret_cast<const uint32_t*>(a.seq + (oi & ~3u)) + wj) : 0u;
                    }
#pragma unroll
                    for (uint32_t j = 0; j < 8; j++)
                        if (ok[j]) reinterpret_cast<uint32_t*>(bufs + (i0 + j) * stride)[wj] = v[j];
                }
            }
            __syncwarp();
            qread = bufs + lane * stride + (o & 3u);
        }
        if (live) {
            if (len < ix.k || len > a.max_len) {
                set_error(a.error, len < ix.k ? -5 : -7, r);  // GROOTGPU_ERR_SHORT_READ / _CAPACITY
            } else {
                if (qread) khf_sketch<S>(qread, len, ix.k, *T, M, sk);                                                              // LDS path, queued reads
                else if (in_smem) khf_sketch<S>(bufs + static_cast<size_t>(cur) * a.tile_bytes + (o - b0), len, ix.k, *T, M, sk);   // LDS path
                else khf_sketch<S>(a.seq + o, len, ix.k, *T, M, sk);                                                          // LDG path
                lp = a.len_params[len];
                valid = true;
                if (a.sketches_out) {
#pragma unroll
                    for (int i = 0; i < S; i++) a.sketches_out[static_cast<size_t>(r) * S + i] = sk[i];
                }
            }
        }
        __syncwarp();
        if (MODE == SEED_PRESCREEN) {
            // ---- first band only: does the key have a bucket? (S == MAXK slots were hashed) ----
            bool found = false;
            if (valid && lp.eq_min <= ix.S && lp.K != 0 && lp.K <= S) {
                uint32_t key[4] = {0, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < MAXK && j < S; j++) key[j] = j < lp.K ? static_cast<uint32_t>(sk[j]) : 0u;
                const LshTable tab = ix.tables[(lp.K - 1) * ix.n_bands];
                uint32_t h = band_key_hash(key) & tab.mask;
                while (true) {
                    const uint4* sp = reinterpret_cast<const uint4*>(tab.slots + h);
                    const uint4 kq = __ldg(sp), rest = __ldg(sp + 1);
                    if (rest.y == 0) break;  // empty
                    if (kq.x == key[0] && kq.y == key[1] && kq.z == key[2] && kq.w == key[3]) { found = true; break; }
                    h = (h + 1) & tab.mask;
                }
            }
            __syncwarp();
            const uint32_t ball = __ballot_sync(0xffffffffu, found);
            uint32_t base = 0;
            if (lane == 0 && ball) base = atomicAdd(a.n_queue, static_cast<uint32_t>(__popc(ball)));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (found) a.queue[base + __popc(ball & ((1u << lane) - 1u))] = r;
            if (live) a.n_hits[r] = 0;
        } else {
            // ---- probe + containment check: the warp's 32 reads pool their candidates (warp_probe) ----
            const uint32_t nh = warp_probe<S, MAXK>(ix, sk, lp, valid, lane, a.stage, r);
            if (live) a.n_hits[r] = nh;
            if (MODE == SEED_QUEUED && a.nw32 != 0 && qread != nullptr && valid && nh != 0)   // a seeded read: its 2-bit copies for the align kernels, while the bases are at hand
                pack_read_from_smem(qread, len, a.nw32, a.reads2 + static_cast<size_t>(r) * 2u * a.nw32, a.read_ok2 + r, a.read_oh + r);
        }
        __syncwarp();  // every lane is done with buffer `cur`
        cur ^= 1;
        tile = tnext;
    }
}

// Sketch-only kernel (window sketching at index time, grootgpu_sketch_batch): one thread per sequence,
// bases read straight from global memory (index windows overlap, so there is no tile to stage).
template <int S>
__global__ void __launch_bounds__(kSeedThreads) sketch_kernel(const uint8_t* __restrict__ seq, const uint64_t* __restrict__ off,
                                                              const uint32_t* __restrict__ lens, uint32_t fixed_len, uint32_t n,
                                                              uint32_t k, MultTable M, uint64_t* __restrict__ out, int* error) {
    __shared__ SeedTabs T;
    build_seed_tabs(&T, k);
    __syncthreads();
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t len = lens ? lens[r] : fixed_len;
        if (len < k) { set_error(error, -5, r); continue; }
        uint64_t sk[S];
        khf_sketch<S>(seq + off[r], len, k, T, M, sk);
#pragma unroll
        for (int i = 0; i < S; i++) out[static_cast<size_t>(r) * S + i] = sk[i];
    }
}

// Exact-size fill: copies the staged hits of every read to hits[hit_off[r]..], redoing the probe for the
// rare read with more than HSTAGE hits, sorts each read's hits ascending (== graph, Node, OffSet order),
// marks (read, graph) segment starts and accumulates theBoss counters.
struct FillArgs {
    const uint8_t* seq;
    const uint32_t* off;
    uint32_t n_reads;
    const LenParam* len_params;
    const uint32_t* n_hits;
    const uint32_t* hit_off;  // exclusive scan of n_hits, [n+1]
    const uint32_t* stage;
    uint32_t* hits;
    uint32_t* hit_read;
    uint8_t* seg_flag;
    unsigned long long* counters;  // [0]=mapped reads, [1]=multimapped reads
    uint32_t* n_overflow;          // device scalar: reads with more than HSTAGE hits, queued by the lean kernel, taken by the refill kernel
    uint32_t* overflow_q;          // [n] those reads (a dense queue: the refill kernel runs with full warps instead of scanning all reads)
    uint32_t* reads2;              // [n * 2 * nw32 + 1] packed copies of the seeded reads (pack_reads_kernel), or nullptr
    uint8_t* read_ok2;             // [n] 1 when reads2 holds the read
    uint4* read_oh;                // [n] one-hot prefixes for the screen: x/y = bases [0,8) / [1,9) of the read, z/w = of its reverse complement
    uint32_t nw32;                 // words per orientation (16 bases each); 0 = packing off
};

// 2-bit copies of the seeded reads for the align kernels, per read r at reads2 + r * 2 * nw32:
//   words [0, nw32)        the read: base i in word i >> 4 at bits 2 * (i & 15), code pack_base2
//   words [nw32, 2 * nw32) its reverse complement shifted up by pad = 16 * nw32 - len bases (base i of the reverse
//                          complement sits at position i + pad)
// read_ok2[r] = 0 when the read does not fit or holds anything but upper-case ACGT — such reads take the byte-wise
// path, which reproduces the reference's handling of 'N' and of bytes RevComplement cannot map (seqio.go:17-23,120-133).
// A GROUP of 8 lanes per read, lane j of the group packing output words j, j + 8, ... (16 bases = four aligned 32-bit
// loads + realignment each; seq must be readable a few bytes past the read, the API asks for 64), so a warp's loads
// run along the reads instead of 32 lanes striding through 32 different reads.
__global__ void __launch_bounds__(256) pack_reads_kernel(FillArgs a) {
    const uint32_t lane = threadIdx.x & 31, gl = lane & 7u, grp = lane >> 3;
    const uint32_t gmask = 0xffu << (grp * 8);
    const uint32_t g0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, gstride = (gridDim.x * blockDim.x) >> 3;
    const uint32_t n_iter = (a.n_reads + gstride - 1) / gstride;          // same trip count for every group of a warp
    for (uint32_t it = 0; it < n_iter; it++) {
        const uint32_t r = g0 + it * gstride;
        const bool live = r < a.n_reads && a.n_hits[r] != 0;
        bool ok = true;
        if (live) {
            const uint32_t o = a.off[r], len = a.off[r + 1] - o;
            uint32_t* out = a.reads2 + static_cast<size_t>(r) * 2u * a.nw32;
            ok = len <= a.nw32 * 16u;
            for (uint32_t j = gl; j < a.nw32 && ok; j += 8) {
                const uint32_t b0 = o + 16u * j;                          // first byte of this output word
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(a.seq + (b0 & ~3u));
                const uint32_t sh = (b0 & 3u) * 8u;
                uint32_t acc = 0;
                if (16u * j < len) {
                    uint32_t prev = __ldg(wp);
#pragma unroll
                    for (uint32_t t = 0; t < 4; t++) {
                        const uint32_t bi = 16u * j + 4u * t;
                        if (bi < len) {
                            const uint32_t next = __ldg(wp + t + 1u);
                            const uint32_t w = __funnelshift_r(prev, next, sh);
                            prev = next;
                            const uint32_t nb = len - bi;
                            const uint32_t bytemask = nb >= 4u ? 0xffffffffu : ((1u << (8u * nb)) - 1u);
                            uint32_t c = (w >> 1) & 0x03030303u;
                            const uint32_t m0 = c & 0x01010101u, m1 = (c >> 1) & 0x01010101u;
                            const uint32_t expect = 0x41414141u ^ (m0 << 1) ^ ((m0 & m1) << 2) ^ ((m1 & ~m0) * 0x15u);   // A 0x41, C 0x43, T 0x54, G 0x47
                            ok = ok && ((expect ^ w) & bytemask) == 0u;
                            c &= bytemask;
                            acc |= ((c | (c >> 6) | (c >> 12) | (c >> 18)) & 0xffu) << (8u * t);
                        }
                    }
                }
                out[j] = acc;
                out[a.nw32 + (a.nw32 - 1u - j)] = rev_pairs16(acc) ^ 0xAAAAAAAAu;
            }
        }
        const uint32_t bad = __ballot_sync(0xffffffffu, live && !ok);
        __syncwarp();                                                   // the group's stores are visible to its reads below
        // what align_screen compares against the allele sets (prefix_pass): the first 9 bases of the read and of its
        // reverse complement one-hot, 4 bits per base (bit pack_base2(b)); lane i of the group converts base i (lane 0
        // also base 8) from the packed words just written, an OR over the group assembles them. Only meaningful for
        // reads of upper-case ACGT (read_ok2): the others build theirs byte-wise in the screen.
        uint64_t f = 0, c = 0;
        if (live && !(bad & gmask)) {
            const uint32_t len = a.off[r + 1] - a.off[r];
            const uint32_t* out = a.reads2 + static_cast<size_t>(r) * 2u * a.nw32;
            const uint32_t pad = a.nw32 * 16u - len, wi = pad >> 4;
            const uint32_t f2 = out[0];
            const uint64_t c2 = (static_cast<uint64_t>(wi + 1 < a.nw32 ? out[a.nw32 + wi + 1] : 0u) << 32 | out[a.nw32 + wi]) >> ((pad & 15u) * 2u);
            for (uint32_t i = gl; i < 9 && i < len; i += 8) {
                f |= static_cast<uint64_t>(1u << ((f2 >> (2 * i)) & 3u)) << (4 * i);
                c |= static_cast<uint64_t>(1u << (static_cast<uint32_t>(c2 >> (2 * i)) & 3u)) << (4 * i);
            }
        }
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) { f |= __shfl_xor_sync(0xffffffffu, f, d); c |= __shfl_xor_sync(0xffffffffu, c, d); }
        if (live && gl == 0) {
            a.read_ok2[r] = (bad & gmask) ? 0 : 1;
            a.read_oh[r] = make_uint4(static_cast<uint32_t>(f), static_cast<uint32_t>(f >> 4), static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 4));
        }
    }
}

// sorts a read's hits ascending (== graph, Node, OffSet order), marks (read, graph) segment starts; returns the number of graphs
__device__ __forceinline__ uint32_t finish_read_hits(const DevIndex& ix, const FillArgs& a, uint32_t r, uint32_t nh, uint32_t base) {
    for (uint32_t i = 1; i < nh; i++) {  // insertion sort, tiny
        uint32_t v = a.hits[base + i];
        uint32_t j = i;
        while (j > 0 && a.hits[base + j - 1] > v) { a.hits[base + j] = a.hits[base + j - 1]; j--; }
        a.hits[base + j] = v;
    }
    uint32_t segs = 0, prev_g = 0xffffffffu;
    for (uint32_t i = 0; i < nh; i++) {
        uint32_t g = ix.wins[a.hits[base + i]].graph;
        a.hit_read[base + i] = r;
        a.seg_flag[base + i] = (g != prev_g);
        segs += (g != prev_g);
        prev_g = g;
    }
    return segs;
}

// REFILL = false: the reads whose hits were all staged (the lean, common kernel). REFILL = true: only the rare reads
// with more than HSTAGE hits, whose probe is redone (kept apart so that its sketch registers do not set the
// occupancy of the common case).
template <int S, int MAXK, bool REFILL>
__global__ void __launch_bounds__(kSeedThreads) fill_kernel(DevIndex ix, FillArgs a, MultTable M) {
    __shared__ SeedTabs T;
    uint32_t n_units = a.n_reads;
    if (REFILL) {
        n_units = *a.n_overflow;
        if (n_units == 0) return;                       // the usual case: nothing to redo, the whole grid leaves at once
        build_seed_tabs(&T, ix.k); __syncthreads();
    }
    unsigned mapped = 0, multi = 0;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        const uint32_t r = REFILL ? a.overflow_q[u] : u;
        const uint32_t nh = a.n_hits[r];
        if (nh == 0) continue;
        if (!REFILL && nh > HSTAGE) { a.overflow_q[atomicAdd(a.n_overflow, 1u)] = r; continue; }
        const uint32_t base = a.hit_off[r];
        if (!REFILL) {
            for (uint32_t i = 0; i < nh; i++) a.hits[base + i] = a.stage[static_cast<size_t>(r) * HSTAGE + i];
        } else {
            const uint32_t o = a.off[r], len = a.off[r + 1] - o;
            uint64_t sk[S];
            khf_sketch<S>(a.seq + o, len, ix.k, T, M, sk);
            uint32_t c = 0;
            lsh_probe<S, MAXK>(ix, sk, a.len_params[len], [&](uint32_t w) { if (c < nh) a.hits[base + c] = w; c++; });
        }
        const uint32_t segs = finish_read_hits(ix, a, r, nh, base);
        mapped++;
        multi += segs > 1;
    }
    mapped = __reduce_add_sync(0xffffffffu, mapped);
    multi = __reduce_add_sync(0xffffffffu, multi);
    if ((threadIdx.x & 31) == 0) {
        if (mapped) atomicAdd(&a.counters[0], static_cast<unsigned long long>(mapped));
        if (multi) atomicAdd(&a.counters[1], static_cast<unsigned long long>(multi));
    }
}

// ---- any sketch size, any maxK -----------------------------------------------------------------------------------
// `groot index` accepts any -s / -y (cmd/index.go:48-49). The kernels above are compiled for the sketch sizes in common
// use with maxK == 4 (sketch and band keys in registers); every other combination runs here: the same arithmetic with
// run-time S and maxK, the sketch in thread-local memory, the probe per thread. Slower per read, bit-identical results.
// Band tables are keyed on the first min(K, 4) hashes of a band (LshSlot holds four); the remaining hashes of a longer
// prefix are compared against the candidate's sketch, so a bucket is exactly lshensemble's prefix match.
constexpr int kMaxSketch = 256;

__device__ inline void khf_sketch_generic(const uint8_t* __restrict__ p, uint32_t len, uint32_t k, const SeedTabs& T, uint32_t S, uint64_t* __restrict__ sk) {
    const uint64_t C = static_cast<uint64_t>(k) * GROOT_MULTI_SEED;
    for (uint32_t i = 0; i < S; i++) sk[i] = ~0ULL;
    uint64_t fh = 0, rh = 0;
    for (uint32_t j = 0; j < k; j++) {
        fh = ((fh << 1) | (fh >> 63)) ^ T.in[p[j]];
        rh = ((rh << 1) | (rh >> 63)) ^ T.c0[p[k - 1 - j] & 7u];
    }
    const uint32_t n = len - k + 1;
    for (uint32_t j = 0; j < n; j++) {
        if (j) {
            const uint32_t bo = p[j - 1], bi = p[j + k - 1];
            fh = ((fh << 1) | (fh >> 63)) ^ T.out[bo] ^ T.in[bi];
            rh = ((rh >> 1) | (rh << 63)) ^ T.cout[bo & 7u] ^ T.cin[bi & 7u];
        }
        const uint64_t h = rh < fh ? rh : fh;
        if (h < sk[0]) sk[0] = h;
        for (uint32_t i = 1; i < S; i++) {                       // khf.go:44-53 / nthash MultiHash
            uint64_t x = h * (static_cast<uint64_t>(i) ^ C);
            x ^= x >> GROOT_MULTI_SHIFT;
            if (x < sk[i]) sk[i] = x;
        }
    }
}

template <class Emit>
__device__ inline void lsh_probe_generic(const DevIndex& ix, const uint64_t* __restrict__ sk, LenParam lp, Emit emit) {
    const uint32_t S = ix.S, maxk = ix.max_k, NB = ix.n_bands, K = lp.K;
    if (lp.eq_min > S || K == 0) return;
    auto band_match = [&](const uint64_t* ws, uint32_t b, uint32_t from) {   // hashes [from, K) of band b equal (low 32 bits: lshensemble's hash key)
        for (uint32_t j = from; j < K; j++)
            if (static_cast<uint32_t>(ws[b * maxk + j]) != static_cast<uint32_t>(sk[b * maxk + j])) return false;
        return true;
    };
    for (uint32_t b = 0; b < NB && b < lp.L; b++) {
        uint32_t key[4] = {0, 0, 0, 0};
        for (uint32_t j = 0; j < 4 && j < K; j++) key[j] = static_cast<uint32_t>(sk[b * maxk + j]);
        const LshTable tab = ix.tables[(K - 1) * NB + b];
        uint32_t h = band_key_hash(key) & tab.mask;
        uint32_t start = 0, count = 0;
        while (true) {
            const uint4* sp = reinterpret_cast<const uint4*>(tab.slots + h);
            const uint4 kq = __ldg(sp), rest = __ldg(sp + 1);
            if (rest.y == 0) break;  // empty
            if (kq.x == key[0] && kq.y == key[1] && kq.z == key[2] && kq.w == key[3]) { start = rest.x; count = rest.y; break; }
            h = (h + 1) & tab.mask;
        }
        for (uint32_t c = 0; c < count; c++) {
            const uint32_t w = __ldg(tab.wins + start + c);
            const uint64_t* ws = ix.sketches + static_cast<size_t>(w) * S;
            if (K > 4 && !band_match(ws, b, 4)) continue;       // the table groups by the first four hashes only
            uint32_t eq = 0;
            for (uint32_t i = 0; i < S; i++) eq += (__ldg(ws + i) == sk[i]);
            bool dup = false;                                    // already reported through an earlier band
            for (uint32_t b2 = 0; b2 < b && !dup; b2++) dup = band_match(ws, b2, 0);
            if (eq >= lp.eq_min && !dup) emit(w);
        }
    }
}

// one thread per read: sketch + probe, hits staged (<= HSTAGE) and counted, like seed_kernel<S, MAXK, SEED_FULL>
__global__ void __launch_bounds__(kSeedThreads) seed_generic_kernel(DevIndex ix, SeedArgs a) {
    __shared__ SeedTabs T;
    build_seed_tabs(&T, ix.k);
    __syncthreads();
    uint64_t sk[kMaxSketch];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += gridDim.x * blockDim.x) {
        const uint32_t o = a.off[r], len = a.off[r + 1] - o;
        if (len < ix.k || len > a.max_len) { set_error(a.error, len < ix.k ? -5 : -7, r); a.n_hits[r] = 0; continue; }
        khf_sketch_generic(a.seq + o, len, ix.k, T, ix.S, sk);
        if (a.sketches_out)
            for (uint32_t i = 0; i < ix.S; i++) a.sketches_out[static_cast<size_t>(r) * ix.S + i] = sk[i];
        uint32_t nh = 0;
        lsh_probe_generic(ix, sk, a.len_params[len], [&](uint32_t w) { if (nh < HSTAGE) a.stage[static_cast<size_t>(r) * HSTAGE + nh] = w; nh++; });
        a.n_hits[r] = nh;
    }
}

// the reads with more than HSTAGE hits: probe redone with the hits written at their final place
__global__ void __launch_bounds__(kSeedThreads) fill_refill_generic_kernel(DevIndex ix, FillArgs a) {
    __shared__ SeedTabs T;
    if (*a.n_overflow == 0) return;
    build_seed_tabs(&T, ix.k);
    __syncthreads();
    uint64_t sk[kMaxSketch];
    unsigned mapped = 0, multi = 0;
    const uint32_t n_units = *a.n_overflow;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        const uint32_t r = a.overflow_q[u];
        const uint32_t nh = a.n_hits[r];
        const uint32_t base = a.hit_off[r];
        const uint32_t o = a.off[r], len = a.off[r + 1] - o;
        khf_sketch_generic(a.seq + o, len, ix.k, T, ix.S, sk);
        uint32_t c = 0;
        lsh_probe_generic(ix, sk, a.len_params[len], [&](uint32_t w) { if (c < nh) a.hits[base + c] = w; c++; });
        const uint32_t segs = finish_read_hits(ix, a, r, nh, base);
        mapped++;
        multi += segs > 1;
    }
    mapped = __reduce_add_sync(0xffffffffu, mapped);
    multi = __reduce_add_sync(0xffffffffu, multi);
    if ((threadIdx.x & 31) == 0) {
        if (mapped) atomicAdd(&a.counters[0], static_cast<unsigned long long>(mapped));
        if (multi) atomicAdd(&a.counters[1], static_cast<unsigned long long>(multi));
    }
}

// window sketching / grootgpu_sketch_batch for any sketch size
__global__ void __launch_bounds__(kSeedThreads) sketch_generic_kernel(const uint8_t* __restrict__ seq, const uint64_t* __restrict__ off,
                                                                      const uint32_t* __restrict__ lens, uint32_t fixed_len, uint32_t n,
                                                                      uint32_t k, uint32_t S, uint64_t* __restrict__ out, int* error) {
    __shared__ SeedTabs T;
    build_seed_tabs(&T, k);
    __syncthreads();
    uint64_t sk[kMaxSketch];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t len = lens ? lens[r] : fixed_len;
        if (len < k) { set_error(error, -5, r); continue; }
        khf_sketch_generic(seq + off[r], len, k, T, S, sk);
        for (uint32_t i = 0; i < S; i++) out[static_cast<size_t>(r) * S + i] = sk[i];
    }
}

}  // namespace groot
