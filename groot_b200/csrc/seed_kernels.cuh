// K1+K2: per-read KHF MinHash sketch (ntHash v1 rolling canonical hash + S multi-hashes, per-slot min)
// fused with the LSH Ensemble probe and the full containment check.
//
// Replaces, per read:
//   Sequence.RunMinHash -> KHFsketch.AddSequence -> nthash MultiHash   (src/seqio/seqio.go:40-68, src/minhash/khf.go:35-56)
//   ContainmentIndex.Query -> lshensemble Query + Containment > t      (src/lshe/lshe.go:153-175)
//   the sketch worker loop of theBoss.mapReads                         (src/pipeline/boss.go:145-201)
//
// Mapping: ONE THREAD PER READ. The S running minima (2*S registers), the two rolling hashes and the
// probe all stay in that thread's registers: no shuffles, no idle lanes (a warp-per-read layout would
// waste 26 of 96 lane-slots on 70 k-mers), and the rolling hash stays a 1-step roll. Read bytes come
// in through a TMA bulk copy (cp.async.bulk, global -> shared, mbarrier completion) of each warp's
// 32-read tile, double buffered; the 100-byte read stride is co-prime with the 32 banks so the
// per-thread byte reads are conflict free. The kernel is INT-ALU bound (~18k integer ops per 100 bp read versus
// ~320 algorithmic bytes), see DESIGN.md "Rooflines".
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
#define GROOT_KHF_VARIANT 0
#endif
template <int S>
__device__ __forceinline__ void khf_update(uint64_t h, const MultTable& M, uint64_t (&sk)[S]) {
    sk[0] = h < sk[0] ? h : sk[0];
#if GROOT_KHF_VARIANT == 0
#pragma unroll
    for (int i = 1; i < S; i++) {
        uint64_t x = h * M.c[i];
        x ^= x >> GROOT_MULTI_SHIFT;
        sk[i] = x < sk[i] ? x : sk[i];
    }
#else
    // x_i = h*c0 + h*low_i  (exact mod 2^64, see MultTable)
    const uint32_t h_lo = static_cast<uint32_t>(h), h_hi = static_cast<uint32_t>(h >> 32);
    const uint64_t A = h * M.c0;
#pragma unroll
    for (int i = 1; i < S; i++) {
        uint32_t x_lo, x_hi;
        asm("{\n .reg .u64 w;\n mad.wide.u32 w, %2, %3, %4;\n mov.b64 {%0, %1}, w;\n}" : "=r"(x_lo), "=r"(x_hi) : "r"(h_lo), "r"(M.low[i]), "l"(A));
        asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x_hi) : "r"(h_hi), "r"(M.low[i]));
        uint64_t x;
#if GROOT_KHF_VARIANT == 1
        x = (static_cast<uint64_t>(x_hi) << 32) | x_lo;
        x ^= x >> GROOT_MULTI_SHIFT;
#else
        // x >> 27 through the FMA pipe: (x_lo >> 27) = mulhi(x_lo, 32); (x_hi << 5) = x_hi * 32; (x_hi >> 27) = mulhi(x_hi, 32)
        const uint32_t a = __umulhi(x_lo, M.m32);
        uint32_t t_lo;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t_lo) : "r"(x_hi), "r"(M.m32), "r"(a));
        const uint32_t t_hi = __umulhi(x_hi, M.m32);
        x = (static_cast<uint64_t>(x_hi ^ t_hi) << 32) | (x_lo ^ t_lo);
#endif
        sk[i] = x < sk[i] ? x : sk[i];
    }
#endif
}

// p: read bases (shared or global), len >= k
template <int S>
__device__ __forceinline__ void khf_sketch(const uint8_t* __restrict__ p, uint32_t len, uint32_t k, const SeedTabs& T,
                                           const MultTable& M, uint64_t (&sk)[S]) {
#pragma unroll
    for (int i = 0; i < S; i++) sk[i] = ~0ULL;
    uint64_t fh = 0, rh = 0;
    for (uint32_t j = 0; j < k; j++) {
        fh = ((fh << 1) | (fh >> 63)) ^ T.in[p[j]];
        rh = ((rh << 1) | (rh >> 63)) ^ T.c0[p[k - 1 - j] & 7u];
    }
    khf_update<S>(rh < fh ? rh : fh, M, sk);
    const uint32_t n = len - k + 1;
    for (uint32_t j = 1; j < n; j++) {
        const uint32_t bo = p[j - 1], bi = p[j + k - 1];
        fh = ((fh << 1) | (fh >> 63)) ^ T.out[bo] ^ T.in[bi];
        rh = ((rh >> 1) | (rh << 63)) ^ T.cout[bo & 7u] ^ T.cin[bi & 7u];
        khf_update<S>(rh < fh ? rh : fh, M, sk);
    }
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
};

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

template <int S, int MAXK>
__global__ void __launch_bounds__(kSeedThreads, GROOT_SEED_MIN_BLOCKS) seed_kernel(DevIndex ix, SeedArgs a, MultTable M) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int kWarps = kSeedThreads / 32;
    SeedTabs* T = reinterpret_cast<SeedTabs*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(SeedTabs));  // 2 mbarriers per warp
    uint8_t* bufs_all = smem_raw + sizeof(SeedTabs) + 64;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = (a.n_reads + kTileReads - 1) / kTileReads;
    const bool staged = a.tile_bytes != 0;
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
    auto process = [&](const uint8_t* p, uint32_t r, uint32_t len) {
        uint32_t nh = 0;
        uint64_t sk[S];
        khf_sketch<S>(p, len, ix.k, *T, M, sk);
        if (a.sketches_out) {
#pragma unroll
            for (int i = 0; i < S; i++) a.sketches_out[static_cast<size_t>(r) * S + i] = sk[i];
        }
        uint32_t* st = a.stage + static_cast<size_t>(r) * HSTAGE;
        lsh_probe<S, MAXK>(ix, sk, a.len_params[len], [&](uint32_t w) { if (nh < HSTAGE) st[nh] = w; nh++; });
        return nh;
    };

    unsigned phase[2] = {0, 0};
    int cur = 0;
    uint32_t tile = next_tile();
    if (staged && lane == 0 && tile < n_tiles) issue(tile, 0);
    while (tile < n_tiles) {
        const uint32_t tnext = next_tile();
        if (staged && lane == 0 && tnext < n_tiles) issue(tnext, cur ^ 1);  // the other buffer was released by the __syncwarp below
        const uint32_t r0 = tile * kTileReads, r1 = min(a.n_reads, r0 + kTileReads);
        const uint32_t r = r0 + lane;
        bool in_smem = false;
        uint32_t b0 = 0;
        if (staged) {
            mbar_wait(&bar[cur], phase[cur]);
            phase[cur] ^= 1u;
            b0 = a.off[r0] & ~15u;
            const uint32_t b1 = (a.off[r1] + 15u) & ~15u;
            in_smem = (b1 - b0) <= a.tile_bytes;
        }
        if (r < a.n_reads) {
            const uint32_t o = a.off[r], len = a.off[r + 1] - o;
            uint32_t nh = 0;
            if (len < ix.k || len > a.max_len) {
                set_error(a.error, len < ix.k ? -5 : -7, r);  // GROOTGPU_ERR_SHORT_READ / _CAPACITY
            } else if (in_smem) {
                nh = process(bufs + static_cast<size_t>(cur) * a.tile_bytes + (o - b0), r, len);   // LDS path
            } else {
                nh = process(a.seq + o, r, len);                                                   // LDG path
            }
            a.n_hits[r] = nh;
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
};

template <int S, int MAXK>
__global__ void __launch_bounds__(kSeedThreads) fill_kernel(DevIndex ix, FillArgs a, MultTable M) {
    __shared__ SeedTabs T;
    build_seed_tabs(&T, ix.k);
    __syncthreads();
    unsigned mapped = 0, multi = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < a.n_reads; r += gridDim.x * blockDim.x) {
        const uint32_t nh = a.n_hits[r];
        if (nh == 0) continue;
        const uint32_t base = a.hit_off[r];
        if (nh <= HSTAGE) {
            for (uint32_t i = 0; i < nh; i++) a.hits[base + i] = a.stage[static_cast<size_t>(r) * HSTAGE + i];
        } else {
            const uint32_t o = a.off[r], len = a.off[r + 1] - o;
            uint64_t sk[S];
            khf_sketch<S>(a.seq + o, len, ix.k, T, M, sk);
            uint32_t c = 0;
            lsh_probe<S, MAXK>(ix, sk, a.len_params[len], [&](uint32_t w) { if (c < nh) a.hits[base + c] = w; c++; });
        }
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

}  // namespace groot
