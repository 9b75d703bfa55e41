// Device-side views shared by the kernels (sm_100a). Integer / byte work only — no tensor cores.
#pragma once
#include <cstdint>

#include "flat_index.h"

namespace groot {

// ntHash v1 constants (will-rowe/nthash v0.2.0; see oracle/nthash.hpp for the provenance notes)
#define GROOT_SEED_A 0x3c8bfbb395c60474ULL
#define GROOT_SEED_C 0x3193c18562a02b4cULL
#define GROOT_SEED_G 0x20323ed082572324ULL
#define GROOT_SEED_T 0x295549f54be24456ULL
#define GROOT_MULTI_SEED 0x90b45d39fb6da1faULL
#define GROOT_MULTI_SHIFT 27

// One 32-byte slot of an open-addressing table over LSH band keys: the "flattened CSR of the
// lshensemble index". key = low 32 bits of K consecutive sketch slots (lshensemble's 32-bit
// hashKeyFunc), [start, start+count) = bucket in LshTable::wins (window ids ascending).
struct LshSlot {
    uint32_t key[4];
    uint32_t start;
    uint32_t count;  // 0 == empty slot
    uint32_t pad[2];
};
static_assert(sizeof(LshSlot) == 32, "LshSlot must be one 32-byte sector");

struct LshTable {
    const LshSlot* slots;
    const uint32_t* wins;
    uint32_t mask;  // capacity - 1 (power of two)
    uint32_t pad;
};

// per query length: what the host optimiser chose (lshe_params.cpp)
struct LenParam {
    uint8_t K;        // hashes per band probed
    uint8_t L;        // bands probed
    uint16_t eq_min;  // equal sketch slots needed to pass the containment threshold; > S == never
};

__host__ __device__ inline uint32_t band_key_hash(const uint32_t key[4]) {
    uint32_t h = key[0] * 0x9E3779B1u;
    h = (h ^ (h >> 15)) + key[1] * 0x85EBCA77u;
    h = (h ^ (h >> 13)) + key[2] * 0xC2B2AE3Du;
    h = (h ^ (h >> 16)) + key[3] * 0x27D4EB2Fu;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}

// 128-bit digest of a whole sketch: the key of the "identical sketch" table (below). Any mixing function works — a
// collision only adds a candidate that the verification rejects — but host and device must agree bit for bit.
__host__ __device__ inline void sketch_digest(const uint64_t* sk, uint32_t S, uint32_t key[4]) {
    uint32_t a[4] = {0x243F6A88u, 0x85A308D3u, 0x13198A2Eu, 0x03707344u};
    for (uint32_t i = 0; i < S; i++) {
        const uint32_t lo = static_cast<uint32_t>(sk[i]), hi = static_cast<uint32_t>(sk[i] >> 32);
        uint32_t& x = a[i & 3u];
        x = ((x << 13) | (x >> 19)) ^ (lo * 0x9E3779B1u + hi);
    }
    for (int j = 0; j < 4; j++) key[j] = a[j];
}

// What one step of the packed walk needs from a node, in ONE 32-byte sector (align_kernels.cuh, dfs_packed): the walk is
// bound by the number of sectors it pulls through L1 / L2 per node step (measured: 134 per pair), and NodeRec + path
// bitset + out-edge were three of them.
struct WalkNode {
    uint32_t seq_off;    // NodeRec::seq_off (position in node_seq2 / node_n2)
    uint32_t seq_len;
    uint32_t edge;       // edge_cnt == 1: the target node itself; otherwise NodeRec::edge_off (index into edges[])
    uint32_t edge_cnt;
    uint32_t mask[4];    // path bitset of the node when its graph has <= 128 paths; otherwise mask[0] = NodeRec::mask_off
};
static_assert(sizeof(WalkNode) == 32, "WalkNode must be one 32-byte sector");
constexpr uint32_t kWalkMaskWords = 4;

struct DevIndex {
    const NodeRec* nodes;
    const uint8_t* node_seq;
    const uint32_t* edges;
    const uint32_t* node_path_id;
    const int32_t* node_path_pos;
    const uint32_t* node_mask;
    const WinRec* wins;
    const uint32_t* cn_node;
    const uint64_t* sketches;
    const uint32_t* graph_mask_words;
    const LshTable* tables;  // [(K-1)*n_bands + band]; slots == nullptr when not built yet
    const uint32_t* win_kmers;    // [n_wins * 32] per window: the 5-base prefixes possible at any of its tries (host/prefix_table.cpp, build_window_kmer_sets)
    const uint32_t* pfxset;  // [node_seq bytes + 1] allele sets of the first 8 steps from a position (host/prefix_table.cpp), position = NodeRec::seq_off + offset
    const uint32_t* node_seq2;    // node_seq packed 2 bits per base (pack_base2), 16 bases per word, same positions as node_seq
    const uint32_t* node_n2;      // same layout: bit 2i of a word set when base i is an 'N' wildcard (alignment.go:212-215)
    const uint8_t* graph_has_n;   // [G] 1 when a node of the graph holds an 'N' (only then node_n2 is consulted)
    const WalkNode* wnodes;       // [n_nodes] the packed walk's view of nodes[] + edges[] + node_mask[]
    // Windows grouped by IDENTICAL sketch (key = sketch_digest). When a query needs every slot equal (eq_min == S: a read as
    // long as the index window at the default threshold) and probes one band, "in the band's bucket and all S slots equal"
    // is the same set as "identical sketch": this table yields ~1 candidate per seeded read where the band bucket holds ~11
    // (neighbouring windows share their first K minima), i.e. a tenth of the verification loads.
    LshTable full;
    uint32_t k, S, max_k, n_bands, n_wins;
};

// 2-bit code of an upper-case base: A0 C1 T2 G3, so that the complement (seqio.go:17-23) is code ^ 2
__host__ __device__ inline uint32_t pack_base2(uint8_t b) { return (b >> 1) & 3u; }

// multi-hash multipliers c_i = i ^ (k * multiSeed), passed by value => they live in the constant bank.
// For i < 32 the xor only touches the low 5 bits: c_i = c0 + low[i] with c0 = C & ~31, low[i] = (C & 31) ^ i,
// so h * c_i = h * c0 + h * low[i]: one 64-bit product per k-mer plus a 32x32->64 multiply-add per hash.
struct MultTable {
    uint64_t c[32];
    uint64_t c0;
    uint32_t low[32];
    uint32_t m32;  // == 32, kept as a runtime value so that >>27 can be issued as multiplies (FMA pipe) instead of shifts (ALU pipe)
    uint32_t one;  // == 1, a runtime value for the same reason: the running-minimum update as predicated multiply-adds
    uint32_t c_hi;      // high word of every c_i (the xor with i < 32 only touches the low word)
    uint32_t key_low;   // low bits of a product's high word that do NOT decide the order of x ^ (x >> 27): 31 (GROOTGPU_KHF_KEYBITS lowers the
                        // number of deciding bits in test runs, so that the exact tie path is exercised)
};

enum : int { HSTAGE = 4 };  // hits per read staged by the seed kernel before the exact-size fill

}  // namespace groot
