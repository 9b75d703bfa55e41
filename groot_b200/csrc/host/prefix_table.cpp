// Prefix table: for every base position of every graph node, the distinct 8-base strings spelled by the
// traversals that start there (following out-edges; shorter when a sink node ends the traversal first).
//
// It is a derived acceleration structure for the align kernel's screen phase: dfsRecursive
// (src/graph/alignment.go:196-254) can only succeed from (node, offset) if the read's first bases agree with one
// of these prefixes (a reference 'N' is a wildcard at its position), so a try is rejected with one or two loads and a
// masked XOR instead of a bounded DFS (8 bases keep the number of distinct prefixes small even across dense SNP
// bubbles and still reject a random try with probability 1 - 4^-8).
// List entries (pfx_off/pfx): bits 0-15 = bases 2 bits each (A0 C1 T2 G3 == pack_base2 of device_types.cuh, base i
// at bits 2i), bits 32-36 = length (1..8), bit 40 = "always pass" (the position has too many distinct prefixes),
// bits 48-55 = wildcard mask (bit i: base i is an 'N').
// pfx1[pos] is the one-load form: bits 0-15 bases, bits 16-19 length, bits 20-27 wildcard mask, bit 31 clear for a
// position with exactly one prefix. With bit 31 set it holds the COMMON prefix of the position's alternatives (up to
// their first divergence or wildcard; length 0 when there is none): a read that fails it fails them all, only
// the others walk the list.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../flat_index.h"

namespace groot {
namespace {
constexpr uint32_t kPfxLen = 8;
constexpr size_t kMaxPerPos = 24;
constexpr uint64_t kWild = 1ull << 40;

inline int code_of(uint8_t b) { return b == 'A' ? 0 : b == 'C' ? 1 : b == 'T' ? 2 : b == 'G' ? 3 : -1; }

struct Builder {
    const FlatIndex& ix;
    std::vector<uint64_t> out;  // entries of the current position
    bool overflow = false;
    void walk(uint32_t node, uint32_t off, uint32_t packed, uint32_t len, uint32_t nmask) {
        if (overflow) return;
        const NodeRec& nd = ix.nodes[node];
        for (uint32_t i = off; i < nd.seq_len && len < kPfxLen; i++) {
            int c = code_of(ix.node_seq[nd.seq_off + i]);
            if (c < 0) { nmask |= 1u << len; c = 0; }
            packed |= static_cast<uint32_t>(c) << (2 * len);
            len++;
        }
        if (len == kPfxLen || nd.edge_cnt == 0) {
            uint64_t e = static_cast<uint64_t>(packed) | (static_cast<uint64_t>(len) << 32) | (static_cast<uint64_t>(nmask) << 48);
            if (std::find(out.begin(), out.end(), e) == out.end()) {
                if (out.size() >= kMaxPerPos) { overflow = true; return; }
                out.push_back(e);
            }
            return;
        }
        for (uint32_t e = 0; e < nd.edge_cnt; e++) walk(ix.edges[nd.edge_off + e], 0, packed, len, nmask);
    }
};
}  // namespace

void build_prefix_table(const FlatIndex& ix, std::vector<uint32_t>& pfx_off, std::vector<uint64_t>& pfx, std::vector<uint32_t>& pfx1) {
    pfx_off.assign(ix.node_seq.size() + 1, 0);
    pfx1.assign(ix.node_seq.size() + 1, 0x80000000u);
    pfx.clear();
    Builder b{ix};
    for (uint32_t n = 0; n < ix.nodes.size(); n++) {
        const NodeRec& nd = ix.nodes[n];
        for (uint32_t off = 0; off < nd.seq_len; off++) {
            b.out.clear(); b.overflow = false;
            b.walk(n, off, 0, 0, 0);
            pfx_off[nd.seq_off + off] = static_cast<uint32_t>(pfx.size());
            if (b.overflow) pfx.push_back(kWild | (1ull << 32));
            else pfx.insert(pfx.end(), b.out.begin(), b.out.end());
            uint32_t& one = pfx1[nd.seq_off + off];
            if (!b.overflow && b.out.size() == 1) {
                const uint64_t e = b.out[0];
                one = static_cast<uint32_t>(e & 0xffffu) | (static_cast<uint32_t>(e >> 32) & 15u) << 16 | (static_cast<uint32_t>(e >> 48) & 0xffu) << 20;
            } else if (!b.overflow) {
                uint32_t lc = kPfxLen;
                for (uint64_t e : b.out) {
                    lc = std::min<uint32_t>(lc, static_cast<uint32_t>(e >> 32) & 15u);
                    const uint32_t nm = static_cast<uint32_t>(e >> 48) & 0xffu;
                    for (uint32_t i = 0; i < lc; i++) {
                        const bool same = ((static_cast<uint32_t>(e) ^ static_cast<uint32_t>(b.out[0])) >> (2 * i) & 3u) == 0;
                        if (!same || (nm >> i & 1u) || (static_cast<uint32_t>(b.out[0] >> 48) >> i & 1u)) { lc = i; break; }
                    }
                }
                one = 0x80000000u | (static_cast<uint32_t>(b.out[0]) & ((1u << (2 * lc)) - 1u)) | lc << 16;
            }
        }
    }
    pfx_off[ix.node_seq.size()] = static_cast<uint32_t>(pfx.size());
}

}  // namespace groot
