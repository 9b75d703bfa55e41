// Prefix table: for every base position of every graph node, the distinct 8-base strings spelled by the
// traversals that start there (following out-edges; shorter when a sink node ends the traversal first).
//
// It is a derived acceleration structure for the align kernel's scan phase: dfsRecursive
// (src/graph/alignment.go:196-254) can only succeed from (node, offset) if the read's first bases agree with one
// of these prefixes (reference 'N' is a wildcard), so a try is rejected with two loads and a masked XOR instead
// of a bounded DFS (8 bases keep the number of distinct prefixes small even across dense SNP bubbles and still
// reject a random try with probability 1 - 4^-8). Entries: bits 0-31 = bases 2 bits each (A0 C1 G2 T3, base i at
// bits 2i), bits 32-36 = length (1..8), bit 40 = wildcard (prefix holds an 'N', or the position has too many distinct prefixes) = "always pass".
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../flat_index.h"

namespace groot {
namespace {
constexpr uint32_t kPfxLen = 8;
constexpr size_t kMaxPerPos = 24;
constexpr uint64_t kWild = 1ull << 40;

inline int code_of(uint8_t b) { return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : -1; }

struct Builder {
    const FlatIndex& ix;
    std::vector<uint64_t> out;  // entries of the current position
    bool overflow = false;
    void walk(uint32_t node, uint32_t off, uint32_t packed, uint32_t len, bool has_n) {
        if (overflow) return;
        const NodeRec& nd = ix.nodes[node];
        for (uint32_t i = off; i < nd.seq_len && len < kPfxLen; i++) {
            int c = code_of(ix.node_seq[nd.seq_off + i]);
            if (c < 0) { has_n = true; c = 0; }
            packed |= static_cast<uint32_t>(c) << (2 * len);
            len++;
        }
        if (len == kPfxLen || nd.edge_cnt == 0) {
            uint64_t e = static_cast<uint64_t>(packed) | (static_cast<uint64_t>(len) << 32) | (has_n ? kWild : 0);
            if (std::find(out.begin(), out.end(), e) == out.end()) {
                if (out.size() >= kMaxPerPos) { overflow = true; return; }
                out.push_back(e);
            }
            return;
        }
        for (uint32_t e = 0; e < nd.edge_cnt; e++) walk(ix.edges[nd.edge_off + e], 0, packed, len, has_n);
    }
};
}  // namespace

void build_prefix_table(const FlatIndex& ix, std::vector<uint32_t>& pfx_off, std::vector<uint64_t>& pfx) {
    pfx_off.assign(ix.node_seq.size() + 1, 0);
    pfx.clear();
    Builder b{ix};
    for (uint32_t n = 0; n < ix.nodes.size(); n++) {
        const NodeRec& nd = ix.nodes[n];
        for (uint32_t off = 0; off < nd.seq_len; off++) {
            b.out.clear(); b.overflow = false;
            b.walk(n, off, 0, 0, false);
            pfx_off[nd.seq_off + off] = static_cast<uint32_t>(pfx.size());
            if (b.overflow) pfx.push_back(kWild | (1ull << 32));
            else pfx.insert(pfx.end(), b.out.begin(), b.out.end());
        }
    }
    pfx_off[ix.node_seq.size()] = static_cast<uint32_t>(pfx.size());
}

}  // namespace groot
