// Prefix sets: for every base position of every graph node, WHICH bases a traversal starting there can spell at each
// of its first 8 steps — eight 4-bit allele sets (bit pack_base2(b) of nibble i = "some traversal has base b at step
// i") in one 32-bit word per position.
//
// It is a derived acceleration structure for the align kernels' screen phase: dfsRecursive
// (src/graph/alignment.go:196-254) can only succeed from (node, offset) if every one of the read's first bases lies in
// the set of its step, so a try is rejected with ONE load and an AND instead of a bounded DFS. The condition is
// necessary, not sufficient (it ignores which alleles occur together), and errs only towards "pass":
//   * a reference 'N' is a wildcard (alignment.go:212-215)            -> nibble 0xF
//   * once some traversal has ended in a sink node, a longer read is accepted on it whatever follows
//     (alignment.go:229)                                              -> every later nibble 0xF
// An earlier form kept the exact list of distinct 8-base prefixes per position; walking those lists (up to 24 entries
// next to dense SNP bubbles) cost a fifth of the screen kernel's issue slots at one active lane, for a rejection rate
// the sets match in practice (profiles/r01_notes.md).
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

#include "../flat_index.h"

namespace groot {
namespace {
constexpr uint32_t kPfxLen = 8;
inline int code_of(uint8_t b) { return b == 'A' ? 0 : b == 'C' ? 1 : b == 'T' ? 2 : b == 'G' ? 3 : -1; }   // == pack_base2
}  // namespace

void build_prefix_sets(const FlatIndex& ix, std::vector<uint32_t>& pset) {
    pset.assign(ix.node_seq.size() + 1, 0xffffffffu);
    std::vector<std::pair<uint32_t, uint32_t>> frontier, next;   // (node, offset)
    for (uint32_t n = 0; n < ix.nodes.size(); n++) {
        const NodeRec& nd = ix.nodes[n];
        for (uint32_t off = 0; off < nd.seq_len; off++) {
            uint32_t sets = 0;
            bool open_end = false;
            frontier.assign(1, {n, off});
            for (uint32_t d = 0; d < kPfxLen; d++) {
                if (open_end || frontier.empty() || frontier.size() > 64) { sets |= 0xFu << (4 * d); open_end = true; continue; }
                uint32_t m = 0;
                next.clear();
                for (auto [fn, fo] : frontier) {
                    const NodeRec& f = ix.nodes[fn];
                    const int c = code_of(ix.node_seq[f.seq_off + fo]);
                    m |= c < 0 ? 0xFu : 1u << c;
                    if (fo + 1 < f.seq_len) next.emplace_back(fn, fo + 1);
                    else if (f.edge_cnt == 0) open_end = true;
                    else for (uint32_t e = 0; e < f.edge_cnt; e++) next.emplace_back(ix.edges[f.edge_off + e], 0u);
                }
                sets |= m << (4 * d);
                std::sort(next.begin(), next.end());
                next.erase(std::unique(next.begin(), next.end()), next.end());
                frontier.swap(next);
            }
            pset[nd.seq_off + off] = sets;
        }
    }
}

}  // namespace groot
