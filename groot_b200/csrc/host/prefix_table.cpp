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

// Window 5-mer sets: for every indexed window, WHICH 5-base prefixes a read can start with at ANY of the window's tries —
// the seed-node offsets of stage 1, offsets 0..10 of the contained nodes of stage 2, the seed position of stages 3 / 4
// (alignment.go:35-103) — as a 1 024-bit set per window (32 words), every combination of the allele sets of the first five
// steps of every try position. A read whose first five bases (and, for the start-clipped stage 3, bases 1..5) are not in
// the set cannot pass the allele-set test at any try of that (window, strand): the screen skips the whole try list
// with one load. This is what a read costs on the strand it does NOT align on — half of all (pair, strand) scans.
void build_window_kmer_sets(const FlatIndex& ix, const std::vector<uint32_t>& pset, std::vector<uint32_t>& wk) {
    wk.assign(ix.wins.size() * 32, 0u);
    auto add_position = [&](uint32_t* row, uint32_t pos) {
        const uint32_t sets = pset[pos];
        uint32_t s[5];
        for (int i = 0; i < 5; i++) s[i] = (sets >> (4 * i)) & 0xFu;
        for (uint32_t c0 = 0; c0 < 4; c0++) if (s[0] >> c0 & 1u)
            for (uint32_t c1 = 0; c1 < 4; c1++) if (s[1] >> c1 & 1u)
                for (uint32_t c2 = 0; c2 < 4; c2++) if (s[2] >> c2 & 1u)
                    for (uint32_t c3 = 0; c3 < 4; c3++) if (s[3] >> c3 & 1u)
                        for (uint32_t c4 = 0; c4 < 4; c4++) if (s[4] >> c4 & 1u) {
                            const uint32_t idx = c0 | c1 << 2 | c2 << 4 | c3 << 6 | c4 << 8;
                            row[idx >> 5] |= 1u << (idx & 31u);
                        }
    };
    for (size_t w = 0; w < ix.wins.size(); w++) {
        const WinRec& wr = ix.wins[w];
        uint32_t* row = &wk[w * 32];
        const NodeRec& sn = ix.nodes[wr.node];
        const uint32_t t1 = wr.merge_span + wr.win_size + 1u;
        for (uint32_t t = 0; t < t1 && wr.offset + t < sn.seq_len; t++) add_position(row, sn.seq_off + wr.offset + t);
        for (uint32_t j = 0; j < wr.cn_cnt; j++) {
            const NodeRec& cn = ix.nodes[ix.cn_node[wr.cn_off + j]];
            for (uint32_t off = 0; off < 11u && off < cn.seq_len; off++) add_position(row, cn.seq_off + off);
        }
    }
}

}  // namespace groot
