// Graph weighting on the host copy of the index: the order-dependent f64 part of `groot align`.
//
// Replaces GrootGraph.IncrementSubPath (src/graph/graph.go:401-451), GrootGraphNode.IncrementKmerFreq
// (src/graph/node.go:25-28), GrootGraph.Prune (graph.go:455-525) and SaveGraphAsGFA (src/graph/graphio.go:19-112).
// The GPU decides WHICH windows get incremented (grootgpu_pair.n_incremented); the additions are replayed
// here in read order because float accumulation order is part of the reference's result.
#include <algorithm>
#include <cstdio>
#include <set>

#include "../flat_index.h"

namespace groot {

void increment_sub_path(FlatIndex& idx, uint32_t win, double num_kmers) {
    const WinRec& w = idx.wins[win];
    if (w.cn_cnt == 1) {  // single segment: all k-mers, KmerTotal untouched (graph.go:409-422)
        idx.kmer_freq[idx.cn_node[w.cn_off]] += num_kmers;
        return;
    }
    double total = 0.0;
    for (uint32_t j = 0; j < w.cn_cnt; j++) total += static_cast<double>(idx.nodes[idx.cn_node[w.cn_off + j]].seq_len);
    for (uint32_t j = 0; j < w.cn_cnt; j++) {
        uint32_t n = idx.cn_node[w.cn_off + j];
        double share = ((static_cast<double>(idx.nodes[n].seq_len) / total) * num_kmers) * static_cast<double>(idx.cn_count[w.cn_off + j]);
        idx.kmer_freq[n] += share;
    }
    idx.kmer_total[w.graph] += static_cast<uint64_t>(num_kmers);
}

// per-node surviving path ids after pruning live in idx.pruned_paths (lazily sized)
bool prune_graph(FlatIndex& idx, uint32_t g, double min_cov) {
    const uint32_t nb = idx.graph_node_base[g], ne = idx.graph_node_base[g + 1], pb = idx.graph_path_base[g];
    std::set<uint32_t> rm_path;
    std::set<uint32_t> rm_node;
    for (uint32_t n = nb; n < ne; n++) {
        const NodeRec& nr = idx.nodes[n];
        double per_base = idx.kmer_freq[n] / static_cast<double>(nr.seq_len);
        if (per_base < min_cov)
            for (uint32_t j = 0; j < nr.path_cnt; j++) { rm_path.insert(idx.node_path_id[nr.path_off + j]); rm_node.insert(n); }
    }
    if (rm_path.size() == idx.n_paths_of(g)) return false;
    if (rm_node.empty()) return true;
    if (idx.pruned_paths.size() != idx.nodes.size()) idx.pruned_paths.assign(idx.nodes.size(), {});
    for (uint32_t n = nb; n < ne; n++) {
        const NodeRec& nr = idx.nodes[n];
        std::vector<uint32_t> keep;
        keep.push_back(UINT32_MAX);  // sentinel: "this node's path list was rewritten"
        for (uint32_t j = 0; j < nr.path_cnt; j++) if (!rm_path.count(idx.node_path_id[nr.path_off + j])) keep.push_back(idx.node_path_id[nr.path_off + j]);
        idx.pruned_paths[n] = keep;
        if (rm_node.count(n)) idx.node_marked[n] = 1;
    }
    for (uint32_t p : rm_path) idx.path_len[pb + p] = 0;
    return true;
}

std::string graph_to_gfa(const FlatIndex& idx, uint32_t g, long long total_kmers) {
    const uint32_t nb = idx.graph_node_base[g], ne = idx.graph_node_base[g + 1], pb = idx.graph_path_base[g];
    bool used = false;
    std::string out = "H\tVN:Z:1\n";
    out += "#\tthis graph is approximately weighted using k-mer frequencies from projected read sketches (total k-mers projected across all graphs: " + std::to_string(total_kmers) + ")\n";
    std::string links;
    char buf[64];
    for (uint32_t n = nb; n < ne; n++) {
        if (idx.node_marked[n]) continue;
        const NodeRec& nr = idx.nodes[n];
        if (idx.kmer_freq[n] > 0) used = true;
        snprintf(buf, sizeof buf, "%lld", static_cast<long long>(idx.kmer_freq[n]));
        out += "S\t" + std::to_string(nr.seg_id) + "\t";
        out.append(reinterpret_cast<const char*>(&idx.node_seq[nr.seq_off]), nr.seq_len);
        out += "\tKC:i:"; out += buf; out += "\n";
        for (uint32_t e = 0; e < nr.edge_cnt; e++) {
            uint32_t t = idx.edges[nr.edge_off + e];
            if (idx.node_marked[t]) continue;  // edges to pruned nodes are dropped (graph.go:506-513)
            links += "L\t" + std::to_string(nr.seg_id) + "\t+\t" + std::to_string(idx.nodes[t].seg_id) + "\t+\t0M\n";
        }
    }
    if (!used) return "";
    out += links;
    const bool pruned = idx.pruned_paths.size() == idx.nodes.size();
    for (uint32_t p = 0; p < idx.n_paths_of(g); p++) {
        if (idx.path_len[pb + p] == 0) continue;
        std::string segs, ovl;
        for (uint32_t n = nb; n < ne; n++) {
            if (idx.node_marked[n]) continue;
            const NodeRec& nr = idx.nodes[n];
            bool has = false;
            if (pruned && !idx.pruned_paths[n].empty()) { for (size_t j = 1; j < idx.pruned_paths[n].size(); j++) if (idx.pruned_paths[n][j] == p) has = true; }
            else has = (idx.node_mask[nr.mask_off + p / 32] >> (p % 32)) & 1u;
            if (!has) continue;
            if (!segs.empty()) { segs += ","; ovl += ","; }
            segs += std::to_string(nr.seg_id) + "+";
            ovl += std::to_string(nr.seq_len) + "M";
        }
        out += "P\t" + idx.path_name[pb + p] + "\t" + segs + "\t" + ovl + "\n";
    }
    return out;
}

}  // namespace groot
