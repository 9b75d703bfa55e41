// Flat (CSR) in-memory layout of the groot index: what the reference keeps as a gob-encoded
// map[uint32]*GrootGraph (groot.gg, src/pipeline/runtime.go:15-27, src/graph/graph.go:18-34,
// src/graph/node.go:13-22) plus WindowLookup map[string]Key (groot.lshe, src/lshe/lshe.go:17-49),
// restated as arrays so the whole thing is one cudaMemcpy per array and every record the kernels
// touch is a single 32-byte sector.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace groot {

// One graph node == GrootGraphNode (node.go:13-22). Global node index = graph_node_base[g] + position
// in SortedNodes (topological order, graph.go:150-218).
struct NodeRec {
    uint32_t seq_off;   // into node_seq
    uint32_t seq_len;   // SegmentLength
    uint32_t edge_off;  // into edges[]: OutEdges as global node indices, reference order = descending SegmentID (graph.go:203)
    uint32_t edge_cnt;
    uint32_t path_off;  // into node_path_id[] / node_path_pos[]: PathIDs ascending + Position[pathID]
    uint32_t path_cnt;
    uint32_t mask_off;  // into node_mask[]: bitset over the graph's path ids (graph_mask_words[g] words)
    uint32_t seg_id;    // SegmentID
};
static_assert(sizeof(NodeRec) == 32, "NodeRec must be one 32-byte sector");

// One indexed window == lshe.Key (lshe.go:17-28) minus Ref/RC/Freq (unused by align).
struct WinRec {
    uint32_t graph;       // GraphID
    uint32_t node;        // global node index of Key.Node
    uint32_t offset;      // Key.OffSet
    uint32_t merge_span;  // Key.MergeSpan
    uint32_t win_size;    // Key.WindowSize
    uint32_t cn_off;      // into cn_node[] / cn_count[]: ContainedNodes, ascending SegmentID
    uint32_t cn_cnt;
    uint32_t seg_id;      // Key.Node (SegmentID) — the graphminion.go:57 sort key
};
static_assert(sizeof(WinRec) == 32, "WinRec must be one 32-byte sector");

struct IndexParams {
    uint32_t k = 31, S = 21, w = 100, num_part = 8, max_k = 4;
};

struct FlatIndex {
    IndexParams p;
    // graphs
    uint32_t n_graphs = 0;
    std::vector<uint32_t> graph_node_base;   // [G+1]
    std::vector<uint32_t> graph_path_base;   // [G+1]
    std::vector<uint32_t> graph_mask_words;  // [G] ceil(n_paths/32)
    std::vector<uint8_t> graph_masked;       // [G] Masked (pipeline/index.go:59-65)
    std::vector<uint64_t> graph_raw_windows; // [G] numWindows (graph.go:238-241)
    // paths (global path index = graph_path_base[g] + pathID)
    std::vector<std::string> path_name;
    std::vector<int32_t> path_len;           // Lengths (set to 0 by Prune for removed paths)
    // nodes
    std::vector<NodeRec> nodes;
    std::vector<uint8_t> node_seq;
    std::vector<uint32_t> edges;
    std::vector<uint32_t> node_path_id;
    std::vector<int32_t> node_path_pos;
    std::vector<uint32_t> node_mask;
    // windows, sorted by (graph, SegmentID, OffSet, arrival)
    std::vector<WinRec> wins;
    std::vector<uint32_t> cn_node;           // global node index
    std::vector<uint32_t> cn_count;          // integer-valued f64 in the reference
    std::vector<uint64_t> sketches;          // [W*S]
    // mutable state of `groot align`
    std::vector<double> kmer_freq;           // per node (node.go:21)
    std::vector<uint64_t> kmer_total;        // per graph (graph.go:26)
    std::vector<uint8_t> node_marked;        // Prune marks (node.go:22)
    std::vector<std::vector<uint32_t>> pruned_paths;  // filled by prune(): per node surviving path ids (empty vector == untouched)

    uint32_t graph_of_node(uint32_t node) const;
    uint32_t n_paths_of(uint32_t g) const { return graph_path_base[g + 1] - graph_path_base[g]; }
};

// ---- host-side construction / persistence (groot_b200/csrc/host/*.cpp) ----

// One window-to-be-sketched: `len` bases at `off` of `seqs`
struct GraphBuild;  // opaque scratch of the builder

// MSA text -> graph g appended to idx (nodes, edges, paths, positions, masks). Throws std::runtime_error.
void append_graph_from_msa(FlatIndex& idx, const std::string& msa_text);

// Sketch callback: sketches n windows of length w starting at offsets off[i] of seqs into out[n*S].
using SketchFn = void (*)(void* ctx, const uint8_t* seqs, size_t seqs_len, const uint64_t* off, uint32_t n, uint32_t w,
                          uint32_t k, uint32_t S, uint64_t* out);
// WindowGraph + SketchIndexer (graph.go:229-396, pipeline/index.go:184-211) for every unmasked graph.
void build_windows(FlatIndex& idx, SketchFn sketch, void* ctx);

void save_index(const FlatIndex& idx, const std::string& path);
void load_index(FlatIndex& idx, const std::string& path);
void validate_index(const FlatIndex& idx);   // range checks of every offset / index; throws std::runtime_error
// groot.gg + groot.lshe (Go gob, host/gob_reader.cpp)
void load_index_gob(FlatIndex& idx, const std::string& gg_path, const std::string& lshe_path);
void save_index_gob(const FlatIndex& idx, const std::string& gg_path, const std::string& lshe_path, const std::string& version);   // host/gob_writer.cpp
// canonical text dump shared (as a FORMAT) with the oracle; sink(line incl. '\n')
void dump_index(const FlatIndex& idx, void (*sink)(void* ctx, const char* data, size_t n), void* ctx);

// derived acceleration structure for the align kernel (host/prefix_table.cpp)
void build_prefix_sets(const FlatIndex& idx, std::vector<uint32_t>& pset);
void build_window_kmer_sets(const FlatIndex& idx, const std::vector<uint32_t>& pset, std::vector<uint32_t>& wk);

// lshensemble parameter optimiser + containment threshold, exact f64 expressions (host, once per query size)
void optimal_kl(int max_k, int max_l, int x, int q, double t, int* K, int* L);
int eq_min_for(int S, int q_size, int x_size, double threshold);

// IncrementSubPath replay / Prune / GFA
void increment_sub_path(FlatIndex& idx, uint32_t win, double num_kmers);
bool prune_graph(FlatIndex& idx, uint32_t g, double min_cov);
std::string graph_to_gfa(const FlatIndex& idx, uint32_t g, long long total_kmers);

}  // namespace groot
