// Persistence and canonical dump of the flat index.
//
// save_index/load_index replace Info.Dump/Load + ContainmentIndex.Dump/Load (src/pipeline/runtime.go:64-91,
// src/lshe/lshe.go:72-146). The file (".grootb200") is this library's own little-endian array container,
// not Go gob: header magic, parameters, then length-prefixed arrays in a fixed order.
// dump_index writes the text form that tests compare against the oracle's dump of the same database.
#include <cstdio>
#include <cstring>
#include <ios>
#include <stdexcept>

#include "../flat_index.h"

namespace groot {
namespace {
const char kMagic[8] = {'G', 'R', 'T', 'B', '2', '0', '0', 1};

struct Writer {
    FILE* f;
    void raw(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("short write"); }
    template <class T> void pod(const T& v) { raw(&v, sizeof v); }
    template <class T> void vec(const std::vector<T>& v) { uint64_t n = v.size(); pod(n); raw(v.data(), n * sizeof(T)); }
};
struct Reader {
    FILE* f;
    void raw(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) throw std::runtime_error("truncated index file"); }
    template <class T> void pod(T& v) { raw(&v, sizeof v); }
    template <class T> void vec(std::vector<T>& v) {
        uint64_t n; pod(n);
        if (n > (1ull << 36) / sizeof(T)) throw std::runtime_error("implausible array length in index file");
        v.resize(n); raw(v.data(), n * sizeof(T));
    }
};
}  // namespace

void save_index(const FlatIndex& idx, const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot create " + path);
    try {
        Writer w{f};
        w.raw(kMagic, 8);
        w.pod(idx.p); w.pod(idx.n_graphs);
        w.vec(idx.graph_node_base); w.vec(idx.graph_path_base); w.vec(idx.graph_mask_words); w.vec(idx.graph_masked); w.vec(idx.graph_raw_windows);
        uint64_t np = idx.path_name.size(); w.pod(np);
        for (auto& s : idx.path_name) { uint32_t n = static_cast<uint32_t>(s.size()); w.pod(n); w.raw(s.data(), n); }
        w.vec(idx.path_len);
        w.vec(idx.nodes); w.vec(idx.node_seq); w.vec(idx.edges); w.vec(idx.node_path_id); w.vec(idx.node_path_pos); w.vec(idx.node_mask);
        w.vec(idx.wins); w.vec(idx.cn_node); w.vec(idx.cn_count); w.vec(idx.sketches);
        w.vec(idx.kmer_freq); w.vec(idx.kmer_total);
    } catch (...) { fclose(f); throw; }
    if (fclose(f) != 0) throw std::runtime_error("cannot write " + path);
}

// Every CSR offset / count and every node, path or window index of a loaded file is range-checked against the array it
// points into, and the parameters against each other: a truncated, corrupt or version-skewed file must fail here
// (GROOTGPU_ERR_FORMAT) instead of reading out of bounds on the host (index_to_device, build_prefix_sets) or in the kernels.
void validate_index(const FlatIndex& idx) {
    auto bad = [](const char* what) { throw std::runtime_error(std::string("inconsistent index file: ") + what); };
    const IndexParams& p = idx.p;
    if (p.k < 1 || p.w < p.k || p.max_k < 1 || p.S < p.max_k) bad("parameters (need 1 <= k <= w, 1 <= maxK <= sketch size)");
    const size_t G = idx.n_graphs, N = idx.nodes.size(), P = idx.path_name.size(), W = idx.wins.size();
    if (idx.graph_node_base.size() != G + 1 || idx.graph_path_base.size() != G + 1 || idx.graph_mask_words.size() != G ||
        idx.graph_masked.size() != G || idx.graph_raw_windows.size() != G) bad("per-graph array sizes");
    if (idx.path_len.size() != P || idx.kmer_freq.size() != N || idx.kmer_total.size() != G) bad("per-path / per-node array sizes");
    if (idx.node_path_pos.size() != idx.node_path_id.size() || idx.cn_count.size() != idx.cn_node.size()) bad("paired array sizes");
    if (idx.sketches.size() != W * static_cast<size_t>(p.S)) bad("sketch array size");
    if (G == 0 || idx.graph_node_base[0] != 0 || idx.graph_path_base[0] != 0 || idx.graph_node_base[G] != N || idx.graph_path_base[G] != P) bad("graph bases");
    for (size_t g = 0; g < G; g++) {
        if (idx.graph_node_base[g] > idx.graph_node_base[g + 1] || idx.graph_path_base[g] > idx.graph_path_base[g + 1]) bad("graph bases not ascending");
        const uint32_t np = idx.graph_path_base[g + 1] - idx.graph_path_base[g];
        if (idx.graph_mask_words[g] != (np + 31) / 32) bad("path bitset width");
        const uint32_t nb = idx.graph_node_base[g], ne = idx.graph_node_base[g + 1], mw = idx.graph_mask_words[g];
        for (uint32_t n = nb; n < ne; n++) {
            const NodeRec& nr = idx.nodes[n];
            if (static_cast<uint64_t>(nr.seq_off) + nr.seq_len > idx.node_seq.size()) bad("node sequence range");
            if (static_cast<uint64_t>(nr.edge_off) + nr.edge_cnt > idx.edges.size()) bad("node edge range");
            if (static_cast<uint64_t>(nr.path_off) + nr.path_cnt > idx.node_path_id.size()) bad("node path range");
            if (static_cast<uint64_t>(nr.mask_off) + mw > idx.node_mask.size()) bad("node bitset range");
            for (uint32_t e = 0; e < nr.edge_cnt; e++) { const uint32_t t = idx.edges[nr.edge_off + e]; if (t < nb || t >= ne) bad("edge target outside its graph"); }
            for (uint32_t j = 0; j < nr.path_cnt; j++) if (idx.node_path_id[nr.path_off + j] >= np) bad("path id outside its graph");
        }
    }
    for (const WinRec& w : idx.wins) {
        if (w.graph >= G) bad("window graph id");
        const uint32_t nb = idx.graph_node_base[w.graph], ne = idx.graph_node_base[w.graph + 1];
        if (w.node < nb || w.node >= ne) bad("window seed node outside its graph");
        if (w.cn_cnt == 0 || static_cast<uint64_t>(w.cn_off) + w.cn_cnt > idx.cn_node.size()) bad("window contained-node range");
        for (uint32_t j = 0; j < w.cn_cnt; j++) { const uint32_t c = idx.cn_node[w.cn_off + j]; if (c < nb || c >= ne) bad("contained node outside its graph"); }
    }
    if (W == 0) throw std::runtime_error("loaded an empty index file");  // lshe.go:103-105
}

void load_index(FlatIndex& idx, const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::ios_base::failure("cannot open " + path);
    try {
        Reader r{f};
        char magic[8]; r.raw(magic, 8);
        if (memcmp(magic, kMagic, 8) != 0) throw std::runtime_error("not a grootb200 index (or a different format version)");
        r.pod(idx.p); r.pod(idx.n_graphs);
        r.vec(idx.graph_node_base); r.vec(idx.graph_path_base); r.vec(idx.graph_mask_words); r.vec(idx.graph_masked); r.vec(idx.graph_raw_windows);
        uint64_t np; r.pod(np);
        if (np > (1u << 28)) throw std::runtime_error("implausible path count");
        idx.path_name.resize(np);
        for (auto& s : idx.path_name) { uint32_t n; r.pod(n); if (n > (1u << 20)) throw std::runtime_error("implausible path name"); s.resize(n); r.raw(&s[0], n); }
        r.vec(idx.path_len);
        r.vec(idx.nodes); r.vec(idx.node_seq); r.vec(idx.edges); r.vec(idx.node_path_id); r.vec(idx.node_path_pos); r.vec(idx.node_mask);
        r.vec(idx.wins); r.vec(idx.cn_node); r.vec(idx.cn_count); r.vec(idx.sketches);
        r.vec(idx.kmer_freq); r.vec(idx.kmer_total);
        idx.node_marked.assign(idx.nodes.size(), 0);
        validate_index(idx);
    } catch (...) { fclose(f); throw; }
    fclose(f);
}

void dump_index(const FlatIndex& idx, void (*sink)(void*, const char*, size_t), void* ctx) {
    std::string line;
    char buf[256];
    auto emit = [&] { line.push_back('\n'); sink(ctx, line.data(), line.size()); line.clear(); };
    snprintf(buf, sizeof buf, "I k=%u S=%u w=%u numPart=%u maxK=%u graphs=%u windows=%zu", idx.p.k, idx.p.S, idx.p.w, idx.p.num_part, idx.p.max_k, idx.n_graphs, idx.wins.size());
    line = buf; emit();
    for (uint32_t g = 0; g < idx.n_graphs; g++) {
        uint32_t nb = idx.graph_node_base[g], ne = idx.graph_node_base[g + 1], pb = idx.graph_path_base[g];
        snprintf(buf, sizeof buf, "G %u masked=%d paths=%u nodes=%u", g, idx.graph_masked[g] ? 1 : 0, idx.n_paths_of(g), ne - nb);
        line = buf; emit();
        for (uint32_t p = 0; p < idx.n_paths_of(g); p++) { snprintf(buf, sizeof buf, "P %u %d ", p, idx.path_len[pb + p]); line = buf; line += idx.path_name[pb + p]; emit(); }
        for (uint32_t n = nb; n < ne; n++) {
            const NodeRec& nr = idx.nodes[n];
            snprintf(buf, sizeof buf, "N %u ", nr.seg_id); line = buf;
            line.append(reinterpret_cast<const char*>(&idx.node_seq[nr.seq_off]), nr.seq_len);
            line += " E";
            for (uint32_t e = 0; e < nr.edge_cnt; e++) { snprintf(buf, sizeof buf, " %u", idx.nodes[idx.edges[nr.edge_off + e]].seg_id); line += buf; }
            line += " P";
            for (uint32_t j = 0; j < nr.path_cnt; j++) { snprintf(buf, sizeof buf, " %u:%d", idx.node_path_id[nr.path_off + j], idx.node_path_pos[nr.path_off + j]); line += buf; }
            emit();
        }
    }
    for (size_t w = 0; w < idx.wins.size(); w++) {
        const WinRec& wr = idx.wins[w];
        snprintf(buf, sizeof buf, "W %u %u %u span=%u w=%u S", wr.graph, wr.seg_id, wr.offset, wr.merge_span, wr.win_size); line = buf;
        for (uint32_t i = 0; i < idx.p.S; i++) { snprintf(buf, sizeof buf, " %016llx", static_cast<unsigned long long>(idx.sketches[w * idx.p.S + i])); line += buf; }
        line += " C";
        for (uint32_t j = 0; j < wr.cn_cnt; j++) { snprintf(buf, sizeof buf, " %u:%u", idx.nodes[idx.cn_node[wr.cn_off + j]].seg_id, idx.cn_count[wr.cn_off + j]); line += buf; }
        emit();
    }
}

}  // namespace groot
