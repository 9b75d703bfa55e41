// Host-side index construction for the B200 path: MSA -> variation graph (flat CSR) -> window keys.
//
// What it replaces in the reference (all offline, once per database):
//   gfa.ReadMSA / gfa.MSA2GFA (third-party will-rowe/gfa; call site src/pipeline/index.go:43-49)
//   graph.CreateGrootGraph + topoSort + GetPaths      (src/graph/graph.go:37-218, 575-622)
//   GrootGraph.WindowGraph                             (src/graph/graph.go:229-396)
//   SketchIndexer.Run                                  (src/pipeline/index.go:184-211)
// The per-window KHF sketches are NOT computed here: build_windows() hands every window of every path
// to a callback, which libgrootgpu wires to the same CUDA sketch kernel the read path uses.
//
// Deterministic choices where the reference's order is undefined (Go map iteration / goroutine
// arrival): MSA column nodes numbered by first occurrence in row order; paths windowed in ascending
// pathID; windows ordered by (graph, SegmentID, OffSet, arrival).
#include <algorithm>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#include "../flat_index.h"

namespace groot {

uint32_t FlatIndex::graph_of_node(uint32_t node) const {
    auto it = std::upper_bound(graph_node_base.begin(), graph_node_base.end(), node);
    return static_cast<uint32_t>(it - graph_node_base.begin()) - 1;
}

namespace {

struct Row { std::string name, seq; };

std::vector<Row> parse_msa(const std::string& text) {
    std::vector<Row> rows;
    size_t i = 0, n = text.size();
    bool have = false;
    Row cur;
    auto flush = [&] { if (have && cur.name != "consensus") rows.push_back(std::move(cur)); cur = Row(); };
    while (i < n) {
        size_t e = text.find('\n', i);
        if (e == std::string::npos) e = n;
        size_t b = i, t = e;
        while (t > b && (text[t - 1] == '\r' || text[t - 1] == ' ' || text[t - 1] == '\t')) t--;
        if (t > b) {
            if (text[b] == '>') {
                flush();
                have = true;
                size_t ne = b + 1;
                while (ne < t && text[ne] != ' ' && text[ne] != '\t') ne++;
                cur.name.assign(text, b + 1, ne - b - 1);
            } else if (have) {
                cur.seq.append(text, b, t - b);
            }
        }
        i = e + 1;
    }
    flush();
    return rows;
}

// column graph before squashing
struct ColNode {
    std::string seq;
    std::vector<uint32_t> rows;  // ascending
    std::vector<uint32_t> out, in;
    int32_t merged_into = -1;
};

inline uint8_t normalise_base(uint8_t c) {  // Sequence.BaseCheck, src/seqio/seqio.go:72-91
    if (c >= 'a' && c <= 'z') c = static_cast<uint8_t>(c - 32);
    return (c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N') ? c : 'N';
}

}  // namespace

void append_graph_from_msa(FlatIndex& idx, const std::string& msa_text) {
    std::vector<Row> rows = parse_msa(msa_text);
    if (rows.empty()) throw std::runtime_error("MSA holds no sequences");
    const size_t ncol = rows[0].seq.size();
    for (auto& r : rows) if (r.seq.size() != ncol) throw std::runtime_error("MSA rows differ in length: " + r.name);
    const uint32_t R = static_cast<uint32_t>(rows.size());

    // ---- MSA2GFA: one node per distinct byte per column, edges between consecutive non-gap nodes ----
    std::vector<ColNode> cn;
    std::vector<int32_t> last(R, -1);
    std::vector<std::vector<uint32_t>> chain(R);
    int32_t slot[256];
    for (size_t c = 0; c < ncol; c++) {
        std::fill(std::begin(slot), std::end(slot), -1);
        for (uint32_t r = 0; r < R; r++) {
            uint8_t b = static_cast<uint8_t>(rows[r].seq[c]);
            if (b == '-') continue;
            if (slot[b] < 0) { slot[b] = static_cast<int32_t>(cn.size()); cn.emplace_back(); cn.back().seq.assign(1, static_cast<char>(b)); }
            uint32_t id = static_cast<uint32_t>(slot[b]);
            cn[id].rows.push_back(r);
            if (last[r] >= 0) {
                auto& o = cn[last[r]].out;
                if (std::find(o.begin(), o.end(), id) == o.end()) { o.push_back(id); cn[id].in.push_back(static_cast<uint32_t>(last[r])); }
            }
            last[r] = static_cast<int32_t>(id);
            chain[r].push_back(id);
        }
    }
    // ---- squash linear chains with identical row sets ----
    for (uint32_t u = 0; u < cn.size(); u++) {
        if (cn[u].merged_into >= 0) continue;
        while (cn[u].out.size() == 1) {
            uint32_t v = cn[u].out[0];
            if (cn[v].in.size() != 1 || cn[v].rows != cn[u].rows) break;
            cn[u].seq += cn[v].seq;
            cn[u].out = cn[v].out;
            for (uint32_t w : cn[u].out) for (uint32_t& p : cn[w].in) if (p == v) p = u;
            cn[v].merged_into = static_cast<int32_t>(u);
        }
    }
    // ---- number survivors 1..N in creation (column) order ----
    std::vector<uint32_t> seg_of(cn.size(), 0);
    std::vector<uint32_t> survivors;
    for (uint32_t u = 0; u < cn.size(); u++) if (cn[u].merged_into < 0) { survivors.push_back(u); seg_of[u] = static_cast<uint32_t>(survivors.size()); }
    const uint32_t N = static_cast<uint32_t>(survivors.size());

    // ---- topological order as graph.go:150-218 produces it: DFS from every segment in GFA order,
    //      out-edges visited in descending SegmentID, finished nodes prepended ----
    std::vector<std::vector<uint32_t>> out_desc(N + 1);  // by seg id
    for (uint32_t u : survivors) {
        auto& o = out_desc[seg_of[u]];
        for (uint32_t v : cn[u].out) o.push_back(seg_of[v]);
        std::sort(o.begin(), o.end(), std::greater<uint32_t>());
    }
    std::vector<uint32_t> finish;
    finish.reserve(N);
    if (N > 1) {
        std::vector<uint8_t> state(N + 1, 0);
        std::vector<std::pair<uint32_t, uint32_t>> st;
        for (uint32_t s = 1; s <= N; s++) {
            if (state[s]) continue;
            st.push_back({s, 0}); state[s] = 1;
            while (!st.empty()) {
                auto& top = st.back();
                if (top.second < out_desc[top.first].size()) {
                    uint32_t v = out_desc[top.first][top.second++];
                    if (state[v] == 0) { state[v] = 1; st.push_back({v, 0}); }
                } else {
                    state[top.first] = 2; finish.push_back(top.first); st.pop_back();
                }
            }
        }
        std::reverse(finish.begin(), finish.end());
    } else {
        finish.push_back(1);  // single-node graph: no sort (graph.go:134-136)
    }
    std::vector<uint32_t> topo_of_seg(N + 1, 0);
    for (uint32_t i = 0; i < N; i++) topo_of_seg[finish[i]] = i;

    // ---- append to the flat index ----
    const uint32_t g = idx.n_graphs++;
    if (idx.graph_node_base.empty()) { idx.graph_node_base.push_back(0); idx.graph_path_base.push_back(0); }
    const uint32_t node_base = idx.graph_node_base.back();
    const uint32_t path_base = idx.graph_path_base.back();
    const uint32_t mw = (R + 31) / 32;
    idx.graph_mask_words.push_back(mw);
    idx.nodes.resize(node_base + N);
    idx.node_mask.resize(idx.node_mask.size() + static_cast<size_t>(N) * mw, 0);
    const size_t mask_base = idx.node_mask.size() - static_cast<size_t>(N) * mw;
    std::vector<uint32_t> surv_by_seg(N + 1);
    for (uint32_t u : survivors) surv_by_seg[seg_of[u]] = u;
    for (uint32_t t = 0; t < N; t++) {
        uint32_t seg = finish[t];
        const ColNode& c = cn[surv_by_seg[seg]];
        NodeRec& nr = idx.nodes[node_base + t];
        nr.seg_id = seg;
        nr.seq_off = static_cast<uint32_t>(idx.node_seq.size());
        nr.seq_len = static_cast<uint32_t>(c.seq.size());
        for (char ch : c.seq) idx.node_seq.push_back(normalise_base(static_cast<uint8_t>(ch)));
        nr.edge_off = static_cast<uint32_t>(idx.edges.size());
        nr.edge_cnt = static_cast<uint32_t>(out_desc[seg].size());
        for (uint32_t v : out_desc[seg]) idx.edges.push_back(node_base + topo_of_seg[v]);
        nr.path_off = static_cast<uint32_t>(idx.node_path_id.size());
        nr.path_cnt = static_cast<uint32_t>(c.rows.size());
        nr.mask_off = static_cast<uint32_t>(mask_base + static_cast<size_t>(t) * mw);
        for (uint32_t r : c.rows) {
            idx.node_path_id.push_back(r);
            idx.node_path_pos.push_back(0);  // filled below
            idx.node_mask[nr.mask_off + r / 32] |= 1u << (r % 32);
        }
    }
    // positions (GetPaths, graph.go:586-610): walk every path through its own node chain
    bool masked = false;
    uint64_t raw = 0;
    for (uint32_t r = 0; r < R; r++) {
        int32_t pos = 0;
        int64_t prev = -1;
        for (uint32_t id : chain[r]) {
            uint32_t head = id;
            while (cn[head].merged_into >= 0) head = static_cast<uint32_t>(cn[head].merged_into);
            if (static_cast<int64_t>(head) == prev) continue;
            prev = head;
            NodeRec& nr = idx.nodes[node_base + topo_of_seg[seg_of[head]]];
            for (uint32_t j = 0; j < nr.path_cnt; j++)
                if (idx.node_path_id[nr.path_off + j] == r) { idx.node_path_pos[nr.path_off + j] = pos; break; }
            pos += static_cast<int32_t>(nr.seq_len);
        }
        idx.path_name.push_back(rows[r].name);
        idx.path_len.push_back(pos);
        if (pos < static_cast<int32_t>(idx.p.w)) masked = true;                 // pipeline/index.go:59-65
        raw += static_cast<uint64_t>(std::max<int64_t>(0, static_cast<int64_t>(pos) - idx.p.w + 1));
    }
    idx.graph_masked.push_back(masked ? 1 : 0);
    idx.graph_raw_windows.push_back(masked ? 0 : raw);
    idx.graph_node_base.push_back(node_base + N);
    idx.graph_path_base.push_back(path_base + R);
    idx.kmer_freq.resize(node_base + N, 0.0);
    idx.node_marked.resize(node_base + N, 0);
    idx.kmer_total.push_back(0);
    (void)g;
}

// ---------------------------------------------------------------------------------------------
namespace {
struct PendingWin {
    uint32_t node, offset, merge_span, seg_id, arrival;
    const uint64_t* sketch;
    std::vector<std::pair<uint32_t, uint32_t>> cn;  // (global node, count), ascending seg id
};
}  // namespace

void build_windows(FlatIndex& idx, SketchFn sketch_fn, void* ctx) {
    const uint32_t w = idx.p.w, k = idx.p.k, S = idx.p.S;
    if (w < k) throw std::runtime_error("window size must be >= k-mer size");
    // A. spell every path of every unmasked graph; window i of a path starts at path_begin + i
    std::vector<uint8_t> seqs;
    std::vector<uint64_t> off;
    struct PathSpan { uint32_t g, p; uint64_t begin; uint32_t len; uint64_t first_win; std::vector<uint32_t> nodes; };
    std::vector<PathSpan> spans;
    for (uint32_t g = 0; g < idx.n_graphs; g++) {
        if (idx.graph_masked[g]) continue;
        uint32_t nb = idx.graph_node_base[g], ne = idx.graph_node_base[g + 1];
        for (uint32_t p = 0; p < idx.n_paths_of(g); p++) {
            PathSpan sp; sp.g = g; sp.p = p; sp.begin = seqs.size(); sp.first_win = off.size();
            for (uint32_t n = nb; n < ne; n++) {
                const NodeRec& nr = idx.nodes[n];
                if (!(idx.node_mask[nr.mask_off + p / 32] >> (p % 32) & 1u)) continue;
                sp.nodes.push_back(n);
                seqs.insert(seqs.end(), idx.node_seq.begin() + nr.seq_off, idx.node_seq.begin() + nr.seq_off + nr.seq_len);
            }
            sp.len = static_cast<uint32_t>(seqs.size() - sp.begin);
            if (static_cast<int32_t>(sp.len) != idx.path_len[idx.graph_path_base[g] + p]) throw std::runtime_error("windowing did not traverse entire path");
            for (uint32_t i = 0; i + w <= sp.len; i++) off.push_back(sp.begin + i);
            spans.push_back(std::move(sp));
        }
    }
    if (off.empty()) throw std::runtime_error("could not create and sketch any graphs");
    // B. sketch all windows in one go (GPU)
    std::vector<uint64_t> sk(off.size() * static_cast<size_t>(S));
    sketch_fn(ctx, seqs.data(), seqs.size(), off.data(), static_cast<uint32_t>(off.size()), w, k, S, sk.data());

    // C/D. per graph: merge runs of identical sketches per path, drop the final run once anything was
    //      emitted (graph.go:285,305,336-338), merge identical (node, offset, sketch) across paths
    idx.wins.clear(); idx.cn_node.clear(); idx.cn_count.clear(); idx.sketches.clear();
    size_t si = 0;
    std::vector<uint32_t> base_node, base_off;
    while (si < spans.size()) {
        const uint32_t g = spans[si].g;
        std::vector<PendingWin> pend;
        std::unordered_map<uint64_t, std::vector<uint32_t>> by_loc;  // (node, offset) -> pend indices
        for (; si < spans.size() && spans[si].g == g; si++) {
            const PathSpan& sp = spans[si];
            base_node.resize(sp.len); base_off.resize(sp.len);
            std::vector<uint32_t> node_begin(sp.nodes.size() + 1, 0);
            uint32_t it = 0;
            for (size_t j = 0; j < sp.nodes.size(); j++) {
                node_begin[j] = it;
                uint32_t L = idx.nodes[sp.nodes[j]].seq_len;
                for (uint32_t o = 0; o < L; o++) { base_node[it] = static_cast<uint32_t>(j); base_off[it] = o; it++; }
            }
            node_begin[sp.nodes.size()] = it;
            const uint32_t nwin = sp.len - w + 1;
            bool sent = false;
            uint32_t i0 = 0;
            auto emit = [&](uint32_t a, uint32_t b) {  // windows a..b (inclusive) share one sketch
                PendingWin pw;
                pw.node = sp.nodes[base_node[a]]; pw.offset = base_off[a]; pw.merge_span = b - a;
                pw.seg_id = idx.nodes[pw.node].seg_id; pw.sketch = &sk[(sp.first_win + a) * S];
                // ContainedNodes[seg]++ per base per merged window (graph.go:326-328)
                uint32_t j0 = base_node[a], j1 = base_node[b + w - 1];
                for (uint32_t j = j0; j <= j1; j++) {
                    uint64_t cnt = 0;
                    uint32_t xb = std::max(node_begin[j], a), xe = std::min(node_begin[j + 1], b + w);
                    for (uint32_t x = xb; x < xe; x++) {
                        uint32_t lo = std::max<int64_t>(static_cast<int64_t>(x) - w + 1, a), hi = std::min(x, b);
                        cnt += hi - lo + 1;
                    }
                    pw.cn.push_back({sp.nodes[j], static_cast<uint32_t>(cnt)});
                }
                std::sort(pw.cn.begin(), pw.cn.end(), [&](auto& x, auto& y) { return idx.nodes[x.first].seg_id < idx.nodes[y.first].seg_id; });
                // cross-path merge at the same node+offset with an identical sketch (graph.go:354-387):
                // frequencies add up; Ref append and MergeSpan max are lost in the reference (copy semantics)
                uint64_t loc = (static_cast<uint64_t>(pw.node) << 32) | pw.offset;
                auto& lst = by_loc[loc];
                for (uint32_t e : lst) {
                    if (memcmp(pend[e].sketch, pw.sketch, 8 * S) == 0) {
                        auto& dst = pend[e].cn;
                        for (auto& kv : pw.cn) {
                            auto pos = std::find_if(dst.begin(), dst.end(), [&](auto& d) { return d.first == kv.first; });
                            if (pos != dst.end()) pos->second += kv.second;
                            else dst.push_back(kv);
                        }
                        std::sort(dst.begin(), dst.end(), [&](auto& x, auto& y) { return idx.nodes[x.first].seg_id < idx.nodes[y.first].seg_id; });
                        return;
                    }
                }
                pw.arrival = static_cast<uint32_t>(pend.size());
                lst.push_back(pw.arrival);
                pend.push_back(std::move(pw));
            };
            for (uint32_t i = 1; i < nwin; i++) {
                if (memcmp(&sk[(sp.first_win + i) * S], &sk[(sp.first_win + i0) * S], 8 * S) != 0) { emit(i0, i - 1); sent = true; i0 = i; }
            }
            if (!sent) emit(i0, nwin - 1);
        }
        std::vector<uint32_t> order(pend.size());
        for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return std::tie(pend[a].seg_id, pend[a].offset, pend[a].arrival) < std::tie(pend[b].seg_id, pend[b].offset, pend[b].arrival);
        });
        for (uint32_t o : order) {
            const PendingWin& pw = pend[o];
            WinRec wr;
            wr.graph = g; wr.node = pw.node; wr.offset = pw.offset; wr.merge_span = pw.merge_span; wr.win_size = w;
            wr.cn_off = static_cast<uint32_t>(idx.cn_node.size()); wr.cn_cnt = static_cast<uint32_t>(pw.cn.size()); wr.seg_id = pw.seg_id;
            for (auto& kv : pw.cn) { idx.cn_node.push_back(kv.first); idx.cn_count.push_back(kv.second); }
            idx.wins.push_back(wr);
            idx.sketches.insert(idx.sketches.end(), pw.sketch, pw.sketch + S);
        }
    }
    if (idx.wins.empty()) throw std::runtime_error("no sketches produced after windowing graph seqs");
}

}  // namespace groot
