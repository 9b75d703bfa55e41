// Writer for the reference's own index files — the other direction of host/gob_reader.cpp: a flat index becomes groot.gg
// (pipeline.Info with the graph Store, src/pipeline/runtime.go:15-27,64-73; graph.GrootGraph / GrootGraphNode,
// src/graph/graph.go:18-34, src/graph/node.go:13-22) and groot.lshe (lshe.ContainmentIndex with WindowLookup
// map[string]Key, src/lshe/lshe.go:17-49,72-92) in Go's encoding/gob, so that an index built on the GPU
// (`groot-b200 index`) can be loaded by the Go `groot align` / `groot haplotype`.
//
// Stream layout (see gob_reader.cpp for the format): for every type that is not built in, one message
// [-id][wireType value] — component types before the types that use them — then ONE message [id][value] with the whole
// Info / ContainmentIndex. Struct fields holding a zero value are omitted, nil slices too, maps are always sent.
// No Go toolchain exists in this image: the output is checked byte for byte against the independent Python encoder
// tests/gob_writer.py (pinned to the documentation's example streams) and through the reader; interop with real Go
// stays "parity unpinned" (DESIGN.md).
#include <cstdio>
#include <cstring>
#include <ios>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../flat_index.h"

namespace groot {
namespace {

enum : int { tBool = 1, tInt = 2, tUint = 3, tFloat = 4, tBytes = 5, tString = 6 };

void put_uint(std::string& o, uint64_t u) {
    if (u < 128) { o.push_back(static_cast<char>(u)); return; }
    char b[8]; int n = 0;
    while (u) { b[n++] = static_cast<char>(u & 0xff); u >>= 8; }
    o.push_back(static_cast<char>(256 - n));                         // the negated byte count, then big-endian bytes
    while (n) o.push_back(b[--n]);
}
void put_int(std::string& o, int64_t i) { put_uint(o, i < 0 ? (~static_cast<uint64_t>(i) << 1) | 1u : static_cast<uint64_t>(i) << 1); }
void put_float(std::string& o, double f) {
    uint64_t u, r = 0;
    memcpy(&u, &f, 8);
    for (int k = 0; k < 8; k++) { r = (r << 8) | (u & 0xff); u >>= 8; }   // byte-reversed IEEE bits: small exponents first
    put_uint(o, r);
}
void put_bytes(std::string& o, const void* p, size_t n) { put_uint(o, n); o.append(static_cast<const char*>(p), n); }
void put_string(std::string& o, const std::string& s) { put_bytes(o, s.data(), s.size()); }

// ---- type descriptors: what the type-definition messages are generated from ------------------------------------------
struct Type {
    enum Kind { Builtin, Struct, Slice, Map } kind = Builtin;
    int builtin = 0;
    std::string name;
    std::vector<std::pair<std::string, const Type*>> fields;
    const Type* elem = nullptr; const Type* key = nullptr;
    mutable int id = 0;
};
struct Types {                                                       // owns the descriptors of one stream
    std::vector<std::unique_ptr<Type>> all;
    const Type* builtin(int b) { all.emplace_back(new Type()); all.back()->builtin = b; all.back()->id = b; return all.back().get(); }
    const Type* slice(const Type* e, const char* name = "") { all.emplace_back(new Type()); Type& t = *all.back(); t.kind = Type::Slice; t.elem = e; t.name = name; return &t; }
    const Type* map(const Type* k, const Type* e, const char* name = "") { all.emplace_back(new Type()); Type& t = *all.back(); t.kind = Type::Map; t.key = k; t.elem = e; t.name = name; return &t; }
    const Type* strct(const char* name, std::vector<std::pair<std::string, const Type*>> f) { all.emplace_back(new Type()); Type& t = *all.back(); t.kind = Type::Struct; t.name = name; t.fields = std::move(f); return &t; }
};

class Stream {
  public:
    // ids are handed out from 65 in the order the types are first met walking the top-level type's fields; a type's
    // definition is sent once everything it refers to has been
    int define(const Type* t) {
        if (t->id) return t->id;
        const int id = next_id_++;
        t->id = id;
        std::string body;
        auto common = [&](std::string& o) {                        // CommonType{Name, Id}
            if (!t->name.empty()) { put_uint(o, 1); put_string(o, t->name); put_uint(o, 1); put_int(o, id); }
            else { put_uint(o, 2); put_int(o, id); }
            o.push_back(0);
        };
        if (t->kind == Type::Struct) {
            std::vector<int> kids;
            for (auto& f : t->fields) kids.push_back(define(f.second));
            put_uint(body, 3); put_uint(body, 1); common(body);     // wireType.StructT { CommonType,
            put_uint(body, 1); put_uint(body, t->fields.size());    //   Field []*fieldType{Name, Id} }
            for (size_t i = 0; i < t->fields.size(); i++) { put_uint(body, 1); put_string(body, t->fields[i].first); put_uint(body, 1); put_int(body, kids[i]); body.push_back(0); }
            body.push_back(0); body.push_back(0);
        } else if (t->kind == Type::Slice) {
            const int e = define(t->elem);
            put_uint(body, 2); put_uint(body, 1); common(body); put_uint(body, 1); put_int(body, e); body.push_back(0); body.push_back(0);   // SliceT{CommonType, Elem}
        } else {
            const int k = define(t->key), e = define(t->elem);
            put_uint(body, 4); put_uint(body, 1); common(body); put_uint(body, 1); put_int(body, k); put_uint(body, 1); put_int(body, e); body.push_back(0); body.push_back(0);   // MapT{CommonType, Key, Elem}
        }
        std::string msg;
        put_int(msg, -id);
        msg += body;
        message(msg);
        return id;
    }
    void message(const std::string& body) { put_uint(out_, body.size()); out_ += body; }
    // everything in front of the encoded value of a top-level struct: the type definitions, the byte count of the value
    // message and its type id (the value itself — hundreds of MB for a large index — is written behind it without a copy)
    const std::string& prefix(const Type* t, size_t encoded_size) {
        std::string id;
        put_int(id, define(t));
        put_uint(out_, id.size() + encoded_size);
        out_ += id;
        return out_;
    }
  private:
    std::string out_;
    int next_id_ = 65;
};

// (field delta, value)* 0 with the zero-valued fields left out
struct StructOut {
    std::string& o; int prev = -1;
    explicit StructOut(std::string& out) : o(out) {}
    void field(int i) { put_uint(o, static_cast<uint64_t>(i - prev)); prev = i; }
    void u(int i, uint64_t v) { if (v) { field(i); put_uint(o, v); } }
    void i64(int i, int64_t v) { if (v) { field(i); put_int(o, v); } }
    void f(int i, double v) { if (v != 0.0) { field(i); put_float(o, v); } }
    void b(int i, bool v) { if (v) { field(i); put_uint(o, 1); } }
    void s(int i, const std::string& v) { if (!v.empty()) { field(i); put_string(o, v); } }
    void end() { o.push_back(0); }
};

void write_file(const std::string& path, const std::string& head, const std::string& data) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::ios_base::failure("cannot create " + path);
    const bool ok = fwrite(head.data(), 1, head.size(), f) == head.size() && fwrite(data.data(), 1, data.size(), f) == data.size();
    if (fclose(f) != 0 || !ok) throw std::ios_base::failure("cannot write " + path);
}

}  // namespace

void save_index_gob(const FlatIndex& ix, const std::string& gg_path, const std::string& lshe_path, const std::string& version) {
    // ---- groot.gg: pipeline.Info ----
    {
        Types T;
        const Type *Bool = T.builtin(tBool), *Int = T.builtin(tInt), *Uint = T.builtin(tUint), *Float = T.builtin(tFloat), *Bytes = T.builtin(tBytes), *String = T.builtin(tString);
        const Type* node = T.strct("GrootGraphNode", {{"SegmentID", Uint}, {"SegmentLength", Float}, {"Sequence", Bytes}, {"OutEdges", T.slice(Uint, "Nodes")},
                                                      {"PathIDs", T.slice(Uint)}, {"Position", T.map(Int, Int)}, {"KmerFreq", Float}, {"Marked", Bool}});
        const Type* graph = T.strct("GrootGraph", {{"GrootVersion", String}, {"GraphID", Uint}, {"SortedNodes", T.slice(node)}, {"Paths", T.map(Uint, Bytes)},
                                                   {"Lengths", T.map(Uint, Int)}, {"NodeLookup", T.map(Uint, Int)}, {"Masked", Bool}, {"KmerTotal", Uint}, {"EMiterations", Int}});
        const Type* aligncmd = T.strct("AlignCmd", {{"Fasta", Bool}, {"BloomFilter", Bool}, {"MinKmerCoverage", Float}, {"BAMout", String}, {"NoExactAlign", Bool}});
        const Type* haplocmd = T.strct("HaploCmd", {{"Cutoff", Float}, {"MinIterations", Int}, {"MaxIterations", Int}, {"TotalKmers", Int}, {"HaploDir", String}});
        const Type* info = T.strct("Info", {{"Version", String}, {"NumProc", Int}, {"Profiling", Bool}, {"KmerSize", Int}, {"SketchSize", Int}, {"WindowSize", Int},
                                            {"NumPart", Int}, {"MaxK", Int}, {"MaxSketchSpan", Int}, {"ContainmentThreshold", Float}, {"IndexDir", String},
                                            {"Store", T.map(Uint, graph, "Store")}, {"Sketch", aligncmd}, {"Haplotype", haplocmd}});
        std::string v;
        StructOut I(v);
        I.s(0, version); I.i64(1, 1); I.i64(3, ix.p.k); I.i64(4, ix.p.S); I.i64(5, ix.p.w); I.i64(6, ix.p.num_part); I.i64(7, ix.p.max_k);
        I.i64(8, 30);                                                // MaxSketchSpan: cmd/index.go's default; `groot align` does not read it
        I.f(9, 0.99); I.s(10, "index");                              // overwritten by `groot align` from its own flags (cmd/align.go:108-118)
        I.field(11);                                                 // Store map[uint32]*GrootGraph
        put_uint(v, ix.n_graphs);
        for (uint32_t g = 0; g < ix.n_graphs; g++) {
            put_uint(v, g);
            StructOut G(v);
            G.s(0, version); G.u(1, g);
            const uint32_t n0 = ix.graph_node_base[g], n1 = ix.graph_node_base[g + 1];
            if (n1 > n0) {
                G.field(2);                                          // SortedNodes
                put_uint(v, n1 - n0);
                for (uint32_t n = n0; n < n1; n++) {
                    const NodeRec& nd = ix.nodes[n];
                    StructOut N(v);
                    N.u(0, nd.seg_id); N.f(1, static_cast<double>(nd.seq_len));
                    if (nd.seq_len) { N.field(2); put_bytes(v, ix.node_seq.data() + nd.seq_off, nd.seq_len); }
                    if (nd.edge_cnt) { N.field(3); put_uint(v, nd.edge_cnt); for (uint32_t e = 0; e < nd.edge_cnt; e++) put_uint(v, ix.nodes[ix.edges[nd.edge_off + e]].seg_id); }
                    if (nd.path_cnt) { N.field(4); put_uint(v, nd.path_cnt); for (uint32_t p = 0; p < nd.path_cnt; p++) put_uint(v, ix.node_path_id[nd.path_off + p]); }
                    N.field(5);                                      // Position map[int]int
                    put_uint(v, nd.path_cnt);
                    for (uint32_t p = 0; p < nd.path_cnt; p++) { put_int(v, ix.node_path_id[nd.path_off + p]); put_int(v, ix.node_path_pos[nd.path_off + p]); }
                    N.f(6, n < ix.kmer_freq.size() ? ix.kmer_freq[n] : 0.0);
                    N.b(7, n < ix.node_marked.size() && ix.node_marked[n]);
                    N.end();
                }
            }
            const uint32_t p0 = ix.graph_path_base[g], p1 = ix.graph_path_base[g + 1];
            G.field(3); put_uint(v, p1 - p0); for (uint32_t p = p0; p < p1; p++) { put_uint(v, p - p0); put_string(v, ix.path_name[p]); }     // Paths
            G.field(4); put_uint(v, p1 - p0); for (uint32_t p = p0; p < p1; p++) { put_uint(v, p - p0); put_int(v, ix.path_len[p]); }         // Lengths
            G.field(5); put_uint(v, n1 - n0); for (uint32_t n = n0; n < n1; n++) { put_uint(v, ix.nodes[n].seg_id); put_int(v, n - n0); }     // NodeLookup
            G.b(6, ix.graph_masked[g] != 0);
            G.u(7, g < ix.kmer_total.size() ? ix.kmer_total[g] : 0);
            G.end();
        }
        I.field(12); { StructOut A(v); A.f(2, 1.0); A.end(); }       // Sketch: AlignCmd{MinKmerCoverage: 1.0}
        I.field(13); v.push_back(0);                                 // Haplotype: HaploCmd{}
        I.end();
        Stream s;
        write_file(gg_path, s.prefix(info, v.size()), v);
    }
    // ---- groot.lshe: lshe.ContainmentIndex ----
    {
        Types T;
        const Type *Bool = T.builtin(tBool), *Int = T.builtin(tInt), *Uint = T.builtin(tUint), *Float = T.builtin(tFloat), *String = T.builtin(tString);
        const Type* key = T.strct("Key", {{"GraphID", Uint}, {"Node", Uint}, {"OffSet", Uint}, {"ContainedNodes", T.map(Uint, Float)}, {"Ref", T.slice(Uint)}, {"RC", Bool},
                                          {"Sketch", T.slice(Uint)}, {"Freq", Float}, {"MergeSpan", Uint}, {"WindowSize", Uint}});
        const Type* cindex = T.strct("ContainmentIndex", {{"NumPart", Int}, {"MaxK", Int}, {"NumWindowKmers", Int}, {"SketchSize", Int}, {"WindowLookup", T.map(String, key)}});
        std::string v;
        StructOut C(v);
        C.i64(0, ix.p.num_part); C.i64(1, ix.p.max_k); C.i64(2, static_cast<int64_t>(ix.p.w) - ix.p.k + 1); C.i64(3, ix.p.S);
        C.field(4);
        put_uint(v, ix.wins.size());
        uint32_t arrival = 0;
        for (size_t w = 0; w < ix.wins.size(); w++) {
            const WinRec& wr = ix.wins[w];
            // "g%dn%do%d-%d" (graph.go:361, pipeline/index.go:197): the last number tells the windows of one (graph, node, offset) apart
            arrival = w > 0 && ix.wins[w - 1].graph == wr.graph && ix.wins[w - 1].seg_id == wr.seg_id && ix.wins[w - 1].offset == wr.offset ? arrival + 1 : 0;
            char name[96];
            snprintf(name, sizeof name, "g%un%uo%u-%u", wr.graph, wr.seg_id, wr.offset, arrival);
            put_string(v, name);
            StructOut K(v);
            K.u(0, wr.graph); K.u(1, wr.seg_id); K.u(2, wr.offset);
            K.field(3);                                              // ContainedNodes map[uint64]float64
            put_uint(v, wr.cn_cnt);
            for (uint32_t j = 0; j < wr.cn_cnt; j++) { put_uint(v, ix.nodes[ix.cn_node[wr.cn_off + j]].seg_id); put_float(v, static_cast<double>(ix.cn_count[wr.cn_off + j])); }
            K.field(4); put_uint(v, 1); put_uint(v, 0);              // Ref: not kept by the flat index, not read by align
            K.field(6); put_uint(v, ix.p.S); for (uint32_t s = 0; s < ix.p.S; s++) put_uint(v, ix.sketches[w * ix.p.S + s]);
            K.u(8, wr.merge_span); K.u(9, wr.win_size);
            K.end();
        }
        C.end();
        Stream s;
        write_file(lshe_path, s.prefix(cindex, v.size()), v);
    }
}

}  // namespace groot
