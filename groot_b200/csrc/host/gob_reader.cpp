// Reader for the reference's own index files: groot.gg (pipeline.Info incl. the graph Store, src/pipeline/runtime.go:15-27,
// 64-91; graph.GrootGraph / GrootGraphNode, src/graph/graph.go:18-34, src/graph/node.go:13-22) and groot.lshe
// (lshe.ContainmentIndex with WindowLookup map[string]Key, src/lshe/lshe.go:17-49, 72-146), both written with Go's
// encoding/gob. With them the library adopts the Go host's graphs as they are — node order inside SortedNodes, segment
// ids, path ids — instead of rebuilding them from the MSAs with its own numbering of bubble nodes.
//
// encoding/gob is a self-describing stream: messages of [byte count][type id][payload]; a negative type id introduces the
// definition (a wireType value) of type -id; ints are zig-zag varints with a length-prefixed big-endian form above 127,
// floats are byte-reversed IEEE bits, structs are (field delta, value)* 0 with zero-valued fields omitted, slices and maps
// are a count followed by the elements. No Go toolchain exists in this image: the decoder follows the published format
// description and is tested against streams produced by an independent Python restatement of the ENCODER
// (tests/gob_writer.py), i.e. gob interop is "parity unpinned" against real Go output.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <ios>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../flat_index.h"

namespace groot {
namespace {

struct GobValue {
    enum Kind { Nil, Bool, Int, Uint, Float, Bytes, List, Nums, NumMap, Map, Struct } kind = Nil;
    uint64_t u = 0;                 // Bool / Uint
    int64_t i = 0;                  // Int
    double f = 0.0;                 // Float
    std::string bytes;              // []byte / string
    std::vector<GobValue> list;     // slice / array of non-scalar elements
    std::vector<int64_t> nums;      // slice of integer scalars; NumMap: key, value, key, value ... (float values as their IEEE bits)
    bool nummap_float = false;      // NumMap: the values are floats
    std::vector<std::pair<GobValue, GobValue>> map;           // map with non-scalar key or value
    std::vector<std::pair<std::string, GobValue>> fields;     // struct: only the fields present in the stream
    const GobValue* field(const char* name) const {
        for (auto& f : fields) if (f.first == name) return &f.second;
        return nullptr;
    }
    int64_t int_field(const char* name, int64_t def = 0) const {
        const GobValue* v = field(name);
        if (!v) return def;
        if (v->kind == Int) return v->i;
        if (v->kind == Uint || v->kind == Bool) return static_cast<int64_t>(v->u);
        throw std::runtime_error(std::string("gob: field ") + name + " is not an integer");
    }
    double float_field(const char* name) const { const GobValue* v = field(name); return v ? v->f : 0.0; }
};

enum : int { tBool = 1, tInt = 2, tUint = 3, tFloat = 4, tBytes = 5, tString = 6, tComplex = 7, tInterface = 8,
             tWireType = 16, tArrayType = 17, tCommonType = 18, tSliceType = 19, tStructType = 20, tFieldType = 21, tFieldTypeSlice = 22, tMapType = 23 };

struct TypeDef {
    enum Kind { Struct, Slice, Array, Map } kind = Struct;
    std::string name;
    int elem = 0, key = 0;
    std::vector<std::pair<std::string, int>> fields;
};

class GobDecoder {
  public:
    explicit GobDecoder(const std::string& path) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) throw std::ios_base::failure("cannot open " + path);
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        buf_.resize(n > 0 ? static_cast<size_t>(n) : 0);
        if (n > 0 && fread(&buf_[0], 1, buf_.size(), f) != buf_.size()) { fclose(f); throw std::ios_base::failure("cannot read " + path); }
        fclose(f);
        if (buf_.empty()) throw std::runtime_error("gob: " + path + " appears empty");   // runtime.go:86-88, lshe.go:101-103
        bootstrap();
    }
    // the first value of the stream (type definitions before it are absorbed)
    GobValue decode_value() {
        while (pos_ < buf_.size()) {
            const uint64_t len = get_uint();
            if (len > buf_.size() - pos_) throw std::runtime_error("gob: message runs past the end of the file");
            const size_t end = pos_ + len;
            const int64_t id = get_int();
            if (id < 0) {
                define_type(static_cast<int>(-id));
            } else {
                const int tid = static_cast<int>(id);
                if (!is_struct(tid)) { if (get_uint() != 0) throw std::runtime_error("gob: missing zero byte before a non-struct value"); }
                GobValue v = value(tid);
                if (pos_ != end) throw std::runtime_error("gob: value does not fill its message");
                return v;
            }
            if (pos_ != end) throw std::runtime_error("gob: type definition does not fill its message");
        }
        throw std::runtime_error("gob: no value in the stream");
    }

  private:
    std::string buf_;
    size_t pos_ = 0;
    std::map<int, TypeDef> types_;

    uint8_t byte() { if (pos_ >= buf_.size()) throw std::runtime_error("gob: unexpected end of file"); return static_cast<uint8_t>(buf_[pos_++]); }
    uint64_t get_uint() {
        const uint8_t b = byte();
        if (b < 128) return b;
        const int n = 256 - b;                       // the byte holds the NEGATED byte count
        if (n < 1 || n > 8) throw std::runtime_error("gob: bad integer length");
        uint64_t v = 0;
        for (int k = 0; k < n; k++) v = (v << 8) | byte();
        return v;
    }
    int64_t get_int() {
        const uint64_t u = get_uint();
        return (u & 1) ? static_cast<int64_t>(~(u >> 1)) : static_cast<int64_t>(u >> 1);
    }
    double get_float() {
        uint64_t u = get_uint(), r = 0;              // byte-reversed IEEE-754 bits
        for (int k = 0; k < 8; k++) { r = (r << 8) | (u & 0xff); u >>= 8; }
        double d;
        memcpy(&d, &r, 8);
        return d;
    }
    void bootstrap() {
        auto st = [&](int id, const char* name, std::initializer_list<std::pair<const char*, int>> fl) {
            TypeDef t; t.kind = TypeDef::Struct; t.name = name;
            for (auto& f : fl) t.fields.push_back({f.first, f.second});
            types_[id] = t;
        };
        st(tWireType, "wireType", {{"ArrayT", tArrayType}, {"SliceT", tSliceType}, {"StructT", tStructType}, {"MapT", tMapType}});
        st(tArrayType, "arrayType", {{"CommonType", tCommonType}, {"Elem", tInt}, {"Len", tInt}});
        st(tCommonType, "CommonType", {{"Name", tString}, {"Id", tInt}});
        st(tSliceType, "sliceType", {{"CommonType", tCommonType}, {"Elem", tInt}});
        st(tStructType, "structType", {{"CommonType", tCommonType}, {"Field", tFieldTypeSlice}});
        st(tFieldType, "fieldType", {{"Name", tString}, {"Id", tInt}});
        st(tMapType, "mapType", {{"CommonType", tCommonType}, {"Key", tInt}, {"Elem", tInt}});
        TypeDef fs; fs.kind = TypeDef::Slice; fs.name = "[]*fieldType"; fs.elem = tFieldType;
        types_[tFieldTypeSlice] = fs;
    }
    bool is_struct(int id) const { auto it = types_.find(id); return it != types_.end() && it->second.kind == TypeDef::Struct; }
    static bool is_int_scalar(int id) { return id == tBool || id == tInt || id == tUint; }

    void define_type(int id) {
        const GobValue w = value(tWireType);
        TypeDef t;
        auto common = [&](const GobValue& v) { if (const GobValue* c = v.field("CommonType")) if (const GobValue* nm = c->field("Name")) t.name = nm->bytes; };
        if (const GobValue* s = w.field("StructT")) {
            t.kind = TypeDef::Struct; common(*s);
            if (const GobValue* fl = s->field("Field"))
                for (const GobValue& f : fl->list) {
                    const GobValue* nm = f.field("Name");
                    t.fields.push_back({nm ? nm->bytes : std::string(), static_cast<int>(f.int_field("Id"))});
                }
        } else if (const GobValue* s2 = w.field("SliceT")) {
            t.kind = TypeDef::Slice; common(*s2); t.elem = static_cast<int>(s2->int_field("Elem"));
        } else if (const GobValue* a = w.field("ArrayT")) {
            t.kind = TypeDef::Array; common(*a); t.elem = static_cast<int>(a->int_field("Elem"));
        } else if (const GobValue* m = w.field("MapT")) {
            t.kind = TypeDef::Map; common(*m); t.key = static_cast<int>(m->int_field("Key")); t.elem = static_cast<int>(m->int_field("Elem"));
        } else throw std::runtime_error("gob: unsupported type definition (GobEncoder / marshaler types are not used by groot)");
        types_[id] = t;
    }

    // Info / Store / ContainmentIndex nest six deep; a damaged type table can make a type contain itself, and every level
    // of a struct costs only one byte of input — so the depth is bounded here, not by the file size
    static constexpr int kMaxDepth = 64;
    int depth_ = 0;
    struct DepthGuard {
        int& d;
        explicit DepthGuard(int& depth) : d(depth) { if (++d > kMaxDepth) { --d; throw std::runtime_error("gob: values nested too deeply"); } }
        ~DepthGuard() { --d; }
    };

    GobValue value(int id) {
        GobValue v;
        const DepthGuard guard(depth_);
        switch (id) {
            case tBool: v.kind = GobValue::Bool; v.u = get_uint(); return v;
            case tInt: v.kind = GobValue::Int; v.i = get_int(); return v;
            case tUint: v.kind = GobValue::Uint; v.u = get_uint(); return v;
            case tFloat: v.kind = GobValue::Float; v.f = get_float(); return v;
            case tBytes: case tString: {
                v.kind = GobValue::Bytes;
                const uint64_t n = get_uint();
                if (n > buf_.size() - pos_) throw std::runtime_error("gob: string runs past the end of the file");
                v.bytes.assign(buf_, pos_, n); pos_ += n;
                return v;
            }
            case tComplex: case tInterface: throw std::runtime_error("gob: complex / interface values are not used by groot");
            default: break;
        }
        auto it = types_.find(id);
        if (it == types_.end()) throw std::runtime_error("gob: value of an undefined type " + std::to_string(id));
        const TypeDef& t = it->second;
        if (t.kind == TypeDef::Struct) {
            v.kind = GobValue::Struct;
            int fieldnum = -1;
            while (true) {
                const uint64_t delta = get_uint();
                if (delta == 0) break;
                fieldnum += static_cast<int>(delta);
                if (fieldnum < 0 || fieldnum >= static_cast<int>(t.fields.size())) throw std::runtime_error("gob: field number out of range in " + t.name);
                v.fields.push_back({t.fields[fieldnum].first, value(t.fields[fieldnum].second)});
            }
            return v;
        }
        if (t.kind == TypeDef::Slice || t.kind == TypeDef::Array) {
            const uint64_t n = get_uint();
            if (n > buf_.size() - pos_) throw std::runtime_error("gob: implausible element count");
            if (is_int_scalar(t.elem)) {
                v.kind = GobValue::Nums; v.nums.resize(n);
                for (uint64_t k = 0; k < n; k++) v.nums[k] = t.elem == tInt ? get_int() : static_cast<int64_t>(get_uint());
            } else {
                v.kind = GobValue::List; v.list.reserve(std::min<uint64_t>(n, 1u << 16));   // n is only bounded by the bytes left
                for (uint64_t k = 0; k < n; k++) v.list.push_back(value(t.elem));
            }
            return v;
        }
        // map
        const uint64_t n = get_uint();
        if (n > buf_.size() - pos_) throw std::runtime_error("gob: implausible element count");
        if (is_int_scalar(t.key) && (is_int_scalar(t.elem) || t.elem == tFloat)) {
            v.kind = GobValue::NumMap; v.nummap_float = t.elem == tFloat; v.nums.resize(2 * n);
            for (uint64_t k = 0; k < n; k++) {
                v.nums[2 * k] = t.key == tInt ? get_int() : static_cast<int64_t>(get_uint());
                if (t.elem == tFloat) { const double d = get_float(); memcpy(&v.nums[2 * k + 1], &d, 8); }
                else v.nums[2 * k + 1] = t.elem == tInt ? get_int() : static_cast<int64_t>(get_uint());
            }
        } else {
            v.kind = GobValue::Map; v.map.reserve(std::min<uint64_t>(n, 1u << 16));
            for (uint64_t k = 0; k < n; k++) { GobValue key = value(t.key); GobValue val = value(t.elem); v.map.push_back({std::move(key), std::move(val)}); }
        }
        return v;
    }
};

int64_t key_as_int(const GobValue& k) {
    if (k.kind == GobValue::Uint || k.kind == GobValue::Bool) return static_cast<int64_t>(k.u);
    if (k.kind == GobValue::Int) return k.i;
    throw std::runtime_error("gob: integer map key expected");
}
double nummap_float(const GobValue& m, size_t k) { double d; memcpy(&d, &m.nums[2 * k + 1], 8); return d; }

}  // namespace

// groot.gg + groot.lshe -> flat index. Graph i of the result is Store[i] (GraphIDs must be 0..G-1, as cmd/index.go numbers them).
void load_index_gob(FlatIndex& idx, const std::string& gg_path, const std::string& lshe_path) {
    idx = FlatIndex();
    const GobValue info = GobDecoder(gg_path).decode_value();
    if (info.kind != GobValue::Struct) throw std::runtime_error("gob: groot.gg does not hold a pipeline.Info");
    idx.p.k = static_cast<uint32_t>(info.int_field("KmerSize")); idx.p.S = static_cast<uint32_t>(info.int_field("SketchSize"));
    idx.p.w = static_cast<uint32_t>(info.int_field("WindowSize")); idx.p.num_part = static_cast<uint32_t>(info.int_field("NumPart"));
    idx.p.max_k = static_cast<uint32_t>(info.int_field("MaxK"));
    const GobValue* store = info.field("Store");
    if (!store || store->kind != GobValue::Map || store->map.empty()) throw std::runtime_error("gob: groot.gg holds no graphs");
    std::vector<const GobValue*> graphs(store->map.size(), nullptr);
    for (auto& kv : store->map) {
        const int64_t g = key_as_int(kv.first);
        if (g < 0 || g >= static_cast<int64_t>(graphs.size()) || graphs[g]) throw std::runtime_error("gob: graph ids are not 0..G-1");
        graphs[g] = &kv.second;
    }
    idx.n_graphs = static_cast<uint32_t>(graphs.size());
    idx.graph_node_base.push_back(0); idx.graph_path_base.push_back(0);
    std::vector<std::map<uint64_t, uint32_t>> node_of_seg(graphs.size());     // per graph: SegmentID -> global node index
    for (size_t g = 0; g < graphs.size(); g++) {
        const GobValue& gr = *graphs[g];
        if (gr.int_field("GraphID") != static_cast<int64_t>(g)) throw std::runtime_error("gob: GraphID does not match its Store key");
        // paths: Paths map[uint32][]byte, Lengths map[uint32]int
        const GobValue* paths = gr.field("Paths");
        const uint32_t np = paths ? static_cast<uint32_t>(paths->map.size()) : 0u;
        std::vector<std::string> names(np);
        std::vector<int32_t> lens(np, 0);
        if (paths) for (auto& kv : paths->map) { const int64_t p = key_as_int(kv.first); if (p < 0 || p >= np) throw std::runtime_error("gob: path ids are not 0..P-1"); names[p] = kv.second.bytes; }
        if (const GobValue* ln = gr.field("Lengths"))
            for (size_t k = 0; k < ln->nums.size() / 2; k++) { const int64_t p = ln->nums[2 * k]; if (p < 0 || p >= np) throw std::runtime_error("gob: Lengths key out of range"); lens[p] = static_cast<int32_t>(ln->nums[2 * k + 1]); }
        for (uint32_t p = 0; p < np; p++) { idx.path_name.push_back(names[p]); idx.path_len.push_back(lens[p]); }
        const uint32_t mw = (np + 31) / 32;
        idx.graph_mask_words.push_back(mw);
        idx.graph_masked.push_back(gr.int_field("Masked") ? 1 : 0);
        idx.graph_raw_windows.push_back(0);                       // numWindows is unexported: not in the file
        idx.kmer_total.push_back(static_cast<uint64_t>(gr.int_field("KmerTotal")));
        // nodes, in SortedNodes order
        const GobValue* sn = gr.field("SortedNodes");
        const uint32_t nb = static_cast<uint32_t>(idx.nodes.size());
        const uint32_t nn = sn ? static_cast<uint32_t>(sn->list.size()) : 0u;
        for (uint32_t n = 0; n < nn; n++) node_of_seg[g][static_cast<uint64_t>(sn->list[n].int_field("SegmentID"))] = nb + n;
        if (node_of_seg[g].size() != nn) throw std::runtime_error("gob: duplicate SegmentID in a graph");
        for (uint32_t n = 0; n < nn; n++) {
            const GobValue& nd = sn->list[n];
            NodeRec r{};
            r.seg_id = static_cast<uint32_t>(nd.int_field("SegmentID"));
            const GobValue* seq = nd.field("Sequence");
            r.seq_off = static_cast<uint32_t>(idx.node_seq.size());
            r.seq_len = seq ? static_cast<uint32_t>(seq->bytes.size()) : 0u;
            if (seq) idx.node_seq.insert(idx.node_seq.end(), seq->bytes.begin(), seq->bytes.end());
            r.edge_off = static_cast<uint32_t>(idx.edges.size());
            if (const GobValue* oe = nd.field("OutEdges"))
                for (int64_t e : oe->nums) {                       // segment ids, in the reference's order (graph.go:203)
                    auto it = node_of_seg[g].find(static_cast<uint64_t>(e));
                    if (it == node_of_seg[g].end()) throw std::runtime_error("gob: out-edge to an unknown segment");
                    idx.edges.push_back(it->second);
                }
            r.edge_cnt = static_cast<uint32_t>(idx.edges.size()) - r.edge_off;
            // PathIDs + Position map[int]int -> parallel arrays, path ids ascending
            std::map<int64_t, int64_t> position;
            if (const GobValue* ps = nd.field("Position")) for (size_t k = 0; k < ps->nums.size() / 2; k++) position[ps->nums[2 * k]] = ps->nums[2 * k + 1];
            std::vector<int64_t> pids;
            if (const GobValue* pi = nd.field("PathIDs")) pids = pi->nums;
            std::sort(pids.begin(), pids.end());
            r.path_off = static_cast<uint32_t>(idx.node_path_id.size());
            r.mask_off = static_cast<uint32_t>(idx.node_mask.size());
            idx.node_mask.resize(idx.node_mask.size() + mw, 0u);
            for (int64_t p : pids) {
                if (p < 0 || p >= np) throw std::runtime_error("gob: PathID out of range");
                idx.node_path_id.push_back(static_cast<uint32_t>(p));
                auto it = position.find(p);
                idx.node_path_pos.push_back(it == position.end() ? 0 : static_cast<int32_t>(it->second));
                idx.node_mask[r.mask_off + p / 32] |= 1u << (p % 32);
            }
            r.path_cnt = static_cast<uint32_t>(pids.size());
            idx.nodes.push_back(r);
            idx.kmer_freq.push_back(nd.float_field("KmerFreq"));
            idx.node_marked.push_back(nd.int_field("Marked") ? 1 : 0);
        }
        idx.graph_node_base.push_back(static_cast<uint32_t>(idx.nodes.size()));
        idx.graph_path_base.push_back(static_cast<uint32_t>(idx.path_name.size()));
    }
    // ---- groot.lshe: WindowLookup map[string]Key ----
    const GobValue ci = GobDecoder(lshe_path).decode_value();
    if (ci.kind != GobValue::Struct) throw std::runtime_error("gob: groot.lshe does not hold an lshe.ContainmentIndex");
    if (ci.int_field("SketchSize") != idx.p.S || ci.int_field("MaxK") != idx.p.max_k || ci.int_field("NumPart") != idx.p.num_part ||
        ci.int_field("NumWindowKmers") != static_cast<int64_t>(idx.p.w) - idx.p.k + 1)
        throw std::runtime_error("gob: groot.lshe and groot.gg disagree on the index parameters");
    const GobValue* wl = ci.field("WindowLookup");
    if (!wl || wl->kind != GobValue::Map || wl->map.empty()) throw std::runtime_error("loaded an empty index file");   // lshe.go:103-105
    struct Win { uint32_t graph, seg, off, arrival; const GobValue* key; };
    std::vector<Win> wins;
    wins.reserve(wl->map.size());
    for (auto& kv : wl->map) {
        const GobValue& k = kv.second;
        Win w;
        w.graph = static_cast<uint32_t>(k.int_field("GraphID")); w.seg = static_cast<uint32_t>(k.int_field("Node")); w.off = static_cast<uint32_t>(k.int_field("OffSet"));
        // the lookup string is "g%dn%do%d-%d" (graph.go:361, pipeline/index.go:197): its last number tells windows of one (graph, node, offset) apart
        const std::string& name = kv.first.bytes;
        const size_t dash = name.rfind('-');
        w.arrival = dash == std::string::npos ? 0u : static_cast<uint32_t>(strtoul(name.c_str() + dash + 1, nullptr, 10));
        w.key = &k;
        if (w.graph >= idx.n_graphs) throw std::runtime_error("gob: window of an unknown graph");
        wins.push_back(w);
    }
    std::sort(wins.begin(), wins.end(), [](const Win& a, const Win& b) {
        if (a.graph != b.graph) return a.graph < b.graph;
        if (a.seg != b.seg) return a.seg < b.seg;
        if (a.off != b.off) return a.off < b.off;
        return a.arrival < b.arrival;
    });
    for (const Win& w : wins) {
        const GobValue& k = *w.key;
        WinRec r{};
        r.graph = w.graph; r.seg_id = w.seg; r.offset = w.off;
        auto it = node_of_seg[w.graph].find(w.seg);
        if (it == node_of_seg[w.graph].end()) throw std::runtime_error("gob: window on an unknown segment");
        r.node = it->second;
        r.merge_span = static_cast<uint32_t>(k.int_field("MergeSpan")); r.win_size = static_cast<uint32_t>(k.int_field("WindowSize"));
        r.cn_off = static_cast<uint32_t>(idx.cn_node.size());
        if (const GobValue* cn = k.field("ContainedNodes")) {
            std::map<uint64_t, double> by_seg;                     // ascending SegmentID
            for (size_t j = 0; j < cn->nums.size() / 2; j++) by_seg[static_cast<uint64_t>(cn->nums[2 * j])] = cn->nummap_float ? nummap_float(*cn, j) : static_cast<double>(cn->nums[2 * j + 1]);
            for (auto& sv : by_seg) {
                auto nt = node_of_seg[w.graph].find(sv.first);
                if (nt == node_of_seg[w.graph].end()) throw std::runtime_error("gob: contained node is an unknown segment");
                idx.cn_node.push_back(nt->second);
                idx.cn_count.push_back(static_cast<uint32_t>(sv.second));
            }
        }
        r.cn_cnt = static_cast<uint32_t>(idx.cn_node.size()) - r.cn_off;
        const GobValue* sk = k.field("Sketch");
        if (!sk || sk->nums.size() != idx.p.S) throw std::runtime_error("gob: a window sketch has the wrong size");
        for (int64_t v : sk->nums) idx.sketches.push_back(static_cast<uint64_t>(v));
        idx.wins.push_back(r);
    }
}

}  // namespace groot
