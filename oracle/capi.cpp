// ORACLE — TEST INFRASTRUCTURE ONLY (see nthash.hpp header).
//
// extern "C" surface of the CPU oracle, loaded with ctypes by oracle/pyoracle.py. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs call it.
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "pipeline.hpp"

using namespace oracle;

namespace {
struct Result {
    std::vector<ReadResult> reads;
    Counts counts;
    std::vector<uint32_t> readLens;
};
struct Fnv {
    uint64_t h = 1469598103934665603ULL;
    void add(const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ULL; } }
};
void set_err(char* err, int errlen, const std::string& m) { if (err && errlen > 0) { snprintf(err, errlen, "%s", m.c_str()); } }

// Canonical text dump of an index, shared FORMAT (not code) with groot_b200's own dump
// (groot_b200/csrc/host/index_dump.cpp): the parity test compares hashes / text of the two.
void dump_index(Index& idx, FILE* f, Fnv* fnv) {
    std::string line;
    auto emit = [&]() { line.push_back('\n'); if (f) fwrite(line.data(), 1, line.size(), f); if (fnv) fnv->add(line.data(), line.size()); line.clear(); };
    char buf[256];
    snprintf(buf, sizeof buf, "I k=%d S=%d w=%d numPart=%d maxK=%d graphs=%zu windows=%zu", idx.kmerSize, idx.sketchSize, idx.windowSize, idx.numPart, idx.maxK, idx.store.size(), idx.db.windows.size());
    line = buf; emit();
    for (auto& gk : idx.store) {
        GrootGraph& g = *gk.second;
        snprintf(buf, sizeof buf, "G %u masked=%d paths=%zu nodes=%zu", g.graphID, g.masked ? 1 : 0, g.paths.size(), g.sortedNodes.size());
        line = buf; emit();
        for (auto& pk : g.paths) { snprintf(buf, sizeof buf, "P %u %d ", pk.first, g.lengths[pk.first]); line = buf; line += pk.second; emit(); }
        for (auto& n : g.sortedNodes) {
            snprintf(buf, sizeof buf, "N %llu ", static_cast<unsigned long long>(n->segmentID)); line = buf; line += n->sequence; line += " E";
            for (uint64_t e : n->outEdges) { snprintf(buf, sizeof buf, " %llu", static_cast<unsigned long long>(e)); line += buf; }
            line += " P";
            for (uint32_t p : n->pathIDs) { auto it = n->position.find(static_cast<int>(p)); snprintf(buf, sizeof buf, " %u:%d", p, it == n->position.end() ? 0 : it->second); line += buf; }
            emit();
        }
    }
    for (size_t w = 0; w < idx.db.windows.size(); w++) {
        const Key& k = idx.db.windows[w];
        snprintf(buf, sizeof buf, "W %u %llu %u span=%u w=%u S", k.graphID, static_cast<unsigned long long>(k.node), k.offSet, k.mergeSpan, k.windowSize); line = buf;
        for (uint64_t s : k.sketch) { snprintf(buf, sizeof buf, " %016llx", static_cast<unsigned long long>(s)); line += buf; }
        line += " C";
        for (auto& cn : k.containedNodes) { snprintf(buf, sizeof buf, " %llu:%.0f", static_cast<unsigned long long>(cn.first), cn.second); line += buf; }
        emit();
    }
}
}  // namespace

extern "C" {

// ---- known-answer helpers ----
uint64_t oracle_ntf64(const uint8_t* s, unsigned k) { return ntf64(s, k); }
uint64_t oracle_ntr64(const uint8_t* s, unsigned k) { return ntr64(s, k); }
int oracle_sketch(const uint8_t* seq, uint64_t len, int k, int S, uint64_t* out) { return khf_sketch(seq, len, k, S, out) ? 0 : -1; }
// all canonical k-mer hashes of a sequence (rolling), for rolling == from-scratch checks
int64_t oracle_kmer_hashes(const uint8_t* seq, uint64_t len, int k, uint64_t* out) {
    NtHasher h(seq, len, k);
    if (!h.ok()) return -1;
    int64_t n = 0; uint64_t v;
    while (h.next(true, &v)) out[n++] = v;
    return n;
}
void oracle_multi_hash(uint64_t h, int k, int S, uint64_t* out) { multi_hash(h, k, S, out); }
int oracle_revcomp(uint8_t* seq, uint8_t* qual, uint64_t len) {
    FASTQread r; r.seq.assign(reinterpret_cast<char*>(seq), len); r.qual.assign(reinterpret_cast<char*>(qual), len);
    try { rev_complement(&r); } catch (std::exception&) { return -1; }
    memcpy(seq, r.seq.data(), len); memcpy(qual, r.qual.data(), len);
    return 0;
}
void oracle_basecheck(uint8_t* seq, uint64_t len) { std::string s(reinterpret_cast<char*>(seq), len); base_check(&s); memcpy(seq, s.data(), len); }
// returns new length; seq/qual are trimmed in place
uint64_t oracle_qualtrim(uint8_t* seq, uint8_t* qual, uint64_t len, int minQual) {
    FASTQread r; r.seq.assign(reinterpret_cast<char*>(seq), len); r.qual.assign(reinterpret_cast<char*>(qual), len);
    qual_trim(&r, minQual);
    memcpy(seq, r.seq.data(), r.seq.size()); memcpy(qual, r.qual.data(), r.qual.size());
    return r.seq.size();
}
void oracle_optimal_kl(int maxK, int maxL, int x, int q, double t, int* K, int* L) { optimal_kl(maxK, maxL, x, q, t, K, L); }
int oracle_eq_min(int S, int qSize, int xSize, double t) { return eq_min_for(S, qSize, xSize, t); }
double oracle_containment(const uint64_t* q, const uint64_t* x, int S, int qSize, int xSize) { return containment(q, x, S, qSize, xSize); }

// MSA -> GFA text (S/L/P lines) for the cluster-139 <-> test.gfa golden comparison
int64_t oracle_msa2gfa_text(const char* msaPath, char* out, int64_t cap) {
    try {
        Gfa g = msa2gfa(read_msa_text(slurp(msaPath)));
        std::string s = "H\tVN:Z:1\n";
        for (auto& seg : g.segments) s += "S\t" + seg.name + "\t" + seg.seq + "\n";
        for (auto& l : g.links) s += "L\t" + l.from + "\t+\t" + l.to + "\t+\t0M\n";
        for (auto& p : g.paths) { s += "P\t" + p.name + "\t"; for (size_t i = 0; i < p.segs.size(); i++) { if (i) s += ","; s += p.segs[i] + "+"; } s += "\n"; }
        if (static_cast<int64_t>(s.size()) + 1 > cap) return -static_cast<int64_t>(s.size()) - 1;
        memcpy(out, s.data(), s.size()); out[s.size()] = 0;
        return static_cast<int64_t>(s.size());
    } catch (std::exception& e) { return 0; }
}

// AlignRead against a graph loaded from a GFA file with a hand-made seed (src/graph/alignment_test.go
// fixtures). cnNodes: ContainedNodes keys (may be empty). Output rows: pathID,pos,flags,startClip,endClip,seqLength.
// names_out receives '\n'-joined path names in pathID order. Returns #records or -1.
int oracle_gfa_align(const char* gfaPath, int graphID, const char* readSeq, uint64_t node, uint32_t offset, uint32_t mergeSpan,
                     uint32_t windowSize, const uint64_t* cnNodes, int nCn, int32_t* out, int cap, char* names_out, int names_cap) {
    try {
        Gfa gfa = read_gfa_text(slurp(gfaPath));
        auto g = create_groot_graph(gfa, graphID);
        Key key; key.graphID = graphID; key.node = node; key.offSet = offset; key.mergeSpan = mergeSpan; key.windowSize = windowSize;
        for (int i = 0; i < nCn; i++) key.containedNodes[cnNodes[i]] = 1.0;
        FASTQread read; read.id = "@r"; read.seq = readSeq; read.qual.assign(read.seq.size(), '+');
        std::vector<AlignRecord> recs;
        for (int s = 0; s < 2 && recs.empty(); s++) {  // graphminion.go:76-95: forward, then reverse complement
            recs = align_read(*g, read, key);
            if (recs.empty()) rev_complement(&read);
        }
        int n = 0;
        for (auto& r : recs) { if (n >= cap) break; int32_t* o = out + 6 * n++; o[0] = r.pathID; o[1] = r.pos; o[2] = r.flags; o[3] = r.startClip; o[4] = r.endClip; o[5] = r.seqLength; }
        std::string names;
        for (auto& p : g->paths) { names += p.second; names += "\n"; }
        snprintf(names_out, names_cap, "%s", names.c_str());
        return n;
    } catch (std::exception& e) { return -1; }
}

// ---- index ----
void* oracle_index_build_files(const char** paths, int n, int k, int S, int w, int numPart, int maxK, char* err, int errlen) {
    try {
        std::vector<std::string> texts;
        for (int i = 0; i < n; i++) texts.push_back(slurp(paths[i]));
        auto idx = build_index(texts, k, S, w, numPart, maxK);
        return new std::shared_ptr<Index>(idx);
    } catch (std::exception& e) { set_err(err, errlen, e.what()); return nullptr; }
}
void* oracle_index_build_dir(const char* dir, int k, int S, int w, int numPart, int maxK, char* err, int errlen) {
    try {
        auto files = list_msas(dir);
        std::vector<const char*> p;
        for (auto& f : files) p.push_back(f.c_str());
        return oracle_index_build_files(p.data(), static_cast<int>(p.size()), k, S, w, numPart, maxK, err, errlen);
    } catch (std::exception& e) { set_err(err, errlen, e.what()); return nullptr; }
}
void oracle_index_free(void* h) { delete static_cast<std::shared_ptr<Index>*>(h); }
static Index& IDX(void* h) { return **static_cast<std::shared_ptr<Index>*>(h); }

// out[0..7] = graphs, masked graphs, paths, nodes, path bases, windows(keys), raw windows, max merge span
void oracle_index_stats(void* h, uint64_t* out) {
    Index& idx = IDX(h);
    uint64_t masked = 0, paths = 0, nodes = 0, bases = 0, raw = 0, span = 0;
    for (auto& gk : idx.store) {
        if (gk.second->masked) masked++;
        paths += gk.second->paths.size(); nodes += gk.second->sortedNodes.size();
        for (auto& l : gk.second->lengths) bases += l.second;
        if (!gk.second->masked) raw += gk.second->numWindows;
    }
    for (auto& w : idx.db.windows) span = std::max<uint64_t>(span, w.mergeSpan);
    out[0] = idx.store.size(); out[1] = masked; out[2] = paths; out[3] = nodes; out[4] = bases; out[5] = idx.db.windows.size(); out[6] = raw; out[7] = span;
}
uint64_t oracle_index_dump_hash(void* h) { Fnv f; dump_index(IDX(h), nullptr, &f); return f.h; }
int oracle_index_dump_file(void* h, const char* path) { FILE* f = fopen(path, "wb"); if (!f) return -1; dump_index(IDX(h), f, nullptr); fclose(f); return 0; }
void oracle_index_window_sketches(void* h, uint64_t* out) {
    Index& idx = IDX(h);
    for (size_t w = 0; w < idx.db.windows.size(); w++) memcpy(out + w * idx.sketchSize, idx.db.windows[w].sketch.data(), 8 * idx.sketchSize);
}
void oracle_index_params(void* h, int x, int q, double t, int* K, int* L) { auto kl = IDX(h).db.params(x, q, t); *K = kl.first; *L = kl.second; }

// node weights in (graph ascending, SortedNodes order); kmerTotal per graph ascending
uint64_t oracle_index_num_nodes(void* h) { uint64_t n = 0; for (auto& gk : IDX(h).store) n += gk.second->sortedNodes.size(); return n; }
void oracle_index_weights(void* h, double* kmerFreq, uint64_t* kmerTotal) {
    size_t i = 0, gi = 0;
    for (auto& gk : IDX(h).store) { for (auto& n : gk.second->sortedNodes) kmerFreq[i++] = n->kmerFreq; kmerTotal[gi++] = gk.second->kmerTotal; }
}
void oracle_index_reset_weights(void* h) { for (auto& gk : IDX(h).store) { for (auto& n : gk.second->sortedNodes) n->kmerFreq = 0; gk.second->kmerTotal = 0; } }

// GraphPruner (src/pipeline/sketch.go:378-430): prune every graph, return '\n'-joined names of the paths
// of surviving graphs (what CollectOutput returns; note: ALL paths of a kept graph are listed —
// the reference never deletes from g.Paths, graph.go:517-523). Destructive.
int64_t oracle_prune_paths(void* h, double minKmerCov, char* out, int64_t cap) {
    std::string s;
    for (auto& gk : IDX(h).store) {
        if (!gk.second->prune(minKmerCov)) continue;
        for (auto& p : gk.second->paths) { s += p.second; s += "\n"; }
    }
    if (static_cast<int64_t>(s.size()) + 1 > cap) return -static_cast<int64_t>(s.size()) - 1;
    memcpy(out, s.data(), s.size()); out[s.size()] = 0;
    return static_cast<int64_t>(s.size());
}
// paths of kept graphs whose length was not zeroed by pruning (== the P lines SaveGraphAsGFA writes, graphio.go:71-100)
int64_t oracle_gfa_text(void* h, uint32_t graphID, long totalKmers, char* out, int64_t cap) {
    auto it = IDX(h).store.find(graphID);
    if (it == IDX(h).store.end()) return 0;
    std::string s = it->second->toGFA(totalKmers);
    if (static_cast<int64_t>(s.size()) + 1 > cap) return -static_cast<int64_t>(s.size()) - 1;
    memcpy(out, s.data(), s.size()); out[s.size()] = 0;
    return static_cast<int64_t>(s.size());
}

// ---- align ----
void* oracle_map_reads(void* h, const uint8_t* seqs, const uint64_t* off, uint32_t n, double threshold, int noAlign, int threads, int keepSketches, char* err, int errlen) {
    try {
        Index& idx = IDX(h);
        std::vector<FASTQread> reads(n);
        auto res = std::make_unique<Result>();
        res->readLens.resize(n);
        for (uint32_t i = 0; i < n; i++) {
            reads[i].seq.assign(reinterpret_cast<const char*>(seqs + off[i]), off[i + 1] - off[i]);
            reads[i].qual.assign(reads[i].seq.size(), 'I');
            res->readLens[i] = static_cast<uint32_t>(reads[i].seq.size());
        }
        map_reads(idx, reads, threshold, noAlign != 0, threads, &res->reads, &res->counts, keepSketches != 0);
        return res.release();
    } catch (std::exception& e) { set_err(err, errlen, e.what()); return nullptr; }
}
void oracle_result_free(void* r) { delete static_cast<Result*>(r); }
void oracle_result_counts(void* r, uint64_t* out) { auto& c = static_cast<Result*>(r)->counts; out[0] = c.received; out[1] = c.mapped; out[2] = c.multimapped; out[3] = c.alignments; }
// sizes: out[0]=hits, out[1]=(read,graph) pairs, out[2]=records
void oracle_result_sizes(void* r, uint64_t* out) {
    out[0] = out[1] = out[2] = 0;
    for (auto& rr : static_cast<Result*>(r)->reads) { out[0] += rr.hits.size(); out[1] += rr.graphs.size(); for (auto& g : rr.graphs) out[2] += g.records.size(); }
}
void oracle_result_hits(void* r, uint64_t* hitOff, uint32_t* hits) {
    uint64_t o = 0; size_t i = 0;
    for (auto& rr : static_cast<Result*>(r)->reads) { hitOff[i++] = o; for (uint32_t hIdx : rr.hits) hits[o++] = hIdx; }
    hitOff[i] = o;
}
void oracle_result_sketches(void* r, int S, uint64_t* out) {
    size_t i = 0;
    for (auto& rr : static_cast<Result*>(r)->reads) { if (rr.sketch.size() == static_cast<size_t>(S)) memcpy(out + i * S, rr.sketch.data(), 8 * S); i++; }
}
// pairs: read, graph, numIncremented, numRecords
void oracle_result_pairs(void* r, uint32_t* out) {
    size_t o = 0; uint32_t ri = 0;
    for (auto& rr : static_cast<Result*>(r)->reads) { for (auto& g : rr.graphs) { out[o++] = ri; out[o++] = g.graphID; out[o++] = g.numIncremented; out[o++] = static_cast<uint32_t>(g.records.size()); } ri++; }
}
// records in (read, graph, emission) order: read, graph, path, pos, flags, startClip, endClip, seqLength
void oracle_result_records(void* r, int32_t* out) {
    size_t o = 0; int32_t ri = 0;
    for (auto& rr : static_cast<Result*>(r)->reads) {
        for (auto& g : rr.graphs) for (auto& a : g.records) {
            out[o++] = ri; out[o++] = static_cast<int32_t>(g.graphID); out[o++] = a.pathID; out[o++] = a.pos; out[o++] = a.flags; out[o++] = a.startClip; out[o++] = a.endClip; out[o++] = a.seqLength;
        }
        ri++;
    }
}
// reference name / length of (graph, path) — header @SQ content (boss.go:45-105, graphio.go:141-154)
int oracle_ref_name(void* h, uint32_t graphID, uint32_t pathID, char* out, int cap, int* length) {
    auto it = IDX(h).store.find(graphID);
    if (it == IDX(h).store.end()) return -1;
    auto p = it->second->paths.find(pathID);
    if (p == it->second->paths.end()) return -1;
    snprintf(out, cap, "%s", p->second.c_str());
    *length = it->second->lengths[pathID];
    return 0;
}

}  // extern "C"
