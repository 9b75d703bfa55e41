// ORACLE — TEST INFRASTRUCTURE ONLY (see nthash.hpp header).
//
// CPU restatement of the two pipelines around the hot path:
//   * `groot index`: src/pipeline/index.go:37-211 (MSAconverter -> GraphSketcher -> SketchIndexer)
//   * `groot align`: src/pipeline/boss.go:108-242 (sketch worker loop) and
//     src/pipeline/graphminion.go:40-103 (per-graph weighting + alignment loop), emulating `-p 1`:
//     reads in input order, graphs in ascending GraphID, mappings in ascending (Node, OffSet, dup idx).
#pragma once
#include <algorithm>
#include <atomic>
#include <dirent.h>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "align.hpp"
#include "lshe.hpp"

namespace oracle {

struct Index {
    int kmerSize = 31, sketchSize = 21, windowSize = 100, numPart = 8, maxK = 4;
    std::map<uint32_t, std::shared_ptr<GrootGraph>> store;  // graph.Store
    ContainmentIndex db;
};

inline std::vector<std::string> list_msas(const std::string& dir) {
    // filepath.Glob(msaDir + "/cluster*.msa") — returned sorted lexicographically (cmd/index.go:143)
    std::vector<std::string> out;
    DIR* d = opendir(dir.c_str());
    if (!d) throw std::runtime_error("cannot open MSA dir " + dir);
    while (dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n.rfind("cluster", 0) == 0 && n.size() > 4 && n.substr(n.size() - 4) == ".msa") out.push_back(dir + "/" + n);
    }
    closedir(d);
    std::sort(out.begin(), out.end());
    return out;
}

// index.go:37-211 — msaTexts[i] becomes graph i
inline std::shared_ptr<Index> build_index(const std::vector<std::string>& msaTexts, int k, int S, int w, int numPart, int maxK) {
    auto idx = std::make_shared<Index>();
    idx->kmerSize = k; idx->sketchSize = S; idx->windowSize = w; idx->numPart = numPart; idx->maxK = maxK;
    idx->db.init(numPart, maxK, w - k + 1, S);
    struct Tmp { uint32_t g; uint64_t node; uint32_t off; int dup; std::string name; Key key; };
    std::vector<Tmp> all;
    for (size_t i = 0; i < msaTexts.size(); i++) {
        Gfa gfa = msa2gfa(read_msa_text(msaTexts[i]));
        auto g = create_groot_graph(gfa, static_cast<int>(i));
        for (auto& kv : g->lengths) if (kv.second < w) { g->masked = true; break; }  // index.go:59-65
        if (!g->masked) {
            auto windows = window_graph(*g, w, k, S);
            for (auto& kv : windows)
                for (size_t j = 0; j < kv.second.size(); j++) {
                    const Key& key = kv.second[j];
                    all.push_back({key.graphID, key.node, key.offSet, static_cast<int>(j), kv.first + "-" + std::to_string(j), key});
                }
        }
        idx->store[g->graphID] = g;
    }
    std::sort(all.begin(), all.end(), [](const Tmp& a, const Tmp& b) {
        return std::tie(a.g, a.node, a.off, a.dup) < std::tie(b.g, b.node, b.off, b.dup);
    });
    for (auto& t : all) idx->db.addWindow(t.name, t.key);
    if (all.empty()) throw std::runtime_error("could not create and sketch any graphs");
    idx->db.bootstrap();
    return idx;
}

// result of the minion loop for one (read, graph)
struct GraphResult {
    uint32_t graphID = 0;
    uint32_t hitBegin = 0, hitCount = 0;  // slice of ReadResult::hits
    uint32_t numIncremented = 0;          // mappings that received IncrementSubPath before the loop ended
    std::vector<AlignRecord> records;
};
struct ReadResult {
    std::vector<uint64_t> sketch;
    std::vector<uint32_t> hits;  // window indices ascending == (graph, node, offset, dup) order
    std::vector<GraphResult> graphs;
};
struct Counts { uint64_t received = 0, mapped = 0, multimapped = 0, alignments = 0; };

// boss.go:145-201 + graphminion.go:46-102 for one read, without touching graph weights
inline void map_read(Index& idx, const FASTQread& readIn, double threshold, bool noAlign, ReadResult* out) {
    out->hits.clear(); out->graphs.clear();
    if (!run_minhash(readIn.seq, idx.kmerSize, idx.sketchSize, &out->sketch))
        throw std::runtime_error("read shorter than k (reference panics at boss.go:164-166)");
    int kmerCount = static_cast<int>(readIn.seq.size()) - idx.kmerSize + 1;  // boss.go:169
    idx.db.query(out->sketch.data(), kmerCount, threshold, &out->hits);
    size_t i = 0;
    while (i < out->hits.size()) {
        GraphResult gr;
        gr.graphID = idx.db.windows[out->hits[i]].graphID;
        gr.hitBegin = static_cast<uint32_t>(i);
        size_t j = i;
        while (j < out->hits.size() && idx.db.windows[out->hits[j]].graphID == gr.graphID) j++;
        gr.hitCount = static_cast<uint32_t>(j - i);
        GrootGraph& g = *idx.store.at(gr.graphID);
        FASTQread read = readIn;  // DeepCopy / struct copy (boss.go:184-191)
        bool found = false;
        for (size_t m = i; m < j; m++) {
            gr.numIncremented++;  // graphminion.go:67 (applied by replay_weights)
            if (noAlign) continue;
            const Key& mapping = idx.db.windows[out->hits[m]];
            for (int s = 0; s < 2; s++) {
                auto recs = align_read(g, read, mapping);
                if (!recs.empty()) { gr.records = recs; found = true; break; }
                rev_complement(&read);  // graphminion.go:94
            }
            if (found) break;
        }
        out->graphs.push_back(std::move(gr));
        i = j;
    }
}

// graphminion.go:60,67 — ordered replay of IncrementSubPath for one read
inline void replay_weights(Index& idx, size_t readLen, const ReadResult& rr) {
    double kmerCount = static_cast<double>(static_cast<int>(readLen) - idx.kmerSize) + 1.0;
    for (auto& gr : rr.graphs) {
        GrootGraph& g = *idx.store.at(gr.graphID);
        for (uint32_t m = 0; m < gr.numIncremented; m++)
            g.incrementSubPath(idx.db.windows[rr.hits[gr.hitBegin + m]].containedNodes, kmerCount);
    }
}

inline void map_reads(Index& idx, const std::vector<FASTQread>& reads, double threshold, bool noAlign, int threads,
                      std::vector<ReadResult>* results, Counts* counts, bool keepSketches = false) {
    results->assign(reads.size(), ReadResult());
    // pre-warm the (K,L) cache so worker threads only read it
    {
        std::map<size_t, bool> lens;
        for (auto& r : reads) lens[r.seq.size()] = true;
        for (auto& kv : lens) for (int x : idx.db.partUpper) idx.db.params(x, static_cast<int>(kv.first) - idx.kmerSize + 1, threshold);
    }
    int T = std::max(1, threads);
    std::atomic<size_t> next{0};
    std::vector<std::string> errors(T);
    auto worker = [&](int t) {
        try {
            const size_t chunk = 256;
            while (true) {
                size_t b = next.fetch_add(chunk);
                if (b >= reads.size()) break;
                size_t e = std::min(reads.size(), b + chunk);
                for (size_t r = b; r < e; r++) {
                    map_read(idx, reads[r], threshold, noAlign, &(*results)[r]);
                    if (!keepSketches) { (*results)[r].sketch.clear(); (*results)[r].sketch.shrink_to_fit(); }
                }
            }
        } catch (std::exception& ex) { errors[t] = ex.what(); }
    };
    if (T == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(worker, t);
        for (auto& x : th) x.join();
    }
    for (auto& e : errors) if (!e.empty()) throw std::runtime_error(e);
    *counts = Counts();
    for (size_t r = 0; r < reads.size(); r++) {
        auto& rr = (*results)[r];
        replay_weights(idx, reads[r].seq.size(), rr);
        counts->received++;
        if (!rr.graphs.empty()) counts->mapped++;
        if (rr.graphs.size() > 1) counts->multimapped++;
        for (auto& gr : rr.graphs) counts->alignments += gr.records.size();
    }
}

}  // namespace oracle
