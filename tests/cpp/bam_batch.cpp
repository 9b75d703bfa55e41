// Test / timing harness for the host driver's BAM output (format_batch_bam + BamWriter, host only — no device): reads a
// fabricated batch (reads, compact pairs, path ids, node -> path table, @SQ list) from a file written by
// tests/test_bam_cpu.py, writes the BAM, prints "<seconds> <uncompressed record bytes> <bam bytes> <delta blocks> <of those: own Huffman code> <zlib blocks>".
//   bam_batch <batch file> <out.bam> <workers> <level> <delta 0|1> [repeats]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../groot_b200/csrc/host/pipeline.h"

using namespace groot_host;

namespace {
struct Reader {
    FILE* f;
    void get(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) throw std::runtime_error("short batch file"); }
    template <class T> T one() { T v; get(&v, sizeof v); return v; }
    template <class V> void vec(V& v, size_t n) { v.resize(n); get(v.data(), n * sizeof(typename V::value_type)); }
};
}  // namespace

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: bam_batch <batch> <out.bam> <workers> <level> <delta> [repeats]\n"); return 1; }
    const unsigned workers = static_cast<unsigned>(atoi(argv[3]));
    const int level = atoi(argv[4]);
    const bool delta = atoi(argv[5]) != 0;
    const int repeats = argc > 6 ? atoi(argv[6]) : 1;
    try {
        Reader r{fopen(argv[1], "rb")};
        if (!r.f) throw std::runtime_error("cannot open batch file");
        char magic[4]; r.get(magic, 4);
        if (memcmp(magic, "BAMT", 4)) throw std::runtime_error("not a batch file");
        const uint32_t n_reads = r.one<uint32_t>(), n_pairs = r.one<uint32_t>();
        const uint64_t n_records = r.one<uint64_t>();
        const uint32_t path_bytes = r.one<uint32_t>(), n_graphs = r.one<uint32_t>(), n_refs = r.one<uint32_t>(), n_nodes = r.one<uint32_t>();
        ReadBatch b;
        r.vec(b.id_off, n_reads + 1); r.vec(b.seq_off, n_reads + 1); r.vec(b.qual_off, n_reads + 1);
        r.vec(b.id, b.id_off.back()); r.vec(b.seq, b.seq_off.back()); r.vec(b.qual, b.qual_off.back());
        std::vector<grootgpu_cpair> cpairs; r.vec(cpairs, n_pairs);
        std::vector<uint8_t> rec_path; r.vec(rec_path, n_records * path_bytes);
        std::vector<uint32_t> graph_ref_base; r.vec(graph_ref_base, n_graphs + 1);
        std::vector<std::pair<std::string, int32_t>> refs(n_refs);
        for (auto& ref : refs) { std::vector<char> nm; r.vec(nm, r.one<uint32_t>()); ref.first.assign(nm.begin(), nm.end()); ref.second = r.one<int32_t>(); }
        std::vector<uint32_t> node_graph(n_nodes), node_off(n_nodes + 1, 0), ids;
        std::vector<int32_t> pos;
        for (uint32_t n = 0; n < n_nodes; n++) {
            node_graph[n] = r.one<uint32_t>();
            const uint32_t k = r.one<uint32_t>();
            std::vector<uint32_t> i; std::vector<int32_t> p;
            r.vec(i, k); r.vec(p, k);
            ids.insert(ids.end(), i.begin(), i.end()); pos.insert(pos.end(), p.begin(), p.end());
            node_off[n + 1] = node_off[n] + k;
        }
        fclose(r.f);

        BamBatch bb;
        bb.reads = &b; bb.cpairs = cpairs.data(); bb.n_pairs = n_pairs; bb.n_records = n_records;
        bb.rec_path_c = rec_path.data(); bb.rec_path_bytes = path_bytes; bb.graph_ref_base = graph_ref_base.data();
        bb.node_paths = [&](uint32_t node, NodePathsView* v) {
            if (node >= n_nodes) return false;
            v->graph = node_graph[node]; v->ids = ids.data() + node_off[node]; v->pos = pos.data() + node_off[node]; v->n = node_off[node + 1] - node_off[node];
            return true;
        };
        uint64_t raw_bytes = 0;
        for (const grootgpu_cpair& p : cpairs) {
            if (!p.rec_count) continue;
            const uint32_t cs = (p.offset_flags & GROOTGPU_CPAIR_CLIP_START) ? 1 : 0, ce = (p.offset_flags & GROOTGPU_CPAIR_CLIP_END) ? 1 : 0;
            const uint64_t il = b.id_off[p.read + 1] - b.id_off[p.read], sl = b.seq_off[p.read + 1] - b.seq_off[p.read];
            raw_bytes += static_cast<uint64_t>(p.rec_count) * BamWriter::record_size(static_cast<uint32_t>(il ? il - 1 : 0), cs, static_cast<uint32_t>(sl) - cs - ce, ce);
        }
        FILE* out = fopen(argv[2], "wb");
        if (!out) throw std::runtime_error("cannot open output");
        double best = 1e30;
        uint64_t bam_bytes = 0;
        BamBlockStats st;
        for (int it = 0; it < repeats; it++) {
            std::vector<std::vector<uint8_t>> outs;
            const auto t0 = std::chrono::steady_clock::now();
            const std::string err = format_batch_bam(bb, workers, level, delta, outs, &st);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (!err.empty()) throw std::runtime_error(err);
            best = std::min(best, dt);
            if (it == 0) {
                BamWriter w(out, "@HD\tVN:1.5\tSO:unknown\n", refs, level);
                for (auto& o : outs) { w.append_blocks(o); bam_bytes += o.size(); }
                w.close();
            }
        }
        fclose(out);
        printf("%.6f %llu %llu %llu %llu %llu\n", best, static_cast<unsigned long long>(raw_bytes), static_cast<unsigned long long>(bam_bytes),
               static_cast<unsigned long long>(st.delta), static_cast<unsigned long long>(st.delta_own_code), static_cast<unsigned long long>(st.zlib));
    } catch (std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    return 0;
}
