// A stand-in for libgrootgpu.so on a machine without a GPU — TEST INFRASTRUCTURE ONLY, never shipped or linked into
// the product. It implements the handful of C-ABI entry points the host driver's ReadMapper / GraphPruner call
// (include/grootgpu.h) with a deterministic fake "alignment": whether a read maps, where, on which strand and to how
// many paths is a pure function of an FNV-1a hash of its bases, which tests/test_mapper_cpu.py restates to build the BAM
// it expects. What this buys: the threading and buffer-lifetime logic of the driver (reader thread, BAM stage, rank
// team, "results are valid until the next align call") runs under pytest and the sanitizers on the CPU.
//   * result arrays live in ONE buffer per handle that every align call overwrites (and poisons first), so a driver
//     that reads them after the next call fails the comparison;
//   * the multi-GPU calls are emulated with threads: grootgpu_gather merges the ranks' shards on rank 0.
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/grootgpu.h"

struct grootgpu_index {
    uint32_t n_graphs, paths_per_graph, n_nodes;
    std::vector<std::string> ref_names;
    std::vector<uint32_t> path_ids;                 // 0..P-1 (every node lies on every path of its graph)
    std::vector<int32_t> positions;                 // [n_nodes * P]
    std::vector<grootgpu_cpair> cpairs;             // the handle's result arrays
    std::vector<uint8_t> rec_path;
    uint64_t kmer_total = 0;
    int calls = 0;
};

namespace {
thread_local std::string g_err;
int fail(int code, const char* msg) { g_err = msg; return code; }
uint64_t fnv1a(const uint8_t* p, size_t n) { uint64_t h = 1469598103934665603ull; for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; } return h; }

struct CommShared {
    std::mutex mu; std::condition_variable cv;
    int world = 0, arrived = 0, handles = 0; uint64_t gen = 0, id = 0;
    std::vector<const grootgpu_batch_result*> local;
    std::vector<grootgpu_cpair> m_cpairs; std::vector<uint8_t> m_rec_path;   // rank 0's merged arrays
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const uint64_t g = gen;
        if (++arrived == world) { arrived = 0; gen++; cv.notify_all(); } else cv.wait(lk, [&] { return gen != g; });
    }
};
std::mutex g_comm_mu;
std::map<uint64_t, CommShared*> g_comms;
uint64_t g_next_id = 1;
}  // namespace

struct grootgpu_comm { CommShared* sh; int rank; grootgpu_index* idx; };

extern "C" {

grootgpu_index* mock_index_create(uint32_t n_graphs, uint32_t paths_per_graph, uint32_t n_nodes) {
    grootgpu_index* ix = new grootgpu_index();
    ix->n_graphs = n_graphs; ix->paths_per_graph = paths_per_graph; ix->n_nodes = n_nodes;
    for (uint32_t g = 0; g < n_graphs; g++) for (uint32_t p = 0; p < paths_per_graph; p++) ix->ref_names.push_back("gene_" + std::to_string(g) + "_" + std::to_string(p));
    for (uint32_t p = 0; p < paths_per_graph; p++) ix->path_ids.push_back(p);
    for (uint32_t n = 0; n < n_nodes; n++) for (uint32_t p = 0; p < paths_per_graph; p++) ix->positions.push_back(static_cast<int32_t>((n * 7u + p * 13u) % 1000u));
    return ix;
}
void grootgpu_index_destroy(grootgpu_index* ix) { delete ix; }
const char* grootgpu_last_error(void) { return g_err.c_str(); }

int grootgpu_index_get_info(const grootgpu_index* ix, grootgpu_index_info* out) {
    memset(out, 0, sizeof *out);
    out->n_graphs = ix->n_graphs; out->n_paths = ix->n_graphs * ix->paths_per_graph; out->n_nodes = ix->n_nodes; out->max_paths_per_graph = ix->paths_per_graph;
    return 0;
}
int grootgpu_index_ref(const grootgpu_index* ix, uint32_t g, uint32_t p, const char** name, int32_t* length) {
    if (g >= ix->n_graphs || p >= ix->paths_per_graph) return fail(GROOTGPU_ERR_ARG, "no such path");
    *name = ix->ref_names[g * ix->paths_per_graph + p].c_str(); *length = 2000;
    return 0;
}
int grootgpu_index_node_paths(const grootgpu_index* ix, uint32_t node, uint32_t* graph, const uint32_t** ids, const int32_t** pos, uint32_t* n) {
    if (node >= ix->n_nodes) return fail(GROOTGPU_ERR_ARG, "no such node");
    *graph = node % ix->n_graphs; *ids = ix->path_ids.data(); *pos = ix->positions.data() + static_cast<size_t>(node) * ix->paths_per_graph; *n = ix->paths_per_graph;
    return 0;
}

int grootgpu_align_batch(grootgpu_index* ix, const uint8_t* seq, const uint64_t* seq_off, uint32_t n_reads, const grootgpu_align_params* prm, grootgpu_batch_result* out) {
    if (!prm->compact_records) return fail(GROOTGPU_ERR_ARG, "the mock only writes compact records");
    if (!seq_off && !prm->fixed_read_len) return fail(GROOTGPU_ERR_ARG, "seq_off == NULL without fixed_read_len");
    // what the previous call returned is gone: poison it before anything else
    for (grootgpu_cpair& p : ix->cpairs) p = {0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu};
    std::fill(ix->rec_path.begin(), ix->rec_path.end(), 0xee);
    ix->cpairs.clear(); ix->rec_path.clear();
    memset(out, 0, sizeof *out);
    const uint32_t P = ix->paths_per_graph;
    static const bool fast = getenv("MOCK_FAST") != nullptr;          // timing runs: hash 16 bases only, the fake device should cost next to nothing
    for (uint32_t r = 0; r < n_reads; r++) {
        const uint64_t a = seq_off ? seq_off[r] : static_cast<uint64_t>(r) * prm->fixed_read_len, b = seq_off ? seq_off[r + 1] : a + prm->fixed_read_len;
        if (prm->fixed_read_len && b - a != prm->fixed_read_len) return fail(GROOTGPU_ERR_ARG, "fixed_read_len does not match the offsets");
        const uint64_t h = fnv1a(seq + a, fast ? std::min<uint64_t>(b - a, 16) : b - a);
        if (h % 100 >= 52) continue;
        out->mapped++;
        if (prm->no_align) continue;
        const uint32_t L = static_cast<uint32_t>(b - a);
        grootgpu_cpair p;
        p.read = r; p.node = static_cast<uint32_t>((h >> 8) % ix->n_nodes);
        p.rec_count = 1u + static_cast<uint32_t>((h >> 24) % std::min(P, 20u));
        p.offset_flags = static_cast<uint32_t>((h >> 44) % 50);
        if ((h >> 40) & 1u) p.offset_flags |= GROOTGPU_CPAIR_REVERSE;
        if (L >= 3 && ((h >> 52) & 1u)) p.offset_flags |= GROOTGPU_CPAIR_CLIP_START;
        if (L >= 3 && ((h >> 53) & 1u)) p.offset_flags |= GROOTGPU_CPAIR_CLIP_END;
        ix->cpairs.push_back(p);
        for (uint32_t j = 0; j < p.rec_count; j++) ix->rec_path.push_back(static_cast<uint8_t>(j));
        ix->kmer_total += L;
    }
    out->n_reads = n_reads; out->received = n_reads;
    out->n_pairs = ix->cpairs.size(); out->n_records = ix->rec_path.size(); out->alignments = out->n_records;
    out->cpairs = ix->cpairs.data(); out->rec_path_c = ix->rec_path.data(); out->rec_path_bytes = 1;
    out->result_set = static_cast<uint32_t>(ix->calls++ & 1);
    return 0;
}

int grootgpu_weights(const grootgpu_index* ix, double* kmer_freq, uint64_t* kmer_total) {
    for (uint32_t n = 0; n < ix->n_nodes; n++) kmer_freq[n] = 0.0;
    for (uint32_t g = 0; g < ix->n_graphs; g++) kmer_total[g] = g == 0 ? ix->kmer_total : 0;
    return 0;
}
int grootgpu_prune(grootgpu_index* ix, double, uint8_t* kept) { for (uint32_t g = 0; g < ix->n_graphs; g++) kept[g] = 1; return 0; }

// ---- the multi-GPU calls, emulated with threads -------------------------------------------------------------------
int grootgpu_comm_id(uint8_t id[GROOTGPU_COMM_ID_BYTES]) {
    memset(id, 0, GROOTGPU_COMM_ID_BYTES);
    std::lock_guard<std::mutex> lk(g_comm_mu);
    const uint64_t v = g_next_id++;
    memcpy(id, &v, 8);
    return 0;
}
int grootgpu_comm_create(grootgpu_index* idx, const uint8_t id[GROOTGPU_COMM_ID_BYTES], int rank, int world, grootgpu_comm** out) {
    uint64_t v; memcpy(&v, id, 8);
    CommShared* sh;
    { std::lock_guard<std::mutex> lk(g_comm_mu); CommShared*& s = g_comms[v]; if (!s) { s = new CommShared(); s->world = world; s->id = v; s->local.assign(world, nullptr); } s->handles++; sh = s; }
    *out = new grootgpu_comm{sh, rank, idx};
    sh->barrier();
    return 0;
}
// every rank hands in its shard's compact result; rank 0 gets the merged batch (read indices rebased by the shard starts)
int grootgpu_gather(grootgpu_comm* c, const grootgpu_batch_result* local, int to_host, grootgpu_batch_result* merged) {
    CommShared* sh = c->sh;
    { std::lock_guard<std::mutex> lk(sh->mu); sh->local[c->rank] = local; }
    sh->barrier();
    if (c->rank == 0) {
        if (to_host != 1 || !merged) return fail(GROOTGPU_ERR_ARG, "the mock's rank 0 wants host results");
        for (grootgpu_cpair& p : sh->m_cpairs) p = {0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu};
        sh->m_cpairs.clear(); sh->m_rec_path.clear();
        memset(merged, 0, sizeof *merged);
        uint32_t base = 0;
        for (int r = 0; r < sh->world; r++) {
            const grootgpu_batch_result* l = sh->local[r];
            for (uint64_t i = 0; i < l->n_pairs; i++) { grootgpu_cpair p = l->cpairs[i]; p.read += base; sh->m_cpairs.push_back(p); }
            const uint8_t* rp = static_cast<const uint8_t*>(l->rec_path_c);
            sh->m_rec_path.insert(sh->m_rec_path.end(), rp, rp + l->n_records);
            base += l->n_reads;
            merged->received += l->received; merged->mapped += l->mapped; merged->multimapped += l->multimapped; merged->alignments += l->alignments;
        }
        merged->n_reads = base; merged->n_pairs = sh->m_cpairs.size(); merged->n_records = sh->m_rec_path.size();
        merged->cpairs = sh->m_cpairs.data(); merged->rec_path_c = sh->m_rec_path.data(); merged->rec_path_bytes = 1;
    }
    sh->barrier();                                  // the shards may be overwritten from here on
    return 0;
}
int grootgpu_comm_sync(grootgpu_comm* c) {          // the weights of all ranks end up on rank 0
    CommShared* sh = c->sh;
    static std::mutex mu; static uint64_t total = 0;
    { std::lock_guard<std::mutex> lk(mu); total += c->idx->kmer_total; }
    sh->barrier();
    { std::lock_guard<std::mutex> lk(mu); c->idx->kmer_total = c->rank == 0 ? total : 0; }
    sh->barrier();
    if (c->rank == 0) { std::lock_guard<std::mutex> lk(mu); total = 0; }
    return 0;
}
void grootgpu_comm_destroy(grootgpu_comm* c) {
    { std::lock_guard<std::mutex> lk(g_comm_mu); if (--c->sh->handles == 0) { g_comms.erase(c->sh->id); delete c->sh; } }
    delete c;
}

}  // extern "C"
