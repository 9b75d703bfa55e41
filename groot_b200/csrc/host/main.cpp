// groot-b200 — command-line driver mirroring the reference's `groot index` / `groot align` (cmd/index.go:45-51,
// cmd/align.go:44-49, cmd/root.go:69-72): same flag names and defaults, BAM on STDOUT, weighted GFAs in --graphDir.
// All sketching / querying / aligning happens in libgrootgpu.so on the GPU; there is no CPU fallback.
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pipeline.h"

using namespace groot_host;

namespace {
struct Args {
    std::vector<std::string> pos;
    std::vector<std::pair<std::string, std::string>> kv;
    bool has(const std::string& a, const std::string& b = "") const { for (auto& p : kv) if (p.first == a || (!b.empty() && p.first == b)) return true; return false; }
    std::string get(const std::string& a, const std::string& b, const std::string& def) const {
        for (auto& p : kv) if (p.first == a || (!b.empty() && p.first == b)) return p.second;
        return def;
    }
};
const char* kBoolFlags[] = {"--fasta", "--noAlign", "--profiling", "--lowCov", nullptr};
Args parse(int argc, char** argv, int from) {
    Args a;
    for (int i = from; i < argc; i++) {
        std::string s = argv[i];
        if (s.size() > 1 && s[0] == '-') {
            std::string key = s, val;
            size_t eq = s.find('=');
            if (eq != std::string::npos) { key = s.substr(0, eq); val = s.substr(eq + 1); }
            bool is_bool = false;
            for (const char** b = kBoolFlags; *b; b++) if (key == *b) is_bool = true;
            if (!is_bool && eq == std::string::npos) { if (i + 1 >= argc) { fprintf(stderr, "flag needs an argument: %s\n", key.c_str()); exit(1); } val = argv[++i]; }
            a.kv.push_back({key, is_bool ? "true" : val});
        } else a.pos.push_back(s);
    }
    return a;
}
[[noreturn]] void fatal(const std::string& m) { fprintf(stderr, "%s\n", m.c_str()); exit(1); }   // misc.ErrorCheck -> log.Fatalf (src/misc/misc.go:17-21)
std::vector<std::string> split_commas(const std::string& s) {
    std::vector<std::string> out; size_t a = 0;
    while (a <= s.size()) { size_t e = s.find(',', a); if (e == std::string::npos) e = s.size(); if (e > a) out.push_back(s.substr(a, e - a)); a = e + 1; }
    return out;
}
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

static int run_index(const Args& a) {
    std::string indexDir = a.get("-i", "--indexDir", ""), msaDir = a.get("-m", "--msaDir", "");
    if (indexDir.empty()) { printf("please specify a directory for the index files (--indexDir)\n"); return 1; }
    if (msaDir.empty()) fatal("required flag(s) \"msaDir\" not set");
    grootgpu_index_params p;
    p.kmer_size = static_cast<uint32_t>(atoi(a.get("-k", "--kmerSize", "31").c_str()));
    p.sketch_size = static_cast<uint32_t>(atoi(a.get("-s", "--sketchSize", "21").c_str()));
    p.window_size = static_cast<uint32_t>(atoi(a.get("-w", "--windowSize", "100").c_str()));
    p.num_part = static_cast<uint32_t>(atoi(a.get("-x", "--numPart", "8").c_str()));
    p.max_k = static_cast<uint32_t>(atoi(a.get("-y", "--maxK", "4").c_str()));
    const double t0 = now_s();
    grootgpu_index* idx = nullptr;
    if (grootgpu_index_build_dir(msaDir.c_str(), &p, atoi(a.get("--device", "", "0").c_str()), &idx)) fatal(grootgpu_last_error());
    mkdir(indexDir.c_str(), 0777);
    if (grootgpu_index_save(idx, (indexDir + "/groot.grootb200").c_str())) fatal(grootgpu_last_error());
    // and the reference's own pair (Info.Dump + ContainmentIndex.Dump, cmd/index.go:181-186), so that the Go `groot align` can load
    // an index built here; `groot-b200 align` itself prefers the flat file. --gob=false skips them.
    if (a.get("--gob", "", "true") != "false" &&
        grootgpu_index_save_gob(idx, (indexDir + "/groot.gg").c_str(), (indexDir + "/groot.lshe").c_str())) fatal(grootgpu_last_error());
    grootgpu_index_info ii; grootgpu_index_get_info(idx, &ii);
    fprintf(stderr, "\tnumber of groot graphs built: %u\n\t\tgraphs sketched: %u\n\t\tgraph windows processed: %llu\n\tnumber of sketches added to the LSH Ensemble index: %u\nfinished in %.3fs\n",
            ii.n_graphs, ii.n_graphs - ii.n_masked_graphs, static_cast<unsigned long long>(ii.n_raw_windows), ii.n_windows, now_s() - t0);
    grootgpu_index_destroy(idx);
    return 0;
}

static int run_align(const Args& a) {
    Info info;
    info.IndexDir = a.get("-i", "--indexDir", "");
    if (info.IndexDir.empty()) { printf("please specify a directory with the index files (--indexDir)\n"); return 1; }
    info.NumProc = atoi(a.get("-p", "--processors", "1").c_str());
    info.ContainmentThreshold = atof(a.get("-t", "--contThresh", "0.99").c_str());
    info.Sketch.MinKmerCoverage = atof(a.get("-c", "--minKmerCov", "1.0").c_str());
    info.Sketch.Fasta = a.has("--fasta");
    info.Sketch.NoExactAlign = a.has("--noAlign");
    info.Sketch.BAMout = a.get("--bamOut", "", "");
    info.GraphDir = a.get("-g", "--graphDir", "./groot-graphs");
    info.Device = atoi(a.get("--device", "", "0").c_str());
    for (const std::string& d : split_commas(a.get("--devices", "", ""))) info.Devices.push_back(atoi(d.c_str()));   // no counterpart: GPUs to shard the reads over
    if (info.Devices.empty()) info.Devices.push_back(info.Device);
    info.BatchReads = static_cast<uint32_t>(atoi(a.get("--batchReads", "", "1048576").c_str()));
    info.BamLevel = atoi(a.get("--bamLevel", "", "-1").c_str());   // no counterpart: deflate level of the BAM (default = zlib's, as biogo's writer)
    info.BamDelta = atoi(a.get("--bamDelta", "", "1").c_str()) != 0;   // no counterpart: 0 = every BGZF block through zlib (host/bgzf.h)
    std::vector<std::string> fastq = split_commas(a.get("-f", "--fastq", ""));
    const double t0 = now_s();
    std::vector<grootgpu_index*> replicas;       // the index is replicated on every GPU of the run; replicas[0] ends up with the results
    // the library's own flat index file when `groot-b200 index` wrote one, else the reference's groot.gg + groot.lshe (cmd/align.go:94-107)
    struct stat sb;
    const bool own = stat((info.IndexDir + "/groot.grootb200").c_str(), &sb) == 0;
    for (int dev : info.Devices) {
        grootgpu_index* r = nullptr;
        const int rc = own ? grootgpu_index_load((info.IndexDir + "/groot.grootb200").c_str(), dev, &r)
                           : grootgpu_index_load_gob((info.IndexDir + "/groot.gg").c_str(), (info.IndexDir + "/groot.lshe").c_str(), dev, &r);
        if (rc) fatal(grootgpu_last_error());
        replicas.push_back(r);
    }
    grootgpu_index* idx = replicas[0];
    auto destroy_all = [&] { for (grootgpu_index* r : replicas) grootgpu_index_destroy(r); };
    try {
        FastqStream stream(fastq, info.Sketch.Fasta);
        ReadMapper mapper(&info, replicas);
        int rc = mapper.Run(stream);
        if (rc) fatal(mapper.error());
        const uint64_t* st = mapper.CollectReadStats();
        fprintf(stderr, "\tnumber of reads received from input: %llu\n\tmean read length: %.0f\n", static_cast<unsigned long long>(stream.rawCount()),
                stream.rawCount() ? static_cast<double>(stream.lengthTotal()) / stream.rawCount() : 0.0);
        fprintf(stderr, "\tseconds waiting for the FASTQ reader / in device calls / in the BAM stage (the three overlap): %.2f / %.2f / %.2f\n",
                mapper.StageSeconds()[0], mapper.StageSeconds()[1], mapper.StageSeconds()[2]);    // no counterpart: where the wall clock goes
        if (st[1] == 0) { fprintf(stderr, "no reads could be mapped to the reference graphs\n"); destroy_all(); return 0; }   // sketch.go:328-334
        fprintf(stderr, "\ttotal number of unmapped reads: %llu\n\ttotal number of mapped reads: %llu\n\t\tmapped to one graph: %llu\n\t\tmapped to multiple graphs: %llu\n"
                        "\ttotal number of exact alignments: %llu\n\ttotal number of k-mers projected onto graphs: %llu\n",
                (unsigned long long)(st[0] - st[1]), (unsigned long long)st[1], (unsigned long long)(st[1] - st[2]), (unsigned long long)st[2],
                (unsigned long long)mapper.alignmentCount(), (unsigned long long)st[3]);
        GraphPruner pruner(&info, idx);
        if (pruner.Run()) fatal(grootgpu_last_error());
        size_t kept = 0;
        for (uint8_t k : pruner.kept()) kept += k;
        if (kept == 0) fprintf(stderr, "\tno graphs remaining after pruning\n");          // sketch.go:421-424
        else fprintf(stderr, "\ttotal number of graphs remaining: %zu\n\ttotal number of possible haplotypes found: %zu\n", kept, pruner.CollectOutput().size());
        {   // cmd/align.go:153-161 saves every graph of info.Store. GraphPruner replaces the store by the kept graphs only
            // when at least one survives (sketch.go:421-428): with none kept the ORIGINAL store is still in place, and
            // since Prune returns false before touching a graph (graph.go:476-478) every graph that received reads is
            // written unpruned (SaveGraphAsGFA itself skips the unused ones, graphio.go:67-69).
            mkdir(info.GraphDir.c_str(), 0777);
            for (uint32_t g = 0; g < pruner.kept().size(); g++) {
                if (kept && !pruner.kept()[g]) continue;
                int written = 0;
                if (grootgpu_graph_save_gfa(idx, g, (info.GraphDir + "/groot-graph-" + std::to_string(g) + ".gfa").c_str(), static_cast<int64_t>(st[3]), &written))
                    fatal(grootgpu_last_error());
            }
        }
    } catch (std::exception& e) { fatal(e.what()); }
    fprintf(stderr, "finished in %.3fs\n", now_s() - t0);
    destroy_all();
    return 0;
}

// cmd/report.go:35-129: --bamFile (STDIN when absent), -c/--covCutoff 0.97, --lowCov (overrides -c)
static int run_report(const Args& a) {
    const std::string bamFile = a.get("--bamFile", "", "");
    double cutoff = atof(a.get("-c", "--covCutoff", "0.97").c_str());
    const bool lowCov = a.has("--lowCov");
    if (!bamFile.empty()) {
        struct stat sb;
        if (stat(bamFile.c_str(), &sb) != 0) fatal("BAM file does not exist: " + bamFile);
        if (bamFile.size() < 4 || bamFile.substr(bamFile.size() - 4) != ".bam") fatal("the BAM file does not have a `.bam` extension: " + bamFile);
    }
    if (cutoff > 1.0) fatal("supplied coverage cutoff exceeds 1.0 (100%): " + std::to_string(cutoff));
    if (lowCov) cutoff = 0.97;
    try {
        for (const ReportLine& l : RunReport(bamFile, cutoff, lowCov)) printf("%s\t%zu\t%d\t%s\n", l.arg.c_str(), l.count, l.length, l.cigar.c_str());
    } catch (std::exception& e) { fatal(e.what()); }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: groot-b200 {index,align,report,version} [flags]\n"); return 1; }
    std::string cmd = argv[1];
    Args a = parse(argc, argv, 2);
    if (cmd == "version") { printf("%s\n", grootgpu_version()); return 0; }
    if (cmd == "index") return run_index(a);
    if (cmd == "align") return run_align(a);
    if (cmd == "report") return run_report(a);
    fprintf(stderr, "unknown command \"%s\" for \"groot-b200\"\n", cmd.c_str());
    return 1;
}
