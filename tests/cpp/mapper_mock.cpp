// Host driver on a machine without a GPU: FastqStream -> ReadMapper::Run -> BAM, linked against tests/cpp/mock_grootgpu.cpp
// instead of libgrootgpu.so (the device side is a deterministic fake). Test / timing infrastructure only.
//   mapper_mock --bam out.bam [-p N] [--batch N] [--devices W] [--level L] [--delta 0|1] [--noAlign] [--graphs G --paths P --nodes N] reads.fq...
// prints "received mapped multimapped alignments kmers seconds"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../groot_b200/csrc/host/pipeline.h"

extern "C" grootgpu_index* mock_index_create(uint32_t n_graphs, uint32_t paths_per_graph, uint32_t n_nodes);

int main(int argc, char** argv) {
    groot_host::Info info;
    std::vector<std::string> files;
    int devices = 1;
    uint32_t graphs = 11, paths = 25, nodes = 500;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto val = [&] { return std::string(argv[++i]); };
        if (a == "--bam") info.Sketch.BAMout = val();
        else if (a == "-p") info.NumProc = atoi(val().c_str());
        else if (a == "--batch") info.BatchReads = static_cast<uint32_t>(atoi(val().c_str()));
        else if (a == "--devices") devices = atoi(val().c_str());
        else if (a == "--level") info.BamLevel = atoi(val().c_str());
        else if (a == "--delta") info.BamDelta = atoi(val().c_str()) != 0;
        else if (a == "--noAlign") info.Sketch.NoExactAlign = true;
        else if (a == "--graphs") graphs = static_cast<uint32_t>(atoi(val().c_str()));
        else if (a == "--paths") paths = static_cast<uint32_t>(atoi(val().c_str()));
        else if (a == "--nodes") nodes = static_cast<uint32_t>(atoi(val().c_str()));
        else files.push_back(a);
    }
    std::vector<grootgpu_index*> replicas;
    for (int d = 0; d < devices; d++) { replicas.push_back(mock_index_create(graphs, paths, nodes)); info.Devices.push_back(d); }
    int rc = 0;
    try {
        const auto t0 = std::chrono::steady_clock::now();
        groot_host::FastqStream stream(files, false);
        groot_host::ReadMapper mapper(&info, replicas);
        rc = mapper.Run(stream);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rc) fprintf(stderr, "%s\n", mapper.error().c_str());
        const uint64_t* st = mapper.CollectReadStats();
        printf("%llu %llu %llu %llu %llu %.4f\n", (unsigned long long)st[0], (unsigned long long)st[1], (unsigned long long)st[2],
               (unsigned long long)mapper.alignmentCount(), (unsigned long long)st[3], dt);
        fprintf(stderr, "stages: reader wait %.2f s, device calls %.2f s, BAM stage %.2f s\n", mapper.StageSeconds()[0], mapper.StageSeconds()[1], mapper.StageSeconds()[2]);
    } catch (std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        rc = 2;
    }
    for (grootgpu_index* r : replicas) grootgpu_index_destroy(r);
    return rc ? 2 : 0;
}
