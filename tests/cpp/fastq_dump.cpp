// Test harness: prints every read the host driver's FastqStream (DataStreamer + FastqHandler + FastqChecker mirror)
// yields, one "id<TAB>seq<TAB>qual" line each, then "#count total_length". argv: [--fasta] [--batch N] [--count] [--threads T] files...
// (--count: only the last line — for timing the reader)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../groot_b200/csrc/host/pipeline.h"

int main(int argc, char** argv) {
    bool fasta = false, count_only = false;
    uint32_t batch = 3;
    unsigned threads = 1;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--fasta")) fasta = true;
        else if (!strcmp(argv[i], "--count")) count_only = true;
        else if (!strcmp(argv[i], "--threads")) threads = static_cast<unsigned>(atoi(argv[++i]));
        else if (!strcmp(argv[i], "--batch")) batch = static_cast<uint32_t>(atoi(argv[++i]));
        else files.push_back(argv[i]);
    }
    try {
        groot_host::FastqStream s(files, fasta);
        s.set_copy_threads(threads);
        groot_host::ReadBatch b;
        while (s.next(b, batch))
            for (uint32_t r = 0; r < b.size() && !count_only; r++)
                printf("%.*s\t%.*s\t%.*s\n", static_cast<int>(b.id_off[r + 1] - b.id_off[r]), reinterpret_cast<const char*>(b.id.data() + b.id_off[r]),
                       static_cast<int>(b.seq_off[r + 1] - b.seq_off[r]), reinterpret_cast<const char*>(b.seq.data() + b.seq_off[r]),
                       static_cast<int>(b.qual_off[r + 1] - b.qual_off[r]), reinterpret_cast<const char*>(b.qual.data() + b.qual_off[r]));
        printf("#%llu %llu\n", static_cast<unsigned long long>(s.rawCount()), static_cast<unsigned long long>(s.lengthTotal()));
    } catch (std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 2;
    }
    return 0;
}
