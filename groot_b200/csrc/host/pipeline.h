// Host-side mirror of the reference's `groot align` pipeline (src/pipeline), in C++ because this image has no Go
// toolchain (INTEGRATION.md shows the cgo stub for the Go host). Same stage names, argument meaning and error
// behaviour as the reference:
//
//   DataStreamer   src/pipeline/sketch.go:24-77     lines from FASTQ/FASTA files (.gz by extension) or STDIN
//   FastqHandler   src/pipeline/sketch.go:153-238   4 lines -> FASTQread (header must start with '@'); --fasta mode
//   FastqChecker   src/pipeline/sketch.go:241-282   read count / mean length; "no fastq reads received" is fatal
//   ReadMapper     src/pipeline/sketch.go:285-351   theBoss.mapReads: sketch -> query -> align -> BAM + graph weights
//   GraphPruner    src/pipeline/sketch.go:354-430   Prune(minKmerCoverage) per graph, surviving paths
//
// The reference connects stages with chan-of-one-item; here the first three stages are fused into a batching reader
// that hands ReadMapper struct-of-arrays batches (bases / qualities / names back to back + offsets), which is what
// crosses the C ABI (grootgpu_align_batch). Nothing in this file computes sketches, queries or alignments.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <new>
#include <utility>
#include <string>
#include <vector>

#include "../../../include/grootgpu.h"

namespace groot_host {

// src/pipeline/runtime.go:15-56 (the fields `groot align` uses)
struct Info {
    std::string Version = GROOTGPU_REFERENCE_VERSION;
    int NumProc = 1;
    double ContainmentThreshold = 0.99;     // -t
    std::string IndexDir;
    struct AlignCmd {
        bool Fasta = false;                 // --fasta
        double MinKmerCoverage = 1.0;       // -c
        std::string BAMout;                 // "" == STDOUT
        bool NoExactAlign = false;          // --noAlign
    } Sketch;
    std::string GraphDir = "./groot-graphs";  // -g
    uint32_t BatchReads = 1u << 20;           // reads per device batch (no counterpart: the reference streams single reads)
    int Device = 0;
    std::vector<int> Devices;                 // --devices 0,1,...: one index replica per GPU, reads sharded, results gathered to the first (empty = {Device})
    int BamLevel = -1;                        // deflate level of the BGZF blocks (-1 = zlib default, like bam.NewWriter; 0..9)
    bool BamDelta = true;                     // repeated records of one read as back-references written directly (bgzf.h); false = zlib only
};

// A byte vector whose resize() leaves new bytes uninitialised: the reader sizes the batch arrays first and then lets
// several threads copy lines into them; zero-filling 200 MB per batch first would double the memory traffic.
template <class T> struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U>&) {}
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new (static_cast<void*>(p)) U; else ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
    }
};
using ByteVec = std::vector<uint8_t, NoInitAlloc<uint8_t>>;

// seqio.FASTQread for a whole batch (src/seqio/seqio.go:26-37), struct-of-arrays
struct ReadBatch {
    ByteVec id, seq, qual;                     // ID lines incl. the leading '@' / bases / raw ASCII qualities
    std::vector<uint64_t> id_off, seq_off, qual_off;
    uint32_t size() const { return seq_off.empty() ? 0 : static_cast<uint32_t>(seq_off.size() - 1); }
    void clear();
};

// DataStreamer + FastqHandler + FastqChecker
class FastqStream {
  public:
    FastqStream(const std::vector<std::string>& files, bool fasta);
    ~FastqStream();
    // fills `b` with up to max_reads reads; false when the input is exhausted and b is empty
    bool next(ReadBatch& b, uint32_t max_reads);
    // helper threads for copying parsed lines into the batch (1 = none; the I/O + line scan thread is separate)
    void set_copy_threads(unsigned n) { copy_threads_ = n < 1 ? 1 : n; }
    uint64_t rawCount() const { return raw_count_; }
    uint64_t lengthTotal() const { return length_total_; }

  private:
    bool getline(const char*& line, size_t& len);   // view into the current block, valid until the next call
    struct Scanner;                                 // reads blocks and finds their lines on a thread of its own (pipeline.cpp)
    std::shared_ptr<Scanner> scan_;
    void* cur_ = nullptr;                           // Scanner::Block being consumed
    size_t line_i_ = 0;
    unsigned copy_threads_ = 1;
    std::vector<std::string> files_;
    bool fasta_;
    std::string pending_header_, fasta_seq_;
    bool fasta_done_ = false;
    uint64_t raw_count_ = 0, length_total_ = 0;
};

// Minimal BAM writer (what the reference gets from biogo/hts: bam.NewWriter + Write + Close, boss.go:45-105,225-241).
// It writes the header; records are formatted and compressed elsewhere — by NumProc worker threads, each on its own
// slice of a batch (format_batch_bam) — and appended as ready-made BGZF blocks (append_blocks): a BAM stream is a
// concatenation of independently deflated blocks, and a record may straddle two of them.
class BamWriter {
  public:
    BamWriter(FILE* out, const std::string& sam_header_text, const std::vector<std::pair<std::string, int32_t>>& refs, int level = -1);
    ~BamWriter();
    void append_blocks(const std::vector<uint8_t>& bgzf);   // flushes the header if it is still pending, then appends the blocks as they are
    void close();  // flushes and appends the BGZF EOF block
    // one sam.Record as AlignRead builds it (src/graph/alignment.go:114-156), written at p (record_size bytes); returns the end
    static size_t record_size(uint32_t name_len, uint32_t clip_start, uint32_t match_len, uint32_t clip_end);
    static uint8_t* format_record_at(uint8_t* p, const uint8_t* name, uint32_t name_len, int32_t ref_id, int32_t pos, uint16_t flag,
                                     uint32_t clip_start, uint32_t match_len, uint32_t clip_end, const uint8_t* seq, const uint8_t* qual);
    // a further record of the same read (another path): a copy of prev with refID / pos / bin / flag replaced
    static void repeat_record_at(uint8_t* p, const uint8_t* prev, size_t len, int32_t ref_id, int32_t pos, uint16_t flag, uint32_t match_len);
    static uint16_t record_bin(int32_t pos, uint32_t match_len);
  private:
    void flush_block();
    FILE* out_;
    std::vector<uint8_t> buf_;
    int level_;
    bool closed_ = false;
};

// One batch of compact results (grootgpu_batch_result with compact_records) -> BGZF blocks. node_paths is
// grootgpu_index_node_paths (graph, sorted path ids and Position[pathID] of a node); false = error (grootgpu_last_error).
struct NodePathsView { uint32_t graph = 0; const uint32_t* ids = nullptr; const int32_t* pos = nullptr; uint32_t n = 0; };
struct BamBatch {
    const ReadBatch* reads = nullptr;
    const grootgpu_cpair* cpairs = nullptr;
    uint64_t n_pairs = 0, n_records = 0;
    const void* rec_path_c = nullptr;
    uint32_t rec_path_bytes = 1;
    const uint32_t* graph_ref_base = nullptr;          // first @SQ of every graph
    std::function<bool(uint32_t node, NodePathsView*)> node_paths;
};
// outs[t] = the blocks of worker t (append in order). Returns "" or the error text. level = deflate level of zlib
// (-1 default); delta = let the block writer encode repeated records as back-references itself (bgzf.h).
struct BamBlockStats { uint64_t zlib = 0, delta = 0, delta_own_code = 0; };   // BGZF blocks by encoder (delta_own_code: of the delta blocks, those with a Huffman code of their own)
std::string format_batch_bam(const BamBatch& in, unsigned workers, int level, bool delta, std::vector<std::vector<uint8_t>>& outs,
                             BamBlockStats* stats = nullptr);

class ReadMapper {
  public:
    // indexes[r] = the replica on GPU r of the run (indexes[0] receives the merged results and the graph weights)
    ReadMapper(Info* info, const std::vector<grootgpu_index*>& indexes);
    // theBoss.mapReads: drains the stream; BAM to info->Sketch.BAMout or STDOUT; returns 0 or a GROOTGPU_ERR_* code
    int Run(FastqStream& reads);
    // corresponds to num. reads, total num. mapped, num. multimapped, total k-mers (sketch.go:289,302-305)
    const uint64_t* CollectReadStats() const { return read_stats_; }
    uint64_t alignmentCount() const { return alignment_count_; }
    // wall-clock seconds of the run spent [0] waiting for the reader, [1] inside the device calls, [2] by the BAM stage
    // (it runs beside the other two: the largest of the three is what bounds the run)
    const double* StageSeconds() const { return stage_seconds_; }
    const std::string& error() const { return err_; }

  private:
    Info* info_;
    std::vector<grootgpu_index*> indexes_;
    grootgpu_index* index_;                   // == indexes_[0]
    uint64_t read_stats_[4] = {0, 0, 0, 0};
    uint64_t alignment_count_ = 0;
    double stage_seconds_[3] = {0, 0, 0};
    std::string err_;
};

// `groot report` (cmd/report.go:104-129 -> reporting.BAMreader.Run, src/reporting/reporting.go:33-173): BAM from a file
// or STDIN -> per-reference pileup -> references whose breadth of coverage reaches the cutoff, one line each:
// gene \t read count \t gene length \t coverage cigar (cigarClean, reporting.go:178-213). The reference prints in Go
// map order; here the lines come in @SQ order.
struct ReportLine { std::string arg; size_t count; int32_t length; std::string cigar; };
// input_file == "" reads STDIN. Throws std::runtime_error on a malformed BAM.
std::vector<ReportLine> RunReport(const std::string& input_file, double coverage_cutoff, bool low_cov);
std::string cigarClean(const std::vector<char>& str, bool* internal_d);

class GraphPruner {
  public:
    GraphPruner(Info* info, grootgpu_index* index) : info_(info), index_(index) {}
    int Run();
    const std::vector<std::string>& CollectOutput() const { return found_paths_; }
    const std::vector<uint8_t>& kept() const { return kept_; }
  private:
    Info* info_;
    grootgpu_index* index_;
    std::vector<std::string> found_paths_;
    std::vector<uint8_t> kept_;
};

}  // namespace groot_host
