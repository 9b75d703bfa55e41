// Implementation of the host-side pipeline mirror (see pipeline.h for the reference file:line of every stage).
#include "pipeline.h"

#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <ctime>
#include <memory>
#include <stdexcept>

namespace groot_host {

void ReadBatch::clear() {
    id.clear(); seq.clear(); qual.clear();
    id_off.assign(1, 0); seq_off.assign(1, 0); qual_off.assign(1, 0);
}

// ---- DataStreamer (sketch.go:41-77) ---------------------------------------------------------------------------
FastqStream::FastqStream(const std::vector<std::string>& files, bool fasta) : files_(files), fasta_(fasta) {
    use_stdin_ = files.empty();   // no input file: scan STDIN (sketch.go:45-53)
}
FastqStream::~FastqStream() { if (gz_) gzclose(static_cast<gzFile>(gz_)); }

bool FastqStream::open_next() {
    if (gz_) { gzclose(static_cast<gzFile>(gz_)); gz_ = nullptr; }
    if (use_stdin_) {
        if (stdin_done_) return false;
        stdin_done_ = true;
        gz_ = gzdopen(0, "rb");
    } else {
        if (file_i_ >= files_.size()) return false;
        gz_ = gzopen(files_[file_i_].c_str(), "rb");       // gz decided by content; the reference keys on the ".gz" extension (sketch.go:60-68)
        if (!gz_) throw std::runtime_error("open " + files_[file_i_] + ": no such file or directory");   // misc.ErrorCheck(err) -> log.Fatal
        file_i_++;
    }
    if (!gz_) throw std::runtime_error("cannot open input");
    gzbuffer(static_cast<gzFile>(gz_), 1 << 20);
    return true;
}

// bufio.Scanner with ScanLines: strips the trailing "\n" and an optional "\r"
bool FastqStream::getline(std::string& line) {
    line.clear();
    while (true) {
        if (!gz_ && !open_next()) return false;
        char buf[1 << 16];
        bool got = false;
        while (gzgets(static_cast<gzFile>(gz_), buf, sizeof buf)) {
            got = true;
            size_t n = strlen(buf);
            line.append(buf, n);
            if (n && buf[n - 1] == '\n') { line.pop_back(); if (!line.empty() && line.back() == '\r') line.pop_back(); return true; }
        }
        if (got) return true;                                   // last line of a file without a trailing newline
        gzclose(static_cast<gzFile>(gz_)); gz_ = nullptr;       // end of this file: the next one is scanned on its own (sketch.go:55-75)
    }
}

// ---- FastqHandler (sketch.go:175-238) + FastqChecker (sketch.go:259-282) ---------------------------------------
bool FastqStream::next(ReadBatch& b, uint32_t max_reads) {
    b.clear();
    std::string l1, l2, l3, l4;
    auto push = [&](const std::string& id, const std::string& seq, const std::string& qual) {
        b.id.insert(b.id.end(), id.begin(), id.end()); b.id_off.push_back(b.id.size());
        b.seq.insert(b.seq.end(), seq.begin(), seq.end()); b.seq_off.push_back(b.seq.size());
        b.qual.insert(b.qual.end(), qual.begin(), qual.end()); b.qual_off.push_back(b.qual.size());
        raw_count_++; length_total_ += seq.size();
    };
    if (fasta_) {
        // '>' starts an entry, sequence lines are concatenated, an empty line ends the input (sketch.go:179-212)
        std::string line;
        while (b.size() < max_reads) {
            if (!getline(line) || line.empty()) {
                if (!pending_header_.empty()) { pending_header_[0] = '@'; push(pending_header_, l2, ""); pending_header_.clear(); }
                break;
            }
            if (line[0] == '>') {
                if (!pending_header_.empty()) { pending_header_[0] = '@'; push(pending_header_, l2, ""); }
                pending_header_ = line; l2.clear();
            } else l2 += line;
        }
        return b.size() > 0;
    }
    while (b.size() < max_reads) {
        if (!getline(l1) || !getline(l2) || !getline(l3) || !getline(l4)) break;   // an incomplete trailing record is dropped (sketch.go:216-236)
        if (l1.empty() || l1[0] != '@')   // seqio.NewFASTQread (seqio.go:178-180) -> log.Fatal
            throw std::runtime_error("read ID in fastq file does not begin with @: " + l1);
        push(l1, l2, l4);
    }
    return b.size() > 0;
}

// ---- BAM / BGZF ------------------------------------------------------------------------------------------------
namespace {
void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back(static_cast<uint8_t>(x >> (8 * i))); }
void put16(std::vector<uint8_t>& v, uint16_t x) { v.push_back(static_cast<uint8_t>(x)); v.push_back(static_cast<uint8_t>(x >> 8)); }
int reg2bin(int64_t beg, int64_t end) {  // SAM spec 5.3
    --end;
    if (beg >> 14 == end >> 14) return static_cast<int>(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return static_cast<int>(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return static_cast<int>(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return static_cast<int>(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return static_cast<int>(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}
const size_t kBgzfBlock = 0xff00;
}  // namespace

BamWriter::BamWriter(FILE* out, const std::string& text, const std::vector<std::pair<std::string, int32_t>>& refs) : out_(out) {
    buf_.insert(buf_.end(), {'B', 'A', 'M', 1});
    put32(buf_, static_cast<uint32_t>(text.size()));
    buf_.insert(buf_.end(), text.begin(), text.end());
    put32(buf_, static_cast<uint32_t>(refs.size()));
    for (auto& r : refs) {
        put32(buf_, static_cast<uint32_t>(r.first.size() + 1));
        buf_.insert(buf_.end(), r.first.begin(), r.first.end()); buf_.push_back(0);
        put32(buf_, static_cast<uint32_t>(r.second));
        while (buf_.size() >= kBgzfBlock) flush_block();
    }
}
BamWriter::~BamWriter() { if (!closed_) close(); }

void BamWriter::flush_block() {
    size_t n = std::min(buf_.size(), kBgzfBlock);
    std::vector<uint8_t> comp(compressBound(n) + 64);
    z_stream zs{};
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("deflateInit2 failed");
    zs.next_in = buf_.data(); zs.avail_in = static_cast<uInt>(n);
    zs.next_out = comp.data(); zs.avail_out = static_cast<uInt>(comp.size());
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); throw std::runtime_error("deflate failed"); }
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
    uint16_t bsize = static_cast<uint16_t>(clen + 25);
    hdr[16] = static_cast<uint8_t>(bsize); hdr[17] = static_cast<uint8_t>(bsize >> 8);
    fwrite(hdr, 1, 18, out_);
    fwrite(comp.data(), 1, clen, out_);
    uint32_t crc = static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), buf_.data(), static_cast<uInt>(n)));
    uint8_t tail[8];
    for (int i = 0; i < 4; i++) { tail[i] = static_cast<uint8_t>(crc >> (8 * i)); tail[4 + i] = static_cast<uint8_t>(static_cast<uint32_t>(n) >> (8 * i)); }
    fwrite(tail, 1, 8, out_);
    buf_.erase(buf_.begin(), buf_.begin() + n);
}

void BamWriter::write(const uint8_t* name, uint32_t name_len, int32_t ref_id, int32_t pos, uint16_t flag, uint32_t clip_start, uint32_t match_len,
                      uint32_t clip_end, const uint8_t* seq, const uint8_t* qual) {
    static const char* codes = "=ACMGRSVTWYHKDBN";
    uint8_t nt16[256];
    memset(nt16, 15, sizeof nt16);
    for (int i = 0; i < 16; i++) { nt16[static_cast<uint8_t>(codes[i])] = static_cast<uint8_t>(i); nt16[static_cast<uint8_t>(codes[i] | 0x20)] = static_cast<uint8_t>(i); }
    const uint32_t n_cigar = 1 + (clip_start ? 1 : 0) + (clip_end ? 1 : 0);
    const uint32_t block = 32 + name_len + 1 + 4 * n_cigar + (match_len + 1) / 2 + match_len;
    put32(buf_, block);
    put32(buf_, static_cast<uint32_t>(ref_id));
    put32(buf_, static_cast<uint32_t>(pos));
    buf_.push_back(static_cast<uint8_t>(name_len + 1));
    buf_.push_back(30);                                               // MapQ (alignment.go:143)
    put16(buf_, static_cast<uint16_t>(reg2bin(pos, pos + std::max<int64_t>(1, match_len))));
    put16(buf_, static_cast<uint16_t>(n_cigar));
    put16(buf_, flag);
    put32(buf_, match_len);
    put32(buf_, 0xffffffffu);                                         // MateRef nil
    put32(buf_, 0xffffffffu);                                         // no mate position
    put32(buf_, 0);                                                   // TempLen
    buf_.insert(buf_.end(), name, name + name_len); buf_.push_back(0);
    if (clip_start) put32(buf_, (clip_start << 4) | 5);               // H (alignment.go:132-134)
    put32(buf_, (match_len << 4) | 0);                                // M (alignment.go:135)
    if (clip_end) put32(buf_, (clip_end << 4) | 5);                   // H (alignment.go:136-138)
    for (uint32_t i = 0; i < match_len; i += 2) {
        uint8_t hi = nt16[seq[i]], lo = i + 1 < match_len ? nt16[seq[i + 1]] : 0;
        buf_.push_back(static_cast<uint8_t>((hi << 4) | lo));
    }
    buf_.insert(buf_.end(), qual, qual + match_len);                  // raw ASCII, not de-offset (alignment.go:121)
    while (buf_.size() >= kBgzfBlock) flush_block();
}

void BamWriter::close() {
    while (!buf_.empty()) flush_block();
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, out_);
    fflush(out_);
    closed_ = true;
}

// ---- ReadMapper (sketch.go:308-351 -> boss.go:45-242, graphminion.go:40-103) ------------------------------------
ReadMapper::ReadMapper(Info* info, grootgpu_index* index) : info_(info), index_(index) {}

int ReadMapper::Run(FastqStream& reads) {
    grootgpu_index_info ii;
    int rc = grootgpu_index_get_info(index_, &ii);
    if (rc) { err_ = grootgpu_last_error(); return rc; }
    // setupBAM (boss.go:45-105): one @SQ per path of every graph (here: graphs ascending, paths ascending), @PG, @RG
    std::vector<std::pair<std::string, int32_t>> refs;
    std::vector<uint32_t> graph_ref_base(ii.n_graphs + 1, 0);
    for (uint32_t g = 0; g < ii.n_graphs; g++) {
        graph_ref_base[g] = static_cast<uint32_t>(refs.size());
        for (uint32_t p = 0;; p++) {
            const char* name; int32_t len;
            if (grootgpu_index_ref(index_, g, p, &name, &len) != 0) break;
            refs.push_back({name, len});
        }
    }
    graph_ref_base[ii.n_graphs] = static_cast<uint32_t>(refs.size());
    FILE* fh = stdout;
    std::unique_ptr<BamWriter> bam;
    if (!info_->Sketch.NoExactAlign) {                                // boss.go:112-116
        if (!info_->Sketch.BAMout.empty()) {
            fh = fopen(info_->Sketch.BAMout.c_str(), "wb");
            if (!fh) { err_ = "could not open file for BAM writing: " + info_->Sketch.BAMout; return GROOTGPU_ERR_IO; }
        }
        std::string text = "@HD\tVN:1.5\tSO:unknown\n";
        for (auto& r : refs) text += "@SQ\tSN:" + r.first + "\tLN:" + std::to_string(r.second) + "\n";
        char dt[64]; time_t now = time(nullptr); strftime(dt, sizeof dt, "%Y-%m-%dT%H:%M:%S%z", localtime(&now));
        text += "@RG\tID:readsID\tDT:" + std::string(dt) + "\tPG:groot align\tPI:1000\tPL:illumina\tSM:sampleID\n";
        text += "@PG\tID:1\tPN:groot\tCL:groot align\tVN:" + info_->Version + "\n";
        bam.reset(new BamWriter(fh, text, refs));
    }
    grootgpu_align_params prm{};
    prm.containment_threshold = info_->ContainmentThreshold;
    prm.no_align = info_->Sketch.NoExactAlign ? 1 : 0;
    prm.project_on_device = 1;   // graphminion.go:67 IncrementSubPath: ordered f64 weighting on the GPU, bit-identical to the host replay
    ReadBatch b;
    std::vector<uint8_t> rc_seq, rc_qual;
    uint8_t ctab[256];                     // complementBases (seqio.go:17-23): everything else maps to 0
    memset(ctab, 0, sizeof ctab);
    ctab['A'] = 'T'; ctab['T'] = 'A'; ctab['C'] = 'G'; ctab['G'] = 'C'; ctab['N'] = 'N';
    while (reads.next(b, info_->BatchReads)) {
        grootgpu_batch_result res;
        rc = grootgpu_align_batch(index_, b.seq.data(), b.seq_off.data(), b.size(), &prm, &res);
        if (rc) { err_ = grootgpu_last_error(); return rc; }
        read_stats_[0] += res.received; read_stats_[1] += res.mapped; read_stats_[2] += res.multimapped;
        alignment_count_ += res.alignments;
        if (!bam) continue;
        for (uint64_t i = 0; i < res.n_pairs; i++) {
            const grootgpu_pair& p = res.pairs[i];
            if (p.rec_count == 0) continue;
            const uint64_t so = b.seq_off[p.read], sl = b.seq_off[p.read + 1] - so;
            const uint64_t qo = b.qual_off[p.read], ql = b.qual_off[p.read + 1] - qo;
            const uint8_t* seq = b.seq.data() + so;
            const uint8_t* qual = b.qual.data() + qo;
            if (ql < sl) {   // FASTA mode / truncated qualities: the reference panics in RevComplement or when slicing Qual (seqio.go:125-127, alignment.go:121)
                err_ = "read without a full quality string reached the BAM writer (the reference panics here)";
                return GROOTGPU_ERR_FORMAT;
            }
            if (p.reverse) {                                           // read.RevComplement() (seqio.go:120-133)
                rc_seq.resize(sl); rc_qual.resize(sl);
                for (uint64_t k = 0; k < sl; k++) { rc_seq[k] = ctab[seq[sl - 1 - k]]; rc_qual[k] = qual[sl - 1 - k]; }
                seq = rc_seq.data(); qual = rc_qual.data();
            }
            const uint32_t match = static_cast<uint32_t>(sl) - p.clip_start - p.clip_end;     // alignment.go:117
            const uint64_t io = b.id_off[p.read], il = b.id_off[p.read + 1] - io;
            for (uint32_t j = 0; j < p.rec_count; j++) {
                uint16_t flag = 0;
                if (p.rec_count > 1 && j != 0) flag |= 0x100;          // sam.Secondary (alignment.go:147-149)
                if (p.reverse) flag |= 0x10;                           // sam.Reverse (alignment.go:150-152)
                bam->write(b.id.data() + io + 1, static_cast<uint32_t>(il ? il - 1 : 0),                  // Name = ID[1:] (alignment.go:119)
                           static_cast<int32_t>(graph_ref_base[p.graph] + res.rec_path[p.rec_begin + j]), res.rec_pos[p.rec_begin + j], flag,
                           p.clip_start, match, p.clip_end, seq, qual);  // Seq/Qual = read[0:seqLength] (alignment.go:120-121)
            }
        }
    }
    if (reads.rawCount() == 0) { err_ = "no fastq reads received"; return GROOTGPU_ERR_EMPTY; }   // sketch.go:275-277
    if (bam) { bam->close(); if (fh != stdout) fclose(fh); }
    std::vector<double> kf(ii.n_nodes);
    std::vector<uint64_t> kt(ii.n_graphs);
    grootgpu_weights(index_, kf.data(), kt.data());
    for (uint64_t v : kt) read_stats_[3] += v;                          // sketch.go:342-345
    return 0;
}

// ---- GraphPruner (sketch.go:378-430) ------------------------------------------------------------------------------
int GraphPruner::Run() {
    grootgpu_index_info ii;
    int rc = grootgpu_index_get_info(index_, &ii);
    if (rc) return rc;
    kept_.assign(ii.n_graphs, 0);
    rc = grootgpu_prune(index_, info_->Sketch.MinKmerCoverage, kept_.data());
    if (rc) return rc;
    for (uint32_t g = 0; g < ii.n_graphs; g++) {
        if (!kept_[g]) continue;
        for (uint32_t p = 0;; p++) {   // every path of a kept graph is listed: the reference never deletes from g.Paths (sketch.go:410-413)
            const char* name; int32_t len;
            if (grootgpu_index_ref(index_, g, p, &name, &len) != 0) break;
            found_paths_.push_back(name);
        }
    }
    return 0;
}

}  // namespace groot_host
