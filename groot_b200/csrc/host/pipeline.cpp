// Implementation of the host-side pipeline mirror (see pipeline.h for the reference file:line of every stage).
#include "pipeline.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cerrno>

#include "bgzf.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <stdexcept>
#include <thread>

namespace groot_host {

void ReadBatch::clear() {
    id.clear(); seq.clear(); qual.clear();
    id_off.assign(1, 0); seq_off.assign(1, 0); qual_off.assign(1, 0);
}

// ---- DataStreamer (sketch.go:41-77): I/O and line scan on a thread of their own ----------------------------------
// The reference's DataStreamer is a goroutine that scans its files line by line and sends every line down a channel.
// Here it is a thread that takes the input a block at a time — a stretch of the mapped file for plain regular files,
// read() / gzread into a buffer for gzip (decided by the magic bytes), FIFOs and STDIN —, finds the line ends (one SSE2
// pass), and hands the block over together with the table of its lines; the consumer (FastqHandler, below) walks the
// table and copies the lines into the batch. A block holds whole lines only: mapped blocks are cut behind a line end, the
// unfinished tail of a buffered block is carried over to the front of the next (a line longer than a block makes the next
// block larger); blocks do not span files; the last line of a file needs no newline, and a "\r" in front of the line end
// is dropped (bufio.ScanLines, sketch.go:55-75).
struct FastqStream::Scanner : std::enable_shared_from_this<FastqStream::Scanner> {
    static constexpr size_t kBlock = 8u << 20;
    struct Mapping {                       // a whole plain file, mapped read-only
        const char* p = nullptr; size_t n = 0;
        ~Mapping() { if (p) munmap(const_cast<char*>(p), n); }
    };
    struct Block {
        const char* base = nullptr;        // the bytes of the block: `own`, or a stretch of a mapped file
        std::vector<char> own;
        std::shared_ptr<Mapping> map;      // keeps the file mapped while the block is in use
        std::vector<uint32_t> off, len;    // the lines of the block
        bool last = false;                 // nothing comes after this block (end of the input, or error set)
        std::exception_ptr error;
        void reset() { off.clear(); len.clear(); last = false; map.reset(); base = nullptr; off.reserve(kBlock / 32); len.reserve(kBlock / 32); }
    };
    std::vector<std::string> files;
    bool use_stdin = false;
    Block blocks[3];
    std::deque<Block*> ready, free_blocks;
    std::mutex mu; std::condition_variable cv;
    bool stop = false, done = false;
    std::thread th;

    // consumer side
    Block* next_block() { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !ready.empty(); }); Block* b = ready.front(); ready.pop_front(); return b; }
    void give_back(Block* b) { b->map.reset(); { std::lock_guard<std::mutex> lk(mu); free_blocks.push_back(b); } cv.notify_all(); }
    // scanner side
    Block* take_free() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || !free_blocks.empty(); });
        if (stop) return nullptr;
        Block* b = free_blocks.front(); free_blocks.pop_front();
        lk.unlock();
        b->reset();
        return b;
    }
    void publish(Block* b) { { std::lock_guard<std::mutex> lk(mu); ready.push_back(b); } cv.notify_all(); }

    struct Input {                         // one open file: mapped, read through its descriptor, or a gzFile
        gzFile gz = nullptr; int fd = -1;
        std::shared_ptr<Mapping> map;
        ~Input() { if (gz) gzclose(gz); if (fd > 0) ::close(fd); }
        size_t read(char* dst, size_t n) {             // 0 = end of file
            if (gz) { const int got = gzread(gz, dst, static_cast<unsigned>(std::min<size_t>(n, 1u << 30))); if (got < 0) throw std::runtime_error("error reading input file"); return static_cast<size_t>(got); }
            ssize_t got;
            do got = ::read(fd, dst, std::min<size_t>(n, 1u << 30)); while (got < 0 && errno == EINTR);
            if (got < 0) throw std::runtime_error("error reading input file");
            return static_cast<size_t>(got);
        }
    };
    void open(Input& in, size_t file_i) const {
        auto as_gz = [&](int fd) {              // zlib's transparent mode takes plain and gzip alike; it owns the descriptor
            in.gz = gzdopen(fd, "rb");
            if (!in.gz) throw std::runtime_error("cannot open input");
            gzbuffer(in.gz, 1 << 20);
        };
        if (use_stdin) { as_gz(0); return; }   // no input file: scan STDIN (sketch.go:45-53); a pipe cannot be rewound after a look at it
        in.fd = ::open(files[file_i].c_str(), O_RDONLY);
        if (in.fd < 0) throw std::runtime_error("open " + files[file_i] + ": no such file or directory");   // misc.ErrorCheck(err) -> log.Fatal
        struct stat sb;
        if (fstat(in.fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { const int fd = in.fd; in.fd = -1; as_gz(fd); return; }   // a FIFO / process substitution: as STDIN
        if (sb.st_size == 0) return;            // nothing to read: the read path meets end-of-file at once
        // A regular file is mapped: its pages are scanned and copied from where the page cache has them, without the
        // read() copy into a block first (a third of the scanner's time). gzip is decided by content (the reference keys
        // on the ".gz" extension, sketch.go:60-68): the magic bytes hand the descriptor to zlib instead.
        void* m = mmap(nullptr, static_cast<size_t>(sb.st_size), PROT_READ, MAP_PRIVATE, in.fd, 0);
        if (m == MAP_FAILED) return;            // cannot be mapped: plain read()
        in.map = std::make_shared<Mapping>();
        in.map->p = static_cast<const char*>(m); in.map->n = static_cast<size_t>(sb.st_size);
        if (in.map->n >= 2 && static_cast<uint8_t>(in.map->p[0]) == 0x1f && static_cast<uint8_t>(in.map->p[1]) == 0x8b) {
            in.map.reset();
            const int fd = in.fd; in.fd = -1;
            as_gz(fd);
            return;
        }
        madvise(m, in.map->n, MADV_SEQUENTIAL);
    }
    static void add_line(Block* b, const char* d, size_t start, size_t n) {
        if (n && d[start + n - 1] == '\r') n--;
        b->off.push_back(static_cast<uint32_t>(start)); b->len.push_back(static_cast<uint32_t>(n));
    }
    // the lines of d[0, n) into the block's table; returns where the unfinished last line starts (n if there is none)
    static size_t scan_lines(Block* b, const char* d, size_t n) {
        size_t p = 0, i = 0;
#if defined(__SSE2__)
        // every line end in one pass, 16 bytes per step (a memchr call per 15-150 byte line costs more in set-up than in scanning)
        const __m128i nlv = _mm_set1_epi8('\n');
        for (; i + 16 <= n; i += 16) {
            unsigned m = static_cast<unsigned>(_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(d + i)), nlv)));
            while (m) {
                const size_t e = i + static_cast<unsigned>(__builtin_ctz(m));
                add_line(b, d, p, e - p);
                p = e + 1;
                m &= m - 1;
            }
        }
#endif
        for (i = std::max(i, p); i < n;) {
            const char* nl = static_cast<const char*>(memchr(d + i, '\n', n - i));
            if (!nl) break;
            add_line(b, d, p, static_cast<size_t>(nl - d) - p);
            p = i = static_cast<size_t>(nl - d) + 1;
        }
        return p;
    }
    // a mapped file: blocks are stretches of the mapping cut behind a line end (no tail to carry over)
    bool run_mapped(const std::shared_ptr<Mapping>& map) {
        const char* base = map->p;
        const size_t size = map->n;
        for (size_t pos = 0; pos < size;) {
            size_t end = std::min(pos + kBlock, size);
            if (end < size) {
                const char* nl = static_cast<const char*>(memchr(base + end - 1, '\n', size - (end - 1)));
                end = nl ? static_cast<size_t>(nl - base) + 1 : size;
            }
            if (end - pos > 0xfff00000u) throw std::runtime_error("a line of the input is longer than 4 GB");
            Block* b = take_free();
            if (!b) return false;
            b->map = map; b->base = base + pos;
            const size_t n = end - pos, p = scan_lines(b, b->base, n);
            if (p < n) add_line(b, b->base, p, n - p);                    // unterminated last line of the file
            publish(b);
            pos = end;
        }
        return true;
    }
    // anything else: blocks are filled by read() / gzread; the unfinished tail of one goes to the front of the next
    bool run_read(Input& in) {
        std::vector<char> carry;
        size_t want = kBlock;
        for (bool eof = false; !eof;) {
            Block* b = take_free();
            if (!b) return false;
            const size_t cap = std::max(want, carry.size() + kBlock / 2);
            if (cap > 0xfff00000u) throw std::runtime_error("a line of the input is longer than 4 GB");
            if (b->own.size() < cap) b->own.resize(cap);
            char* d = b->own.data();
            b->base = d;
            size_t n = carry.size();
            if (n) memcpy(d, carry.data(), n);
            carry.clear();
            const size_t got = in.fd >= 0 || in.gz ? in.read(d + n, b->own.size() - n) : 0;   // neither: an empty file
            if (got == 0) eof = true;
            n += got;
            const size_t p = scan_lines(b, d, n);
            if (p < n) {
                if (eof) add_line(b, d, p, n - p);                            // unterminated last line of the file
                else { carry.assign(d + p, d + n); if (p == 0) want = 2 * n + kBlock; }   // a line longer than the block: a larger one next
            }
            publish(b);
        }
        return true;
    }
    void run() {
        try {
            const size_t n_files = use_stdin ? 1 : files.size();
            for (size_t fi = 0; fi < n_files; fi++) {
                Input in;
                open(in, fi);
                if (!(in.map ? run_mapped(in.map) : run_read(in))) return;
            }
            if (Block* b = take_free()) { b->last = true; publish(b); }
        } catch (...) {
            if (Block* b = take_free()) { b->last = true; b->error = std::current_exception(); publish(b); }
        }
        { std::lock_guard<std::mutex> lk(mu); done = true; }
    }
    void start() {
        for (Block& b : blocks) free_blocks.push_back(&b);
        th = std::thread([self = shared_from_this()] { self->run(); });   // the thread keeps the scanner alive (see ~FastqStream)
    }
};

FastqStream::FastqStream(const std::vector<std::string>& files, bool fasta) : files_(files), fasta_(fasta) {}

FastqStream::~FastqStream() {
    if (!scan_) return;
    bool finished;
    { std::lock_guard<std::mutex> lk(scan_->mu); scan_->stop = true; finished = scan_->done; }
    scan_->cv.notify_all();
    // A scanner that has run to the end of its input is joined. One that is still going — the consumer gave up early, the
    // scanner may sit in a read() on a pipe that delivers nothing — is left to notice `stop` on its own: it holds the only
    // other reference to its state.
    if (finished) scan_->th.join(); else scan_->th.detach();
}

// the next line of the input: a view into the current block, valid until the next call
bool FastqStream::getline(const char*& line, size_t& len) {
    if (!scan_) {
        scan_ = std::make_shared<Scanner>();
        scan_->files = files_; scan_->use_stdin = files_.empty();
        scan_->start();
    }
    while (true) {
        Scanner::Block* cur = static_cast<Scanner::Block*>(cur_);
        if (cur) {
            if (line_i_ < cur->off.size()) { line = cur->base + cur->off[line_i_]; len = cur->len[line_i_]; line_i_++; return true; }
            if (cur->last) { if (cur->error) std::rethrow_exception(cur->error); return false; }
            scan_->give_back(cur);
        }
        cur_ = scan_->next_block();
        line_i_ = 0;
    }
}

// ---- FastqHandler (sketch.go:175-238) + FastqChecker (sketch.go:259-282) ---------------------------------------
namespace {
// a line goes into the batch with one memmove: iterators of the vector's own element type (a `const char*` range is
// converted element by element, which was 80 % of the reader's time)
inline void append(ByteVec& v, const char* p, size_t n) {
    const uint8_t* q = reinterpret_cast<const uint8_t*>(p);
    v.insert(v.end(), q, q + n);
}
}  // namespace

bool FastqStream::next(ReadBatch& b, uint32_t max_reads) {
    b.clear();
    auto push = [&](const char* id, size_t id_n, const char* seq, size_t seq_n, const char* qual, size_t qual_n) {
        append(b.id, id, id_n); b.id_off.push_back(b.id.size());
        append(b.seq, seq, seq_n); b.seq_off.push_back(b.seq.size());
        append(b.qual, qual, qual_n); b.qual_off.push_back(b.qual.size());
        raw_count_++; length_total_ += seq_n;
    };
    const char* p; size_t n;
    if (fasta_) {
        // '>' starts an entry, sequence lines are concatenated, an empty line ends the input (sketch.go:179-212)
        auto flush = [&] { pending_header_[0] = '@'; push(pending_header_.data(), pending_header_.size(), fasta_seq_.data(), fasta_seq_.size(), "", 0); };
        while (b.size() < max_reads && !fasta_done_) {
            if (!getline(p, n) || n == 0) {                       // an empty line ends the input for good (sketch.go:180-182)
                if (!pending_header_.empty()) { flush(); pending_header_.clear(); }
                fasta_done_ = true;
                break;
            }
            if (p[0] == '>') {
                if (!pending_header_.empty()) flush();
                pending_header_.assign(p, n); fasta_seq_.clear();
            } else fasta_seq_.append(p, n);
        }
        return b.size() > 0;
    }
    auto next_line = [&](const char*& q, size_t& m) { while (getline(q, m)) if (m != 0) return true; return false; };
    std::vector<uint32_t> plan;                      // per planned record: the table entries of its ID, bases and qualities
    while (b.size() < max_reads) {
        // Fast path: the records that lie completely inside the current block. Their lines are already in the block's
        // table, so the batch offsets are a running sum (no data touched but the '@'), and the copies — what the reader's
        // time goes into — are independent per record: with copy_threads_ > 1 they are spread over helper threads.
        Scanner::Block* cur = static_cast<Scanner::Block*>(cur_);
        if (cur && line_i_ < cur->off.size()) {
            const char* d = cur->base;
            const uint32_t *off = cur->off.data(), *len = cur->len.data();
            const size_t n_lines = cur->off.size(), first = b.size();
            size_t li = line_i_;
            uint64_t id_at = b.id.size(), seq_at = b.seq.size(), qual_at = b.qual.size(), bases = 0;
            plan.clear();
            while (first + plan.size() / 3 < max_reads) {
                size_t q = li; uint32_t idx[4]; int got = 0;
                while (q < n_lines && got < 4) { if (len[q]) idx[got++] = static_cast<uint32_t>(q); q++; }
                if (got < 4 || d[off[idx[0]]] != '@') break;      // runs into the next block / not a header: the slow path decides
                plan.insert(plan.end(), {idx[0], idx[1], idx[3]});
                id_at += len[idx[0]]; seq_at += len[idx[1]]; qual_at += len[idx[3]]; bases += len[idx[1]];
                b.id_off.push_back(id_at); b.seq_off.push_back(seq_at); b.qual_off.push_back(qual_at);
                li = q;
            }
            if (!plan.empty()) {
                b.id.resize(id_at); b.seq.resize(seq_at); b.qual.resize(qual_at);
                const size_t n_rec = plan.size() / 3;
                auto copy = [&](size_t r0, size_t r1) {
                    for (size_t r = r0; r < r1; r++) {
                        const uint32_t* e = &plan[3 * r];
                        memcpy(b.id.data() + b.id_off[first + r], d + off[e[0]], len[e[0]]);
                        memcpy(b.seq.data() + b.seq_off[first + r], d + off[e[1]], len[e[1]]);
                        memcpy(b.qual.data() + b.qual_off[first + r], d + off[e[2]], len[e[2]]);
                    }
                };
                const unsigned T = n_rec >= 8192 ? copy_threads_ : 1u;
                if (T > 1) {
                    std::vector<std::thread> th;
                    for (unsigned t = 1; t < T; t++) th.emplace_back(copy, n_rec * t / T, n_rec * (t + 1) / T);
                    copy(0, n_rec / T);
                    for (auto& x : th) x.join();
                } else copy(0, n_rec);
                raw_count_ += n_rec; length_total_ += bases;
                line_i_ = li;
                continue;
            }
        }
        // Slow path, one record: its four lines are copied into the batch as they are found (a view does not survive the
        // next getline). An EMPTY line never fills a slot: the DataStreamer forwards it as a nil slice (append([]byte(nil), ...)
        // of nothing, sketch.go:49,70) and the handler's `l1 == nil / l2 == nil / ...` tests take the next line for the same
        // slot (sketch.go:216-236), so blank lines anywhere in a FASTQ stream are skipped.
        if (!next_line(p, n)) break;
        const size_t id0 = b.id.size(), seq0 = b.seq.size(), qual0 = b.qual.size();
        append(b.id, p, n);
        bool ok = next_line(p, n);
        if (ok) { append(b.seq, p, n); ok = next_line(p, n); }        // line 3 ('+') is dropped
        if (ok) ok = next_line(p, n);
        if (!ok) { b.id.resize(id0); b.seq.resize(seq0); b.qual.resize(qual0); break; }   // an incomplete trailing record is dropped (sketch.go:216-236)
        if (b.id.size() == id0 || b.id[id0] != '@')   // only a complete record reaches seqio.NewFASTQread (seqio.go:178-180) -> log.Fatal
            throw std::runtime_error("read ID in fastq file does not begin with @: " + std::string(b.id.begin() + id0, b.id.end()));
        append(b.qual, p, n);
        b.id_off.push_back(b.id.size()); b.seq_off.push_back(b.seq.size()); b.qual_off.push_back(b.qual.size());
        raw_count_++; length_total_ += b.seq.size() - seq0;
    }
    return b.size() > 0;
}

// ---- BAM / BGZF ------------------------------------------------------------------------------------------------
namespace {
void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back(static_cast<uint8_t>(x >> (8 * i))); }
int reg2bin(int64_t beg, int64_t end) {  // SAM spec 5.3
    --end;
    if (beg >> 14 == end >> 14) return static_cast<int>(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return static_cast<int>(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return static_cast<int>(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return static_cast<int>(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return static_cast<int>(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}
}  // namespace

BamWriter::BamWriter(FILE* out, const std::string& text, const std::vector<std::pair<std::string, int32_t>>& refs, int level) : out_(out), level_(level) {
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
    const size_t n = std::min(buf_.size(), kBgzfBlock);
    std::vector<uint8_t> blk;
    BgzfDeflater::compress_plain(buf_.data(), n, level_, blk);
    fwrite(blk.data(), 1, blk.size(), out_);
    buf_.erase(buf_.begin(), buf_.begin() + n);
}

void BamWriter::append_blocks(const std::vector<uint8_t>& bgzf) {
    while (!buf_.empty()) flush_block();
    fwrite(bgzf.data(), 1, bgzf.size(), out_);
}

size_t BamWriter::record_size(uint32_t name_len, uint32_t clip_start, uint32_t match_len, uint32_t clip_end) {
    const uint32_t n_cigar = 1 + (clip_start ? 1 : 0) + (clip_end ? 1 : 0);
    return 36u + name_len + 1u + 4u * n_cigar + (match_len + 1u) / 2u + match_len;
}

namespace {
inline void st32(uint8_t* p, uint32_t x) { memcpy(p, &x, 4); }     // little-endian hosts only (x86-64 / aarch64), as the rest of the library
inline void st16(uint8_t* p, uint16_t x) { memcpy(p, &x, 2); }
}  // namespace

uint8_t* BamWriter::format_record_at(uint8_t* p, const uint8_t* name, uint32_t name_len, int32_t ref_id, int32_t pos, uint16_t flag,
                                     uint32_t clip_start, uint32_t match_len, uint32_t clip_end, const uint8_t* seq, const uint8_t* qual) {
    static const struct Nt16 {
        uint8_t t[256];
        Nt16() {
            static const char* codes = "=ACMGRSVTWYHKDBN";
            memset(t, 15, sizeof t);
            for (int i = 0; i < 16; i++) { t[static_cast<uint8_t>(codes[i])] = static_cast<uint8_t>(i); t[static_cast<uint8_t>(codes[i] | 0x20)] = static_cast<uint8_t>(i); }
        }
    } nt16;
    const uint32_t n_cigar = 1 + (clip_start ? 1 : 0) + (clip_end ? 1 : 0);
    st32(p, static_cast<uint32_t>(record_size(name_len, clip_start, match_len, clip_end) - 4));   // block_size
    st32(p + 4, static_cast<uint32_t>(ref_id));
    st32(p + 8, static_cast<uint32_t>(pos));
    p[12] = static_cast<uint8_t>(name_len + 1);
    p[13] = 30;                                                      // MapQ (alignment.go:143)
    st16(p + 14, record_bin(pos, match_len));
    st16(p + 16, static_cast<uint16_t>(n_cigar));
    st16(p + 18, flag);
    st32(p + 20, match_len);
    st32(p + 24, 0xffffffffu);                                       // MateRef nil
    st32(p + 28, 0xffffffffu);                                       // no mate position
    st32(p + 32, 0);                                                 // TempLen
    p += 36;
    memcpy(p, name, name_len); p[name_len] = 0; p += name_len + 1;
    if (clip_start) { st32(p, (clip_start << 4) | 5); p += 4; }      // H (alignment.go:132-134)
    st32(p, (match_len << 4) | 0); p += 4;                           // M (alignment.go:135)
    if (clip_end) { st32(p, (clip_end << 4) | 5); p += 4; }          // H (alignment.go:136-138)
    for (uint32_t i = 0; i + 1 < match_len; i += 2) *p++ = static_cast<uint8_t>((nt16.t[seq[i]] << 4) | nt16.t[seq[i + 1]]);
    if (match_len & 1u) *p++ = static_cast<uint8_t>(nt16.t[seq[match_len - 1]] << 4);
    memcpy(p, qual, match_len);                                      // raw ASCII, not de-offset (alignment.go:121)
    return p + match_len;
}

uint16_t BamWriter::record_bin(int32_t pos, uint32_t match_len) { return static_cast<uint16_t>(reg2bin(pos, pos + std::max<int64_t>(1, match_len))); }

// the next record of the same read against another path: everything but refID, pos, bin and the flag is the same
// (alignment.go:296-315 builds them from one read in a loop over the paths of the start node)
void BamWriter::repeat_record_at(uint8_t* p, const uint8_t* prev, size_t len, int32_t ref_id, int32_t pos, uint16_t flag, uint32_t match_len) {
    memcpy(p, prev, len);
    st32(p + 4, static_cast<uint32_t>(ref_id));
    st32(p + 8, static_cast<uint32_t>(pos));
    st16(p + 14, record_bin(pos, match_len));
    st16(p + 18, flag);
}

void BamWriter::close() {
    while (!buf_.empty()) flush_block();
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, out_);
    fflush(out_);
    closed_ = true;
}

// ---- one batch of compact results -> BGZF blocks (the writer loop of boss.go:225-242 over alignment.go:114-156 records) ----
// NumProc workers each format and deflate a contiguous slice of the pairs (slices of roughly equal record count) into
// ready-made BGZF blocks; the slices are appended in order. One worker reproduces the serial writer byte for byte up
// to block boundaries. The first record of a pair is formatted from the read (reverse-complemented when the pair is on
// the reverse strand, read.RevComplement(), seqio.go:120-133); its other records are copies with refID / pos / bin / flag
// patched, and the block writer is told so (bgzf.h).
std::string format_batch_bam(const BamBatch& in, unsigned workers, int level, bool delta, std::vector<std::vector<uint8_t>>& outs, BamBlockStats* stats) {
    const ReadBatch& b = *in.reads;
    workers = std::max(1u, workers);
    uint8_t ctab[256];                     // complementBases (seqio.go:17-23): everything else maps to 0
    memset(ctab, 0, sizeof ctab);
    ctab['A'] = 'T'; ctab['T'] = 'A'; ctab['C'] = 'G'; ctab['G'] = 'C'; ctab['N'] = 'N';
    // first record of every pair (the compact output carries counts only)
    std::vector<uint64_t> rec_begin(in.n_pairs + 1, 0);
    for (uint64_t i = 0; i < in.n_pairs; i++) rec_begin[i + 1] = rec_begin[i] + in.cpairs[i].rec_count;
    const uint8_t* path8 = in.rec_path_bytes == 1 ? static_cast<const uint8_t*>(in.rec_path_c) : nullptr;
    const uint16_t* path16 = in.rec_path_bytes == 2 ? static_cast<const uint16_t*>(in.rec_path_c) : nullptr;
    std::vector<uint64_t> cut(workers + 1, in.n_pairs);
    cut[0] = 0;
    {
        uint64_t i = 0;
        for (unsigned t = 1; t < workers; t++) {
            const uint64_t want = in.n_records * t / workers;
            while (i < in.n_pairs && rec_begin[i] < want) i++;
            cut[t] = i;
        }
    }
    outs.assign(workers, {});
    std::vector<std::string> errs(workers);
    std::vector<BamBlockStats> wstats(workers);
    auto work = [&](unsigned t) {
        std::vector<uint8_t> rc_seq, rc_qual;
        std::vector<uint8_t>& out = outs[t];
        try {
            BgzfDeflater z(level, delta);
            size_t prev_len = 0;               // length of the record in front of the next one (0: none pending)
            for (uint64_t i = cut[t]; i < cut[t + 1]; i++) {
                const grootgpu_cpair& p = in.cpairs[i];
                if (p.rec_count == 0) continue;
                const bool reverse = (p.offset_flags & GROOTGPU_CPAIR_REVERSE) != 0;
                const uint32_t clip_start = (p.offset_flags & GROOTGPU_CPAIR_CLIP_START) ? 1u : 0u, clip_end = (p.offset_flags & GROOTGPU_CPAIR_CLIP_END) ? 1u : 0u;
                const int32_t offset = static_cast<int32_t>(p.offset_flags & GROOTGPU_CPAIR_OFFSET_MASK);
                NodePathsView np;
                if (!in.node_paths(p.node, &np)) { errs[t] = grootgpu_last_error(); return; }
                const uint64_t so = b.seq_off[p.read], sl = b.seq_off[p.read + 1] - so;
                const uint64_t qo = b.qual_off[p.read], ql = b.qual_off[p.read + 1] - qo;
                const uint8_t* seq = b.seq.data() + so;
                const uint8_t* qual = b.qual.data() + qo;
                if (ql < sl) {   // FASTA mode / truncated qualities: the reference panics in RevComplement or when slicing Qual (seqio.go:125-127, alignment.go:121)
                    errs[t] = "read without a full quality string reached the BAM writer (the reference panics here)";
                    return;
                }
                if (reverse) {                                             // read.RevComplement() (seqio.go:120-133)
                    rc_seq.resize(sl); rc_qual.resize(sl);
                    for (uint64_t k = 0; k < sl; k++) { rc_seq[k] = ctab[seq[sl - 1 - k]]; rc_qual[k] = qual[sl - 1 - k]; }
                    seq = rc_seq.data(); qual = rc_qual.data();
                }
                const uint32_t match = static_cast<uint32_t>(sl) - clip_start - clip_end;         // alignment.go:117
                const uint64_t io = b.id_off[p.read], il = b.id_off[p.read + 1] - io;
                const uint32_t name_len = static_cast<uint32_t>(il ? il - 1 : 0);                 // Name = ID[1:] (alignment.go:119)
                const size_t rl = BamWriter::record_size(name_len, clip_start, match, clip_end);
                for (uint32_t j = 0; j < p.rec_count; j++) {
                    uint16_t flag = 0;
                    if (p.rec_count > 1 && j != 0) flag |= 0x100;          // sam.Secondary (alignment.go:147-149)
                    if (reverse) flag |= 0x10;                             // sam.Reverse (alignment.go:150-152)
                    const uint32_t path = path8 ? path8[rec_begin[i] + j] : path16[rec_begin[i] + j];
                    const uint32_t* it = std::lower_bound(np.ids, np.ids + np.n, path);   // Position[pathID] of the start node (alignment.go:296)
                    const int32_t pos = (it != np.ids + np.n && *it == path ? np.pos[it - np.ids] : 0) + offset;
                    const int32_t ref_id = static_cast<int32_t>(in.graph_ref_base[np.graph] + path);
                    uint8_t* dst = z.reserve(rl);
                    if (j == 0) {
                        BamWriter::format_record_at(dst, b.id.data() + io + 1, name_len, ref_id, pos, flag, clip_start, match, clip_end, seq, qual);   // Seq/Qual = read[0:seqLength] (alignment.go:120-121)
                        z.commit(rl, static_cast<uint32_t>(prev_len));     // the record of another read in front: header fields and the name's prefix usually match
                    } else {
                        BamWriter::repeat_record_at(dst, dst - rl, rl, ref_id, pos, flag, match);
                        z.commit(rl, static_cast<uint32_t>(rl));
                    }
                    prev_len = rl;
                }
                if (z.pending() >= (1u << 20)) { z.drain(false, out); if (z.pending() < rl) prev_len = 0; }   // whole blocks; the remainder stays pending
            }
            z.drain(true, out);
            wstats[t].zlib = z.zlib_blocks(); wstats[t].delta = z.delta_blocks(); wstats[t].delta_own_code = z.dynamic_blocks();
        } catch (std::exception& e) { errs[t] = e.what(); }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 1; t < workers; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    if (stats) { *stats = BamBlockStats(); for (const BamBlockStats& w : wstats) { stats->zlib += w.zlib; stats->delta += w.delta; stats->delta_own_code += w.delta_own_code; } }
    for (unsigned t = 0; t < workers; t++) if (!errs[t].empty()) return errs[t];
    return "";
}

namespace {
double now_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

// ---- ReadMapper (sketch.go:308-351 -> boss.go:45-242, graphminion.go:40-103) ------------------------------------
ReadMapper::ReadMapper(Info* info, const std::vector<grootgpu_index*>& indexes) : info_(info), indexes_(indexes), index_(indexes.at(0)) {}

namespace {
// The ranks of a multi-GPU run inside one process: rank 0 is the caller's thread, ranks 1.. are helper threads that live
// for the whole stream. Per batch every rank maps its contiguous shard (grootgpu_align_batch, results kept on the
// device) and takes part in ONE gather to rank 0 (grootgpu_gather, NCCL over NVLink); the graph weights travel round
// the ranks inside the library. A failure on any rank is agreed on at a barrier BEFORE the collective, so nobody hangs.
struct RankTeam {
    std::vector<grootgpu_index*> idx;
    std::vector<grootgpu_comm*> comm;
    std::vector<std::thread> th;
    std::mutex mu; std::condition_variable cv;
    uint64_t gen = 0;                       // batch generation handed to the helpers
    int arrived = 0; uint64_t bar_gen = 0;  // barrier
    bool stop = false;
    const uint8_t* seq = nullptr; const uint64_t* seq_off = nullptr; uint32_t n = 0;
    grootgpu_align_params prm{};
    uint32_t rec_path_bytes = 1;
    std::vector<int> rc; std::vector<std::string> err;
    int world() const { return static_cast<int>(idx.size()); }

    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const uint64_t g = bar_gen;
        if (++arrived == world()) { arrived = 0; bar_gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return bar_gen != g; });
    }
    static void shard(uint32_t n, int world, int r, uint32_t* lo, uint32_t* hi) {
        const uint32_t base = n / world, rem = n % world;
        *lo = r * base + std::min<uint32_t>(r, rem); *hi = *lo + base + (static_cast<uint32_t>(r) < rem ? 1u : 0u);
    }
    // one batch on rank r; merged is filled on rank 0
    void rank_batch(int r, grootgpu_batch_result* merged) {
        uint32_t lo, hi; shard(n, world(), r, &lo, &hi);
        grootgpu_batch_result res{};
        rc[r] = 0;
        if (hi > lo) {
            rc[r] = grootgpu_align_batch(idx[r], seq, seq_off + lo, hi - lo, &prm, &res);
            if (rc[r]) err[r] = grootgpu_last_error();
        } else res.rec_path_bytes = rec_path_bytes;   // an empty shard still takes part in the gather (same format, no arrays)
        barrier();
        bool any = false;
        for (int x : rc) any = any || x != 0;
        if (any) return;
        rc[r] = grootgpu_gather(comm[r], &res, r == 0 ? 1 : 0, merged);
        if (rc[r]) err[r] = grootgpu_last_error();
    }
    int start(std::string* e) {
        const int W = world();
        comm.assign(W, nullptr); rc.assign(W, 0); err.assign(W, "");
        { grootgpu_index_info ii; grootgpu_index_get_info(idx[0], &ii); rec_path_bytes = ii.max_paths_per_graph <= 256 ? 1u : 2u; }
        uint8_t id[GROOTGPU_COMM_ID_BYTES];
        if (int x = grootgpu_comm_id(id)) { *e = grootgpu_last_error(); return x; }
        std::vector<std::thread> init;
        for (int r = 0; r < W; r++) init.emplace_back([&, r] { rc[r] = grootgpu_comm_create(idx[r], id, r, W, &comm[r]); if (rc[r]) err[r] = grootgpu_last_error(); });
        for (auto& t : init) t.join();
        for (int r = 0; r < W; r++) if (rc[r]) { *e = err[r]; return rc[r]; }
        for (int r = 1; r < W; r++)
            th.emplace_back([this, r] {
                uint64_t seen = 0;
                while (true) {
                    { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return stop || gen != seen; }); if (stop && gen == seen) break; seen = gen; }
                    if (seq) rank_batch(r, nullptr);
                    else { rc[r] = grootgpu_comm_sync(comm[r]); if (rc[r]) err[r] = grootgpu_last_error(); }   // end of stream: the collective sync
                    barrier();
                }
            });
        return 0;
    }
    // rank 0's side of one batch (seq != nullptr) or of the final sync (seq == nullptr)
    int run(const uint8_t* s, const uint64_t* so, uint32_t nn, grootgpu_batch_result* merged, std::string* e) {
        { std::lock_guard<std::mutex> lk(mu); seq = s; seq_off = so; n = nn; gen++; }
        cv.notify_all();
        if (s) rank_batch(0, merged);
        else { rc[0] = grootgpu_comm_sync(comm[0]); if (rc[0]) err[0] = grootgpu_last_error(); }
        barrier();
        for (int r = 0; r < world(); r++) if (rc[r]) { *e = err[r]; return rc[r]; }
        return 0;
    }
    ~RankTeam() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& t : th) if (t.joinable()) t.join();
        for (grootgpu_comm* c : comm) if (c) grootgpu_comm_destroy(c);
    }
};
}  // namespace

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
        bam.reset(new BamWriter(fh, text, refs, info_->BamLevel));
    }
    grootgpu_align_params prm{};
    prm.containment_threshold = info_->ContainmentThreshold;
    prm.no_align = info_->Sketch.NoExactAlign ? 1 : 0;
    prm.project_on_device = 1;   // graphminion.go:67 IncrementSubPath: ordered f64 weighting on the GPU, bit-identical to the host replay
    prm.compact_records = 1;     // the BAM writer needs (read, start locus, strand, clips, path ids): a fifth of the full output's bytes
    RankTeam team;
    if (indexes_.size() > 1) {
        team.idx = indexes_;
        team.prm = prm; team.prm.results_on_device = 1;
        if ((rc = team.start(&err_))) return rc;
    }
    // Three stages, as the reference's pipeline has them (pipeline.go:36-45; boss.go:225-242 writes the BAM from its own
    // goroutine): DataStreamer/FastqHandler parse batch i+2 on the reader thread while batch i+1 is on the GPU (this thread)
    // and the BAM stage formats, deflates and writes batch i. Three batch buffers go round; a buffer is released by the
    // last stage that needs it (the BAM stage reads names, bases and qualities from it).
    // the reader's line copies may use helper threads: freely when there is no BAM stage to feed (--noAlign), sparingly beside it
    reads.set_copy_threads(info_->Sketch.NoExactAlign ? static_cast<unsigned>(std::min(std::max(1, info_->NumProc), 4)) : (info_->NumProc >= 8 ? 2u : 1u));
    constexpr int kSlots = 3;
    struct BatchFeed {
        FastqStream& reads; uint32_t batch_reads;
        ReadBatch slot[kSlots];
        int state[kSlots] = {0, 0, 0};     // 0 = free, 1 = filled, 2 = end of input
        bool stop = false;
        std::exception_ptr error;
        std::mutex mu; std::condition_variable cv;
        std::thread th;
        BatchFeed(FastqStream& r, uint32_t n) : reads(r), batch_reads(n) {
            th = std::thread([this] {
                try {
                    for (int i = 0;; i = (i + 1) % kSlots) {
                        { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return stop || state[i] == 0; }); if (stop) return; }
                        const bool ok = reads.next(slot[i], batch_reads);
                        { std::lock_guard<std::mutex> lk(mu); state[i] = ok ? 1 : 2; }
                        cv.notify_all();
                        if (!ok) return;
                    }
                } catch (...) {
                    { std::lock_guard<std::mutex> lk(mu); error = std::current_exception(); for (int& st : state) st = 2; }
                    cv.notify_all();
                }
            });
        }
        ReadBatch* take(int i) {           // nullptr at the end of the input; rethrows the reader's error
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return state[i] != 0; });
            if (error) std::rethrow_exception(error);
            return state[i] == 1 ? &slot[i] : nullptr;
        }
        void release(int i) { { std::lock_guard<std::mutex> lk(mu); state[i] = 0; } cv.notify_all(); }
        ~BatchFeed() { { std::lock_guard<std::mutex> lk(mu); stop = true; } cv.notify_all(); if (th.joinable()) th.join(); }
    } feed(reads, info_->BatchReads);
    // The BAM stage: one batch at a time, in order. The library's result arrays are only valid until the next align call
    // on the handle (include/grootgpu.h), which this thread makes while the stage is still working: the compact records
    // (16 bytes per pair + 1-2 bytes per record) are copied into the stage's own buffers at the hand-over.
    struct BamStage {
        BatchFeed& feed; BamWriter* bam; Info* info; grootgpu_index* index; const uint32_t* graph_ref_base;
        std::vector<grootgpu_cpair> cpairs; std::vector<uint8_t> rec_path;
        uint64_t n_records = 0; uint32_t path_bytes = 1;
        ReadBatch* batch = nullptr; int slot = 0;
        bool busy = false, stop = false;
        double seconds = 0;                // time spent formatting and deflating
        std::string error;
        std::mutex mu; std::condition_variable cv;
        std::thread th;
        // the finished blocks of a batch go to the file from a thread of their own: the workers are already on the next
        // batch while the previous one is written (a BAM of a 1 M-read batch is ~70 MB: 0.1-0.5 s on a disk)
        std::vector<std::vector<uint8_t>> to_write;
        bool writing = false;
        std::thread wth;
        BamStage(BatchFeed& f, BamWriter* w, Info* inf, grootgpu_index* ix, const uint32_t* grb) : feed(f), bam(w), info(inf), index(ix), graph_ref_base(grb) {
            wth = std::thread([this] {
                while (true) {
                    std::vector<std::vector<uint8_t>> outs;
                    { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return stop || writing; }); if (!writing) return; outs.swap(to_write); }
                    for (auto& o : outs) bam->append_blocks(o);
                    { std::lock_guard<std::mutex> lk(mu); writing = false; }
                    cv.notify_all();
                }
            });
            th = std::thread([this] {
                while (true) {
                    { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return stop || busy; }); if (!busy) return; }
                    std::string e;
                    const double t0 = now_seconds();
                    try {
                        BamBatch bb;
                        bb.reads = batch; bb.cpairs = cpairs.data(); bb.n_pairs = cpairs.size(); bb.n_records = n_records;
                        bb.rec_path_c = rec_path.data(); bb.rec_path_bytes = path_bytes; bb.graph_ref_base = graph_ref_base;
                        bb.node_paths = [this](uint32_t node, NodePathsView* v) {
                            return grootgpu_index_node_paths(index, node, &v->graph, &v->ids, &v->pos, &v->n) == 0;
                        };
                        const unsigned workers = static_cast<unsigned>(std::max(1, info->NumProc));
                        std::vector<std::vector<uint8_t>> outs;
                        e = format_batch_bam(bb, workers, info->BamLevel, info->BamDelta, outs);
                        if (e.empty()) {                             // batches are written in order: wait for the one in front
                            std::unique_lock<std::mutex> lk(mu);
                            cv.wait(lk, [&] { return !writing; });
                            to_write.swap(outs); writing = true;
                            lk.unlock();
                            cv.notify_all();
                        }
                    } catch (std::exception& ex) { e = ex.what(); }
                    feed.release(slot);
                    { std::lock_guard<std::mutex> lk(mu); busy = false; seconds += now_seconds() - t0; if (error.empty()) error = e; }
                    cv.notify_all();
                }
            });
        }
        // waits for the batch in front, then takes this one; returns the first error of the stage so far ("" = none)
        std::string submit(int slot_i, ReadBatch* b, const grootgpu_batch_result& res) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return !busy; });
            if (!error.empty()) return error;
            cpairs.assign(res.cpairs, res.cpairs + res.n_pairs);
            const uint8_t* rp = static_cast<const uint8_t*>(res.rec_path_c);
            rec_path.assign(rp, rp + res.n_records * res.rec_path_bytes);
            n_records = res.n_records; path_bytes = res.rec_path_bytes; batch = b; slot = slot_i;
            busy = true;
            lk.unlock();
            cv.notify_all();
            return "";
        }
        std::string drain(double* busy_seconds) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !busy && !writing; }); *busy_seconds = seconds; return error; }
        ~BamStage() {
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !busy && !writing; }); stop = true; }
            cv.notify_all();
            if (th.joinable()) th.join();
            if (wth.joinable()) wth.join();
        }
    };
    std::unique_ptr<BamStage> stage;
    if (bam) stage.reset(new BamStage(feed, bam.get(), info_, index_, graph_ref_base.data()));
    for (int slot_i = 0;; slot_i = (slot_i + 1) % kSlots) {
        const double t_take = now_seconds();
        ReadBatch* bp = feed.take(slot_i);
        stage_seconds_[0] += now_seconds() - t_take;
        if (!bp) break;
        bool handed_over = false;
        struct Release { BatchFeed& f; int i; bool& handed; ~Release() { if (!handed) f.release(i); } } release_slot{feed, slot_i, handed_over};
        ReadBatch& b = *bp;
        grootgpu_batch_result res;
        {   // reads of one length (the usual Illumina run): the library then needs no offsets on the device
            const uint64_t L0 = b.size() ? b.seq_off[1] - b.seq_off[0] : 0;
            bool same = L0 > 0 && L0 < (1ull << 31);
            for (uint32_t i = 1; i < b.size() && same; i++) same = b.seq_off[i + 1] - b.seq_off[i] == L0;
            prm.fixed_read_len = same ? static_cast<uint32_t>(L0) : 0u;
            team.prm.fixed_read_len = prm.fixed_read_len;
        }
        const double t_dev = now_seconds();
        if (indexes_.size() > 1) {
            if ((rc = team.run(b.seq.data(), b.seq_off.data(), b.size(), &res, &err_))) return rc;
        } else {
            rc = grootgpu_align_batch(index_, b.seq.data(), b.seq_off.data(), b.size(), &prm, &res);
            if (rc) { err_ = grootgpu_last_error(); return rc; }
        }
        stage_seconds_[1] += now_seconds() - t_dev;
        read_stats_[0] += res.received; read_stats_[1] += res.mapped; read_stats_[2] += res.multimapped;
        alignment_count_ += res.alignments;
        if (!stage) continue;
        const std::string werr = stage->submit(slot_i, bp, res);
        if (!werr.empty()) { err_ = werr; return GROOTGPU_ERR_FORMAT; }
        handed_over = true;
    }
    if (stage) {
        const std::string werr = stage->drain(&stage_seconds_[2]);
        stage.reset();
        if (!werr.empty()) { err_ = werr; return GROOTGPU_ERR_FORMAT; }
    }
    if (reads.rawCount() == 0) { err_ = "no fastq reads received"; return GROOTGPU_ERR_EMPTY; }   // sketch.go:275-277
    if (bam) { bam->close(); if (fh != stdout) fclose(fh); }
    if (indexes_.size() > 1 && (rc = team.run(nullptr, nullptr, 0, nullptr, &err_))) return rc;   // collect the graph weights on the first GPU
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

// ---- groot report (src/reporting/reporting.go) ---------------------------------------------------------------------
// reporting.go:178-213, kept statement for statement (including what it does with the last element)
std::string cigarClean(const std::vector<char>& str, bool* internal_d) {
    int counter = 1;
    char preVal = str.empty() ? 'M' : str[0];
    std::string cigar;
    int nD = 0, nM = 0;
    auto bump = [&](char v) { if (v == 'D') nD++; else nM++; };
    for (size_t i = 0; i < str.size(); i++) {
        const char val = str[i];
        if (i == 0) continue;
        if (i == str.size() - 1) {
            if (val == preVal) { counter++; cigar += std::to_string(counter) + val; bump(val); }
            else { cigar += std::to_string(counter) + preVal + "1" + val; bump(val); }
            break;
        }
        if (val == preVal) counter++;
        else { bump(preVal); cigar += std::to_string(counter) + preVal; preVal = val; counter = 1; }
    }
    *internal_d = !(((nD + nM) <= 2) || (nD == 2 && nM == 1));
    return cigar;
}

std::vector<ReportLine> RunReport(const std::string& input_file, double coverage_cutoff, bool low_cov) {
    gzFile gz = input_file.empty() ? gzdopen(0, "rb") : gzopen(input_file.c_str(), "rb");   // BGZF == concatenated gzip members
    if (!gz) throw std::runtime_error("could not open BAM file " + input_file);
    gzbuffer(gz, 1 << 20);
    auto need = [&](void* dst, size_t n) {
        size_t got = 0;
        while (got < n) {
            const int r = gzread(gz, static_cast<char*>(dst) + got, static_cast<unsigned>(std::min<size_t>(n - got, 1u << 30)));
            if (r <= 0) return got;
            got += static_cast<size_t>(r);
        }
        return got;
    };
    auto fail = [&](const char* m) { gzclose(gz); throw std::runtime_error(std::string("could not read BAM file: ") + m); };
    char magic[4];
    if (need(magic, 4) != 4 || memcmp(magic, "BAM\1", 4) != 0) fail("bad magic");
    int32_t l_text = 0, n_ref = 0;
    if (need(&l_text, 4) != 4 || l_text < 0) fail("truncated header");
    std::string text(static_cast<size_t>(l_text), 0);
    if (need(&text[0], text.size()) != text.size() || need(&n_ref, 4) != 4 || n_ref < 0) fail("truncated header");
    std::vector<std::pair<std::string, int32_t>> refs(static_cast<size_t>(n_ref));
    for (auto& r : refs) {
        int32_t l_name = 0;
        if (need(&l_name, 4) != 4 || l_name < 1) fail("truncated reference list");
        r.first.assign(static_cast<size_t>(l_name), 0);
        if (need(&r.first[0], r.first.size()) != r.first.size() || need(&r.second, 4) != 4) fail("truncated reference list");
        r.first.pop_back();
    }
    // records: (start, reference length of the alignment) per reference; Flags == 4 (unaligned) are skipped (reporting.go:79-83)
    std::vector<std::vector<std::pair<int32_t, int32_t>>> recs(refs.size());
    std::vector<uint8_t> rec;
    while (true) {
        int32_t block = 0;
        const size_t g = need(&block, 4);
        if (g == 0) break;
        if (g != 4 || block < 32) fail("truncated record");
        rec.resize(static_cast<size_t>(block));
        if (need(rec.data(), rec.size()) != rec.size()) fail("truncated record");
        int32_t ref_id, pos; uint16_t n_cig, flag; uint8_t l_name;
        memcpy(&ref_id, &rec[0], 4); memcpy(&pos, &rec[4], 4); l_name = rec[8]; memcpy(&n_cig, &rec[12], 2); memcpy(&flag, &rec[14], 2);
        if (flag == 4) continue;
        if (ref_id < 0 || static_cast<size_t>(ref_id) >= refs.size() || 32u + l_name + 4u * n_cig > rec.size()) fail("bad record");
        int32_t len = 0;                                       // sam.Record.Len(): reference bases consumed (M, D, N, =, X)
        for (uint16_t c = 0; c < n_cig; c++) {
            uint32_t op; memcpy(&op, &rec[32 + l_name + 4 * c], 4);
            const uint32_t t = op & 15u;
            if (t == 0 || t == 2 || t == 3 || t == 7 || t == 8) len += static_cast<int32_t>(op >> 4);
        }
        recs[static_cast<size_t>(ref_id)].push_back({pos, len});
    }
    gzclose(gz);
    std::vector<ReportLine> out;
    for (size_t r = 0; r < refs.size(); r++) {
        if (recs[r].empty() || refs[r].second <= 0) continue;
        std::vector<int> pileup(static_cast<size_t>(refs[r].second), 0);
        for (auto& pr : recs[r]) {
            int start = pr.first, end = pr.first + pr.second;                 // inclusive end, as the reference iterates (reporting.go:108-121)
            if (end > static_cast<int>(pileup.size()) - 1) end = static_cast<int>(pileup.size()) - 1;
            for (int i = std::max(start, 0); i <= end; i++) pileup[static_cast<size_t>(i)]++;
        }
        size_t covered = 0;
        for (int v : pileup) covered += v != 0;
        if (static_cast<double>(covered) / static_cast<double>(pileup.size()) < coverage_cutoff) continue;
        std::vector<char> cig(pileup.size());
        for (size_t i = 0; i < pileup.size(); i++) cig[i] = pileup[i] == 0 ? 'D' : 'M';
        bool internal_d = false;
        std::string clean = cigarClean(cig, &internal_d);
        if (internal_d && low_cov) continue;
        std::string name = refs[r].first;
        if (!name.empty() && name[0] == '*') name.erase(0, 1);                // cluster representative marker (reporting.go:130-133)
        out.push_back({name, recs[r].size(), refs[r].second, clean});
    }
    return out;
}

}  // namespace groot_host
