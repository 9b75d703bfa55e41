// ORACLE — TEST INFRASTRUCTURE ONLY (see nthash.hpp header).
//
// CPU restatement of the reference's src/minhash/khf.go (KHF sketch) and src/seqio/seqio.go
// (FASTQread, RunMinHash, BaseCheck, RevComplement, QualTrim).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "nthash.hpp"

namespace oracle {

// src/minhash/khf.go:18-56 — K-Hash-Functions MinHash: slot i keeps min over k-mers of multihash i.
// Returns false when the hasher cannot be built (len < k), mirroring the AddSequence error
// (khf.go:38-41); the sketch is then left at MaxUint64 exactly like the reference's.
inline bool khf_sketch(const uint8_t* seq, size_t len, unsigned k, unsigned S, uint64_t* sketch) {
    for (unsigned i = 0; i < S; i++) sketch[i] = UINT64_MAX;  // khf.go:21-24
    NtHasher hasher(seq, len, k);
    if (!hasher.ok()) return false;
    std::vector<uint64_t> mh(S);
    uint64_t h;
    while (hasher.next(true, &h)) {                           // khf.go:44 (CANONICAL)
        multi_hash(h, k, S, mh.data());
        for (unsigned i = 0; i < S; i++)
            if (mh[i] < sketch[i]) sketch[i] = mh[i];          // khf.go:47-51
    }
    return true;
}

// src/seqio/seqio.go:26-37
struct FASTQread {
    std::string id;    // line 1 including the leading '@'
    std::string seq;
    std::string misc;
    std::string qual;  // raw ASCII, never de-offset
    bool rc = false;
};

// src/seqio/seqio.go:40-68 with kmv=false (the only mode any command uses: boss.go:163, graph.go:293)
inline bool run_minhash(const std::string& seq, unsigned k, unsigned S, std::vector<uint64_t>* sketch) {
    sketch->assign(S, 0);
    return khf_sketch(reinterpret_cast<const uint8_t*>(seq.data()), seq.size(), k, S, sketch->data());
}

// src/seqio/seqio.go:72-91 — upper-case, everything outside ACGTN becomes N
inline void base_check(std::string* seq) {
    for (auto& ch : *seq) {
        unsigned char c = static_cast<unsigned char>(ch);
        if (c >= 'a' && c <= 'z') c = static_cast<unsigned char>(c - 32);
        switch (c) {
            case 'A': case 'C': case 'T': case 'G': case 'N': ch = static_cast<char>(c); break;
            default: ch = 'N';
        }
    }
}

// src/seqio/seqio.go:17-23 — complementBases is a Go slice literal indexed by byte: length 'T'+1,
// zero everywhere except A,T,C,G,N. Indexing with a byte > 'T' panics in Go (index out of range).
struct BadBase : std::runtime_error { BadBase() : std::runtime_error("RevComplement: base > 'T' (Go index panic)") {} };
inline uint8_t complement_base(uint8_t b) {
    if (b > 'T') throw BadBase();
    switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'N': return 'N'; }
    return 0;
}

// src/seqio/seqio.go:120-133. NOTE the reference swaps Qual alongside Seq in the reversal loop and
// therefore requires len(Qual) >= len(Seq) (Go would panic otherwise); FASTA-mode reads (Qual nil,
// sketch.go:190) can therefore never be reverse complemented by the reference without a panic.
inline void rev_complement(FASTQread* r) {
    for (auto& ch : r->seq) ch = static_cast<char>(complement_base(static_cast<uint8_t>(ch)));
    if (!r->seq.empty() && r->qual.size() < r->seq.size()) throw std::runtime_error("RevComplement: qual shorter than seq (Go index panic)");
    for (long i = 0, j = static_cast<long>(r->seq.size()) - 1; i <= j; i++, j--) {
        std::swap(r->seq[i], r->seq[j]);
        std::swap(r->qual[i], r->qual[j]);
    }
    r->rc = !r->rc;
}

// src/seqio/seqio.go:141-170 (unused by the align pipeline; kept for the seqio_test.go goldens)
inline void qual_trim(FASTQread* r, int minQual) {
    const int encoding = 33;
    int start = 0, qualSum = 0, qualMax = 0;
    int end = static_cast<int>(r->qual.size());
    for (int i = 0; i < static_cast<int>(r->qual.size()); i++) {
        qualSum += minQual - (static_cast<int>(static_cast<uint8_t>(r->qual[i])) - encoding);
        if (qualSum < 0) break;
        if (qualSum > qualMax) { qualMax = qualSum; start = i + 1; }
    }
    qualSum = 0; qualMax = 0;
    for (int i = 0, j = static_cast<int>(r->qual.size()) - 1; j >= i; j--) {
        qualSum += minQual - (static_cast<int>(static_cast<uint8_t>(r->qual[j])) - encoding);
        if (qualSum < 0) break;
        if (qualSum > qualMax) { qualMax = qualSum; end = j; }
    }
    if (start >= end) { start = 0; end = 0; }
    r->seq = r->seq.substr(start, end - start);
    r->qual = r->qual.substr(start, end - start);
}

}  // namespace oracle
