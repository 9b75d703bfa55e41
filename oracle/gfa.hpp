// ORACLE — TEST INFRASTRUCTURE ONLY (see nthash.hpp header).
//
// CPU restatement of what the reference uses from github.com/will-rowe/gfa
// (v0.0.0-20190502084819-05c93955478b, go.mod:15; call sites src/pipeline/index.go:43,49 and
// src/graph/graph.go:51,91,111, src/graph/graphio.go:115-138): ReadMSA, MSA2GFA and a GFA v1 reader.
//
// PARITY STATUS: module not vendored -> restated from its published behaviour and pinned by the
// reference's golden pair  db/clustered-ARG-databases/1.1/arg-annot.90.tar:cluster-139.msa <->
// src/graph/test.gfa (133 S / 176 L / 6 P; checked in tests/test_oracle_kat.py up to the
// within-bubble segment numbering, which follows Go map iteration order in the reference and is
// therefore arbitrary there).  Deterministic choice made here: nodes of one MSA column are numbered
// by first occurrence in MSA row order.  Column bases are compared as raw bytes (no case folding);
// case folding happens later in CreateGrootGraph's BaseCheck (src/graph/graph.go:64-67).
#pragma once
#include <cstdint>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

struct GfaSegment { std::string name; std::string seq; double kc = 0.0; };
struct GfaLink { std::string from, to; };
struct GfaPath { std::string name; std::vector<std::string> segs; };
struct Gfa {
    std::vector<GfaSegment> segments;
    std::vector<GfaLink> links;
    std::vector<GfaPath> paths;
};

struct MsaRow { std::string name; std::string seq; };

// gfa.ReadMSA: FASTA-formatted alignment; the record named "consensus" is dropped; names keep a
// leading '*' (the seed sequence marker); name = header up to the first whitespace.
inline std::vector<MsaRow> read_msa_text(const std::string& text) {
    std::vector<MsaRow> rows;
    std::istringstream in(text);
    std::string line;
    bool have = false;
    MsaRow cur;
    auto flush = [&]() { if (have && cur.name != "consensus") rows.push_back(cur); };
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            flush();
            have = true;
            size_t e = line.find_first_of(" \t");
            cur.name = line.substr(1, e == std::string::npos ? std::string::npos : e - 1);
            cur.seq.clear();
        } else if (have) {
            cur.seq += line;
        }
    }
    flush();
    return rows;
}

inline std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss; ss << f.rdbuf();
    return ss.str();
}

// gfa.MSA2GFA: one node per distinct non-gap base per alignment column (holding the set of rows
// that use it); an edge between consecutive non-gap nodes of each row; squash maximal chains
// u->v where u has exactly one out-edge, v exactly one in-edge and both hold identical row sets;
// surviving nodes are numbered 1..N in column order; L lines "a + b + 0M"; P lines in MSA order.
inline Gfa msa2gfa(const std::vector<MsaRow>& rows) {
    if (rows.empty()) throw std::runtime_error("empty MSA");
    size_t ncol = rows[0].seq.size();
    for (auto& r : rows) if (r.seq.size() != ncol) throw std::runtime_error("MSA rows differ in length: " + r.name);
    struct N { std::string seq; std::vector<uint32_t> rows; std::vector<int> out; std::vector<int> in; bool dead = false; };
    std::vector<N> nodes;
    std::vector<int> last(rows.size(), -1);               // last node per row
    std::vector<std::vector<int>> rowNodes(rows.size());  // node chain per row
    for (size_t c = 0; c < ncol; c++) {
        int colNode[256];
        for (int& x : colNode) x = -1;
        for (size_t r = 0; r < rows.size(); r++) {
            uint8_t b = static_cast<uint8_t>(rows[r].seq[c]);
            if (b == '-') continue;
            if (colNode[b] < 0) { colNode[b] = static_cast<int>(nodes.size()); nodes.emplace_back(); nodes.back().seq.assign(1, static_cast<char>(b)); }
            int id = colNode[b];
            nodes[id].rows.push_back(static_cast<uint32_t>(r));
            if (last[r] >= 0) {
                auto& o = nodes[last[r]].out;
                bool seen = false;
                for (int x : o) if (x == id) { seen = true; break; }
                if (!seen) { o.push_back(id); nodes[id].in.push_back(last[r]); }
            }
            last[r] = id;
            rowNodes[r].push_back(id);
        }
    }
    // squash
    for (size_t u = 0; u < nodes.size(); u++) {
        if (nodes[u].dead) continue;
        while (nodes[u].out.size() == 1) {
            int v = nodes[u].out[0];
            if (nodes[v].in.size() != 1 || nodes[v].rows != nodes[u].rows) break;
            nodes[u].seq += nodes[v].seq;
            nodes[u].out = nodes[v].out;
            for (int w : nodes[u].out) for (int& p : nodes[w].in) if (p == v) p = static_cast<int>(u);
            nodes[v].dead = true;
            nodes[v].seq.clear();  // marker: merged away
            nodes[v].out.clear();
            // remember where v went so row chains can be rewritten
            nodes[v].in.assign(1, static_cast<int>(u));
        }
    }
    std::vector<int> newId(nodes.size(), 0);
    int next = 1;
    for (size_t u = 0; u < nodes.size(); u++) if (!nodes[u].dead) newId[u] = next++;
    Gfa g;
    for (size_t u = 0; u < nodes.size(); u++) {
        if (nodes[u].dead) continue;
        g.segments.push_back({std::to_string(newId[u]), nodes[u].seq, 0.0});
    }
    for (size_t u = 0; u < nodes.size(); u++) {
        if (nodes[u].dead) continue;
        for (int v : nodes[u].out) g.links.push_back({std::to_string(newId[u]), std::to_string(newId[v])});
    }
    for (size_t r = 0; r < rows.size(); r++) {
        GfaPath p; p.name = rows[r].name;
        int prev = -1;
        for (int id : rowNodes[r]) {
            if (nodes[id].dead) continue;  // merged into the chain head, which is already listed
            if (id != prev) p.segs.push_back(std::to_string(newId[id]));
            prev = id;
        }
        g.paths.push_back(std::move(p));
    }
    return g;
}

// Minimal GFA v1 reader (H/S/L/P lines; KC:i tag on segments as read by segment.GetKmerCount,
// src/graph/graph.go:70). Path segment names keep no orientation suffix.
inline Gfa read_gfa_text(const std::string& text) {
    Gfa g;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
        if (line.empty()) continue;
        std::vector<std::string> f;
        size_t s = 0;
        while (true) { size_t e = line.find('\t', s); f.push_back(line.substr(s, e == std::string::npos ? std::string::npos : e - s)); if (e == std::string::npos) break; s = e + 1; }
        if (f[0] == "S" && f.size() >= 3) {
            GfaSegment seg{f[1], f[2], 0.0};
            for (size_t i = 3; i < f.size(); i++) if (f[i].rfind("KC:i:", 0) == 0) seg.kc = std::stod(f[i].substr(5));
            g.segments.push_back(seg);
        } else if (f[0] == "L" && f.size() >= 5) {
            g.links.push_back({f[1], f[3]});
        } else if (f[0] == "P" && f.size() >= 3) {
            GfaPath p; p.name = f[1];
            std::string segs = f[2];
            size_t a = 0;
            while (a < segs.size()) {
                size_t e = segs.find(',', a);
                std::string tok = segs.substr(a, e == std::string::npos ? std::string::npos : e - a);
                if (!tok.empty() && (tok.back() == '+' || tok.back() == '-')) tok.pop_back();
                if (!tok.empty()) p.segs.push_back(tok);
                if (e == std::string::npos) break;
                a = e + 1;
            }
            g.paths.push_back(std::move(p));
        }
    }
    return g;
}

}  // namespace oracle
