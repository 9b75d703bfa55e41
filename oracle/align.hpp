// ORACLE — TEST INFRASTRUCTURE ONLY (see nthash.hpp header).
//
// CPU restatement of src/graph/alignment.go: AlignRead (4-stage hierarchical exact alignment),
// performAlignment, dfsRecursive, processTraversal. No scoring, no DP: exact match DFS in which a
// reference 'N' matches anything and a read may overhang a sink node (alignment.go:229).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "graph.hpp"

namespace oracle {

// what AlignRead puts into each sam.Record (alignment.go:114-156), before BAM materialisation
struct AlignRecord {
    int pathID = 0;        // record.Ref = references[ID]
    int pos = 0;           // record.Pos (0-based)
    int startClip = 0;     // leading H
    int endClip = 0;       // trailing H
    int seqLength = 0;     // M length; Seq/Qual = read[0:seqLength] (prefix even for a start clip)
    uint16_t flags = 0;    // 0x100 secondary, 0x10 reverse
};

// alignment.go:196-254
inline bool dfs_recursive(GrootGraph& g, GrootGraphNode* node, const char* read, int distance, std::vector<uint64_t>& path,
                          std::vector<std::vector<uint64_t>>& sendPath, int readLength, int offset) {
    if (offset >= static_cast<int>(node->sequence.size())) return false;
    for (size_t i = offset; i < node->sequence.size(); i++) {
        char base = node->sequence[i];
        if (distance == readLength) break;
        if (base == 'N') { distance++; continue; }
        if (base == read[distance]) distance++;
        else return false;
    }
    path.push_back(node->segmentID);
    bool result = false;
    if (distance == readLength || node->outEdges.empty()) {
        sendPath.push_back(path);
        result = true;
    } else {
        for (uint64_t nb : node->outEdges) {
            auto it = g.nodeLookup.find(nb);
            if (it == g.nodeLookup.end()) throw std::runtime_error("could not perform node lookup during alignment - possible incorrect seed");
            if (dfs_recursive(g, g.sortedNodes[it->second].get(), read, distance, path, sendPath, readLength, 0)) result = true;
        }
    }
    path.pop_back();  // Go passes the slice by value; popping restores the caller's view
    return result;
}

// alignment.go:263-317. IDs of one traversal are emitted in ascending pathID order (Go: map order).
inline void process_traversal(GrootGraph& g, const std::vector<std::vector<uint64_t>>& paths, int offset,
                              std::vector<int>* IDs, std::map<int, int>* startPositions) {
    IDs->clear(); startPositions->clear();
    for (auto& p : paths) {
        std::map<int, int> nodeIDs, startPos;
        int pathLength = static_cast<int>(p.size());
        for (int i = 0; i < pathLength; i++) {
            GrootGraphNode* node = g.sortedNodes[g.nodeLookup.at(p[i])].get();
            for (uint32_t id : node->pathIDs) {
                nodeIDs[static_cast<int>(id)]++;
                if (i == 0) {
                    auto pit = node->position.find(static_cast<int>(id));  // Go map read: 0 when absent
                    startPos[static_cast<int>(id)] = (pit == node->position.end() ? 0 : pit->second) + offset;
                }
            }
        }
        for (auto& kv : nodeIDs) if (kv.second >= pathLength) IDs->push_back(kv.first);
        for (auto& kv : startPos) if (!startPositions->count(kv.first)) (*startPositions)[kv.first] = kv.second;
    }
}

// alignment.go:162-193
inline void perform_alignment(GrootGraph& g, int nodeLookup, const char* read, int readLength, int offset,
                              std::vector<int>* IDs, std::map<int, int>* startPos) {
    IDs->clear(); startPos->clear();
    std::vector<std::vector<uint64_t>> paths;
    std::vector<uint64_t> path;
    dfs_recursive(g, g.sortedNodes[nodeLookup].get(), read, 0, path, paths, readLength, offset);
    if (!paths.empty()) process_traversal(g, paths, offset, IDs, startPos);
}

// alignment.go:13-159. `mapping` is taken by value: the reference mutates mapping.OffSet while
// shuffling and always restores it.
inline std::vector<AlignRecord> align_read(GrootGraph& g, const FASTQread& read, Key mapping) {
    const int MaxClip = 1;
    auto nl = g.nodeLookup.find(mapping.node);
    if (nl == g.nodeLookup.end()) throw std::runtime_error("could not perform node lookup during alignment - possible incorrect seed");
    int nodeLookup = nl->second;
    std::vector<int> IDs;
    std::map<int, int> startPos;
    int startClippedBases = 0, endClippedBases = 0;
    uint32_t origOffSet = mapping.offSet;
    const char* seq = read.seq.data();
    int L = static_cast<int>(read.seq.size());

    // 1. exact alignment and seed offset shuffling
    for (int shuffles = 0; shuffles <= static_cast<int>(mapping.mergeSpan + mapping.windowSize); shuffles++) {
        perform_alignment(g, nodeLookup, seq, L, static_cast<int>(mapping.offSet), &IDs, &startPos);
        if (!IDs.empty()) break;
        mapping.offSet++;
    }
    mapping.offSet = origOffSet;

    // 2. exact alignment and seed node shuffling (ContainedNodes in ascending segment-ID order)
    if (IDs.empty()) {
        for (auto& cn : mapping.containedNodes) {
            mapping.offSet = 0;
            for (int shuffles = 0; shuffles <= 10; shuffles++) {
                auto it = g.nodeLookup.find(cn.first);
                if (it == g.nodeLookup.end()) throw std::runtime_error("could not perform node lookup during alignment - possible incorrect seed");
                perform_alignment(g, it->second, seq, L, static_cast<int>(mapping.offSet), &IDs, &startPos);
                if (!IDs.empty()) break;
                mapping.offSet++;
            }
            if (!IDs.empty()) break;
        }
        mapping.offSet = origOffSet;
    }

    // 3. hard clipping the start of the read
    if (IDs.empty()) {
        for (int i = 1; i <= MaxClip; i++) {
            if (L - i < 0) break;  // Go would panic slicing an empty read; unreachable for len >= k
            perform_alignment(g, nodeLookup, seq + i, L - i, static_cast<int>(mapping.offSet), &IDs, &startPos);
            startClippedBases++;
            if (!IDs.empty()) break;
        }
    }

    // 4. hard clipping the end of the read
    if (IDs.empty()) {
        startClippedBases = 0;
        for (int i = MaxClip; i > 0; i--) {
            perform_alignment(g, nodeLookup, seq, L - 1, static_cast<int>(mapping.offSet), &IDs, &startPos);
            endClippedBases++;
            if (!IDs.empty()) break;
        }
    }

    std::vector<AlignRecord> out;
    if (IDs.empty()) return out;
    for (size_t c = 0; c < IDs.size(); c++) {
        AlignRecord r;
        r.seqLength = L - endClippedBases - startClippedBases;
        r.pathID = IDs[c];
        r.pos = startPos[IDs[c]];
        r.startClip = startClippedBases;
        r.endClip = endClippedBases;
        if (IDs.size() > 1 && c != 0) r.flags |= 0x100;
        if (read.rc) r.flags |= 0x10;
        out.push_back(r);
    }
    return out;
}

}  // namespace oracle
