// ORACLE — TEST INFRASTRUCTURE ONLY (see nthash.hpp header).
//
// CPU restatement of the reference's variation graph: src/graph/graph.go (CreateGrootGraph,
// topoSort, GetPaths, Graph2Seqs, WindowGraph, IncrementSubPath, Prune), src/graph/node.go,
// src/graph/graphio.go (SaveGraphAsGFA, GetSAMrefs) and the window Key of src/lshe/lshe.go:17-28.
//
// Go maps are restated as std::map so that every "range over a map" of the reference (undefined
// order there) becomes ascending-key order here — the documented deterministic tie-break shared
// with the CUDA path (DESIGN.md "Reference nondeterminism").
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "gfa.hpp"
#include "seqio.hpp"

namespace oracle {

// src/lshe/lshe.go:17-28
struct Key {
    uint32_t graphID = 0;
    uint64_t node = 0;     // segment ID of the first node in the window
    uint32_t offSet = 0;   // offset of the window within that node
    std::map<uint64_t, double> containedNodes;
    std::vector<uint32_t> ref;
    std::vector<uint64_t> sketch;
    uint32_t mergeSpan = 0;
    uint32_t windowSize = 0;
};

// src/graph/node.go:13-22
struct GrootGraphNode {
    uint64_t segmentID = 0;
    double segmentLength = 0;
    std::string sequence;
    std::vector<uint64_t> outEdges;
    std::vector<uint32_t> pathIDs;
    std::map<int, int> position;  // pathID -> 0-based start of this segment in that path
    double kmerFreq = 0;
    bool marked = false;
};

// src/graph/graph.go:18-34
struct GrootGraph {
    uint32_t graphID = 0;
    std::vector<std::shared_ptr<GrootGraphNode>> sortedNodes;
    std::map<uint32_t, std::string> paths;
    std::map<uint32_t, int> lengths;
    std::map<uint64_t, int> nodeLookup;
    bool masked = false;
    uint64_t kmerTotal = 0;
    int numWindows = 0;
    int numDistinctSketches = 0;

    GrootGraphNode* getNode(uint64_t id) {  // graph.go:529-535
        auto it = nodeLookup.find(id);
        if (it == nodeLookup.end()) throw std::runtime_error("can't find node in graph");
        return sortedNodes[it->second].get();
    }

    // graph.go:193-218
    void traverse(std::shared_ptr<GrootGraphNode> node, std::map<uint64_t, std::shared_ptr<GrootGraphNode>>& nodeMap,
                  std::set<uint64_t>& seen, std::vector<std::shared_ptr<GrootGraphNode>>& reversed) {
        if (seen.count(node->segmentID)) return;
        if (!nodeMap.count(node->segmentID)) return;
        seen.insert(node->segmentID);
        std::sort(node->outEdges.begin(), node->outEdges.end(), std::greater<uint64_t>());  // sort.Reverse, graph.go:203
        for (size_t i = 0; i < node->outEdges.size(); i++) {
            auto it = nodeMap.find(node->outEdges[i]);
            if (it == nodeMap.end()) continue;
            traverse(it->second, nodeMap, seen, reversed);
        }
        nodeMap.erase(node->segmentID);
        seen.erase(node->segmentID);
        reversed.push_back(node);  // reference prepends; we append and reverse once at the end
        nodeLookup[node->segmentID] = static_cast<int>(nodeMap.size());
    }

    // graph.go:150-190. seenPaths is never written in the reference, so every node is appended to
    // toposortStart once per path it belongs to (graph.go:160-164) — restated as such.
    void topoSort() {
        std::map<uint64_t, std::shared_ptr<GrootGraphNode>> nodeMap;
        std::vector<uint64_t> toposortStart;
        for (auto& node : sortedNodes) {
            if (paths.empty()) break;  // len(seenPaths)==len(Paths) only when there are no paths
            for (size_t i = 0; i < node->pathIDs.size(); i++) toposortStart.push_back(node->segmentID);
            if (nodeMap.count(node->segmentID)) throw std::runtime_error("graph contains duplicate nodes (identical segment IDs)");
            nodeMap[node->segmentID] = node;
        }
        sortedNodes.clear();
        nodeLookup.clear();
        std::set<uint64_t> seen;
        std::vector<std::shared_ptr<GrootGraphNode>> reversed;
        while (nodeMap.size() > 1) {
            size_t before = nodeMap.size();
            for (uint64_t start : toposortStart) {
                auto it = nodeMap.find(start);
                if (it == nodeMap.end()) continue;
                traverse(it->second, nodeMap, seen, reversed);
            }
            if (nodeMap.size() == before) throw std::runtime_error("topological sort cannot make progress (reference would spin)");
        }
        if (!nodeMap.empty()) throw std::runtime_error("topological sort failed - too many nodes remaining in the pre-sort list");
        sortedNodes.assign(reversed.rbegin(), reversed.rend());
    }

    // graph.go:575-622 (position bookkeeping part) + graph.go:625-644
    std::map<uint32_t, std::string> graph2Seqs() {
        if (paths.empty()) throw std::runtime_error("no paths recorded in current graph");
        std::map<uint32_t, std::string> seqs;
        for (auto& kv : paths) {
            uint32_t pathID = kv.first;
            int refLength = 0;
            std::string s;
            for (auto& node : sortedNodes) {
                for (uint32_t id : node->pathIDs) {
                    if (id == pathID) {
                        node->position[static_cast<int>(pathID)] = refLength;
                        refLength += static_cast<int>(node->sequence.size());
                        s += node->sequence;
                    }
                }
            }
            seqs[pathID] = s;
        }
        return seqs;
    }

    // graph.go:401-451
    void incrementSubPath(const std::map<uint64_t, double>& containedNodes, double numKmers) {
        if (containedNodes.empty()) throw std::runtime_error("ContainedNodes encountered that does not include any segments");
        if (containedNodes.size() == 1) {
            getNode(containedNodes.begin()->first)->kmerFreq += numKmers;  // no KmerTotal bump (early return, graph.go:409-422)
            return;
        }
        double totalLength = 0.0;
        for (auto& kv : containedNodes) totalLength += getNode(kv.first)->segmentLength;
        for (auto& kv : containedNodes) {
            GrootGraphNode* node = getNode(kv.first);
            double kmerShare = ((node->segmentLength / totalLength) * numKmers) * kv.second;
            node->kmerFreq += kmerShare;
        }
        kmerTotal += static_cast<uint64_t>(numKmers);
    }

    // graph.go:455-525
    bool prune(double minKmerCoverage) {
        std::set<uint32_t> removePathID;
        std::set<uint64_t> removeNode;
        for (auto& node : sortedNodes) {
            double perbase = node->kmerFreq / node->segmentLength;
            if (perbase < minKmerCoverage)
                for (uint32_t id : node->pathIDs) { removePathID.insert(id); removeNode.insert(node->segmentID); }
        }
        if (removePathID.size() == paths.size()) return false;
        if (removeNode.empty()) return true;
        for (auto& node : sortedNodes) {
            std::vector<uint32_t> up;
            for (uint32_t id : node->pathIDs) if (!removePathID.count(id)) up.push_back(id);
            node->pathIDs = up;
            if (removeNode.count(node->segmentID)) { node->marked = true; nodeLookup.erase(node->segmentID); }
            std::vector<uint64_t> ue;
            for (uint64_t e : node->outEdges) if (!removeNode.count(e)) ue.push_back(e);
            node->outEdges = ue;
        }
        for (uint32_t id : removePathID) if (paths.count(id)) lengths[id] = 0;
        return true;
    }

    // graphio.go:19-112 minus the timestamp comment line (graphio.go:22-23), which can never be
    // reproduced. Returns "" when no node carries weight (graph not written, graphio.go:67-69).
    std::string toGFA(long totalKmers) const {
        bool used = false;
        std::string out = "H\tVN:Z:1\n";
        out += "#\tthis graph is approximately weighted using k-mer frequencies from projected read sketches (total k-mers projected across all graphs: " + std::to_string(totalKmers) + ")\n";
        std::string links;
        for (auto& node : sortedNodes) {
            if (node->marked) continue;
            if (node->kmerFreq > 0) used = true;
            char buf[64];
            snprintf(buf, sizeof buf, "%lld", static_cast<long long>(node->kmerFreq));  // int(node.KmerFreq)
            out += "S\t" + std::to_string(node->segmentID) + "\t" + node->sequence + "\tKC:i:" + buf + "\n";
            for (uint64_t e : node->outEdges)
                links += "L\t" + std::to_string(node->segmentID) + "\t+\t" + std::to_string(e) + "\t+\t0M\n";
        }
        if (!used) return "";
        out += links;
        for (auto& kv : paths) {
            auto lit = lengths.find(kv.first);
            if (lit != lengths.end() && lit->second == 0) continue;
            std::string segs, ovl;
            for (auto& node : sortedNodes) {
                if (node->marked) continue;
                for (uint32_t id : node->pathIDs) if (id == kv.first) {
                    if (!segs.empty()) { segs += ","; ovl += ","; }
                    segs += std::to_string(node->segmentID) + "+";
                    ovl += std::to_string(node->sequence.size()) + "M";
                    break;
                }
            }
            out += "P\t" + kv.second + "\t" + segs + "\t" + ovl + "\n";
        }
        return out;
    }
};

// graph.go:37-147
inline std::shared_ptr<GrootGraph> create_groot_graph(const Gfa& gfa, int id) {
    auto g = std::make_shared<GrootGraph>();
    g->graphID = static_cast<uint32_t>(id);
    int it = 0;
    for (auto& seg : gfa.segments) {
        size_t pos = 0;
        long segID = 0;
        try { segID = std::stol(seg.name, &pos); } catch (...) { pos = 0; }
        if (pos != seg.name.size() || seg.name.empty()) throw std::runtime_error("could not convert segment name from GFA into an int for groot graph: " + seg.name);
        auto n = std::make_shared<GrootGraphNode>();
        n->segmentID = static_cast<uint64_t>(segID);
        n->sequence = seg.seq;
        base_check(&n->sequence);
        n->segmentLength = static_cast<double>(n->sequence.size());
        n->kmerFreq = seg.kc;
        g->sortedNodes.push_back(n);
        g->nodeLookup[n->segmentID] = it++;
        g->kmerTotal += static_cast<uint64_t>(seg.kc);
    }
    for (auto& l : gfa.links) {
        uint64_t from = std::stoull(l.from), to = std::stoull(l.to);
        g->sortedNodes[g->nodeLookup.at(from)]->outEdges.push_back(to);
    }
    for (uint32_t p = 0; p < gfa.paths.size(); p++) {
        g->paths[p] = gfa.paths[p].name;
        for (auto& s : gfa.paths[p].segs) {
            uint64_t segID = std::stoull(s);
            g->sortedNodes[g->nodeLookup.at(segID)]->pathIDs.push_back(p);
        }
    }
    if (g->sortedNodes.size() > 1) g->topoSort();
    auto seqs = g->graph2Seqs();
    for (auto& kv : seqs) g->lengths[kv.first] = static_cast<int>(kv.second.size());
    return g;
}

// graph.go:229-396. Deterministic restatement of the goroutine fan-out: paths are windowed in
// ascending pathID order and their windows arrive in path order (the reference's arrival order
// is whatever the scheduler produces). Returns map "g%dn%do%d" -> Keys.
inline std::map<std::string, std::vector<Key>> window_graph(GrootGraph& g, int windowSize, int kmerSize, int sketchSize) {
    auto pathSeqs = g.graph2Seqs();
    g.numWindows = 0;
    for (auto& kv : g.lengths) g.numWindows += kv.second - windowSize + 1;
    std::vector<Key> pathWindows;
    for (auto& pk : g.paths) {
        uint32_t pathID = pk.first;
        int pathLength = g.lengths[pathID];
        if (pathLength < windowSize) throw std::runtime_error("graph contains sequence < window size");
        const std::string& pathSequence = pathSeqs[pathID];
        std::vector<uint64_t> segs(pathLength);
        std::vector<uint32_t> offSets(pathLength);
        int iterator = 0;
        for (auto& node : g.sortedNodes)
            for (uint32_t id : node->pathIDs)
                if (id == pathID)
                    for (uint32_t off = 0; off < node->sequence.size(); off++) { segs[iterator] = node->segmentID; offSets[iterator] = off; iterator++; }
        if (iterator != pathLength) throw std::runtime_error("windowing did not traverse entire path");
        Key holder;
        bool sketchSent = false;
        int numWindows = pathLength - windowSize + 1;
        std::vector<uint64_t> sketch;
        for (int i = 0; i < numWindows; i++) {
            if (!run_minhash(pathSequence.substr(i, windowSize), kmerSize, sketchSize, &sketch)) throw std::runtime_error("window shorter than k");
            bool merge = false;
            if (i != 0) {
                if (holder.sketch != sketch) { pathWindows.push_back(holder); sketchSent = true; }
                else merge = true;
            }
            if (!merge) {
                holder = Key();
                holder.graphID = g.graphID; holder.node = segs[i]; holder.offSet = offSets[i];
                holder.ref = {pathID}; holder.sketch = sketch; holder.mergeSpan = 0; holder.windowSize = static_cast<uint32_t>(windowSize);
            }
            for (int j = i; j < i + windowSize; j++) holder.containedNodes[segs[j]] += 1.0;
            if (merge) holder.mergeSpan++;
            // quirk (graph.go:336-338): the last window group of a path is only emitted when
            // nothing was emitted before it
            if (!sketchSent && i == numWindows - 1) pathWindows.push_back(holder);
        }
    }
    std::map<std::string, std::vector<Key>> windowLookup;
    for (auto& window : pathWindows) {
        char kb[96];
        snprintf(kb, sizeof kb, "g%un%lluo%u", window.graphID, static_cast<unsigned long long>(window.node), window.offSet);
        auto it = windowLookup.find(kb);
        if (it != windowLookup.end()) {
            bool dup = false;
            for (auto& existing : it->second) {
                if (existing.sketch == window.sketch) {
                    // graph.go:361-376: ContainedNodes is a map (shared by the range copy) so the
                    // frequencies accumulate; the Ref append and the MergeSpan max are applied to a
                    // copy of the struct and are LOST in the reference — restated as lost.
                    for (auto& kv : window.containedNodes) existing.containedNodes[kv.first] += kv.second;
                    dup = true;
                    break;
                }
            }
            if (!dup) { it->second.push_back(window); g.numDistinctSketches++; }
        } else {
            windowLookup[kb] = {window};
            g.numDistinctSketches++;
        }
    }
    if (g.numDistinctSketches == 0) throw std::runtime_error("no sketches produced after windowing graph seqs");
    return windowLookup;
}

}  // namespace oracle
