// a10 on the device: the ordered graph weighting of `groot align`.
//
// Replaces GrootGraph.IncrementSubPath as driven by the graph minion (src/graph/graph.go:401-451,
// src/pipeline/graphminion.go:60,67). The reference adds, per mapping and in read order, a k-mer share to
// KmerFreq of every contained node: an ORDER-DEPENDENT f64 accumulation per node. Bit-exact reproduction:
//   1. project_count / project_expand  one thread per (read, graph) pair: every (node, increment) item of the
//      pair's first n_incremented mappings, written at an exclusive-scan offset, i.e. in read order. The
//      increment is computed with IEEE-754 round-to-nearest double intrinsics in the reference's expression
//      order ((segLen/total)*numKmers)*count, so every addend equals the host's bit for bit.
//   2. cub::DeviceRadixSort::SortPairs by node id — a STABLE sort, so each node's items stay in read order.
//   3. project_accumulate  one thread per node segment adds its items to KmerFreq one after the other (the
//      dependent DADD chain IS the reference's order); KmerTotal is an integer sum (atomics are exact).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "align_kernels.cuh"
#include "device_types.cuh"

namespace groot {

struct ProjectArgs {
    const uint32_t* off;          // read offsets (lengths)
    const uint32_t* hits;
    const PairOut* pairs;
    const uint32_t* n_segs_ptr;
    const uint32_t* cn_count;     // per contained node: count (integer-valued f64 in the reference)
    uint32_t* item_cnt;           // [n_segs]
    const uint32_t* item_off;     // exclusive scan of item_cnt
    uint32_t* keys;               // [n_items] global node index
    double* vals;                 // [n_items]
    unsigned long long* kmer_total;   // [n_graphs]
    uint32_t k;
};

__global__ void __launch_bounds__(256) project_count_kernel(DevIndex ix, ProjectArgs a) {
    const uint32_t n_segs = *a.n_segs_ptr;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += gridDim.x * blockDim.x) {
        const PairOut p = a.pairs[s];
        uint32_t c = 0;
        for (uint32_t m = 0; m < p.n_incremented; m++) c += ix.wins[a.hits[p.hit_begin + m]].cn_cnt;
        a.item_cnt[s] = c;
    }
}

__global__ void __launch_bounds__(256) project_expand_kernel(DevIndex ix, ProjectArgs a) {
    const uint32_t n_segs = *a.n_segs_ptr;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += gridDim.x * blockDim.x) {
        const PairOut p = a.pairs[s];
        const uint32_t len = a.off[p.read + 1] - a.off[p.read];
        const double kmers = static_cast<double>(static_cast<int>(len) - static_cast<int>(a.k)) + 1.0;   // graphminion.go:60
        uint32_t o = a.item_off[s];
        for (uint32_t m = 0; m < p.n_incremented; m++) {
            const WinRec w = ix.wins[a.hits[p.hit_begin + m]];
            if (w.cn_cnt == 1) {                                    // graph.go:409-422: all k-mers, KmerTotal untouched
                a.keys[o] = ix.cn_node[w.cn_off]; a.vals[o] = kmers; o++;
                continue;
            }
            double total = 0.0;                                     // graph.go:427-434 (integers: exact in any order)
            for (uint32_t j = 0; j < w.cn_cnt; j++) total += static_cast<double>(ix.nodes[ix.cn_node[w.cn_off + j]].seq_len);
            for (uint32_t j = 0; j < w.cn_cnt; j++) {
                const uint32_t n = ix.cn_node[w.cn_off + j];
                const double ratio = __ddiv_rn(static_cast<double>(ix.nodes[n].seq_len), total);
                a.keys[o] = n;
                a.vals[o] = __dmul_rn(__dmul_rn(ratio, kmers), static_cast<double>(a.cn_count[w.cn_off + j]));   // graph.go:442
                o++;
            }
            atomicAdd(&a.kmer_total[w.graph], static_cast<unsigned long long>(kmers));   // graph.go:449 uint64(numKmers)
        }
    }
}

// keys sorted (stable): thread i owns the segment that starts at i, if any. The additions of one node form a
// dependent DADD chain by definition (that IS the reference's order); everything around it is taken off the chain:
// the segment end comes from a galloping + binary search, and the addends are loaded eight at a time ahead of use.
__global__ void __launch_bounds__(256) project_accumulate_kernel(const uint32_t* __restrict__ keys, const double* __restrict__ vals,
                                                                 const uint32_t* __restrict__ n_items_ptr, double* __restrict__ kmer_freq) {
    const uint32_t n = *n_items_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t key = keys[i];
        if (i > 0 && keys[i - 1] == key) continue;
        uint32_t lo = i, step = 1;                         // keys[lo] == key
        while (lo + step < n && keys[lo + step] == key) { lo += step; step <<= 1; }
        uint32_t hi = lo + step < n ? lo + step : n;       // keys[hi] != key or hi == n
        while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (keys[mid] == key) lo = mid; else hi = mid; }
        const uint32_t end = hi;
        double acc = kmer_freq[key];
        uint32_t j = i;
        for (; j + 8 <= end; j += 8) {
            const double v0 = vals[j], v1 = vals[j + 1], v2 = vals[j + 2], v3 = vals[j + 3];
            const double v4 = vals[j + 4], v5 = vals[j + 5], v6 = vals[j + 6], v7 = vals[j + 7];
            acc = __dadd_rn(acc, v0); acc = __dadd_rn(acc, v1); acc = __dadd_rn(acc, v2); acc = __dadd_rn(acc, v3);   // node.go:25-28, in read order
            acc = __dadd_rn(acc, v4); acc = __dadd_rn(acc, v5); acc = __dadd_rn(acc, v6); acc = __dadd_rn(acc, v7);
        }
        for (; j < end; j++) acc = __dadd_rn(acc, vals[j]);
        kmer_freq[key] = acc;
    }
}

}  // namespace groot
