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
//   3. project_accumulate  one warp per node adds the node's items to KmerFreq one after the other (the dependent
//      DADD chain IS the reference's order); KmerTotal is an integer sum (atomics are exact).
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
    const uint32_t* order;        // pair ids sorted by first window (or nullptr): the pairs a warp expands share their windows
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
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_segs; q += gridDim.x * blockDim.x) {
        const uint32_t s = a.order ? a.order[q] : q;
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

// keys sorted (stable): ONE WARP PER NODE. The additions of one node form a dependent DADD chain by definition (that IS
// the reference's order), and the hottest node's chain is the critical path of the whole kernel — so everything else
// is taken off it: lanes 0/1 find the node's segment [lower_bound(node), lower_bound(node + 1)) by binary search, the
// 32 lanes load 32 addends at a time (coalesced, the next 32 prefetched while the current ones are added) and every
// lane runs the same chain, fetching addend l from lane l by shuffle (shuffles do not depend on the accumulator, so
// they pipeline under the DADD latency). One thread per segment with loads ahead of use took ~90 cycles per addend
// (profiles/r01_ncu_summary.md); this form is bounded by the DADD latency alone.
__global__ void __launch_bounds__(256) project_accumulate_kernel(const uint32_t* __restrict__ keys, const double* __restrict__ vals,
                                                                 const uint32_t* __restrict__ n_items_ptr, uint32_t n_nodes,
                                                                 double* __restrict__ kmer_freq) {
    const uint32_t n = *n_items_ptr;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, total_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t node = gwarp; node < n_nodes; node += total_warps) {
        uint32_t lo = 0, hi = n;                              // lower_bound(node + (lane & 1))
        const uint32_t want = node + (lane & 1u);
        while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
        const uint32_t start = __shfl_sync(0xffffffffu, lo, 0), end = __shfl_sync(0xffffffffu, lo, 1);
        if (start >= end) continue;
        double acc = kmer_freq[node];
        double v = start + lane < end ? vals[start + lane] : 0.0;
        for (uint32_t base = start; base < end; base += 32) {
            const uint32_t nxt = base + 32 + lane;
            const double vn = nxt < end ? vals[nxt] : 0.0;
            const uint32_t cnt = end - base < 32u ? end - base : 32u;
#pragma unroll
            for (uint32_t l = 0; l < 32; l++) {
                const double x = __shfl_sync(0xffffffffu, v, l);
                if (l < cnt) acc = __dadd_rn(acc, x);        // node.go:25-28, in read order
            }
            v = vn;
        }
        if (lane == 0) kmer_freq[node] = acc;
    }
}

}  // namespace groot
