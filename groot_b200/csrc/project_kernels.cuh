// a10 on the device: the ordered graph weighting of `groot align`.
//
// Replaces GrootGraph.IncrementSubPath as driven by the graph minion (src/graph/graph.go:401-451,
// src/pipeline/graphminion.go:60,67). The reference adds, per mapping and in read order, a k-mer share to
// KmerFreq of every contained node: an ORDER-DEPENDENT f64 accumulation per node. Bit-exact reproduction:
//   1. project_count / project_expand  every (node, increment) item of each pair's first n_incremented mappings,
//      written at an exclusive-scan offset over the mappings, i.e. in read order. The increment is computed with
//      IEEE-754 round-to-nearest double operations in the reference's expression order
//      ((segLen/total)*numKmers)*count, so every addend equals the host's bit for bit.
//   2. cub::DeviceRadixSort::SortPairs by node id — a STABLE sort, so each node's items stay in read order.
//   3. project_accumulate  one warp per node adds the node's items to KmerFreq one after the other (the dependent
//      DADD chain IS the reference's order); KmerTotal is an integer sum (atomics are exact). The chains run on their
//      own stream behind the batch (capi.cu, acc_enqueue); on N GPUs the weight vector travels rank to rank between them.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "align_kernels.cuh"
#include "device_types.cuh"

namespace groot {

struct ProjectArgs {
    const uint32_t* off;          // read offsets (lengths)
    const uint32_t* hits;         // [n_hits] window ids == the mappings, in (read, graph, Node, OffSet) order
    const uint32_t* hit_read;     // [n_hits]
    const PairOut* pairs;
    const uint32_t* n_segs_ptr;
    const uint32_t* n_hits_ptr;
    const uint32_t* cn_count;     // per contained node: count (integer-valued f64 in the reference)
    const double* cn_ratio;       // per contained node: f64(SegmentLength) / f64(sum of the window's SegmentLengths), graph.go:427-441
    uint32_t* item_cnt;           // [n_hits] weight increments of the mapping (0 when the minion loop stopped before it)
    const uint32_t* item_off;     // exclusive scan of item_cnt
    uint32_t* keys;               // [n_items] global node index
    double* vals;                 // [n_items]
    unsigned long long* kmer_total;   // [n_graphs]
    uint32_t k;
};

// one thread per (read, graph) pair: the first n_incremented mappings of the pair are weighted (graphminion.go:64-98)
__global__ void __launch_bounds__(256) project_count_kernel(DevIndex ix, ProjectArgs a) {
    const uint32_t n_segs = *a.n_segs_ptr;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += gridDim.x * blockDim.x) {
        const PairOut p = a.pairs[s];
        for (uint32_t m = 0; m < p.hit_count; m++)
            a.item_cnt[p.hit_begin + m] = m < p.n_incremented ? ix.wins[a.hits[p.hit_begin + m]].cn_cnt : 0u;
    }
}

// ONE WARP PER 32 MAPPINGS, their (node, increment) items pooled: an inclusive scan of the item counts turns the
// 32 ContainedNodes lists into one run of items which the lanes expand 32 at a time — consecutive items go to
// consecutive slots of keys[] / vals[] (the scan offsets are contiguous), so the stores are coalesced, and no lane
// idles on a neighbour's longer list (one thread per pair ran at ~10 active lanes). The increment is
// ((segLen / total) * numKmers) * count in the reference's expression order (graph.go:442), segLen / total being the
// same correctly rounded f64 quotient whether the host (cn_ratio) or the device computes it.
__global__ void __launch_bounds__(256) project_expand_kernel(DevIndex ix, ProjectArgs a) {
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t n_hits = *a.n_hits_ptr;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, total_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t hb = gwarp * 32; hb < n_hits; hb += total_warps * 32) {
        const uint32_t h = hb + lane;
        uint32_t cnt = 0, cn_off = 0;
        double kmers = 0.0;
        if (h < n_hits) {
            cnt = a.item_cnt[h];
            if (cnt) {
                const WinRec w = ix.wins[a.hits[h]];
                const uint32_t r = a.hit_read[h];
                cn_off = w.cn_off;
                kmers = static_cast<double>(static_cast<int>(a.off[r + 1] - a.off[r]) - static_cast<int>(a.k)) + 1.0;   // graphminion.go:60
                if (cnt > 1) atomicAdd(&a.kmer_total[w.graph], static_cast<unsigned long long>(kmers));                  // graph.go:449 uint64(numKmers)
            }
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); if (lane >= static_cast<uint32_t>(d)) incl += v; }
        const uint32_t excl = incl - cnt, total = __shfl_sync(FULL, incl, 31);
        const uint32_t out0 = a.item_off[hb];                         // items of this warp's mappings start here
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t item = base + lane;
            uint32_t o = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t v = __shfl_sync(FULL, incl, (o + step - 1) & 31u);
                if (v <= item) o += step;
            }
            o &= 31u;
            const uint32_t j = item - __shfl_sync(FULL, excl, o);
            const uint32_t o_cn = __shfl_sync(FULL, cn_off, o), o_cnt = __shfl_sync(FULL, cnt, o);
            const double o_kmers = __shfl_sync(FULL, kmers, o);
            if (item < total) {
                const uint32_t c = o_cn + j;
                a.keys[out0 + item] = ix.cn_node[c];
                a.vals[out0 + item] = o_cnt == 1 ? o_kmers                                            // graph.go:409-422: all k-mers, KmerTotal untouched
                                                 : __dmul_rn(__dmul_rn(a.cn_ratio[c], o_kmers), static_cast<double>(a.cn_count[c]));
            }
        }
    }
}

// keys sorted (stable): ONE WARP PER NODE. The additions of one node form a dependent DADD chain by definition (that IS
// the reference's order), and the hottest node's chain is the critical path of the whole kernel — on N GPUs of the whole
// job, since the chains of the ranks run one after the other — so everything else is taken off it: lanes 0/1 find the
// node's segment [lower_bound(node), lower_bound(node + 1)) by binary search, the 32 lanes load 64 addends at a time
// (coalesced 16-byte loads, the next 64 requested before the current ones are added) and park them in the warp's slice of
// shared memory; the chain itself is then one 16-byte shared-memory load per two additions. 1.5 issue slots per addend
// (a shuffle-fed chain needs 3): the kernel shares its SMs with the mapping kernels of the following batch, and every
// instruction of the chain has to win an issue slot against them.
constexpr int kAccWarps = 8;     // warps (nodes in flight) per block
constexpr int kAccChunk = 64;    // addends per shared-memory refill

__global__ void __launch_bounds__(kAccWarps * 32) project_accumulate_kernel(const uint32_t* __restrict__ keys, const double* __restrict__ vals,
                                                                            uint32_t n, uint32_t n_nodes,
                                                                            double* __restrict__ kmer_freq) {
    __shared__ double2 park[kAccWarps][kAccChunk / 2];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, total_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t node = gwarp; node < n_nodes; node += total_warps) {
        uint32_t lo = 0, hi = n;                              // lower_bound(node + (lane & 1))
        const uint32_t want = node + (lane & 1u);
        while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
        const uint32_t start = __shfl_sync(0xffffffffu, lo, 0), end = __shfl_sync(0xffffffffu, lo, 1);
        if (start >= end) continue;
        double acc = kmer_freq[node];
        auto fetch = [&](uint32_t base) {                     // this lane's two addends of the chunk starting at `base` (0.0 past the end: never added)
            const uint32_t i = base + 2 * lane;
            double2 v;
            v.x = i < end ? vals[i] : 0.0;
            v.y = i + 1 < end ? vals[i + 1] : 0.0;
            return v;
        };
        double2 mine = fetch(start);
        for (uint32_t base = start; base < end; base += kAccChunk) {
            __syncwarp();                                     // the previous chunk has been consumed by every lane
            park[wib][lane] = mine;
            __syncwarp();
            mine = fetch(base + kAccChunk);                   // in flight while this chunk is added
            const uint32_t cnt = end - base < static_cast<uint32_t>(kAccChunk) ? end - base : static_cast<uint32_t>(kAccChunk);
            if (cnt == kAccChunk) {
#pragma unroll
                for (uint32_t j = 0; j < kAccChunk / 2; j++) {
                    const double2 v = park[wib][j];
                    acc = __dadd_rn(acc, v.x);                // node.go:25-28, in read order
                    acc = __dadd_rn(acc, v.y);
                }
            } else {
                for (uint32_t j = 0; j < cnt; j++) acc = __dadd_rn(acc, reinterpret_cast<const double*>(park[wib])[j]);
            }
        }
        if (lane == 0) kmer_freq[node] = acc;
    }
}

}  // namespace groot
