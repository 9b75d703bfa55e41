// K3: hierarchical exact alignment of seeded reads against graph traversals.
//
// Replaces, per (read, graph) pair:
//   the graphMinion loop                         (src/pipeline/graphminion.go:46-102)
//   GrootGraph.AlignRead (4-stage hierarchy)     (src/graph/alignment.go:13-159)
//   performAlignment / dfsRecursive / processTraversal (alignment.go:162-193, 196-254, 263-317)
// It is an exact-match DFS (reference 'N' is a wildcard, a read may overhang a sink node), not an
// edit-distance DP: the reference has no scoring at all.
//
// Mapping: ONE WARP PER (read, graph) PAIR. The reference tries start positions one after the other
// (up to MergeSpan+WindowSize+1 offsets on the seed node, then 11 offsets on every contained node,
// then two 1-base hard clips, all of that again on the reverse complement); here the 32 lanes try 32
// consecutive candidates of that list at once and a ballot picks the lowest-numbered success, which
// is exactly the candidate the sequential loop would have stopped at. Each lane runs the DFS with an
// explicit stack (global workspace, touched only by the rare deep traversal). Path membership is a
// bitset per node, so "a path id is assigned iff it occurs in every node of the traversal"
// (alignment.go:301-307) is an AND over the stack.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.cuh"

namespace groot {

constexpr int kAlignWarps = 8;  // warps per block

struct DfsFrame {
    uint32_t node;
    uint16_t edge_i;
    uint16_t dist;  // read bases consumed after this node
};

struct PairOut {  // == grootgpu_pair (include/grootgpu.h); kept in sync by a static_assert in capi.cu
    uint32_t read, graph, hit_begin, hit_count, n_incremented, rec_begin, rec_count;
    uint8_t reverse, clip_start, clip_end, stage;
};

// read views
struct SmemRead {
    const uint8_t* p;
    __device__ __forceinline__ uint8_t operator()(uint32_t i) const { return p[i]; }
};
__device__ __forceinline__ uint8_t complement_base(uint8_t b) {  // src/seqio/seqio.go:17-23
    switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'N': return 'N'; default: return 0; }
}
struct GlobalRead {  // forward or reverse-complement view of a read in global memory, optional 1-base start clip
    const uint8_t* p;
    uint32_t len;    // full read length
    uint32_t shift;  // 1 for the start-clipped view
    bool rc;
    __device__ __forceinline__ uint8_t operator()(uint32_t i) const {
        i += shift;
        return rc ? complement_base(p[len - 1 - i]) : p[i];
    }
};

enum { DFS_COUNT = 0, DFS_EMIT = 1 };

// dfsRecursive + processTraversal with an explicit stack. Returns the number of (path, pos) records
// the reference would emit for a start at (node0, off0): for every successful traversal in DFS order
// (out-edges in the reference's descending-SegmentID order), every path id present in all its nodes,
// ascending. In DFS_EMIT mode the records are also written.
template <int MODE, class RD>
__device__ uint32_t dfs_align(const DevIndex& ix, uint32_t node0, uint32_t off0, RD rd, uint32_t rlen, uint32_t mw,
                              DfsFrame* __restrict__ stack, uint32_t max_depth, uint32_t* out_path, int32_t* out_pos) {
    uint32_t nrec = 0, depth = 0;
    uint32_t cur = node0, off = off0, dist = 0;
    while (true) {
        const NodeRec nd = ix.nodes[cur];
        bool ok = off < nd.seq_len;  // alignment.go:199-201
        if (ok) {
            const uint8_t* s = ix.node_seq + nd.seq_off;
            for (uint32_t i = off; i < nd.seq_len; i++) {
                if (dist == rlen) break;                  // alignment.go:207-209
                const uint8_t b = s[i];
                if (b == 'N') { dist++; continue; }       // alignment.go:212-215
                if (b == rd(dist)) dist++;
                else { ok = false; break; }               // alignment.go:220-222
            }
        }
        if (ok && depth < max_depth) {
            stack[depth].node = cur; stack[depth].edge_i = 0; stack[depth].dist = static_cast<uint16_t>(dist);
            depth++;
            if (dist == rlen || nd.edge_cnt == 0) {       // alignment.go:229: full read matched OR sink node
                for (uint32_t wi = 0; wi < mw; wi++) {
                    uint32_t m = 0xffffffffu;
                    for (uint32_t d = 0; d < depth && m; d++) m &= ix.node_mask[ix.nodes[stack[d].node].mask_off + wi];
                    if (MODE == DFS_COUNT) {
                        nrec += __popc(m);
                    } else {
                        const NodeRec n0 = ix.nodes[node0];
                        while (m) {
                            const uint32_t pid = wi * 32 + (__ffs(m) - 1);
                            m &= m - 1;
                            int32_t pos = 0;
                            for (uint32_t j = 0; j < n0.path_cnt; j++)
                                if (ix.node_path_id[n0.path_off + j] == pid) { pos = ix.node_path_pos[n0.path_off + j]; break; }
                            out_path[nrec] = pid;
                            out_pos[nrec] = pos + static_cast<int32_t>(off0);  // alignment.go:296
                            nrec++;
                        }
                    }
                }
                depth--;
            }
        }
        bool advanced = false;
        while (depth > 0) {
            DfsFrame& top = stack[depth - 1];
            const NodeRec tn = ix.nodes[top.node];
            if (top.edge_i < tn.edge_cnt) {
                cur = ix.edges[tn.edge_off + top.edge_i];
                top.edge_i++;
                off = 0; dist = top.dist;
                advanced = true;
                break;
            }
            depth--;
        }
        if (!advanced) break;
    }
    return nrec;
}

struct AlignArgs {
    const uint8_t* seq;
    const uint32_t* off;
    const uint32_t* hits;
    const uint32_t* hit_read;
    const uint32_t* seg_begin;     // [n_segs] index into hits of each (read, graph) segment start
    const uint32_t* n_segs_ptr;    // device scalar
    const uint32_t* n_hits_ptr;    // device scalar (total hits)
    PairOut* pairs;                // [n_segs]
    uint32_t* seg_nrec;            // [n_segs]
    uint2* seg_locus;              // [n_segs] (node, offset) of the successful start
    DfsFrame* stack_ws;            // [threads * (max_len + 2)]
    uint32_t max_len;
    int no_align;
    int* error;
};

// One warp per pair; lanes try 32 candidate starts at a time.
__global__ void __launch_bounds__(kAlignWarps * 32) align_search_kernel(DevIndex ix, AlignArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t stride = (a.max_len + 16) & ~15u;
    uint8_t* fwd = smem_raw + static_cast<size_t>(warp) * 2 * stride;
    uint8_t* rcb = fwd + stride;
    const uint32_t n_segs = *a.n_segs_ptr, n_hits = *a.n_hits_ptr;
    const uint32_t gwarp = blockIdx.x * kAlignWarps + warp, total_warps = gridDim.x * kAlignWarps;
    const uint32_t depth_cap = a.max_len + 2;
    DfsFrame* stack = a.stack_ws + (static_cast<size_t>(gwarp) * 32 + lane) * depth_cap;

    for (uint32_t s = gwarp; s < n_segs; s += total_warps) {
        const uint32_t hb = a.seg_begin[s];
        const uint32_t he = (s + 1 < n_segs) ? a.seg_begin[s + 1] : n_hits;
        const uint32_t r = a.hit_read[hb];
        // segments tile hits[]: the segment ends where the next one starts (same read => next graph)
        const uint32_t o = a.off[r], len = a.off[r + 1] - o;
        __syncwarp();
        for (uint32_t i = lane; i < len; i += 32) fwd[i] = a.seq[o + i];
        __syncwarp();
        bool rc_ready = false;
        const uint32_t graph = ix.wins[a.hits[hb]].graph;
        const uint32_t mw = ix.graph_mask_words[graph];
        uint32_t ninc = 0, nrec = 0, lnode = 0, loff = 0;
        uint32_t found_stage = 0, reverse = 0;
        for (uint32_t m = hb; m < he && !found_stage; m++) {
            ninc++;                                            // graphminion.go:67 IncrementSubPath (replayed on the host)
            if (a.no_align) continue;                          // graphminion.go:70-72
            const WinRec wr = ix.wins[a.hits[m]];
            for (uint32_t strand = 0; strand < 2 && !found_stage; strand++) {
                if (strand == 1 && !rc_ready) {                // graphminion.go:94 RevComplement
                    bool bad = false;
                    for (uint32_t i = lane; i < len; i += 32) {
                        uint8_t b = fwd[len - 1 - i];
                        if (b > 'T') bad = true;               // Go: index out of range on complementBases
                        rcb[i] = complement_base(b);
                    }
                    if (__any_sync(0xffffffffu, bad)) { if (lane == 0) { if (atomicCAS(a.error, 0, -6) == 0) a.error[1] = static_cast<int>(r); } }
                    __syncwarp();
                    rc_ready = true;
                }
                const uint8_t* rd = strand ? rcb : fwd;
                // stage 1: seed offset shuffling (alignment.go:35-45)
                const uint32_t t1 = wr.merge_span + wr.win_size + 1;
                for (uint32_t base = 0; base < t1 && !found_stage; base += 32) {
                    const uint32_t t = base + lane;
                    uint32_t cnt = 0;
                    if (t < t1) cnt = dfs_align<DFS_COUNT>(ix, wr.node, wr.offset + t, SmemRead{rd}, len, mw, stack, depth_cap, nullptr, nullptr);
                    const uint32_t ball = __ballot_sync(0xffffffffu, cnt > 0);
                    if (ball) {
                        const int wl = __ffs(ball) - 1;
                        nrec = __shfl_sync(0xffffffffu, cnt, wl);
                        lnode = wr.node; loff = wr.offset + base + wl; found_stage = 1;
                    }
                }
                // stage 2: seed node shuffling over ContainedNodes x offsets 0..10 (alignment.go:48-70)
                const uint32_t t2 = wr.cn_cnt * 11u;
                for (uint32_t base = 0; base < t2 && !found_stage; base += 32) {
                    const uint32_t t = base + lane;
                    uint32_t cnt = 0, node = 0, sh = 0;
                    if (t < t2) {
                        node = ix.cn_node[wr.cn_off + t / 11u]; sh = t % 11u;
                        cnt = dfs_align<DFS_COUNT>(ix, node, sh, SmemRead{rd}, len, mw, stack, depth_cap, nullptr, nullptr);
                    }
                    const uint32_t ball = __ballot_sync(0xffffffffu, cnt > 0);
                    if (ball) {
                        const int wl = __ffs(ball) - 1;
                        nrec = __shfl_sync(0xffffffffu, cnt, wl);
                        lnode = __shfl_sync(0xffffffffu, node, wl); loff = __shfl_sync(0xffffffffu, sh, wl); found_stage = 2;
                    }
                }
                // stages 3 and 4: 1-base hard clip of the start, then of the end (alignment.go:73-103)
                if (!found_stage) {
                    uint32_t cnt = 0;
                    if (lane == 0 && len >= 1) cnt = dfs_align<DFS_COUNT>(ix, wr.node, wr.offset, SmemRead{rd + 1}, len - 1, mw, stack, depth_cap, nullptr, nullptr);
                    if (lane == 1 && len >= 1) cnt = dfs_align<DFS_COUNT>(ix, wr.node, wr.offset, SmemRead{rd}, len - 1, mw, stack, depth_cap, nullptr, nullptr);
                    const uint32_t ball = __ballot_sync(0xffffffffu, cnt > 0);
                    if (ball) {
                        const int wl = __ffs(ball) - 1;
                        nrec = __shfl_sync(0xffffffffu, cnt, wl);
                        lnode = wr.node; loff = wr.offset; found_stage = 3 + wl;
                    }
                }
                if (found_stage) reverse = strand;
            }
        }
        if (lane == 0) {
            PairOut p;
            p.read = r; p.graph = graph; p.hit_begin = hb; p.hit_count = he - hb; p.n_incremented = ninc;
            p.rec_begin = 0; p.rec_count = nrec;
            p.reverse = static_cast<uint8_t>(reverse); p.clip_start = found_stage == 3; p.clip_end = found_stage == 4;
            p.stage = static_cast<uint8_t>(found_stage);
            a.pairs[s] = p;
            a.seg_nrec[s] = nrec;
            a.seg_locus[s] = make_uint2(lnode, loff);
        }
    }
}

struct EmitArgs {
    const uint8_t* seq;
    const uint32_t* off;
    const uint32_t* n_segs_ptr;
    PairOut* pairs;
    const uint32_t* rec_off;   // exclusive scan of seg_nrec
    const uint2* seg_locus;
    uint32_t* rec_path;
    int32_t* rec_pos;
    DfsFrame* stack_ws;
    uint32_t max_len;
    unsigned long long* counters;  // [2] += records
};

// One thread per pair: re-walk the single successful start and write its records at the scanned offset.
__global__ void __launch_bounds__(128) align_emit_kernel(DevIndex ix, EmitArgs a) {
    const uint32_t n_segs = *a.n_segs_ptr;
    const uint32_t gthread = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x;
    const uint32_t depth_cap = a.max_len + 2;
    DfsFrame* stack = a.stack_ws + static_cast<size_t>(gthread) * depth_cap;
    unsigned long long recs = 0;
    for (uint32_t s = gthread; s < n_segs; s += total) {
        PairOut p = a.pairs[s];
        const uint32_t rb = a.rec_off[s];
        a.pairs[s].rec_begin = rb;
        if (p.rec_count == 0) continue;
        const uint32_t o = a.off[p.read], len = a.off[p.read + 1] - o;
        GlobalRead rd{a.seq + o, len, p.clip_start ? 1u : 0u, p.reverse != 0};
        const uint32_t rlen = len - p.clip_start - p.clip_end;
        const uint2 loc = a.seg_locus[s];
        const uint32_t mw = ix.graph_mask_words[p.graph];
        dfs_align<DFS_EMIT>(ix, loc.x, loc.y, rd, rlen, mw, stack, depth_cap, a.rec_path + rb, a.rec_pos + rb);
        recs += p.rec_count;
    }
    recs = __reduce_add_sync(0xffffffffu, static_cast<unsigned>(recs)) ;
    if ((threadIdx.x & 31) == 0 && recs) atomicAdd(&a.counters[2], recs);
}

}  // namespace groot
