// K3: hierarchical exact alignment of seeded reads against graph traversals.
//
// Replaces, per (read, graph) pair:
//   the graphMinion loop                         (src/pipeline/graphminion.go:46-102)
//   GrootGraph.AlignRead (4-stage hierarchy)     (src/graph/alignment.go:13-159)
//   performAlignment / dfsRecursive / processTraversal (alignment.go:162-193, 196-254, 263-317)
// It is an exact-match DFS (reference 'N' is a wildcard, a read may overhang a sink node), not an
// edit-distance DP: the reference has no scoring at all.
//
// The reference walks an ordered list of candidate starts per mapping and strand ("tries"): offsets
// OffSet..OffSet+MergeSpan+WindowSize on the seed node, then offsets 0..10 on every contained node, then a
// 1-base start clip and a 1-base end clip; forward strand first, then the reverse complement; mappings in
// (Node, OffSet) order. It stops at the first try whose DFS yields at least one path id. Three kernels:
//
//   align_screen_kernel  ONE WARP PER PAIR. The 32 lanes evaluate 32 consecutive tries of that list at once
//                        with a DFS bounded to the first kScreenBases read bases (a necessary condition for
//                        the full match); a ballot picks the lowest-numbered survivor == the first try the
//                        sequential loop could possibly stop at. Almost every try dies on its first base.
//   align_verify_kernel  ONE THREAD PER PAIR. Runs the full DFS on that try (32 independent walks per warp
//                        instead of one lane walking while 31 wait); if it fails — low-complexity sequence —
//                        the thread simply continues the reference's sequential enumeration from there.
//                        Path membership is a bitset per node, so "a path id is assigned iff it occurs in
//                        every node of the traversal" (alignment.go:301-307) is an AND over the DFS stack.
//   align_emit_kernel    ONE THREAD PER PAIR, after an exclusive scan of the record counts: expands the
//                        traversal's path bitset into (path, pos) records at the pair's exact offset.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.cuh"

namespace groot {

constexpr int kAlignWarps = 8;      // warps per block in the screen kernel
constexpr int kScreenBases = 16;    // read bases a try must match to survive the screen
constexpr int kMaskWordsInline = 8; // path bitsets of up to 256 paths travel from verify to emit without a second DFS

struct DfsFrame {
    uint32_t node;
    uint16_t edge_i;
    uint16_t dist;  // read bases consumed after this node
};

struct PairOut {  // == grootgpu_pair (include/grootgpu.h); kept in sync by a static_assert in capi.cu
    uint32_t read, graph, hit_begin, hit_count, n_incremented, rec_begin, rec_count;
    uint8_t reverse, clip_start, clip_end, stage;
};

__device__ __forceinline__ uint8_t complement_base(uint8_t b) {  // src/seqio/seqio.go:17-23
    switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'N': return 'N'; default: return 0; }
}
// read views
struct SmemRead {
    const uint8_t* p;
    __device__ __forceinline__ uint8_t operator()(uint32_t i) const { return p[i]; }
};
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

// ---- the ordered list of tries of one (mapping, strand) ------------------------------------------
__device__ __forceinline__ uint32_t tries_per_strand(const WinRec& wr) {
    return (wr.merge_span + wr.win_size + 1u) + wr.cn_cnt * 11u + 2u;
}
// try t -> start node, start offset, hierarchy stage (1..4)
__device__ __forceinline__ void decode_try(const DevIndex& ix, const WinRec& wr, uint32_t t, uint32_t* node, uint32_t* off, uint32_t* stage) {
    const uint32_t t1 = wr.merge_span + wr.win_size + 1u;          // alignment.go:35-45
    if (t < t1) { *node = wr.node; *off = wr.offset + t; *stage = 1; return; }
    t -= t1;
    const uint32_t t2 = wr.cn_cnt * 11u;                           // alignment.go:48-70 (ContainedNodes ascending SegmentID)
    if (t < t2) { *node = ix.cn_node[wr.cn_off + t / 11u]; *off = t % 11u; *stage = 2; return; }
    t -= t2;
    *node = wr.node; *off = wr.offset; *stage = 3 + t;             // alignment.go:73-85 (start clip), 88-103 (end clip)
}

enum { DFS_EXISTS = 0, DFS_COUNT = 1, DFS_EMIT = 2 };

struct DfsResult {
    uint32_t nrec;   // records (path ids over all successful traversals)
    uint32_t ntrav;  // successful traversals with at least one path id
    uint32_t mask[kMaskWordsInline];  // path bitset of the last such traversal (valid when mw <= kMaskWordsInline)
};

// dfsRecursive + processTraversal with an explicit stack. Traversals are visited in the reference's DFS
// order (out-edges in descending SegmentID order); the records of one traversal are its path ids ascending.
//   DFS_EXISTS: true as soon as any traversal completes (no path bookkeeping) — used by the bounded screen
//   DFS_COUNT : counts records / traversals and keeps the path bitset
//   DFS_EMIT  : additionally writes (path, pos) records
template <int MODE, class RD>
__device__ uint32_t dfs_align(const DevIndex& ix, uint32_t node0, uint32_t off0, RD rd, uint32_t rlen, uint32_t mw,
                              DfsFrame* __restrict__ stack, uint32_t max_depth, DfsResult* res, uint32_t* out_path, int32_t* out_pos) {
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
                if (MODE == DFS_EXISTS) return 1;
                uint32_t trav_recs = 0;
                uint32_t tm[kMaskWordsInline];
                const NodeRec n0 = ix.nodes[node0];
                uint32_t j = 0;
                auto word = [&](uint32_t wi) {
                    uint32_t m = 0xffffffffu;
                    for (uint32_t d = 0; d < depth && m; d++) m &= ix.node_mask[ix.nodes[stack[d].node].mask_off + wi];
                    trav_recs += __popc(m);
                    if (MODE == DFS_EMIT) {
                        while (m) {
                            const uint32_t pid = wi * 32 + (__ffs(m) - 1);
                            m &= m - 1;
                            while (j < n0.path_cnt && ix.node_path_id[n0.path_off + j] < pid) j++;   // both ascending
                            const int32_t pos = (j < n0.path_cnt && ix.node_path_id[n0.path_off + j] == pid) ? ix.node_path_pos[n0.path_off + j] : 0;
                            out_path[nrec] = pid;
                            out_pos[nrec] = pos + static_cast<int32_t>(off0);  // alignment.go:296
                            nrec++;
                        }
                        return 0u;
                    }
                    return m;
                };
#pragma unroll
                for (uint32_t wi = 0; wi < kMaskWordsInline; wi++) tm[wi] = wi < mw ? word(wi) : 0u;
                for (uint32_t wi = kMaskWordsInline; wi < mw; wi++) word(wi);
                if (MODE == DFS_COUNT && trav_recs) {
#pragma unroll
                    for (uint32_t wi = 0; wi < kMaskWordsInline; wi++) res->mask[wi] = tm[wi];
                }
                if (MODE != DFS_EMIT) nrec += trav_recs;
                if (trav_recs) res->ntrav++;
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
    if (MODE != DFS_EXISTS) res->nrec = nrec;
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
    uint2* seg_cand;               // [n_segs] screen result: x = mapping index inside the pair (0xffffffff = none), y = strand<<31 | try
    PairOut* pairs;                // [n_segs]
    uint32_t* seg_nrec;            // [n_segs]
    uint2* seg_locus;              // [n_segs] (node, offset) of the successful start
    uint32_t* seg_mask;            // [n_segs * kMaskWordsInline] path bitset when exactly one traversal carried ids
    uint32_t* seg_ntrav;           // [n_segs]
    DfsFrame* stack_ws;            // [threads * (max_len + 2)]
    uint32_t max_len;
    int no_align;
    int* error;
    unsigned long long* counters;  // [3] += pairs that needed the sequential continuation (diagnostic)
};

// ---- screen: one warp per pair, 32 tries at a time ------------------------------------------------
__global__ void __launch_bounds__(kAlignWarps * 32) align_screen_kernel(DevIndex ix, AlignArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t stride = (a.max_len + 16) & ~15u;
    uint8_t* fwd = smem_raw + static_cast<size_t>(warp) * 2 * stride;
    uint8_t* rcb = fwd + stride;
    const uint32_t n_segs = *a.n_segs_ptr, n_hits = *a.n_hits_ptr;
    const uint32_t gwarp = blockIdx.x * kAlignWarps + warp, total_warps = gridDim.x * kAlignWarps;
    DfsFrame stack[kScreenBases + 2];

    for (uint32_t s = gwarp; s < n_segs; s += total_warps) {
        // segments tile hits[]: a segment ends where the next one starts
        const uint32_t hb = a.seg_begin[s];
        const uint32_t he = (s + 1 < n_segs) ? a.seg_begin[s + 1] : n_hits;
        uint2 cand = make_uint2(0xffffffffu, 0u);
        if (!a.no_align) {                                     // graphminion.go:70-72
            const uint32_t r = a.hit_read[hb];
            const uint32_t o = a.off[r], len = a.off[r + 1] - o;
            __syncwarp();
            for (uint32_t i = lane; i < len; i += 32) fwd[i] = a.seq[o + i];
            __syncwarp();
            bool rc_ready = false, found = false;
            for (uint32_t m = hb; m < he && !found; m++) {
                const WinRec wr = ix.wins[a.hits[m]];
                const uint32_t T = tries_per_strand(wr);
                for (uint32_t strand = 0; strand < 2 && !found; strand++) {
                    if (strand == 1 && !rc_ready) {            // graphminion.go:94 RevComplement (forward found nothing)
                        bool bad = false;
                        for (uint32_t i = lane; i < len; i += 32) {
                            uint8_t b = fwd[len - 1 - i];
                            if (b > 'T') bad = true;           // Go: index out of range on complementBases
                            rcb[i] = complement_base(b);
                        }
                        if (__any_sync(0xffffffffu, bad) && lane == 0) { if (atomicCAS(a.error, 0, -6) == 0) a.error[1] = static_cast<int>(r); }
                        __syncwarp();
                        rc_ready = true;
                    }
                    const uint8_t* rd = strand ? rcb : fwd;
                    for (uint32_t base = 0; base < T && !found; base += 32) {
                        const uint32_t t = base + lane;
                        bool ok = false;
                        if (t < T) {
                            uint32_t node, off0, stage;
                            decode_try(ix, wr, t, &node, &off0, &stage);
                            const uint32_t view_len = stage >= 3 ? len - 1 : len;
                            const uint32_t pre = view_len < kScreenBases ? view_len : kScreenBases;
                            ok = dfs_align<DFS_EXISTS>(ix, node, off0, SmemRead{rd + (stage == 3 ? 1 : 0)}, pre, 0, stack, kScreenBases + 2, nullptr, nullptr, nullptr) != 0;
                        }
                        const uint32_t ball = __ballot_sync(0xffffffffu, ok);
                        if (ball) {
                            cand = make_uint2(m - hb, (strand << 31) | (base + (__ffs(ball) - 1)));
                            found = true;
                        }
                    }
                }
            }
        }
        if (lane == 0) a.seg_cand[s] = cand;
    }
}

// ---- verify: one thread per pair -------------------------------------------------------------------
__global__ void __launch_bounds__(128) align_verify_kernel(DevIndex ix, AlignArgs a) {
    const uint32_t n_segs = *a.n_segs_ptr, n_hits = *a.n_hits_ptr;
    const uint32_t gthread = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x;
    const uint32_t depth_cap = a.max_len + 2;
    DfsFrame* stack = a.stack_ws + static_cast<size_t>(gthread) * depth_cap;
    unsigned slow = 0;
    for (uint32_t s = gthread; s < n_segs; s += total) {
        const uint32_t hb = a.seg_begin[s];
        const uint32_t he = (s + 1 < n_segs) ? a.seg_begin[s + 1] : n_hits;
        const uint32_t r = a.hit_read[hb];
        const uint32_t o = a.off[r], len = a.off[r + 1] - o;
        const uint32_t graph = ix.wins[a.hits[hb]].graph;
        const uint32_t mw = ix.graph_mask_words[graph];
        const uint2 cand = a.seg_cand[s];
        PairOut p;
        p.read = r; p.graph = graph; p.hit_begin = hb; p.hit_count = he - hb;
        p.n_incremented = he - hb;                         // every mapping is weighted when none aligns (graphminion.go:64-98)
        p.rec_begin = 0; p.rec_count = 0; p.reverse = 0; p.clip_start = 0; p.clip_end = 0; p.stage = 0;
        DfsResult res;
        res.nrec = 0; res.ntrav = 0;
        uint32_t lnode = 0, loff = 0;
        if (cand.x != 0xffffffffu) {
            uint32_t m = hb + cand.x, strand = cand.y >> 31, t = cand.y & 0x7fffffffu;
            bool first = true;
            while (m < he) {
                const WinRec wr = ix.wins[a.hits[m]];
                const uint32_t T = tries_per_strand(wr);
                bool done = false;
                for (; strand < 2 && !done; strand++, t = 0) {
                    if (strand == 1 && !first && t == 0) {     // sequential continuation reached RevComplement on its own
                        for (uint32_t i = 0; i < len; i++) if (a.seq[o + i] > 'T') { if (atomicCAS(a.error, 0, -6) == 0) a.error[1] = static_cast<int>(r); break; }
                    }
                    for (; t < T; t++) {
                        uint32_t node, off0, stage;
                        decode_try(ix, wr, t, &node, &off0, &stage);
                        GlobalRead rd{a.seq + o, len, stage == 3 ? 1u : 0u, strand != 0};
                        const uint32_t rlen = stage >= 3 ? len - 1 : len;
                        res.nrec = 0; res.ntrav = 0;
                        dfs_align<DFS_COUNT>(ix, node, off0, rd, rlen, mw, stack, depth_cap, &res, nullptr, nullptr);
                        if (res.nrec > 0) {
                            p.n_incremented = m - hb + 1; p.rec_count = res.nrec; p.reverse = static_cast<uint8_t>(strand);
                            p.clip_start = stage == 3; p.clip_end = stage == 4; p.stage = static_cast<uint8_t>(stage);
                            lnode = node; loff = off0;
                            done = true;
                            break;
                        }
                        if (first) { slow++; first = false; }
                    }
                    if (done) break;
                }
                if (done) break;
                m++; strand = 0; t = 0;
            }
        }
        a.pairs[s] = p;
        a.seg_nrec[s] = p.rec_count;
        a.seg_locus[s] = make_uint2(lnode, loff);
        a.seg_ntrav[s] = res.nrec > 0 ? res.ntrav : 0;
        if (res.nrec > 0 && res.ntrav == 1 && mw <= kMaskWordsInline)
            for (uint32_t wi = 0; wi < mw; wi++) a.seg_mask[static_cast<size_t>(s) * kMaskWordsInline + wi] = res.mask[wi];
    }
    slow = __reduce_add_sync(0xffffffffu, slow);
    if ((threadIdx.x & 31) == 0 && slow) atomicAdd(&a.counters[3], static_cast<unsigned long long>(slow));
}

struct EmitArgs {
    const uint8_t* seq;
    const uint32_t* off;
    const uint32_t* n_segs_ptr;
    PairOut* pairs;
    const uint32_t* rec_off;   // exclusive scan of seg_nrec
    const uint2* seg_locus;
    const uint32_t* seg_mask;
    const uint32_t* seg_ntrav;
    uint32_t* rec_path;
    int32_t* rec_pos;
    DfsFrame* stack_ws;
    uint32_t max_len;
};

// One thread per pair: write the pair's records at the scanned offset. The common case (exactly one
// traversal with ids, <= 256 paths in the graph) expands the stored bitset; otherwise the DFS is re-run.
__global__ void __launch_bounds__(128) align_emit_kernel(DevIndex ix, EmitArgs a) {
    const uint32_t n_segs = *a.n_segs_ptr;
    const uint32_t gthread = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x;
    const uint32_t depth_cap = a.max_len + 2;
    DfsFrame* stack = a.stack_ws + static_cast<size_t>(gthread) * depth_cap;
    for (uint32_t s = gthread; s < n_segs; s += total) {
        PairOut p = a.pairs[s];
        const uint32_t rb = a.rec_off[s];
        a.pairs[s].rec_begin = rb;
        if (p.rec_count == 0) continue;
        const uint2 loc = a.seg_locus[s];
        const uint32_t mw = ix.graph_mask_words[p.graph];
        if (a.seg_ntrav[s] == 1 && mw <= kMaskWordsInline) {
            const NodeRec n0 = ix.nodes[loc.x];
            uint32_t j = 0, n = 0;
            for (uint32_t wi = 0; wi < mw; wi++) {
                uint32_t m = a.seg_mask[static_cast<size_t>(s) * kMaskWordsInline + wi];
                while (m) {
                    const uint32_t pid = wi * 32 + (__ffs(m) - 1);
                    m &= m - 1;
                    while (j < n0.path_cnt && ix.node_path_id[n0.path_off + j] < pid) j++;
                    const int32_t pos = (j < n0.path_cnt && ix.node_path_id[n0.path_off + j] == pid) ? ix.node_path_pos[n0.path_off + j] : 0;
                    a.rec_path[rb + n] = pid;
                    a.rec_pos[rb + n] = pos + static_cast<int32_t>(loc.y);
                    n++;
                }
            }
        } else {
            const uint32_t o = a.off[p.read], len = a.off[p.read + 1] - o;
            GlobalRead rd{a.seq + o, len, p.clip_start ? 1u : 0u, p.reverse != 0};
            const uint32_t rlen = len - p.clip_start - p.clip_end;
            DfsResult res;
            res.nrec = 0; res.ntrav = 0;
            dfs_align<DFS_EMIT>(ix, loc.x, loc.y, rd, rlen, mw, stack, depth_cap, &res, a.rec_path + rb, a.rec_pos + rb);
        }
    }
}

}  // namespace groot
