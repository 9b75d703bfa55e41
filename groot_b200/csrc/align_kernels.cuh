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
// (Node, OffSet) order. It stops at the first try whose DFS yields at least one path id.
//
//   align_init_kernel    one thread per pair: default (unaligned) result, cursor at the first try; the host radix-sorts
//                        the pair queue by window id, so that the pairs a warp handles sit on the same graph region.
//   align_screen_kernel  ONE WARP PER QUEUED PAIR. From the pair's cursor the lanes test the tries that can exist at
//                        all (offsets beyond a node are never enumerated; the offsets of up to 32 contained nodes are
//                        pooled) against the allele sets of host/prefix_table.cpp: each of the read's first 8 bases
//                        must lie in the set of its step — a necessary condition for dfsRecursive to succeed there,
//                        one load and an AND per try, no DFS; a ballot picks the lowest-numbered survivor == the next
//                        try the sequential loop could stop at.
//   align_walk_kernel    ONE THREAD PER QUEUED PAIR: the DFS on 2-bit data (dfs_packed: 16 bases per XOR, path bitset
//                        as a running AND, a stack of branch nodes only, the bitsets of the traversals kept for the
//                        emit). A pair whose walk yields no path id is re-queued with its cursor just past that try;
//                        a pair that cannot take the packed walk goes to the slow queue. The host runs two screen/walk
//                        rounds over the shrinking queue (no host sync: counts live on the device), then
//   align_finish_kernel  one warp per leftover / slow pair alternates screen steps and walks (byte-wise walk
//                        available: dfs_masked, dfs_align) to the end of the reference's sequential enumeration.
//                        (Earlier layouts — a warp per pair doing everything, one thread per pair doing everything,
//                        in-kernel warp-synchronous rounds, byte-wise walks with a frame per node — lost 5-10x to
//                        lanes idling on each other; see profiles/r01_notes.md.)
//   align_emit_kernel    8 lanes per pair, after an exclusive scan of the record counts: expands the kept traversal
//                        bitsets into (path, pos) records at the pair's exact offset; align_emit_classify_kernel +
//                        align_emit_multi_kernel walk the few pairs again whose traversals did not fit.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.cuh"

namespace groot {

constexpr int kMaskWordsInline = 8; // path bitsets of up to 256 paths travel from verify to emit without a second DFS
constexpr int kTravWords = 24;       // words of seg_mask per pair: the path bitsets of the pair's first traversals (24 / mw of them)
constexpr uint32_t kTravRewalk = 0x80000000u;  // seg_ntrav flag: the traversals' bitsets did not fit seg_mask, the emit walks again

struct DfsFrame {
    uint32_t node;
    uint16_t edge_i;
    uint16_t dist;  // read bases consumed after this node
};

struct PairOut {  // == grootgpu_pair (include/grootgpu.h); kept in sync by a static_assert in capi.cu
    uint32_t read, graph, hit_begin, hit_count, n_incremented, rec_begin, rec_count;
    uint8_t reverse, clip_start, clip_end, stage;
};

// == grootgpu_cpair (include/grootgpu.h): the compact, BAM-oriented form of a pair; kept in sync by a static_assert in capi.cu
struct CPairOut { uint32_t read, node, offset_flags, rec_count; };
constexpr uint32_t kCPairReverse = 1u << 28, kCPairClipStart = 1u << 29, kCPairClipEnd = 1u << 30, kCPairOffsetMask = (1u << 28) - 1u;

__device__ __forceinline__ uint8_t complement_base(uint8_t b) {  // src/seqio/seqio.go:17-23
    switch (b) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'N': return 'N'; default: return 0; }
}
// read views
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
// PT = element type of out_path (uint32_t, or uint8_t / uint16_t for the compact output, which carries no positions:
// out_pos == nullptr).
template <int MODE, class RD, class PT = uint32_t>
__device__ uint32_t dfs_align(const DevIndex& ix, uint32_t node0, uint32_t off0, RD rd, uint32_t rlen, uint32_t mw,
                              DfsFrame* __restrict__ stack, uint32_t max_depth, DfsResult* res, PT* out_path, int32_t* out_pos) {
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
                            out_path[nrec] = static_cast<PT>(pid);
                            if (out_pos) {
                                while (j < n0.path_cnt && ix.node_path_id[n0.path_off + j] < pid) j++;   // both ascending
                                const int32_t pos = (j < n0.path_cnt && ix.node_path_id[n0.path_off + j] == pid) ? ix.node_path_pos[n0.path_off + j] : 0;
                                out_pos[nrec] = pos + static_cast<int32_t>(off0);  // alignment.go:296
                            }
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

struct ScreenDesc;
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
    uint32_t* seg_mask;            // [n_segs * kTravWords] path bitsets of the pair's first traversals, mw words each
    uint32_t* seg_ntrav;           // [n_segs]
    DfsFrame* stack_ws;            // [threads * (max_len + 2)]
    uint32_t* mask_ws;             // [threads * (max_len + 2) * kMaskWordsInline]
    uint32_t max_len;
    int no_align;
    int* error;
    unsigned long long* counters;  // [3] += pairs that needed the sequential continuation (diagnostic)
    const uint32_t* reads2;        // 2-bit copies of the seeded reads, both orientations (seed_kernels.cuh, pack_reads_kernel)
    const uint8_t* read_ok2;       // [n_reads] 1 when reads2 holds the read
    const uint4* read_oh;          // [n_reads] one-hot 8-base prefixes of the read / its reverse complement (pack_reads_kernel)
    ScreenDesc* sdesc;             // [n_segs]
    uint32_t nw32;                 // words per orientation; 0 = no packed copies
};

// Read view of one thread: forward or reverse complement (computed on the fly through a shared LUT),
// optionally skipping the first base (start clip).
struct ReadView {
    const uint8_t* p;
    const uint8_t* lut;  // shared-memory complement table (src/seqio/seqio.go:17-23: everything else -> 0)
    uint32_t len;        // full read length
    uint32_t shift;
    bool rc;
    __device__ __forceinline__ uint8_t operator()(uint32_t i) const {
        i += shift;
        return rc ? lut[p[len - 1 - i]] : p[i];
    }
};

// DFS with the path bitset carried as a running AND (mw <= kMaskWordsInline). A branch whose bitset becomes
// empty is abandoned: it could only produce traversals without path ids, which the reference discards
// (alignment.go:301-309) and which do not count as a successful alignment (alignment.go:38,58,81,99).
// mask_ws: per-thread save area, one bitset per stack level, written only at branch nodes.
//
// FLAT control flow: one loop whose every iteration compares at most kChunk bases of the current node, so that
// the 32 walks of a warp stay in step no matter where their node boundaries fall (with a per-node inner loop a
// 50-base node in one lane stalls the 1-base nodes of the others: measured 3 active lanes per instruction).
// Node entry / exit (bitset AND, success test, push, next edge, backtrack) are per-lane events of that loop.
constexpr uint32_t kChunk = 8;

template <class RD>
__device__ void dfs_masked(const DevIndex& ix, uint32_t node0, uint32_t off0, RD rd, uint32_t rlen, uint32_t mw,
                           DfsFrame* __restrict__ stack, uint32_t* __restrict__ mask_ws, uint32_t max_depth, DfsResult* res) {
    uint32_t nrec = 0, ntrav = 0, depth = 0;
    uint32_t cur = node0, off = off0, dist = 0;
    uint32_t cm[kMaskWordsInline];
#pragma unroll
    for (int wi = 0; wi < kMaskWordsInline; wi++) cm[wi] = 0xffffffffu;
    NodeRec nd = ix.nodes[cur];
    bool active = true;
    bool ok = off < nd.seq_len;                               // alignment.go:199-201
    while (active) {
        // ---- compare up to kChunk bases of the current node ----
        bool node_done = !ok;
        if (ok) {
            const uint32_t left_node = nd.seq_len - off, left_read = rlen - dist;
            uint32_t n = left_node < left_read ? left_node : left_read;
            n = n < kChunk ? n : kChunk;
            const uint8_t* sq = ix.node_seq + nd.seq_off + off;
#pragma unroll
            for (uint32_t i = 0; i < kChunk; i++) {
                if (i < n) {
                    const uint8_t b = sq[i];
                    if (b != 'N' && b != rd(dist + i)) ok = false;   // alignment.go:212-222
                }
            }
            off += n;
            // a reference 'N' consumes a read base too, so dist advances by n on success (alignment.go:212-219)
            dist += n;
            node_done = !ok || off == nd.seq_len || dist == rlen;     // alignment.go:204-209
        }
        if (!node_done) continue;
        // ---- node finished: membership, success, descend ----
        if (ok) {
            uint32_t any = 0;
#pragma unroll
            for (int wi = 0; wi < kMaskWordsInline; wi++)
                if (wi < mw) { cm[wi] &= ix.node_mask[nd.mask_off + wi]; any |= cm[wi]; }
            ok = any != 0 && depth < max_depth;
        }
        if (ok) {
            if (dist == rlen || nd.edge_cnt == 0) {               // alignment.go:229: full read matched OR sink node
                uint32_t c = 0;
#pragma unroll
                for (int wi = 0; wi < kMaskWordsInline; wi++) if (wi < mw) { c += __popc(cm[wi]); res->mask[wi] = cm[wi]; }
                nrec += c; ntrav++;
            } else {
                stack[depth].node = cur; stack[depth].edge_i = 0; stack[depth].dist = static_cast<uint16_t>(dist);
                if (nd.edge_cnt > 1) {
#pragma unroll
                    for (int wi = 0; wi < kMaskWordsInline; wi++) if (wi < mw) mask_ws[depth * kMaskWordsInline + wi] = cm[wi];
                }
                depth++;
            }
        }
        // ---- next node: first untried edge of the deepest frame that has one ----
        active = false;
        while (depth > 0) {
            DfsFrame& top = stack[depth - 1];
            const NodeRec tn = ix.nodes[top.node];
            if (top.edge_i < tn.edge_cnt) {
                if (top.edge_i > 0) {                             // coming back to a branch node: restore its bitset
#pragma unroll
                    for (int wi = 0; wi < kMaskWordsInline; wi++) if (wi < mw) cm[wi] = mask_ws[(depth - 1) * kMaskWordsInline + wi];
                }
                cur = ix.edges[tn.edge_off + top.edge_i];
                top.edge_i++;
                off = 0; dist = top.dist;
                nd = ix.nodes[cur];
                ok = nd.seq_len > 0;
                active = true;
                break;
            }
            depth--;
        }
    }
    res->nrec = nrec; res->ntrav = ntrav;
}

// ---- packed walk ----------------------------------------------------------------------------------------
// The same DFS as dfs_masked for the common case — read of upper-case ACGT only that fits the packed copy, graph
// of <= 256 paths — on 2-bit data: 16 bases are compared with one XOR (node_seq2 against the
// oriented packed read; two aligned loads + a funnel shift each), and the stack holds BRANCH nodes only, each frame
// carrying its own cursor into edges[] (a linear chain needs no frame, and backtracking never reloads a node: in
// the byte-wise walk the unwind loop over one frame per visited node ran with 2 active lanes and was the hot spot
// of the whole path, profiles/r01_ncu_summary.md).
// Frame fields are reused: node = index into edges[] of the next untried edge, edge_i = edges left, dist as before.
__device__ __forceinline__ uint32_t extract16(const uint32_t* __restrict__ w, uint32_t pos) {
    const uint32_t wi = pos >> 4;
    return __funnelshift_r(w[wi], w[wi + 1], (pos & 15u) * 2u);
}

// EMIT: additionally writes the (path, pos) records of every successful traversal, traversals in DFS order, path ids
// ascending inside one (processTraversal, alignment.go:263-317).
template <bool EMIT, class PT = uint32_t>
__device__ __forceinline__ void dfs_packed(const DevIndex& ix, uint32_t node0, uint32_t off0, const uint32_t* __restrict__ rd2, uint32_t base0,
                                           uint32_t rlen, uint32_t mw, bool has_n, DfsFrame* __restrict__ stack, uint32_t* __restrict__ mask_ws,
                                           uint32_t max_depth, DfsResult* res, uint32_t* __restrict__ trav_masks,
                                           PT* __restrict__ out_path = nullptr, int32_t* __restrict__ out_pos = nullptr) {
    uint32_t nrec = 0, ntrav = 0, depth = 0;
    uint32_t p0_off = 0, p0_cnt = 0;
    if (EMIT) { p0_off = ix.nodes[node0].path_off; p0_cnt = ix.nodes[node0].path_cnt; }
    uint32_t cur = node0, off = off0, dist = 0;
    uint32_t cm[kMaskWordsInline];
#pragma unroll
    for (int wi = 0; wi < kMaskWordsInline; wi++) cm[wi] = 0xffffffffu;
    // one 32-byte record per node step: sequence range, path bitset (graphs of <= 128 paths) and the out-edge of a linear chain
    auto load_node = [&](uint32_t n) {
        const uint4* q = reinterpret_cast<const uint4*>(ix.wnodes + n);
        const uint4 a = __ldg(q), b = __ldg(q + 1);
        WalkNode w;
        w.seq_off = a.x; w.seq_len = a.y; w.edge = a.z; w.edge_cnt = a.w; w.mask[0] = b.x; w.mask[1] = b.y; w.mask[2] = b.z; w.mask[3] = b.w;
        return w;
    };
    const bool inline_mask = mw <= kWalkMaskWords;
    WalkNode nd = load_node(cur);
    if (off >= nd.seq_len || rlen == 0) { res->nrec = 0; res->ntrav = 0; return; }   // alignment.go:199-201
    while (true) {
        // ---- up to 16 bases of the current node ----
        const uint32_t left_node = nd.seq_len - off, left_read = rlen - dist;
        uint32_t n = left_node < left_read ? left_node : left_read;
        n = n < 16u ? n : 16u;
        const uint32_t x = extract16(ix.node_seq2, nd.seq_off + off) ^ extract16(rd2, base0 + dist);
        uint32_t mism = (x | (x >> 1)) & 0x55555555u;
        if (has_n) mism &= ~extract16(ix.node_n2, nd.seq_off + off);   // a reference 'N' matches any read base (alignment.go:212-215)
        if (n < 16u) mism &= (1u << (2u * n)) - 1u;
        if (left_node == 0u) mism = 1u;                           // empty node: dfsRecursive fails on entry
        off += n; dist += n;
        if (mism == 0u && off != nd.seq_len && dist != rlen) continue;
        // ---- node finished (or mismatch): membership, success, descend / backtrack ----
        bool descend = false;
        if (mism == 0u) {
            uint32_t any = 0;
            if (inline_mask) {
#pragma unroll
                for (int wi = 0; wi < static_cast<int>(kWalkMaskWords); wi++)
                    if (wi < mw) { cm[wi] &= nd.mask[wi]; any |= cm[wi]; }
            } else {
#pragma unroll
                for (int wi = 0; wi < kMaskWordsInline; wi++)
                    if (wi < mw) { cm[wi] &= ix.node_mask[nd.mask[0] + wi]; any |= cm[wi]; }
            }
            if (any != 0u) {
                if (dist == rlen || nd.edge_cnt == 0) {           // alignment.go:229: full read matched OR sink node
                    if (EMIT) {
                        uint32_t j = 0;                           // walks the start node's path list: both ascending
#pragma unroll
                        for (int wi = 0; wi < kMaskWordsInline; wi++) {
                            uint32_t m = wi < mw ? cm[wi] : 0u;
                            while (m) {
                                const uint32_t pid = wi * 32 + (__ffs(m) - 1);
                                m &= m - 1;
                                out_path[nrec] = static_cast<PT>(pid);
                                if (out_pos) {
                                    while (j < p0_cnt && ix.node_path_id[p0_off + j] < pid) j++;
                                    const int32_t pos = (j < p0_cnt && ix.node_path_id[p0_off + j] == pid) ? ix.node_path_pos[p0_off + j] : 0;
                                    out_pos[nrec] = pos + static_cast<int32_t>(off0);   // alignment.go:296
                                }
                                nrec++;
                            }
                        }
                    } else {
                        // the path bitsets of the first traversals are kept (as many as fit the pair's kTravWords
                        // words): the emit kernel expands them without walking again
                        const bool keep = (ntrav + 1) * mw <= static_cast<uint32_t>(kTravWords);
                        uint32_t c = 0;
#pragma unroll
                        for (int wi = 0; wi < kMaskWordsInline; wi++)
                            if (wi < mw) { c += __popc(cm[wi]); if (keep) trav_masks[ntrav * mw + wi] = cm[wi]; }
                        nrec += c;
                    }
                    ntrav++;
                } else if (nd.edge_cnt == 1) {
                    cur = nd.edge; descend = true;                // the target itself: no edges[] access on a linear chain
                } else if (depth < max_depth) {
                    stack[depth].node = nd.edge + 1; stack[depth].edge_i = static_cast<uint16_t>(nd.edge_cnt - 1);
                    stack[depth].dist = static_cast<uint16_t>(dist);
#pragma unroll
                    for (int wi = 0; wi < kMaskWordsInline; wi++) if (wi < mw) mask_ws[depth * kMaskWordsInline + wi] = cm[wi];
                    depth++;
                    cur = ix.edges[nd.edge]; descend = true;
                }
            }
        }
        if (!descend) {
            if (depth == 0) break;
            DfsFrame& top = stack[depth - 1];
            cur = ix.edges[top.node]; dist = top.dist;
#pragma unroll
            for (int wi = 0; wi < kMaskWordsInline; wi++) if (wi < mw) cm[wi] = mask_ws[(depth - 1) * kMaskWordsInline + wi];
            top.node++;
            if (--top.edge_i == 0) depth--;
        }
        nd = load_node(cur);
        off = 0;
    }
    res->nrec = nrec; res->ntrav = ntrav;
}

constexpr uint32_t kPrefixBases = 8;  // == kPfxLen of host/prefix_table.cpp
// true unless NO traversal starting at graph position `pos` can spell the read's first min(8, rlen) bases: every base
// must lie in the allele set of its step (host/prefix_table.cpp). oh = the read's bases one-hot, 4 bits per base
// (read_onehot); a read byte that is not upper-case ACGT has an empty nibble and passes (the DFS decides).
__device__ __forceinline__ bool prefix_pass(const DevIndex& ix, uint32_t pos, uint32_t oh, uint32_t rlen) {
    const uint32_t lim = rlen >= kPrefixBases ? 0xffffffffu : (1u << (4 * rlen)) - 1u;
    return (oh & ~__ldg(ix.pfxset + pos) & lim) == 0;
}
// pk = 8 bases 2 bits each (pack_base2), bad = positions that are not upper-case ACGT
__device__ __forceinline__ uint32_t read_onehot(uint32_t pk, uint32_t bad) {
    uint32_t oh = 0;
#pragma unroll
    for (uint32_t i = 0; i < kPrefixBases; i++) oh |= (((bad >> i) & 1u) ? 0u : (1u << ((pk >> (2 * i)) & 3u))) << (4 * i);
    return oh;
}

// ---- the ordered try list of one (mapping, strand) ---------------------------------------------------
// Order (graphminion.go:64-98, alignment.go:35-103): for each mapping, for strand in (forward, reverse
// complement): tries [0, t1) = stage 1 offsets OffSet+t on the seed node (t1 = MergeSpan+WindowSize+1), then 11
// per contained node (ascending SegmentID) = stage 2 offsets 0..10, then stage 3 (1-base start clip) and stage 4
// (1-base end clip).
__device__ __forceinline__ uint32_t tries_per_strand(const WinRec& wr) {
    return (wr.merge_span + wr.win_size + 1u) + wr.cn_cnt * 11u + 2u;
}
__device__ __forceinline__ void decode_try(const DevIndex& ix, const WinRec& wr, uint32_t t, uint32_t* node, uint32_t* off, uint32_t* stage) {
    const uint32_t t1 = wr.merge_span + wr.win_size + 1u;
    if (t < t1) { *node = wr.node; *off = wr.offset + t; *stage = 1; return; }
    t -= t1;
    const uint32_t t2 = wr.cn_cnt * 11u;
    if (t < t2) { *node = ix.cn_node[wr.cn_off + t / 11u]; *off = t % 11u; *stage = 2; return; }
    t -= t2;
    *node = wr.node; *off = wr.offset; *stage = 3 + t;
}

struct PairCursor { uint32_t m_strand; uint32_t t; };   // m_strand = (mapping index inside the pair) << 1 | strand
constexpr uint32_t kNoCand = 0xffffffffu;

// What the screen needs to start on a pair, gathered once by align_init_kernel (one thread per pair, latency hidden by
// sheer parallelism) into 80 contiguous bytes: the screen is one warp per pair, and more than half of its time went
// into a chain of six dependent, mostly DRAM-missing loads (queue -> pair -> read offsets / hits -> window -> seed
// node -> read prefix) before the first try was tested (profiles/r01_notes.md).
struct ScreenDesc {
    uint32_t read, len, hit_begin, hit_count;
    uint4 oh;                    // read_oh[read] when packed != 0
    WinRec w0;                   // the first mapping's window
    uint32_t sn_seq_off, sn_seq_len;   // its seed node
    uint32_t packed;             // the read has a packed copy (upper-case ACGT, fits)
    uint32_t win0;               // window id of the first mapping
};
static_assert(sizeof(ScreenDesc) == 80, "ScreenDesc is five 16-byte words");

struct RoundArgs {
    AlignArgs a;
    PairCursor* cursor;        // [n_segs] next try to examine
    uint2* cand;               // [n_segs] screen result: (m_strand, t) or (kNoCand, 0)
    const uint32_t* queue;     // pairs to process this round
    uint32_t* queue_next;      // pairs re-queued for the next round
    const uint32_t* n_queue;   // device scalar
    uint32_t* n_queue_next;    // device scalar (atomic)
    uint32_t* slow_queue;      // pairs that cannot take the packed walk: align_finish_kernel walks them byte-wise
    uint32_t* n_slow;          // device scalar (atomic)
};

__global__ void __launch_bounds__(256) align_init_kernel(DevIndex ix, AlignArgs a, PairCursor* cursor, uint32_t* queue, uint32_t* qkey, uint32_t* n_queue) {
    const uint32_t n_segs = *a.n_segs_ptr, n_hits = *a.n_hits_ptr;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += gridDim.x * blockDim.x) {
        const uint32_t hb = a.seg_begin[s];                             // segments tile hits[]
        const uint32_t he = (s + 1 < n_segs) ? a.seg_begin[s + 1] : n_hits;
        PairOut p;
        p.read = a.hit_read[hb]; p.graph = ix.wins[a.hits[hb]].graph; p.hit_begin = hb; p.hit_count = he - hb;
        p.n_incremented = he - hb;                  // every mapping is weighted when none aligns (graphminion.go:64-98)
        p.rec_begin = 0; p.rec_count = 0; p.reverse = 0; p.clip_start = 0; p.clip_end = 0; p.stage = 0;
        a.pairs[s] = p;
        a.seg_nrec[s] = 0; a.seg_ntrav[s] = 0; a.seg_locus[s] = make_uint2(0, 0);
        cursor[s] = PairCursor{0u, 0u};
        queue[s] = s;
        qkey[s] = a.hits[hb];                       // first mapping's window: the queue is sorted by it (pairs of one warp walk the same graph region)
        if (!a.no_align) {
            ScreenDesc d;
            d.read = p.read; d.len = a.off[p.read + 1] - a.off[p.read]; d.hit_begin = hb; d.hit_count = he - hb;
            d.w0 = ix.wins[a.hits[hb]];
            const NodeRec sn = ix.nodes[d.w0.node];
            d.sn_seq_off = sn.seq_off; d.sn_seq_len = sn.seq_len;
            d.packed = a.nw32 && a.read_ok2[p.read] ? 1u : 0u;
            d.oh = d.packed ? a.read_oh[p.read] : make_uint4(0, 0, 0, 0);
            d.win0 = a.hits[hb];
            a.sdesc[s] = d;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_queue = a.no_align ? 0u : n_segs;   // graphminion.go:70-72 (--noAlign)
}

// the first kPrefixBases + 1 oriented read bases, 2 bits each, as the two 8-base views the screen needs (v = 0: from
// base 0; v = 1: from base 1, for the start clip). From the packed copy when there is one, else byte-wise with the
// whole warp (lane i handles base i).
__device__ __forceinline__ void warp_read_prefix(const AlignArgs& a, uint32_t r, const uint8_t* __restrict__ rp, uint32_t len, bool rc, uint32_t lane,
                                                 uint32_t (&oh)[2]) {
    constexpr uint32_t m2 = (1u << (2 * kPrefixBases)) - 1u, m1 = (1u << kPrefixBases) - 1u;
    if (a.nw32 && a.read_ok2[r]) {
        const uint4 q = __ldg(a.read_oh + r);
        oh[0] = rc ? q.z : q.x; oh[1] = rc ? q.w : q.y;
        return;
    }
    uint32_t code = 4;
    if (lane <= kPrefixBases && lane < len) {
        const uint8_t c = rc ? complement_base(rp[len - 1 - lane]) : rp[lane];
        code = (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? pack_base2(c) : 4u;
    }
    const uint32_t b0 = __ballot_sync(0xffffffffu, code & 1u), b1 = __ballot_sync(0xffffffffu, (code >> 1) & 1u);
    const uint32_t bb = __ballot_sync(0xffffffffu, code > 3u && lane < len);
    uint32_t p = 0;
#pragma unroll
    for (uint32_t i = 0; i <= kPrefixBases; i++) p |= (((b0 >> i) & 1u) | (((b1 >> i) & 1u) << 1)) << (2 * i);
    oh[0] = read_onehot(p & m2, bb & m1);            // prefix_pass never looks past the read's end
    oh[1] = read_onehot((p >> 2) & m2, (bb >> 1) & m1);
}

// Warp-cooperative screen of one (mapping, strand): the lowest try index t >= t0 whose start position passes the
// prefix filter, or kNoCand. Tries that dfsRecursive rejects on entry (offset beyond the node, alignment.go:199-201)
// are never enumerated: stage 1 covers only the offsets that exist on the seed node, stage 2 pools the existing
// offsets 0..10 of up to 32 contained nodes at a time (inclusive scan + owner search, as in warp_probe); the try
// NUMBERING stays the reference's (decode_try), so cursors and results are unchanged. wr, sn, t0, len are uniform.
__device__ __forceinline__ uint32_t screen_strand(const DevIndex& ix, const WinRec& wr, uint32_t win, uint32_t sn_seq_off, uint32_t sn_seq_len, uint32_t t0,
                                                  const uint32_t (&oh)[2], uint32_t len, uint32_t lane) {
    constexpr uint32_t FULL = 0xffffffffu;
    {   // can ANY try of this (window, strand) pass? The window's set of possible 5-base prefixes answers with one load
        // (lane 0: the read as it is — stages 1, 2, 4; lane 1: without its first base — stage 3). A base that is not ACGT has
        // an empty one-hot nibble and passes the allele-set test, so such a read is not filtered.
        bool may = false;
        if (lane < 2u) {
            const uint32_t o = oh[lane];
            uint32_t idx = 0;
            bool known = true;
#pragma unroll
            for (uint32_t i = 0; i < 5u; i++) {
                const uint32_t nib = (o >> (4u * i)) & 0xFu;
                known = known && nib != 0u;
                idx |= ((static_cast<uint32_t>(__ffs(static_cast<int>(nib))) - 1u) & 3u) << (2u * i);
            }
            may = !known || ((__ldg(ix.win_kmers + static_cast<size_t>(win) * 32u + (idx >> 5)) >> (idx & 31u)) & 1u) != 0u;
        }
        if (!__any_sync(FULL, may)) return kNoCand;
    }
    const uint32_t T1 = wr.merge_span + wr.win_size + 1u;
    // stage 1: offsets OffSet + t on the seed node
    const uint32_t room = sn_seq_len > wr.offset ? sn_seq_len - wr.offset : 0u;
    const uint32_t n1 = T1 < room ? T1 : room;
    for (uint32_t base = t0 & ~31u; base < n1; base += 32) {
        const uint32_t t = base + lane;
        const bool ok = t >= t0 && t < n1 && prefix_pass(ix, sn_seq_off + wr.offset + t, oh[0], len);
        const uint32_t ball = __ballot_sync(FULL, ok);
        if (ball) return base + (__ffs(ball) - 1);
    }
    // stage 2: offsets 0..10 on every contained node
    const uint32_t ci0 = t0 > T1 ? (t0 - T1) / 11u : 0u;
    for (uint32_t cb = ci0 & ~31u; cb < wr.cn_cnt; cb += 32) {
        const uint32_t ci = cb + lane;
        uint32_t cnt = 0, so = 0;
        if (ci < wr.cn_cnt) {
            const NodeRec cn = ix.nodes[ix.cn_node[wr.cn_off + ci]];
            so = cn.seq_off; cnt = cn.seq_len < 11u ? cn.seq_len : 11u;
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); if (lane >= static_cast<uint32_t>(d)) incl += v; }
        const uint32_t excl = incl - cnt, total = __shfl_sync(FULL, incl, 31);
        for (uint32_t ib = 0; ib < total; ib += 32) {
            const uint32_t item = ib + lane;
            uint32_t o = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t v = __shfl_sync(FULL, incl, (o + step - 1) & 31u);
                if (v <= item) o += step;
            }
            o &= 31u;
            const uint32_t off0 = item - __shfl_sync(FULL, excl, o);
            const uint32_t pos = __shfl_sync(FULL, so, o) + off0;
            const uint32_t t = T1 + 11u * (cb + o) + off0;
            const bool ok = item < total && t >= t0 && prefix_pass(ix, pos, oh[0], len);
            const uint32_t ball = __ballot_sync(FULL, ok);
            if (ball) return __shfl_sync(FULL, t, __ffs(ball) - 1);
        }
    }
    // stage 3 (1-base start clip: read[1:]) and stage 4 (1-base end clip: read[:len-1]) at the seed position (alignment.go:73-103)
    const uint32_t tA = T1 + 11u * wr.cn_cnt;
    {
        const uint32_t t = tA + lane;
        const bool ok = lane < 2 && t >= t0 && room > 0 && prefix_pass(ix, sn_seq_off + wr.offset, lane == 0 ? oh[1] : oh[0], len - 1);
        const uint32_t ball = __ballot_sync(FULL, ok);
        if (ball) return tA + (__ffs(ball) - 1);
    }
    return kNoCand;
}

// graphminion.go:94: the read is reverse complemented for the second strand; Go panics for a byte > 'T' (seqio.go:122)
__device__ __forceinline__ void check_revcomp_bytes(const AlignArgs& a, uint32_t r, const uint8_t* __restrict__ rp, uint32_t len, uint32_t lane) {
    if (a.nw32 && a.read_ok2[r]) return;   // upper-case ACGT only
    bool badb = false;
    for (uint32_t i = lane; i < len; i += 32) badb |= rp[i] > 'T';
    if (__any_sync(0xffffffffu, badb) && lane == 0) { if (atomicCAS(a.error, 0, -6) == 0) a.error[1] = static_cast<int>(r); }
}

__global__ void __launch_bounds__(256, 4) align_screen_kernel(DevIndex ix, RoundArgs ra) {
    const AlignArgs& a = ra.a;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_queue = *ra.n_queue;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, total_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t q = gwarp; q < n_queue; q += total_warps) {
        const uint32_t s = ra.queue[q];
        const ScreenDesc d = a.sdesc[s];                               // everything the first mapping needs, five 16-byte loads
        const PairCursor cur = ra.cursor[s];
        const uint32_t hb = d.hit_begin, he = hb + d.hit_count, r = d.read, len = d.len;
        const uint8_t* rp = d.packed ? a.seq : a.seq + a.off[r];       // read bytes: only reads without a packed copy look at them
        uint32_t m = hb + (cur.m_strand >> 1), strand = cur.m_strand & 1u, t0 = cur.t;
        uint2 cand = make_uint2(kNoCand, 0u);
        while (m < he) {
            WinRec wr = d.w0;
            uint32_t win = d.win0;
            uint32_t sn_off = d.sn_seq_off, sn_len = d.sn_seq_len;
            if (m != hb) {                                             // further mappings of the pair (rare): fetched on demand
                win = a.hits[m];
                wr = ix.wins[win];
                const NodeRec sn = ix.nodes[wr.node];
                sn_off = sn.seq_off; sn_len = sn.seq_len;
            }
            uint32_t oh[2];
            if (d.packed) { oh[0] = strand ? d.oh.z : d.oh.x; oh[1] = strand ? d.oh.w : d.oh.y; }
            else warp_read_prefix(a, r, rp, len, strand != 0, lane, oh);
            const uint32_t t = screen_strand(ix, wr, win, sn_off, sn_len, t0, oh, len, lane);
            if (t != kNoCand) { cand = make_uint2(((m - hb) << 1) | strand, t); break; }
            t0 = 0;
            if (strand == 0) { strand = 1; if (!d.packed) check_revcomp_bytes(a, r, rp, len, lane); }
            else { strand = 0; m++; }
        }
        if (lane == 0) ra.cand[s] = cand;
    }
}


// walk one try of pair s: 1 = aligned (the pair's outputs are filled), 0 = no path id from this try, -1 (PACKED_ONLY)
// = the pair cannot take the packed walk (read with other bytes than ACGT or longer than the packed copy, graph of
// more than 256 paths) and is left to the byte-wise walk of align_finish_kernel.
// srow: 9 words of shared memory owned by this thread (or nullptr): reads of up to 128 bases are walked from a copy of their
// packed words there — the walk touches them at every node step, and from global memory that was a sector per step.
constexpr uint32_t kWalkReadWords = 9;
template <bool PACKED_ONLY>
__device__ __forceinline__ int walk_try(const DevIndex& ix, const AlignArgs& a, const uint8_t* lut, uint32_t s, uint32_t hb,
                                        uint32_t m, uint32_t strand, uint32_t t, DfsFrame* stack, uint32_t* mask_ws, uint32_t depth_cap,
                                        uint32_t* srow = nullptr) {
    const WinRec wr = ix.wins[a.hits[m]];
    const uint32_t r = a.hit_read[hb];
    const uint32_t o = a.off[r], len = a.off[r + 1] - o;
    const uint32_t mw = ix.graph_mask_words[wr.graph];
    uint32_t node, off0, stage;
    decode_try(ix, wr, t, &node, &off0, &stage);
    const uint32_t rlen = stage >= 3 ? len - 1 : len;
    DfsResult res;
    res.nrec = 0; res.ntrav = 0;
    bool inline_masks = false;   // seg_mask holds the path bitset of every traversal
    if (a.nw32 && mw <= kMaskWordsInline && a.read_ok2[r]) {
        const uint32_t* rd2 = a.reads2 + static_cast<size_t>(r) * 2u * a.nw32 + (strand ? a.nw32 : 0u);
        const uint32_t base0 = (strand ? a.nw32 * 16u - len : 0u) + (stage == 3 ? 1u : 0u);
        if (srow != nullptr && a.nw32 == 8u) {               // one sector in, then shared memory only
            const uint4 lo = __ldg(reinterpret_cast<const uint4*>(rd2)), hi = __ldg(reinterpret_cast<const uint4*>(rd2) + 1);
            srow[0] = lo.x; srow[1] = lo.y; srow[2] = lo.z; srow[3] = lo.w; srow[4] = hi.x; srow[5] = hi.y; srow[6] = hi.z; srow[7] = hi.w;
            srow[8] = 0u;                                      // only ever shifted into bases past the end of the read
            rd2 = srow;
        }
        dfs_packed<false>(ix, node, off0, rd2, base0, rlen, mw, ix.graph_has_n[wr.graph] != 0, stack, mask_ws, depth_cap, &res,
                          a.seg_mask + static_cast<size_t>(s) * kTravWords);
        inline_masks = res.ntrav * mw <= static_cast<uint32_t>(kTravWords);
    } else if (PACKED_ONLY) {
        return -1;
    } else {
        ReadView rd{a.seq + o, lut, len, stage == 3 ? 1u : 0u, strand != 0};
        if (mw <= kMaskWordsInline) {
            dfs_masked(ix, node, off0, rd, rlen, mw, stack, mask_ws, depth_cap, &res);
            if (res.ntrav == 1) {
                inline_masks = true;
                for (uint32_t wi = 0; wi < mw; wi++) a.seg_mask[static_cast<size_t>(s) * kTravWords + wi] = res.mask[wi];
            }
        } else dfs_align<DFS_COUNT>(ix, node, off0, rd, rlen, mw, stack, depth_cap, &res, static_cast<uint32_t*>(nullptr), nullptr);
    }
    if (res.nrec == 0) return 0;
    PairOut p = a.pairs[s];
    p.n_incremented = m - hb + 1;
    p.rec_count = res.nrec; p.reverse = static_cast<uint8_t>(strand);
    p.clip_start = stage == 3; p.clip_end = stage == 4; p.stage = static_cast<uint8_t>(stage);
    a.pairs[s] = p;
    a.seg_nrec[s] = res.nrec;
    a.seg_locus[s] = make_uint2(node, off0);
    a.seg_ntrav[s] = inline_masks ? res.ntrav : (res.ntrav | kTravRewalk);
    return 1;
}

__global__ void __launch_bounds__(128, 8) align_walk_kernel(DevIndex ix, RoundArgs ra) {
    __shared__ uint32_t s_read[128 * kWalkReadWords];          // row stride 9 words: conflict-free
    uint32_t* srow = s_read + threadIdx.x * kWalkReadWords;
    const AlignArgs& a = ra.a;
    const uint32_t n_queue = *ra.n_queue;
    const uint32_t gthread = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x;
    const uint32_t depth_cap = a.max_len + 2;
    DfsFrame* stack = a.stack_ws + static_cast<size_t>(gthread) * depth_cap;
    uint32_t* mask_ws = a.mask_ws + static_cast<size_t>(gthread) * depth_cap * kMaskWordsInline;
    unsigned failed = 0;
    for (uint32_t q = gthread; q < n_queue; q += total) {
        const uint32_t s = ra.queue[q];
        const uint2 cand = ra.cand[s];
        if (cand.x == kNoCand) continue;                        // try list exhausted: the default (unaligned) result stands
        const uint32_t hb = a.seg_begin[s];
        const int rc = walk_try<true>(ix, a, nullptr, s, hb, hb + (cand.x >> 1), cand.x & 1u, cand.y, stack, mask_ws, depth_cap, srow);
        if (rc == 0) {
            ra.cursor[s] = PairCursor{cand.x, cand.y + 1};      // resume just past this try (the screen handles t == T)
            ra.queue_next[atomicAdd(ra.n_queue_next, 1u)] = s;
            failed++;
        } else if (rc < 0) {
            ra.cursor[s] = PairCursor{cand.x, cand.y};          // the byte-wise walk starts from this very try
            ra.slow_queue[atomicAdd(ra.n_slow, 1u)] = s;
        }
    }
    failed = __reduce_add_sync(0xffffffffu, failed);
    if ((threadIdx.x & 31) == 0 && failed) atomicAdd(&a.counters[3], static_cast<unsigned long long>(failed));
}

// Leftovers after the fixed number of rounds (pairs that keep producing tries which pass the filter but yield no
// path id — low-complexity sequence, 'N' wildcards): ONE WARP PER PAIR alternates screen steps and walks (lane 0)
// until one aligns or the list is exhausted.
__global__ void __launch_bounds__(128, 5) align_finish_kernel(DevIndex ix, RoundArgs ra) {
    __shared__ uint8_t lut[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = complement_base(static_cast<uint8_t>(i));
    __syncthreads();
    const AlignArgs& a = ra.a;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_queue = *ra.n_queue;
    const uint32_t gthread = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gwarp = gthread >> 5, total_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t depth_cap = a.max_len + 2;
    DfsFrame* stack = a.stack_ws + static_cast<size_t>(gthread) * depth_cap;
    uint32_t* mask_ws = a.mask_ws + static_cast<size_t>(gthread) * depth_cap * kMaskWordsInline;
    for (uint32_t q = gwarp; q < n_queue; q += total_warps) {
        const uint32_t s = ra.queue[q];
        const PairOut pp = a.pairs[s];                                 // one sector: read, hit range (filled by align_init_kernel)
        const uint32_t hb = pp.hit_begin, he = hb + pp.hit_count, r = pp.read;
        const uint32_t o = a.off[r], len = a.off[r + 1] - o;
        const uint8_t* rp = a.seq + o;
        const PairCursor cur = ra.cursor[s];
        uint32_t m = hb + (cur.m_strand >> 1), strand = cur.m_strand & 1u, t0 = cur.t;
        bool done = false;
        while (m < he && !done) {
            const uint32_t win = a.hits[m];
            const WinRec wr = ix.wins[win];
            const NodeRec sn = ix.nodes[wr.node];
            uint32_t oh[2];
            warp_read_prefix(a, r, rp, len, strand != 0, lane, oh);
            while (!done) {
                const uint32_t t = screen_strand(ix, wr, win, sn.seq_off, sn.seq_len, t0, oh, len, lane);
                if (t == kNoCand) break;
                uint32_t okw = 0;
                if (lane == 0) okw = walk_try<false>(ix, a, lut, s, hb, m, strand, t, stack, mask_ws, depth_cap) > 0 ? 1u : 0u;
                done = __shfl_sync(0xffffffffu, okw, 0) != 0;
                t0 = t + 1;
            }
            if (done) break;
            t0 = 0;
            if (strand == 0) { strand = 1; check_revcomp_bytes(a, r, rp, len, lane); }
            else { strand = 0; m++; }
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
    const uint32_t* seg_mask;
    const uint32_t* seg_ntrav;
    uint32_t* rec_path;
    int32_t* rec_pos;
    void* rec_c;               // compact output: path ids only, 1 or 2 bytes each (RECW); rec_path / rec_pos unused then
    CPairOut* cpairs;          // compact output: [n_segs]
    DfsFrame* stack_ws;
    uint32_t* mask_ws;
    uint32_t max_len;
    const uint32_t* reads2;
    const uint8_t* read_ok2;
    uint32_t nw32;
    const uint32_t* order;     // pair ids sorted by first window (or nullptr)
    uint32_t* multi_queue;     // pairs whose records need a second DFS (several traversals, > 256 paths)
    uint32_t* n_multi;         // device scalar, zeroed before align_emit_kernel
};

// A GROUP OF 8 LANES PER PAIR (four pairs per warp: the kernel is bound by the chain of dependent loads per pair — order,
// pair, locus, start node, path list — so more pairs in flight per warp is what speeds it up): write the pair's
// records at the scanned offset. The walk kept the path bitset of every traversal
// that fits the pair's kTravWords words (24 traversals in a graph of <= 32 paths, 4 in a graph of 161..192, 3 in one of 225..256):
// each is expanded with the lanes striding over the start node's path list (path ids ascending == record order).
// Pairs with more traversals than that (kTravRewalk) are left to align_emit_multi_kernel.
// RECW selects the record format: 0 = (u32 path id, i32 position) pairs; 1 / 2 = the compact output — the path id alone
// in 1 or 2 bytes plus one CPairOut per pair carrying the start locus, from which the host derives every position as
// Position[path] of the start node + offset (alignment.go:296): an eighth of the record bytes written here and moved to
// the host.
template <int RECW>
__global__ void __launch_bounds__(256) align_emit_kernel(DevIndex ix, EmitArgs a) {
    const uint32_t n_segs = *a.n_segs_ptr;
    const uint32_t lane = threadIdx.x & 31, gl = lane & 7u, gshift = lane & 24u;
    const uint32_t gmask = 0xffu << gshift;                         // the group's lanes: every vote below is among them only
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, total_groups = (gridDim.x * blockDim.x) >> 3;
    for (uint32_t q = group; q < n_segs; q += total_groups) {
        const uint32_t s = a.order ? a.order[q] : q;   // window order: neighbouring groups expand the same start nodes (L1 hits)
        const PairOut p = a.pairs[s];
        const uint32_t rb = a.rec_off[s];
        if (gl == 0) a.pairs[s].rec_begin = rb;
        if (RECW != 0 && gl == 0) {
            const uint2 l = p.rec_count ? a.seg_locus[s] : make_uint2(0xffffffffu, 0u);
            a.cpairs[s] = CPairOut{p.read, l.x, (l.y & kCPairOffsetMask) | (p.reverse ? kCPairReverse : 0u) | (p.clip_start ? kCPairClipStart : 0u) | (p.clip_end ? kCPairClipEnd : 0u),
                                   p.rec_count};
        }
        if (p.rec_count == 0) continue;
        const uint32_t ntrav = a.seg_ntrav[s];
        if (!(ntrav & kTravRewalk)) {
            const uint32_t mw = ix.graph_mask_words[p.graph];
            if (RECW != 0) {
                // compact output: the records of a traversal ARE the set bits of its path bitset, ascending (the bitset is an
                // AND over the traversal's nodes, the start node included, so it is a subset of that node's paths) — no start
                // node, no path list, no positions. Lane j of the group owns word j of the bitset (mw <= 8 here: larger
                // graphs re-walk), an exclusive scan of the popcounts over the group places its bits.
                uint32_t written = 0;
                for (uint32_t t = 0; t < ntrav; t++) {                 // traversals in DFS order, path ids ascending inside one
                    uint32_t m = gl < mw ? a.seg_mask[static_cast<size_t>(s) * kTravWords + t * mw + gl] : 0u;
                    const uint32_t c = __popc(m);
                    uint32_t incl = c;
#pragma unroll
                    for (uint32_t d = 1; d < 8; d <<= 1) { const uint32_t v = __shfl_up_sync(gmask, incl, d, 8); if (gl >= d) incl += v; }
                    uint32_t slot = rb + written + incl - c;
                    written += __shfl_sync(gmask, incl, 7, 8);
                    while (m) {
                        const uint32_t pid = gl * 32u + static_cast<uint32_t>(__ffs(static_cast<int>(m))) - 1u;
                        m &= m - 1u;
                        if (RECW == 1) static_cast<uint8_t*>(a.rec_c)[slot] = static_cast<uint8_t>(pid);
                        else static_cast<uint16_t*>(a.rec_c)[slot] = static_cast<uint16_t>(pid);
                        slot++;
                    }
                }
                continue;
            }
            const uint2 loc = a.seg_locus[s];
            const NodeRec n0 = ix.nodes[loc.x];
            uint32_t written = 0;
            for (uint32_t t = 0; t < ntrav; t++) {                     // traversals in DFS order, path ids ascending inside one
                const uint32_t* mk = a.seg_mask + static_cast<size_t>(s) * kTravWords + t * mw;
                for (uint32_t j0 = 0; j0 < n0.path_cnt; j0 += 8) {
                    const uint32_t j = j0 + gl;
                    uint32_t pid = 0; bool on = false;
                    if (j < n0.path_cnt) { pid = ix.node_path_id[n0.path_off + j]; on = (mk[pid >> 5] >> (pid & 31)) & 1u; }
                    const uint32_t ball = (__ballot_sync(gmask, on) >> gshift) & 0xffu;
                    if (on) {
                        const uint32_t slot = rb + written + __popc(ball & ((1u << gl) - 1u));
                        a.rec_path[slot] = pid;
                        a.rec_pos[slot] = ix.node_path_pos[n0.path_off + j] + static_cast<int32_t>(loc.y);   // alignment.go:296
                    }
                    written += __popc(ball);
                }
            }
        }
    }
}

// Pairs whose records need a second DFS (more traversals spell the read than seg_mask can hold — several are common
// where the MSA places the gaps of two sequences differently — or a graph of more than 256 paths), collected IN WINDOW ORDER so that the 32 walks a
// warp of align_emit_multi_kernel runs together are walks over the same graph region.
__global__ void __launch_bounds__(256) align_emit_classify_kernel(EmitArgs a) {
    const uint32_t n_segs = *a.n_segs_ptr;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t qb = i0 - lane; qb < n_segs; qb += stride) {
        const uint32_t q = qb + lane;
        uint32_t s = 0; bool multi = false;
        if (q < n_segs) { s = a.order ? a.order[q] : q; multi = (a.seg_ntrav[s] & kTravRewalk) != 0; }
        const uint32_t ball = __ballot_sync(0xffffffffu, multi);
        if (!ball) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.n_multi, static_cast<uint32_t>(__popc(ball)));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (multi) a.multi_queue[base + __popc(ball & ((1u << lane) - 1u))] = s;
    }
}

// ONE THREAD PER QUEUED PAIR: re-runs the DFS from the pair's successful start in emit mode — several traversals
// spell the read (typically through an 'N' node next to the read's own allele), each contributes its own records.
template <class PT>
__device__ __forceinline__ void emit_rewalk(const DevIndex& ix, const EmitArgs& a, uint32_t s, DfsFrame* stack, uint32_t* mask_ws, uint32_t depth_cap,
                                            PT* out_path, int32_t* out_pos) {
    const PairOut p = a.pairs[s];
    const uint2 loc = a.seg_locus[s];
    const uint32_t mw = ix.graph_mask_words[p.graph];
    const uint32_t o = a.off[p.read], len = a.off[p.read + 1] - o;
    const uint32_t rlen = len - p.clip_start - p.clip_end;
    DfsResult res;
    res.nrec = 0; res.ntrav = 0;
    if (a.nw32 && mw <= kMaskWordsInline && a.read_ok2[p.read]) {
        const uint32_t* rd2 = a.reads2 + static_cast<size_t>(p.read) * 2u * a.nw32 + (p.reverse ? a.nw32 : 0u);
        const uint32_t base0 = (p.reverse ? a.nw32 * 16u - len : 0u) + (p.clip_start ? 1u : 0u);
        dfs_packed<true, PT>(ix, loc.x, loc.y, rd2, base0, rlen, mw, ix.graph_has_n[p.graph] != 0, stack, mask_ws, depth_cap, &res, nullptr, out_path, out_pos);
    } else {
        GlobalRead rd{a.seq + o, len, p.clip_start ? 1u : 0u, p.reverse != 0};
        dfs_align<DFS_EMIT, GlobalRead, PT>(ix, loc.x, loc.y, rd, rlen, mw, stack, depth_cap, &res, out_path, out_pos);
    }
}

template <int RECW>
__global__ void __launch_bounds__(128) align_emit_multi_kernel(DevIndex ix, EmitArgs a) {
    const uint32_t n = *a.n_multi;
    const uint32_t gthread = blockIdx.x * blockDim.x + threadIdx.x, total = gridDim.x * blockDim.x;
    const uint32_t depth_cap = a.max_len + 2;
    DfsFrame* stack = a.stack_ws + static_cast<size_t>(gthread) * depth_cap;
    uint32_t* mask_ws = a.mask_ws + static_cast<size_t>(gthread) * depth_cap * kMaskWordsInline;
    for (uint32_t q = gthread; q < n; q += total) {
        const uint32_t s = a.multi_queue[q];
        const uint32_t rb = a.rec_off[s];
        if (RECW == 0) emit_rewalk<uint32_t>(ix, a, s, stack, mask_ws, depth_cap, a.rec_path + rb, a.rec_pos + rb);
        else if (RECW == 1) emit_rewalk<uint8_t>(ix, a, s, stack, mask_ws, depth_cap, static_cast<uint8_t*>(a.rec_c) + rb, nullptr);
        else emit_rewalk<uint16_t>(ix, a, s, stack, mask_ws, depth_cap, static_cast<uint16_t*>(a.rec_c) + rb, nullptr);
    }
}

}  // namespace groot
