/*
 * grootgpu.h — C ABI of libgrootgpu.so, the B200-native (sm_100a) implementation of the
 * `groot align` hot path of will-rowe/groot v1.1.2.
 *
 * The reference is pure Go and has no FFI: the narrowest seam with stable types is the pipeline
 * stage ReadMapper.Run -> theBoss.mapReads (src/pipeline/sketch.go:308-351, src/pipeline/boss.go:108-242).
 * Every entry point below names the reference interface it replaces; INTEGRATION.md shows the cgo
 * stub a maintainer would add on the Go side.
 *
 * Conventions: plain C types only; every function returns 0 on success and a negative GROOTGPU_ERR_*
 * code otherwise (grootgpu_last_error() gives the message of the calling thread's last failure);
 * nothing throws or aborts across the boundary. The caller owns every host buffer it passes in;
 * the library owns device memory and the result buffers behind its opaque handles. One handle per
 * GPU; calls on one handle must be serialised by the caller, different handles are independent.
 * There is NO CPU fallback: every compute entry point fails with GROOTGPU_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef GROOTGPU_H
#define GROOTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GROOTGPU_VERSION "0.2.0"
#define GROOTGPU_REFERENCE_VERSION "1.1.2" /* src/version/version.go:6-12; enforced at cmd/align.go:96-98 */

enum {
    GROOTGPU_OK = 0,
    GROOTGPU_ERR_ARG = -1,         /* bad argument */
    GROOTGPU_ERR_CUDA = -2,        /* CUDA runtime failure / no usable device */
    GROOTGPU_ERR_IO = -3,          /* file could not be read / written */
    GROOTGPU_ERR_FORMAT = -4,      /* malformed MSA / index file */
    GROOTGPU_ERR_SHORT_READ = -5,  /* a read is shorter than k: the reference panics (boss.go:164-166) */
    GROOTGPU_ERR_BAD_BASE = -6,    /* RevComplement hit a byte > 'T': the reference panics (seqio.go:17-23,122) */
    GROOTGPU_ERR_CAPACITY = -7,    /* a documented limit was exceeded (see DESIGN.md "Limits") */
    GROOTGPU_ERR_EMPTY = -8,       /* nothing to index / empty index (lshe.go:103-105) */
    GROOTGPU_ERR_COMM = -9         /* NCCL failure / NCCL library not found (multi-GPU entry points only) */
};

typedef struct grootgpu_index grootgpu_index; /* graph store + containment index + device copies + workspaces */

/* Index parameters == the fields `groot index` stores in groot.gg (src/pipeline/runtime.go:15-27;
 * flag defaults cmd/index.go:45-49: k=31 s=21 w=100 x=8 y=4). */
typedef struct {
    uint32_t kmer_size;
    uint32_t sketch_size;
    uint32_t window_size;
    uint32_t num_part;
    uint32_t max_k;
} grootgpu_index_params;

typedef struct {
    grootgpu_index_params params;
    uint32_t n_graphs, n_masked_graphs, n_paths, n_nodes;
    uint64_t n_path_bases, n_raw_windows;
    uint32_t n_windows;      /* distinct (graph,node,offset,sketch) keys == WindowLookup entries */
    uint32_t max_merge_span;
    uint32_t max_paths_per_graph;
} grootgpu_index_info;

/* ---- index bring-up ------------------------------------------------------------------------- */

/* Replaces the `groot index` pipeline MSAconverter -> GraphSketcher -> SketchIndexer
 * (src/pipeline/index.go:37-211; graph.CreateGrootGraph src/graph/graph.go:37-218; WindowGraph
 * src/graph/graph.go:229-396). msa_paths[i] becomes graph i (cmd/index.go:143-153 passes the sorted
 * glob). Window sketching runs on the GPU with the same KHF kernel the read path uses. */
int grootgpu_index_build(const char* const* msa_paths, uint32_t n_msa, const grootgpu_index_params* params,
                         int device, grootgpu_index** out);
/* Same, taking every cluster*.msa of a directory in lexicographic order (cmd/index.go:143). */
int grootgpu_index_build_dir(const char* msa_dir, const grootgpu_index_params* params, int device, grootgpu_index** out);

/* Replace Info.Dump + ContainmentIndex.Dump (src/pipeline/runtime.go:64-73, src/lshe/lshe.go:72-92) and
 * Info.Load + ContainmentIndex.Load (runtime.go:75-91, lshe.go:95-146; cmd/align.go:94-107). The file
 * is this library's own flat little-endian format (".grootb200"), not Go gob. */
int grootgpu_index_save(const grootgpu_index* idx, const char* path);
int grootgpu_index_load(const char* path, int device, grootgpu_index** out);
/* The same from the reference's OWN index files: groot.gg (Info + graph Store) and groot.lshe (ContainmentIndex), Go gob
 * streams (src/pipeline/runtime.go:64-91, src/lshe/lshe.go:72-146; cmd/align.go:94-107). The graphs are adopted as the Go
 * host wrote them — node order, segment ids, path ids — so results index straight into the host's info.Store. */
int grootgpu_index_load_gob(const char* gg_path, const char* lshe_path, int device, grootgpu_index** out);
/* Host-only (no device): decodes and validates the two gob files and writes the canonical dump and/or its hash. */
int grootgpu_gob_dump(const char* gg_path, const char* lshe_path, const char* dump_path, uint64_t* hash);
/* The other direction: writes the index as the reference's own groot.gg + groot.lshe (Info.Dump + ContainmentIndex.Dump,
 * src/pipeline/runtime.go:64-73, src/lshe/lshe.go:72-92), so that an index built here can be loaded by the Go
 * `groot align` / `groot haplotype`. Graph weights (KmerFreq, KmerTotal, Marked) are written as they currently are. */
int grootgpu_index_save_gob(const grootgpu_index* idx, const char* gg_path, const char* lshe_path);
/* Host-only (no device) conversions between the library's flat index file and the reference's pair of gob files. */
int grootgpu_flat_to_gob(const char* flat_path, const char* gg_path, const char* lshe_path);
int grootgpu_gob_to_flat(const char* gg_path, const char* lshe_path, const char* flat_path);
void grootgpu_index_destroy(grootgpu_index* idx);
int grootgpu_index_get_info(const grootgpu_index* idx, grootgpu_index_info* out);

/* Canonical text dump (graphs, paths, nodes, windows) / its FNV-1a-64 hash: the parity surface
 * against the oracle's restatement of groot.gg + groot.lshe. */
int grootgpu_index_dump_file(const grootgpu_index* idx, const char* path);
int grootgpu_index_dump_hash(const grootgpu_index* idx, uint64_t* hash);

/* Host-only helper (no device needed): runs just the MSA -> variation graph step of grootgpu_index_build
 * (gfa.MSA2GFA + graph.CreateGrootGraph, src/pipeline/index.go:43-53) and writes the graph section of the
 * canonical dump and/or its hash. */
int grootgpu_graphs_dump(const char* const* msa_paths, uint32_t n_msa, const grootgpu_index_params* params,
                         const char* dump_path, uint64_t* hash);

/* @SQ content of the BAM header: name/length of path `path_id` of graph `graph_id`
 * (Store.GetSAMrefs, src/graph/graphio.go:141-154). The pointer stays valid until the index is destroyed. */
int grootgpu_index_ref(const grootgpu_index* idx, uint32_t graph_id, uint32_t path_id, const char** name, int32_t* length);

/* (K, L) the LSH Ensemble optimiser picks and the smallest number of equal sketch slots that passes
 * lshensemble.Containment(...) > threshold, for a query of `query_kmers` k-mers (lshe.go:153-171). */
int grootgpu_index_query_params(grootgpu_index* idx, uint32_t query_kmers, double threshold, uint32_t* K, uint32_t* L, uint32_t* eq_min);

/* Same computation from the index parameters alone (host only, no device, no handle). */
int grootgpu_query_params_host(const grootgpu_index_params* params, uint32_t query_kmers, double threshold,
                               uint32_t* K, uint32_t* L, uint32_t* eq_min);

/* ---- the hot path --------------------------------------------------------------------------- */

typedef struct {
    double containment_threshold; /* -t / --contThresh, default 0.99 (cmd/align.go:47) */
    int32_t no_align;             /* --noAlign (cmd/align.go:46): weight graphs from seeds only */
    int32_t keep_sketches;        /* also return the per-read KHF sketches (tests / debugging) */
    int32_t project_on_device;    /* 1: also run the ordered graph weighting (IncrementSubPath, graph.go:401-451) on the device,
                                     bit-identical to grootgpu_project_batch; do NOT call grootgpu_project_batch for that batch */
    int32_t results_on_device;    /* 1: skip the device->host copy of the result arrays; the d_* pointers of the
                                     result are set instead (multi-GPU gather, kernel-side timing) */
    int32_t compact_records;      /* 1: BAM-oriented compact output — cpairs[] + rec_path_c[] instead of hit_off / hits / pairs /
                                     rec_path / rec_pos (about a fifth of the bytes: what crosses PCIe and NVLink). Needs
                                     project_on_device or no weighting: grootgpu_project_batch wants the full arrays */
    uint32_t fixed_read_len;      /* > 0: every read of the batch has this many bases, back to back; grootgpu_align_batch then
                                     accepts seq_off == NULL and does not move 8 bytes of offset per read to the device */
} grootgpu_align_params;

/* One (read, graph) unit == one graphMinionPair (src/pipeline/graphminion.go:14-17). */
typedef struct {
    uint32_t read;           /* read index inside the batch */
    uint32_t graph;          /* GraphID */
    uint32_t hit_begin;      /* slice [hit_begin, hit_begin+hit_count) of hits[]: the mappings, already in */
    uint32_t hit_count;      /*   graphminion.go:57 order (Node, then OffSet, then window order) */
    uint32_t n_incremented;  /* mappings that received IncrementSubPath before the loop stopped (graphminion.go:64-98) */
    uint32_t rec_begin;      /* slice of rec_path[] / rec_pos[] */
    uint32_t rec_count;      /* 0 when no exact alignment was found */
    uint8_t reverse;         /* 1: aligned as reverse complement -> sam.Reverse (alignment.go:150-152) */
    uint8_t clip_start;      /* 1H at the start (alignment.go:73-85,132-134) */
    uint8_t clip_end;        /* 1H at the end   (alignment.go:88-103,136-138) */
    uint8_t stage;           /* 1..4: hierarchy stage that produced the alignment; 0 = none */
} grootgpu_pair;

/* Compact form of a pair (params->compact_records): everything the BAM writer needs (boss.go:225-242, alignment.go:114-156).
 * Record j of the pair has path id rec_path_c[first + j] (first = sum of rec_count of the pairs before it) and position
 * Position[path] of node `node` + offset (alignment.go:296: grootgpu_index_node_paths gives the node's path table);
 * the graph is the node's. Unaligned pairs are kept (rec_count == 0, node == 0xffffffff). */
typedef struct {
    uint32_t read;           /* read index inside the batch */
    uint32_t node;           /* global node index of the start node (graphs ascending, SortedNodes order) */
    uint32_t offset_flags;   /* bits 0..27 offset on that node | GROOTGPU_CPAIR_REVERSE / _CLIP_START / _CLIP_END */
    uint32_t rec_count;
} grootgpu_cpair;
#define GROOTGPU_CPAIR_OFFSET_MASK 0x0fffffffu
#define GROOTGPU_CPAIR_REVERSE     0x10000000u   /* sam.Reverse (alignment.go:150-152) */
#define GROOTGPU_CPAIR_CLIP_START  0x20000000u   /* 1H at the start (alignment.go:73-85,132-134) */
#define GROOTGPU_CPAIR_CLIP_END    0x40000000u   /* 1H at the end (alignment.go:88-103,136-138) */

/* Result of one batch. All pointers are HOST memory owned by the index handle, valid until the next
 * align call on that handle (or its destruction). Record j of a pair is (rec_path[j], rec_pos[j]):
 * sam.Record{Ref: references[path], Pos: pos}; the first record of a pair is primary, the others
 * carry sam.Secondary (alignment.go:147-149); Name/Seq/Qual/CIGAR follow from the read and the
 * pair's clip fields (alignment.go:114-139). */
typedef struct {
    uint32_t n_reads;
    uint64_t n_hits, n_pairs, n_records;
    const uint32_t* hit_off;   /* [n_reads+1] per-read slice of hits[] */
    const uint32_t* hits;      /* [n_hits] window ids, ascending per read (== graph, Node, OffSet order) */
    const grootgpu_pair* pairs;/* [n_pairs] ordered by (read, graph) */
    const uint32_t* rec_path;  /* [n_records] path id inside the pair's graph */
    const int32_t* rec_pos;    /* [n_records] 0-based start */
    const uint64_t* sketches;  /* [n_reads * sketch_size] when keep_sketches, else NULL */
    /* theBoss counters (boss.go:22-27): received, mapped (>=1 graph), multimapped (>1 graph), alignment records */
    uint64_t received, mapped, multimapped, alignments;
    /* device time of the batch in ms, CUDA events on the library's stream: [0] whole batch incl. copies,
     * [1] seed kernel (sketch+probe+verify), [2] align kernel (all pairs), [3] everything else on the device */
    float ms[4];
    /* device time per kernel family, summed over its launches (CUDA events around every launch):
     * [0] seed_kernel, [1] fill_kernel, [2] align_init + align_screen, [3] align_walk, [4] align_finish, [5] align_emit,
     * [6] project_count + project_expand + project_accumulate */
    float kernel_ms[8];
    uint32_t kernel_launches;  /* number of this library's own kernels launched for the batch (CUB scans/selects not counted) */
    uint64_t slow_path_pairs;  /* diagnostic: full DFS walks that yielded no path id (filter false positives) */
    /* device copies of the arrays above (same layouts), valid until the next align call on the handle */
    const uint32_t* d_hit_off;
    const uint32_t* d_hits;
    const grootgpu_pair* d_pairs;
    const uint32_t* d_rec_path;
    const int32_t* d_rec_pos;
    /* compact output (params->compact_records): host arrays (NULL with results_on_device) and their device copies; the
     * full-format host arrays above are NULL then, and so are d_rec_path / d_rec_pos */
    const grootgpu_cpair* cpairs;      /* [n_pairs] ordered by (read, graph) */
    const void* rec_path_c;            /* [n_records] path ids, rec_path_bytes each (1 when no graph has more than 256 paths, else 2) */
    uint32_t rec_path_bytes;
    const grootgpu_cpair* d_cpairs;
    const void* d_rec_path_c;
    uint32_t result_set;               /* which of the handle's two alternating sets of result arrays this batch wrote (used by grootgpu_gather) */
} grootgpu_batch_result;

/* Replaces the per-read loop of theBoss.mapReads (src/pipeline/boss.go:134-203: RunMinHash ->
 * db.Query -> dispatch) fused with the graphMinion loop (src/pipeline/graphminion.go:46-102: sort
 * mappings, AlignRead forward then reverse complement, stop at the first mapping that aligns).
 * seq = concatenated read bases (raw FASTQ line 2 bytes), seq_off[n_reads+1] = byte offsets (NULL allowed with
 * params->fixed_read_len).
 * Host buffers in, host results out (H2D / D2H inside); pinned buffers from grootgpu_host_alloc make
 * the copies asynchronous. Any batch size: the batch is streamed through the device in chunks on two
 * lanes (copy-in, kernels and copy-out of neighbouring chunks overlap; one helper thread is started and
 * joined inside the call), and the result arrays are batch-wide, exactly as if the batch had run in one
 * piece. Graph weights: untouched unless params->project_on_device is set (then the ordered f64 weighting
 * runs on the device, chained across chunks); otherwise call grootgpu_project_batch for the batch.
 * With params->results_on_device the batch runs in one piece (at most 4 GiB of bases). */
int grootgpu_align_batch(grootgpu_index* idx, const uint8_t* seq, const uint64_t* seq_off, uint32_t n_reads,
                         const grootgpu_align_params* params, grootgpu_batch_result* out);

/* Same computation with the reads already resident in HBM (kernel-side throughput measurement and
 * pipelines that keep reads on the device). d_seq must be readable 64 bytes past its end; d_seq_off
 * has n_reads+1 u32 entries; [min_len, max_len] bound the read lengths. stream = a cudaStream_t
 * (NULL = the library's own stream). */
int grootgpu_align_batch_device(grootgpu_index* idx, const uint8_t* d_seq, const uint32_t* d_seq_off, uint32_t n_reads,
                                uint32_t min_len, uint32_t max_len, const grootgpu_align_params* params,
                                void* stream, grootgpu_batch_result* out);

/* Path table of a node: PathIDs ascending and Position[pathID] (src/graph/node.go:13-22), and the node's GraphID — what
 * turns a compact record (node, offset, path id) into sam.Record{Ref, Pos}. Pointers stay valid until the index is destroyed. */
int grootgpu_index_node_paths(const grootgpu_index* idx, uint32_t node, uint32_t* graph, const uint32_t** path_ids,
                              const int32_t** positions, uint32_t* n_paths);

/* Replaces GrootGraph.IncrementSubPath as driven by the minion loop (src/graph/graph.go:401-451,
 * graphminion.go:60,67): replays, in read order, the weight increments of the first n_incremented
 * mappings of every pair in f64 on the host copy of the graphs (order-dependent float accumulation,
 * partitioned by graph). seq_off = the offsets given to grootgpu_align_batch (read lengths). */
int grootgpu_project_batch(grootgpu_index* idx, const grootgpu_batch_result* res, const uint64_t* seq_off);

/* KmerFreq of every node in (graph ascending, SortedNodes order) and KmerTotal per graph
 * (src/graph/node.go:21, src/graph/graph.go:26). n_nodes / n_graphs from grootgpu_index_get_info. */
int grootgpu_weights(const grootgpu_index* idx, double* kmer_freq, uint64_t* kmer_total);
int grootgpu_reset_weights(grootgpu_index* idx);

/* Replaces Sequence.RunMinHash(k, s, false, nil) for a batch of sequences (src/seqio/seqio.go:40-68 ->
 * src/minhash/khf.go:35-56): out[i*sketch_size + j]. Fails with GROOTGPU_ERR_SHORT_READ if any is < k. */
int grootgpu_sketch_batch(int device, const uint8_t* seq, const uint64_t* seq_off, uint32_t n_seqs,
                          uint32_t kmer_size, uint32_t sketch_size, uint64_t* out);

/* ---- multi-GPU ------------------------------------------------------------------------------ */
/* Reads are independent units (the reference already fans them over NumProc workers, boss.go:134-203): every rank
 * (one index handle per GPU, index replicated) maps a contiguous shard of the batch — rank 0 the first reads, rank 1 the
 * next, ... — with grootgpu_align_batch[_device](results_on_device = 1); no collective on the data path. What crosses
 * NVLink (NCCL, loaded at run time):
 *   - ONE gather of the per-rank result arrays to rank 0 per batch (grootgpu_gather), merged there in global read order
 *     — the single ordered BAM writer of boss.go:225-234;
 *   - the graph weights: IncrementSubPath is an order-dependent f64 accumulation (graph.go:401-451), so with
 *     project_on_device every rank expands and sorts its own increments and the per-node chains run rank after rank on
 *     ONE weight vector that travels round the ranks (send/recv of n_nodes doubles per batch) — bit-identical to one GPU
 *     mapping the whole batch. The chains run on their own stream, behind the mapping of the following batches. */
typedef struct grootgpu_comm grootgpu_comm;
#define GROOTGPU_COMM_ID_BYTES 256
/* rank 0 creates the id; the host hands it to the other ranks (MPI, torch.distributed, a file, another thread) */
int grootgpu_comm_id(uint8_t id[GROOTGPU_COMM_ID_BYTES]);
/* Collective over all ranks (blocks until every rank has called it). The communicator is attached to idx: from now on
 * the device-side graph weighting of that index takes part in the ring. One rank per index handle per process thread. */
int grootgpu_comm_create(grootgpu_index* idx, const uint8_t id[GROOTGPU_COMM_ID_BYTES], int rank, int world_size, grootgpu_comm** out);
/* Collective: `local` is the result of this rank's last align call (results_on_device = 1; either output format, the same
 * on all ranks). On rank 0 `merged` receives the batch-wide result: read indices global (rank 0's reads first), hit /
 * record offsets rebased, counters summed; d_* pointers always (complete once the communicator's stream has run: after
 * the next grootgpu_gather or grootgpu_comm_sync), valid until the next grootgpu_gather. to_host = 1: host arrays too, the
 * call returns after the copy; to_host = 2: host arrays too, copied asynchronously — complete after the next
 * grootgpu_gather / grootgpu_comm_sync and valid until the gather after that (two alternating sets). Other ranks may pass
 * merged = NULL. The transfer overlaps the next align call of every rank; the call itself waits for every rank's sizes
 * (one small all-gather), so a host that does not want to wait for the slowest rank may issue it from a second thread
 * while the first already runs the next align call on the same handle — the one exception to "calls on a handle are
 * serialised" (a gather must have returned before the align call after next starts). */
int grootgpu_gather(grootgpu_comm* comm, const grootgpu_batch_result* local, int to_host, grootgpu_batch_result* merged);
/* Collective: waits for the gathers and the weight ring; afterwards rank 0's index holds the graph weights of everything
 * mapped so far on all ranks (grootgpu_weights / _prune / _graph_save_gfa on rank 0), the other ranks' weights are zero. */
int grootgpu_comm_sync(grootgpu_comm* comm);
void grootgpu_comm_destroy(grootgpu_comm* comm);

/* ---- after the stream ends ------------------------------------------------------------------- */

/* Replaces GrootGraph.Prune over the store (src/graph/graph.go:455-525 via GraphPruner,
 * src/pipeline/sketch.go:378-430). kept[g] = 1 when graph g survives. */
int grootgpu_prune(grootgpu_index* idx, double min_kmer_coverage, uint8_t* kept);
/* Replaces GrootGraph.SaveGraphAsGFA (src/graph/graphio.go:19-112) minus the timestamp comment.
 * *written = 0 when the graph carries no weight (graphio.go:67-69). */
int grootgpu_graph_save_gfa(const grootgpu_index* idx, uint32_t graph_id, const char* path, int64_t total_kmers, int* written);

/* ---- utilities ------------------------------------------------------------------------------- */
int grootgpu_host_alloc(void** ptr, size_t bytes); /* pinned host memory */
int grootgpu_host_free(void* ptr);
int grootgpu_device_count(int* n);
/* Measurement aid (no reference counterpart): the integer-instruction issue rate the device sustains, in warp
 * instructions per second, for [0] multiply-adds only (FMA pipe), [1] shift/xor only (ALU pipe), [2] the 1:1 mix of the
 * two that the KHF hash loop (src/minhash/khf.go:44-53) compiles to — the roofline that bounds the sketch kernel. */
int grootgpu_int_issue_peak(int device, double* warp_inst_per_s);
const char* grootgpu_last_error(void);
const char* grootgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GROOTGPU_H */
