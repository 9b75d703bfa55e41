// libgrootgpu.so — C ABI implementation (include/grootgpu.h): index bring-up on the device, the
// batch pipeline  seed (sketch+probe+verify) -> scan -> fill -> segment -> align search -> scan -> emit,
// the ordered graph weighting (count / expand / sort beside the emit, the f64 chains behind the batch on st_acc),
// result transfer (full or compact arrays; host, or batch-wide device arrays), the chunked three-lane host path,
// the multi-GPU gather and weight ring over NCCL, and the host-side ordered weight replay.
//
// There is no CPU fallback anywhere in this file: every compute entry point needs a CUDA device and
// fails with GROOTGPU_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <dirent.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <fstream>
#include <functional>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/grootgpu.h"
#include "align_kernels.cuh"
#include "device_types.cuh"
#include "flat_index.h"
#include "nccl_dyn.h"
#include "project_kernels.cuh"
#include "seed_kernels.cuh"

using namespace groot;

static_assert(sizeof(PairOut) == sizeof(grootgpu_pair), "PairOut and grootgpu_pair must match");
static_assert(offsetof(PairOut, rec_count) == offsetof(grootgpu_pair, rec_count), "PairOut layout");
static_assert(offsetof(PairOut, stage) == offsetof(grootgpu_pair, stage), "PairOut layout");
static_assert(sizeof(CPairOut) == sizeof(grootgpu_cpair) && sizeof(CPairOut) == 16, "CPairOut and grootgpu_cpair must match");
static_assert(kCPairReverse == GROOTGPU_CPAIR_REVERSE && kCPairClipStart == GROOTGPU_CPAIR_CLIP_START && kCPairClipEnd == GROOTGPU_CPAIR_CLIP_END &&
              kCPairOffsetMask == GROOTGPU_CPAIR_OFFSET_MASK, "compact pair flags");

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct CommError : std::runtime_error { using std::runtime_error::runtime_error; };
#define NK(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess) throw CommError(std::string(#call) + ": " + nccl().GetErrorString(r_));   \
    } while (0)
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

// grow-only device / pinned-host buffers
struct DBuf {
    void* p = nullptr; size_t cap = 0;
    DBuf() = default;
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    void need(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
    ~DBuf() { if (p) cudaFree(p); }
};
struct HBuf {
    void* p = nullptr; size_t cap = 0;
    HBuf() = default;
    HBuf(const HBuf&) = delete;
    HBuf& operator=(const HBuf&) = delete;
    void need(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        CK(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        cap = want;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
    ~HBuf() { if (p) cudaFreeHost(p); }
};

template <class T>
T* upload(const std::vector<T>& v, std::vector<void*>& owned) {
    void* d = nullptr;
    CK(cudaMalloc(&d, std::max<size_t>(16, v.size() * sizeof(T))));
    owned.push_back(d);
    if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return static_cast<T*>(d);
}

MultTable make_mult(uint32_t k) {
    MultTable m;
    const uint64_t C = static_cast<uint64_t>(k) * GROOT_MULTI_SEED;
    for (uint64_t i = 0; i < 32; i++) { m.c[i] = i ^ C; m.low[i] = static_cast<uint32_t>((C & 31u) ^ i); }
    m.c0 = C & ~31ull; m.m32 = 32; m.one = 1;
    m.c_hi = static_cast<uint32_t>(C >> 32);
    int keybits = 27;
    if (const char* e = getenv("GROOTGPU_KHF_KEYBITS")) { const int v = atoi(e); if (v >= 1 && v <= 27) keybits = v; }
    m.key_low = (1u << (32 - keybits)) - 1u;
    return m;
}

// ---- scalar traffic without the copy engines ---------------------------------------------------------------
// The batch pipeline needs a handful of device scalars on the host (hit / pair / record totals, error flags) and
// pushes a few back. As cudaMemcpyAsync these 4-byte transfers queue on the DMA engines BEHIND the 200 MB chunk
// copies of the chunked host path and stall the compute stream for milliseconds; as tiny kernels writing mapped
// pinned memory (peek) or taking the value as an argument (poke / zero) they depend on the compute stream alone.
struct PeekArgs { const uint32_t* src[12]; uint32_t n; };
__global__ void peek_kernel(PeekArgs a, volatile uint32_t* dst) {
    if (threadIdx.x < a.n) dst[threadIdx.x] = *a.src[threadIdx.x];
    __threadfence_system();
}
struct PokeArgs { uint32_t* dst[8]; uint32_t val[8]; uint32_t n; };
__global__ void poke_kernel(PokeArgs a) {
    if (threadIdx.x < a.n) *a.dst[threadIdx.x] = a.val[threadIdx.x];
}
struct ZeroArgs { uint32_t* p[6]; uint32_t words[6]; uint32_t n; };
__global__ void zero_kernel(ZeroArgs a) {
    for (uint32_t b = 0; b < a.n; b++)
        for (uint32_t i = threadIdx.x; i < a.words[b]; i += blockDim.x) a.p[b][i] = 0u;
}

// ---- measured integer-issue peak ("measure the INT peak with a micro-benchmark rather than assuming lane count") ----
// 16 independent chains per thread of the instruction classes seed_kernel's hash loop is made of: MODE 0 multiply-adds
// only (IMAD: FMA pipe), MODE 1 shift + xor only (SHF / LOP3: ALU pipe), MODE 2 both in the 1:1 pipe mix that lets a
// scheduler issue every cycle. Full occupancy, no memory traffic in the loop; the result is warp instructions per second.
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* __restrict__ out, uint32_t iters, uint32_t a, uint32_t b, uint32_t sh) {
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 16u + i + blockIdx.x;
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (MODE == 0) { x[i] = x[i] * a + b; x[i] = x[i] * b + a; }
                else if (MODE == 1) { x[i] ^= x[i] >> sh; x[i] ^= x[i] << sh; }
                else { x[i] = x[i] * a + b; x[i] ^= x[i] >> sh; x[i] = x[i] * b + a; }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= x[i];
    if (acc == 0x12345678u) out[0] = acc;   // keeps the chains alive, practically never taken
}
constexpr uint32_t kIntPeakInstPerIter[3] = {4 * 16 * 2, 4 * 16 * 4, 4 * 16 * 4};   // body instructions per thread and iteration: IMAD+IMAD / SHF+LOP3+SHF+LOP3 / IMAD+SHF+LOP3+IMAD

int pick_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) throw CudaError("no CUDA device available (libgrootgpu has no CPU fallback)");
    if (device < 0 || device >= n) throw CudaError("device ordinal out of range");
    CK(cudaSetDevice(device));
    return device;
}

// ---- sketch-only launcher (index windows, grootgpu_sketch_batch) --------------------------------
template <int S>
void launch_sketch(const uint8_t* d_seq, const uint64_t* d_off, const uint32_t* d_lens, uint32_t fixed_len, uint32_t n, uint32_t k,
                   uint64_t* d_out, int* d_err, cudaStream_t st) {
    int blocks = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n) + kSeedThreads - 1) / kSeedThreads, 148ull * 16));
    sketch_kernel<S><<<std::max(blocks, 1), kSeedThreads, 0, st>>>(d_seq, d_off, d_lens, fixed_len, n, k, make_mult(k), d_out, d_err);
}
#define GROOT_S_LIST(X) X(8) X(10) X(16) X(20) X(21) X(24) X(30) X(32)
bool sketch_dispatch(uint32_t S, const uint8_t* d_seq, const uint64_t* d_off, const uint32_t* d_lens, uint32_t fixed_len, uint32_t n,
                     uint32_t k, uint64_t* d_out, int* d_err, cudaStream_t st) {
    switch (S) {
#define X(s) case s: launch_sketch<s>(d_seq, d_off, d_lens, fixed_len, n, k, d_out, d_err, st); return true;
        GROOT_S_LIST(X)
#undef X
        default: return false;
    }
}

// sketches n sequences (host in, host out) on the current device
void sketch_host(const uint8_t* seqs, size_t seqs_len, const uint64_t* off, const uint32_t* lens, uint32_t fixed_len, uint32_t n,
                 uint32_t k, uint32_t S, uint64_t* out) {
    DBuf d_seq, d_off, d_lens, d_out, d_err;
    d_seq.need(seqs_len + 64); d_off.need(sizeof(uint64_t) * n); d_out.need(sizeof(uint64_t) * n * S); d_err.need(8);
    CK(cudaMemcpy(d_seq.p, seqs, seqs_len, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_off.p, off, sizeof(uint64_t) * n, cudaMemcpyHostToDevice));
    if (lens) { d_lens.need(sizeof(uint32_t) * n); CK(cudaMemcpy(d_lens.p, lens, sizeof(uint32_t) * n, cudaMemcpyHostToDevice)); }
    CK(cudaMemset(d_err.p, 0, 8));
    if (!sketch_dispatch(S, d_seq.as<uint8_t>(), d_off.as<uint64_t>(), lens ? d_lens.as<uint32_t>() : nullptr, fixed_len, n, k,
                         d_out.as<uint64_t>(), d_err.as<int>(), 0)) {
        // any other sketch size: the run-time-S kernel (seed_kernels.cuh, "any sketch size, any maxK")
        if (S < 1 || S > static_cast<uint32_t>(kMaxSketch)) throw std::length_error("sketch size must be in 1.." + std::to_string(kMaxSketch) + " (documented limit)");
        const int blocks = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n) + kSeedThreads - 1) / kSeedThreads, 148ull * 16));
        sketch_generic_kernel<<<std::max(blocks, 1), kSeedThreads>>>(d_seq.as<uint8_t>(), d_off.as<uint64_t>(), lens ? d_lens.as<uint32_t>() : nullptr, fixed_len, n, k, S,
                                                                      d_out.as<uint64_t>(), d_err.as<int>());
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int err[2];
    CK(cudaMemcpy(err, d_err.p, 8, cudaMemcpyDeviceToHost));
    if (err[0] != 0) throw std::runtime_error("sequence " + std::to_string(err[1]) + " is shorter than k");
    CK(cudaMemcpy(out, d_out.p, sizeof(uint64_t) * n * S, cudaMemcpyDeviceToHost));
}
void sketch_cb(void*, const uint8_t* seqs, size_t seqs_len, const uint64_t* off, uint32_t n, uint32_t w, uint32_t k, uint32_t S, uint64_t* out) {
    sketch_host(seqs, seqs_len, off, nullptr, w, n, k, S, out);
}

std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss; ss << f.rdbuf();
    return ss.str();
}

}  // namespace

// =================================================================================================
// Everything one batch in flight needs besides the (read-only) index: device scratch, result arrays, streams, events.
// The handle owns two of them ("lanes").
// The sorted (node, increment) items of one batch / chunk, waiting for their turn on the accumulate stream.
struct AccSlot {
    DBuf keys, vals;
    cudaEvent_t ev_sorted = nullptr, ev_done = nullptr;   // items complete (side stream) / the accumulate that read them has finished
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;         // around the accumulate kernel (read back by a later batch)
    bool timed = false;
};

struct Workspace {
    cudaStream_t stream = nullptr;     // the lane's compute stream
    cudaStream_t st_side = nullptr;    // the graph weighting's count / expand / sort run here, next to the record emit on the compute stream
    cudaEvent_t ev[6] = {};
    cudaEvent_t kev[64] = {};          // per-launch timing: pairs (begin, end) tagged with a category
    int kev_cat[32] = {};
    int kev_n = 0;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done = nullptr;
    cudaEvent_t ev_out[2] = {};        // copy-out of the chunk that last used result set 0 / 1 has completed
    int rset = 0;                      // result set in use: the lane alternates so that a chunk never waits for the previous copy-out
    uint32_t* h_peek = nullptr;        // mapped pinned words the device writes scalars into (peek()): no copy engine involved
    uint32_t* d_peek = nullptr;
    DBuf seq, off, n_hits, hit_off, stage, hits, hit_read, seg_flag, seg_begin, scalars, pairs, seg_nrec, seg_locus, rec_off, seg_mask, seg_ntrav, mask_ws, cursor, cand, queue_a, queue_b, qcount,
        rec_path, rec_pos, stack_ws, cub_tmp, cub_tmp2, sketches, tile_counter, error, reads2, read_ok2, read_oh, sdesc, qkey, qkey2, order, slow_q, len_minmax, seed_q,
        item_cnt, item_off, pkeys, pvals, cpairs, rec_c;
    DBuf alt_hits, alt_pairs, alt_rec_path, alt_rec_pos, alt_hit_off, alt_sketches, alt_cpairs, alt_rec_c;   // the other result set
    AccSlot acc[2];                    // the accumulate runs behind the batch (st_acc): two sets of sorted items per lane
    int acc_i = 0;
    bool ring_first = true, ring_last = true;   // this batch / chunk opens / closes the call's turn in the multi-GPU weight ring
    // ordering of the accumulates across lanes (chunked host path): called on the host around the enqueue on st_acc
    std::function<void()> acc_before, acc_after;

    void swap_result_sets() {          // DBuf owns its pointer: swap fields, not objects
        auto sw = [](DBuf& x, DBuf& y) { std::swap(x.p, y.p); std::swap(x.cap, y.cap); };
        sw(hits, alt_hits); sw(pairs, alt_pairs); sw(rec_path, alt_rec_path); sw(rec_pos, alt_rec_pos); sw(hit_off, alt_hit_off); sw(sketches, alt_sketches);
        sw(cpairs, alt_cpairs); sw(rec_c, alt_rec_c);
        rset ^= 1;
    }
    void create() {
        // priorities: the ordered f64 chains (st_acc of the index: short, latency-bound, and on N GPUs the one serial
        // resource of the whole job) run first, the mapping kernels next, the multi-GPU gather (st_gather of the
        // communicator) takes what is left — it has a whole batch time to finish
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        const int prio_map = prio_hi < prio_lo ? prio_lo - 1 : prio_lo;
        CK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_map));
        CK(cudaStreamCreateWithPriority(&st_side, cudaStreamNonBlocking, prio_map));
        for (auto& e : ev) CK(cudaEventCreate(&e));
        for (auto& e : kev) CK(cudaEventCreate(&e));
        for (cudaEvent_t* e : {&ev_fork, &ev_join, &ev_done, &ev_out[0], &ev_out[1], &acc[0].ev_sorted, &acc[0].ev_done, &acc[1].ev_sorted, &acc[1].ev_done})
            CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (AccSlot& a : acc) { CK(cudaEventCreate(&a.ev_t0)); CK(cudaEventCreate(&a.ev_t1)); }
        CK(cudaHostAlloc(reinterpret_cast<void**>(&h_peek), 64 * sizeof(uint32_t), cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_peek), h_peek, 0));
    }
    ~Workspace() {
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& e : kev) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : {ev_fork, ev_join, ev_done, ev_out[0], ev_out[1]}) if (e) cudaEventDestroy(e);
        for (AccSlot& a : acc) for (cudaEvent_t e : {a.ev_sorted, a.ev_done, a.ev_t0, a.ev_t1}) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
        if (st_side) cudaStreamDestroy(st_side);
        if (h_peek) cudaFreeHost(h_peek);
    }
};

struct grootgpu_index {
    FlatIndex h;
    int device = 0;
    DevIndex d{};
    std::vector<void*> owned;          // device allocations freed on destroy
    // LSH tables, built lazily per K (all bands)
    std::vector<LshTable> h_tables;    // [(K-1)*n_bands + band]
    LshTable* d_tables = nullptr;
    bool tables_built[17] = {};
    // per (q, threshold) parameter cache
    std::map<std::pair<uint32_t, double>, LenParam> param_cache;
    DBuf len_params;                   // per read length: (K, L, eq_min), shared by the lanes (prepare_params, under params_mu)
    std::mutex params_mu;
    HBuf r_hit_off, r_hits, r_pairs, r_rec_path, r_rec_pos, r_sketches, r_cpairs, r_rec_c;   // batch-wide host result arrays
    static constexpr uint32_t kLanes = 3, kInBufs = 2 * kLanes;
    Workspace ws[kLanes];              // lanes: the chunked host path runs consecutive chunks on the lanes in turn
    cudaStream_t st_in = nullptr, st_out = nullptr;   // copy-in / copy-out of the chunked host path
    cudaStream_t st_acc = nullptr;     // the ordered f64 chains of the graph weighting (and the multi-GPU weight ring): runs behind the batches
    std::mutex acc_mu;                 // enqueues on st_acc come from both lanes' host threads
    float acc_ms_pending = 0.f;        // accumulate kernel time of earlier batches, not yet reported
    grootgpu_comm* comm = nullptr;     // attached by grootgpu_comm_create
    uint32_t rec_width = 1;            // bytes per path id of the compact output
    // batch-wide result arrays on the DEVICE (chunked host path with results_on_device): two sets, alternating per call
    struct DevResult { DBuf hit_off, hits, pairs, rec_path, rec_pos, cpairs, rec_c; } bres[2];
    int call_parity = 0;               // flips with every align call: which result set (workspace or bres) the call writes
    DBuf in_seq[kInBufs], in_off64[kInBufs], in_off32[kInBufs];
    cudaEvent_t ev_in[kInBufs] = {}, ev_t0 = nullptr, ev_t1 = nullptr;
    // graph weights live on the device once a batch was projected there; the host copy is refreshed lazily
    double* d_kmer_freq = nullptr;
    unsigned long long* d_kmer_total = nullptr;
    const uint32_t* d_cn_count = nullptr;
    const double* d_cn_ratio = nullptr;
    bool weights_on_device = false;
    std::vector<LenParam> h_len_params;
    double lp_threshold = -1; uint32_t lp_min = 1, lp_max = 0;

    ~grootgpu_index() {
        for (void* p : owned) cudaFree(p);
        if (st_in) cudaStreamDestroy(st_in);
        if (st_out) cudaStreamDestroy(st_out);
        if (st_acc) cudaStreamDestroy(st_acc);
        for (auto& e : ev_in) if (e) cudaEventDestroy(e);
        if (ev_t0) cudaEventDestroy(ev_t0);
        if (ev_t1) cudaEventDestroy(ev_t1);
    }
};

constexpr int kCountWords = 8;   // per rank, exchanged by every gather: n_reads, n_hits, n_pairs, n_records, mapped, multimapped, slow_path_pairs, format

// One rank of a multi-GPU run (include/grootgpu.h "multi-GPU"). Two NCCL communicators, so that the weight ring (stream
// st_acc of the index) and the result gather (st_gather) never serialise behind each other.
struct grootgpu_comm {
    grootgpu_index* ix = nullptr;
    int rank = 0, world = 1;
    ncclComm_t ring = nullptr, gath = nullptr;
    cudaStream_t st_gather = nullptr;
    cudaEvent_t ev_sent[2] = {};           // the gather that read result set 0 / 1 of the index has finished with it
    bool ring_pending = false;             // rank 0: the last rank has sent (or will send) a weight vector that was not received yet
    DBuf d_counts;                         // [kCountWords * (world + 2)] u64: the same on the device (all-gather in place)
    uint64_t* h_counts = nullptr;          // pinned, [kCountWords * (world + 2)]: this rank's words, then everybody's
    DBuf m_hit_off, m_hits, m_pairs, m_rec_path, m_rec_pos, m_cpairs, m_rec_c;   // rank 0: the merged batch
    struct HostSet { HBuf hit_off, hits, pairs, rec_path, rec_pos, cpairs, rec_c; } hs[2];   // host copies of the merged batch: alternating,
    int hs_i = 0;                                                                             // so that an asynchronous copy never lands in the arrays the caller is still reading
    ~grootgpu_comm() {
        if (ring) nccl().CommDestroy(ring);
        if (gath) nccl().CommDestroy(gath);
        if (st_gather) cudaStreamDestroy(st_gather);
        for (cudaEvent_t e : {ev_sent[0], ev_sent[1]}) if (e) cudaEventDestroy(e);
        if (h_counts) cudaFreeHost(h_counts);
    }
};

namespace {

// peek: device words -> host through mapped memory (returns after synchronising the stream); poke / zero: host values -> device words
const uint32_t* peek(Workspace* w, cudaStream_t st, std::initializer_list<const uint32_t*> src, int slot = 0) {
    PeekArgs a{};
    for (const uint32_t* p : src) a.src[a.n++] = p;
    peek_kernel<<<1, 32, 0, st>>>(a, w->d_peek + 16 * slot);     // slot: one per stream that may have a peek in flight
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return w->h_peek + 16 * slot;
}
void poke(cudaStream_t st, std::initializer_list<std::pair<uint32_t*, uint32_t>> w) {
    PokeArgs a{};
    for (auto& x : w) { a.dst[a.n] = x.first; a.val[a.n] = x.second; a.n++; }
    poke_kernel<<<1, 32, 0, st>>>(a);
    CK(cudaGetLastError());
}
void zero_words(cudaStream_t st, std::initializer_list<std::pair<void*, uint32_t>> bufs) {
    ZeroArgs a{};
    for (auto& x : bufs) { a.p[a.n] = static_cast<uint32_t*>(x.first); a.words[a.n] = x.second; a.n++; }
    zero_kernel<<<1, 64, 0, st>>>(a);
    CK(cudaGetLastError());
}

void index_to_device(grootgpu_index* ix) {
    pick_device(ix->device);
    FlatIndex& h = ix->h;
    if (h.p.max_k < 1 || h.p.max_k > 16) throw std::length_error("maxK must be in 1..16 (documented limit)");
    if (h.p.S / h.p.max_k < 1) throw std::runtime_error("sketch size must be >= max_k");
    if (h.p.S > static_cast<uint32_t>(kMaxSketch)) throw std::length_error("sketch size must be <= " + std::to_string(kMaxSketch) + " (documented limit)");
    DevIndex& d = ix->d;
    d.nodes = upload(h.nodes, ix->owned);
    std::vector<uint8_t> seq_padded = h.node_seq; seq_padded.resize(seq_padded.size() + 16, 0);
    d.node_seq = upload(seq_padded, ix->owned);
    d.edges = upload(h.edges, ix->owned);
    d.node_path_id = upload(h.node_path_id, ix->owned);
    d.node_path_pos = upload(h.node_path_pos, ix->owned);
    d.node_mask = upload(h.node_mask, ix->owned);
    d.wins = upload(h.wins, ix->owned);
    d.cn_node = upload(h.cn_node, ix->owned);
    d.sketches = upload(h.sketches, ix->owned);
    d.graph_mask_words = upload(h.graph_mask_words, ix->owned);
    ix->d_cn_count = upload(h.cn_count, ix->owned);
    {   // IncrementSubPath's per-node share of a window (graph.go:427-441): segLen / total, both f64, total = sum of integers (exact)
        std::vector<double> ratio(h.cn_node.size(), 1.0);
        for (const WinRec& w : h.wins) {
            double total = 0.0;
            for (uint32_t j = 0; j < w.cn_cnt; j++) total += static_cast<double>(h.nodes[h.cn_node[w.cn_off + j]].seq_len);
            for (uint32_t j = 0; j < w.cn_cnt; j++) ratio[w.cn_off + j] = static_cast<double>(h.nodes[h.cn_node[w.cn_off + j]].seq_len) / total;
        }
        ix->d_cn_ratio = upload(ratio, ix->owned);
    }
    {
        std::vector<uint32_t> pset;
        build_prefix_sets(h, pset);
        d.pfxset = upload(pset, ix->owned);
        std::vector<uint32_t> wk;
        build_window_kmer_sets(h, pset, wk);
        d.win_kmers = upload(wk, ix->owned);
    }
    {   // 2-bit copy of the node sequences + per-graph 'N' flag for the packed walk (align_kernels.cuh, dfs_packed)
        std::vector<uint32_t> seq2((h.node_seq.size() + 15) / 16 + 2, 0u);
        std::vector<uint32_t> n2(seq2.size(), 0u);
        for (size_t i = 0; i < h.node_seq.size(); i++) {
            const uint8_t b = h.node_seq[i];
            seq2[i >> 4] |= pack_base2(b) << (2 * (i & 15));
            if (b != 'A' && b != 'C' && b != 'G' && b != 'T') n2[i >> 4] |= 1u << (2 * (i & 15));
        }
        std::vector<uint8_t> has_n(std::max<uint32_t>(h.n_graphs, 1), 0);
        for (uint32_t g = 0; g < h.n_graphs; g++)
            for (uint32_t n = h.graph_node_base[g]; n < h.graph_node_base[g + 1] && !has_n[g]; n++)
                for (uint32_t i = 0; i < h.nodes[n].seq_len; i++) {
                    const uint8_t b = h.node_seq[h.nodes[n].seq_off + i];
                    if (b != 'A' && b != 'C' && b != 'G' && b != 'T') { has_n[g] = 1; break; }
                }
        d.node_seq2 = upload(seq2, ix->owned); d.node_n2 = upload(n2, ix->owned); d.graph_has_n = upload(has_n, ix->owned);
    }
    {   // one-sector node records for the packed walk (device_types.cuh, WalkNode)
        std::vector<WalkNode> wn(h.nodes.size());
        for (uint32_t g = 0; g < h.n_graphs; g++) {
            const uint32_t mw = h.graph_mask_words[g];
            for (uint32_t n = h.graph_node_base[g]; n < h.graph_node_base[g + 1]; n++) {
                const NodeRec& nr = h.nodes[n];
                WalkNode& x = wn[n];
                x.seq_off = nr.seq_off; x.seq_len = nr.seq_len; x.edge_cnt = nr.edge_cnt;
                x.edge = nr.edge_cnt == 1 ? h.edges[nr.edge_off] : nr.edge_off;
                for (uint32_t j = 0; j < kWalkMaskWords; j++) x.mask[j] = 0;
                if (mw <= kWalkMaskWords) { for (uint32_t j = 0; j < mw; j++) x.mask[j] = h.node_mask[nr.mask_off + j]; }
                else x.mask[0] = nr.mask_off;
            }
        }
        d.wnodes = upload(wn, ix->owned);
    }
    d.k = h.p.k; d.S = h.p.S; d.max_k = h.p.max_k; d.n_bands = h.p.S / h.p.max_k; d.n_wins = static_cast<uint32_t>(h.wins.size());
    {   // windows grouped by identical sketch (DevIndex::full): key = sketch_digest, bucket = window ids ascending
        const uint32_t S = h.p.S, W = static_cast<uint32_t>(h.wins.size());
        std::vector<std::array<uint32_t, 4>> dig(W);
        for (uint32_t w = 0; w < W; w++) sketch_digest(&h.sketches[static_cast<size_t>(w) * S], S, dig[w].data());
        std::vector<uint32_t> order(W);
        for (uint32_t i = 0; i < W; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return dig[x] != dig[y] ? dig[x] < dig[y] : x < y; });
        uint32_t uniq = 0;
        for (uint32_t i = 0; i < W; i++) if (i == 0 || dig[order[i]] != dig[order[i - 1]]) uniq++;
        uint32_t cap = 16;
        while (cap < 2 * uniq) cap <<= 1;
        std::vector<LshSlot> slots(cap);
        memset(slots.data(), 0, slots.size() * sizeof(LshSlot));
        for (uint32_t i = 0; i < W;) {
            uint32_t j = i + 1;
            while (j < W && dig[order[j]] == dig[order[i]]) j++;
            uint32_t hsh = band_key_hash(dig[order[i]].data()) & (cap - 1);
            while (slots[hsh].count != 0) hsh = (hsh + 1) & (cap - 1);
            memcpy(slots[hsh].key, dig[order[i]].data(), 16); slots[hsh].start = i; slots[hsh].count = j - i;
            i = j;
        }
        d.full.slots = upload(slots, ix->owned); d.full.wins = upload(order, ix->owned); d.full.mask = cap - 1; d.full.pad = 0;
    }
    ix->h_tables.assign(static_cast<size_t>(h.p.max_k) * d.n_bands, LshTable{nullptr, nullptr, 0, 0});
    void* dt = nullptr;
    CK(cudaMalloc(&dt, ix->h_tables.size() * sizeof(LshTable)));
    ix->owned.push_back(dt);
    ix->d_tables = static_cast<LshTable*>(dt);
    CK(cudaMemcpy(dt, ix->h_tables.data(), ix->h_tables.size() * sizeof(LshTable), cudaMemcpyHostToDevice));
    d.tables = ix->d_tables;
    for (Workspace& w : ix->ws) w.create();
    ix->len_params.need(60002 * sizeof(LenParam));   // never reallocated: the lanes' kernels read it while prepare_params extends it
    CK(cudaStreamCreateWithFlags(&ix->st_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ix->st_out, cudaStreamNonBlocking));
    {
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaStreamCreateWithPriority(&ix->st_acc, cudaStreamNonBlocking, prio_hi));   // see Workspace::create
    }
    {
        uint32_t max_paths = 0;
        for (uint32_t g = 0; g < h.n_graphs; g++) max_paths = std::max(max_paths, h.n_paths_of(g));
        if (max_paths > 65536) throw std::length_error("a graph of more than 65536 paths (documented limit)");
        ix->rec_width = max_paths <= 256 ? 1u : 2u;
    }
    for (auto& e : ix->ev_in) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventCreate(&ix->ev_t0)); CK(cudaEventCreate(&ix->ev_t1));
    if (h.kmer_freq.size() != h.nodes.size()) h.kmer_freq.assign(h.nodes.size(), 0.0);
    if (h.kmer_total.size() != h.n_graphs) h.kmer_total.assign(h.n_graphs, 0);
    if (h.node_marked.size() != h.nodes.size()) h.node_marked.assign(h.nodes.size(), 0);
    ix->d_kmer_freq = upload(h.kmer_freq, ix->owned);
    {
        std::vector<unsigned long long> kt(h.kmer_total.begin(), h.kmer_total.end());
        ix->d_kmer_total = upload(kt, ix->owned);
    }
}

// Flattened CSR of the lshensemble index for prefix length K: for every band, group the windows by the
// low 32 bits of their first K band hashes (LshForest with 32-bit hash values; the forest's binary-search
// prefix probe returns exactly one such group) into an open-addressing table of 32-byte slots.
void build_tables(grootgpu_index* ix, uint32_t K) {
    if (ix->tables_built[K]) return;
    FlatIndex& h = ix->h;
    const uint32_t S = h.p.S, maxk = h.p.max_k, nb = ix->d.n_bands, W = static_cast<uint32_t>(h.wins.size());
    for (uint32_t b = 0; b < nb; b++) {
        std::vector<uint32_t> order(W);
        for (uint32_t i = 0; i < W; i++) order[i] = i;
        auto keyof = [&](uint32_t w, uint32_t key[4]) {
            for (uint32_t j = 0; j < 4; j++) key[j] = j < K ? static_cast<uint32_t>(h.sketches[static_cast<size_t>(w) * S + b * maxk + j]) : 0u;   // K > 4: the first four (lsh_probe_generic checks the rest)
        };
        std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
            uint32_t kx[4], ky[4]; keyof(x, kx); keyof(y, ky);
            for (int j = 0; j < 4; j++) if (kx[j] != ky[j]) return kx[j] < ky[j];
            return x < y;  // window ids ascending inside a bucket
        });
        uint32_t uniq = 0;
        for (uint32_t i = 0; i < W; i++) {
            uint32_t a[4], c[4];
            if (i == 0) { uniq++; continue; }
            keyof(order[i], a); keyof(order[i - 1], c);
            if (memcmp(a, c, 16) != 0) uniq++;
        }
        uint32_t cap = 16;
        while (cap < 2 * uniq) cap <<= 1;
        std::vector<LshSlot> slots(cap);
        memset(slots.data(), 0, slots.size() * sizeof(LshSlot));
        uint32_t i = 0;
        while (i < W) {
            uint32_t key[4]; keyof(order[i], key);
            uint32_t j = i + 1;
            while (j < W) { uint32_t k2[4]; keyof(order[j], k2); if (memcmp(key, k2, 16) != 0) break; j++; }
            uint32_t hsh = band_key_hash(key) & (cap - 1);
            while (slots[hsh].count != 0) hsh = (hsh + 1) & (cap - 1);
            memcpy(slots[hsh].key, key, 16); slots[hsh].start = i; slots[hsh].count = j - i;
            i = j;
        }
        LshTable t;
        t.slots = upload(slots, ix->owned);
        t.wins = upload(order, ix->owned);
        t.mask = cap - 1; t.pad = 0;
        ix->h_tables[(K - 1) * nb + b] = t;
    }
    CK(cudaMemcpy(ix->d_tables, ix->h_tables.data(), ix->h_tables.size() * sizeof(LshTable), cudaMemcpyHostToDevice));
    ix->tables_built[K] = true;
}

LenParam param_for(grootgpu_index* ix, uint32_t q, double t) {
    auto key = std::make_pair(q, t);
    auto it = ix->param_cache.find(key);
    if (it != ix->param_cache.end()) return it->second;
    const FlatIndex& h = ix->h;
    int K = 0, L = 0;
    int x = static_cast<int>(h.p.w - h.p.k + 1);  // NumWindowKmers == every partition's Upper (lshe.go:108-146)
    optimal_kl(static_cast<int>(h.p.max_k), static_cast<int>(h.p.S / h.p.max_k), x, static_cast<int>(q), t, &K, &L);
    LenParam lp;
    lp.K = static_cast<uint8_t>(K); lp.L = static_cast<uint8_t>(L);
    lp.eq_min = static_cast<uint16_t>(eq_min_for(static_cast<int>(h.p.S), static_cast<int>(q), x, t));
    ix->param_cache[key] = lp;
    return lp;
}

// make sure len_params[len] is on the device for every len in [min_len, max_len] and the tables exist
void prepare_params(grootgpu_index* ix, uint32_t min_len, uint32_t max_len, double t) {
    std::lock_guard<std::mutex> lock(ix->params_mu);   // the lanes of the chunked host path share the tables
    const uint32_t k = ix->h.p.k;
    if (ix->lp_threshold == t && min_len >= ix->lp_min && max_len <= ix->lp_max) return;
    uint32_t lo = std::min(min_len, ix->lp_threshold == t ? ix->lp_min : min_len);
    uint32_t hi = std::max(max_len, ix->lp_threshold == t ? ix->lp_max : max_len);
    ix->h_len_params.assign(static_cast<size_t>(hi) + 1, LenParam{0, 0, 0xffff});
    for (uint32_t len = std::max(lo, k); len <= hi; len++) {
        LenParam lp = param_for(ix, len - k + 1, t);
        ix->h_len_params[len] = lp;
        if (lp.eq_min <= ix->h.p.S && lp.K >= 1) build_tables(ix, lp.K);
    }
    if (ix->h_len_params.size() * sizeof(LenParam) > ix->len_params.cap) throw std::length_error("read longer than 60000 bases (documented limit)");
    CK(cudaMemcpy(ix->len_params.p, ix->h_len_params.data(), ix->h_len_params.size() * sizeof(LenParam), cudaMemcpyHostToDevice));
    ix->lp_threshold = t; ix->lp_min = lo; ix->lp_max = hi;
}

// ---- kernel dispatch on the compile-time sketch size ---------------------------------------------
struct SeedLaunch {
    template <int S>
    static void seed(const DevIndex& d, const SeedArgs& a, uint32_t k, size_t smem, int blocks, cudaStream_t st) {
        CK(cudaFuncSetAttribute(seed_kernel<S, 4, SEED_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        seed_kernel<S, 4, SEED_FULL><<<blocks, kSeedThreads, smem, st>>>(d, a, make_mult(k));
    }
    template <int S>
    static void seed_queued(const DevIndex& d, const SeedArgs& a, uint32_t k, size_t smem, int blocks, cudaStream_t st) {
        CK(cudaFuncSetAttribute(seed_kernel<S, 4, SEED_QUEUED>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        seed_kernel<S, 4, SEED_QUEUED><<<blocks, kSeedThreads, smem, st>>>(d, a, make_mult(k));
    }
    template <int S>
    static void fill(const DevIndex& d, const FillArgs& a, uint32_t k, int blocks, cudaStream_t st) {
        fill_kernel<S, 4, false><<<blocks, kSeedThreads, 0, st>>>(d, a, make_mult(k));
        fill_kernel<S, 4, true><<<blocks, kSeedThreads, 0, st>>>(d, a, make_mult(k));   // reads with more than HSTAGE hits (rare)
    }
    template <int S>
    static int seed_occupancy(size_t smem) {
        int nb = 0;
        cudaFuncSetAttribute(seed_kernel<S, 4, SEED_FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, seed_kernel<S, 4, SEED_FULL>, kSeedThreads, smem);
        return nb;
    }
};
// MAXK is a template parameter of the probe (static register indexing); the slot layout always holds 4
// key words, so max_k < 4 indexes run the MAXK=4 code with the unused words zero... only when the band
// stride matches: bands are max_k apart, hence a separate instantiation per max_k would be needed.
// Round 1 compiles max_k == 4 (the reference default, cmd/index.go:49); other values are rejected.
bool seed_dispatch(uint32_t S, const DevIndex& d, const SeedArgs& a, uint32_t k, size_t smem, int blocks, cudaStream_t st) {
    switch (S) {
#define X(s) case s: SeedLaunch::seed<s>(d, a, k, smem, blocks, st); return true;
        GROOT_S_LIST(X)
#undef X
        default: return false;
    }
}
bool seed_queued_dispatch(uint32_t S, const DevIndex& d, const SeedArgs& a, uint32_t k, size_t smem, int blocks, cudaStream_t st) {
    switch (S) {
#define X(s) case s: SeedLaunch::seed_queued<s>(d, a, k, smem, blocks, st); return true;
        GROOT_S_LIST(X)
#undef X
        default: return false;
    }
}
bool fill_dispatch(uint32_t S, const DevIndex& d, const FillArgs& a, uint32_t k, int blocks, cudaStream_t st) {
    switch (S) {
#define X(s) case s: SeedLaunch::fill<s>(d, a, k, blocks, st); return true;
        GROOT_S_LIST(X)
#undef X
        default: return false;
    }
}
int seed_occupancy_dispatch(uint32_t S, size_t smem) {
    switch (S) {
#define X(s) case s: return SeedLaunch::seed_occupancy<s>(smem);
        GROOT_S_LIST(X)
#undef X
        default: return 0;
    }
}

int g_num_sms(int device) {
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n;
}

// keep the host copy of the weights authoritative before anything on the host touches them (the chains of earlier
// batches may still be running on st_acc)
void sync_weights_to_host(grootgpu_index* ix) {
    if (!ix->weights_on_device) return;
    CK(cudaStreamSynchronize(ix->st_acc));
    FlatIndex& h = ix->h;
    CK(cudaMemcpy(h.kmer_freq.data(), ix->d_kmer_freq, h.kmer_freq.size() * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<unsigned long long> kt(h.kmer_total.size());
    CK(cudaMemcpy(kt.data(), ix->d_kmer_total, kt.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < kt.size(); i++) h.kmer_total[i] = kt[i];
    ix->weights_on_device = false;
}
void push_weights_to_device(grootgpu_index* ix) {
    if (ix->weights_on_device) return;
    CK(cudaStreamSynchronize(ix->st_acc));
    FlatIndex& h = ix->h;
    CK(cudaMemcpy(ix->d_kmer_freq, h.kmer_freq.data(), h.kmer_freq.size() * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<unsigned long long> kt(h.kmer_total.begin(), h.kmer_total.end());
    CK(cudaMemcpy(ix->d_kmer_total, kt.data(), kt.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    ix->weights_on_device = true;
}

// a10 on the device (project_kernels.cuh). Per batch / chunk, on the lane's side stream `st`:
//   project_begin   per-mapping item counts + their exclusive scan (no host round trip)
//   project_finish  item total -> host, expand, stable sort by node into one of the lane's two AccSlots
// (the caller runs the record emit on the main stream between the two), and then, on the index-wide stream st_acc,
//   acc_enqueue     one ordered f64 chain per node over the slot's items.
// The chains are the only part that is serial across batches (and, multi-GPU, across ranks: the weight vector arrives
// from the previous rank before the call's first chain and leaves for the next rank after its last), so they run BEHIND
// the pipeline: nothing on the lanes waits for them except the reuse of a slot two batches later.
ProjectArgs project_args(grootgpu_index* ix, Workspace* w, const uint32_t* d_off, uint32_t n) {
    ProjectArgs pa{};
    pa.off = d_off; pa.hits = w->hits.as<uint32_t>(); pa.hit_read = w->hit_read.as<uint32_t>(); pa.pairs = w->pairs.as<PairOut>();
    pa.n_segs_ptr = w->scalars.as<uint32_t>(); pa.n_hits_ptr = w->hit_off.as<uint32_t>() + n;
    pa.cn_count = ix->d_cn_count; pa.cn_ratio = ix->d_cn_ratio; pa.item_cnt = w->item_cnt.as<uint32_t>(); pa.item_off = w->item_off.as<uint32_t>();
    pa.kmer_total = ix->d_kmer_total; pa.k = ix->h.p.k;
    return pa;
}
template <class KB, class KE>
void project_begin(grootgpu_index* ix, Workspace* w, const uint32_t* d_off, uint32_t n, uint32_t n_segs, uint32_t H, int sms, cudaStream_t st, KB kbegin, KE kend, uint32_t& launches) {
    w->item_cnt.need(4ull * H); w->item_off.need(4ull * (H + 1));
    ProjectArgs pa = project_args(ix, w, d_off, n);
    const int blocks = std::max(1, std::min<int>((n_segs + 255) / 256, sms * 8));
    kbegin(6); project_count_kernel<<<blocks, 256, 0, st>>>(ix->d, pa); launches++; kend();
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, pa.item_cnt, w->item_off.as<uint32_t>(), static_cast<int>(H), st);
    w->cub_tmp2.need(tmp + 16);
    cub::DeviceScan::ExclusiveSum(w->cub_tmp2.p, tmp, pa.item_cnt, w->item_off.as<uint32_t>(), static_cast<int>(H), st);
}

// the multi-GPU weight ring, on st_acc: the vector of n_nodes doubles travels rank 0 -> 1 -> ... -> last -> 0
void ring_recv(grootgpu_index* ix) {
    grootgpu_comm* c = ix->comm;
    if (!c || c->world < 2) return;
    if (c->rank == 0 && !c->ring_pending) return;                 // nothing has been round yet
    NK(nccl().Recv(ix->d_kmer_freq, ix->h.nodes.size(), ncclDouble, (c->rank + c->world - 1) % c->world, c->ring, ix->st_acc));
    if (c->rank == 0) c->ring_pending = false;
}
void ring_send(grootgpu_index* ix) {
    grootgpu_comm* c = ix->comm;
    if (!c || c->world < 2) return;
    NK(nccl().Send(ix->d_kmer_freq, ix->h.nodes.size(), ncclDouble, (c->rank + 1) % c->world, c->ring, ix->st_acc));
    if (c->rank == 0) c->ring_pending = true;                     // it comes back from the last rank
}

// slot == nullptr: the batch / chunk has nothing to add, but still takes its turn (ordering hooks, ring)
void acc_enqueue(grootgpu_index* ix, Workspace* w, AccSlot* slot, uint32_t n_items, int sms, uint32_t& launches) {
    if (w->acc_before) w->acc_before();                            // chunk order across the lanes (host side)
    {
        std::lock_guard<std::mutex> lock(ix->acc_mu);
        cudaStream_t sa = ix->st_acc;
        if (slot) CK(cudaStreamWaitEvent(sa, slot->ev_sorted, 0));
        if (w->ring_first) ring_recv(ix);
        if (slot) {
            const uint32_t n_nodes = static_cast<uint32_t>(ix->h.nodes.size());
            const int ablocks = std::max(1, std::min<int>(static_cast<int>((n_nodes + 7) / 8), sms * 8));
            CK(cudaEventRecord(slot->ev_t0, sa));
            project_accumulate_kernel<<<ablocks, kAccWarps * 32, 0, sa>>>(slot->keys.as<uint32_t>(), slot->vals.as<double>(), n_items, n_nodes, ix->d_kmer_freq); launches++;
            CK(cudaGetLastError());
            CK(cudaEventRecord(slot->ev_t1, sa));
            slot->timed = true;
        }
        if (w->ring_last) ring_send(ix);
        if (slot) CK(cudaEventRecord(slot->ev_done, sa));
    }
    if (w->acc_after) w->acc_after();
}

template <class KB, class KE>
void project_finish(grootgpu_index* ix, Workspace* w, const uint32_t* d_off, uint32_t n, uint32_t H, int sms, cudaStream_t st, KB kbegin, KE kend, uint32_t& launches) {
    ProjectArgs pa = project_args(ix, w, d_off, n);
    const uint32_t* pk = peek(w, st, {w->item_off.as<uint32_t>() + (H - 1), w->item_cnt.as<uint32_t>() + (H - 1)}, 1);
    const uint64_t n_items = static_cast<uint64_t>(pk[0]) + pk[1];
    if (n_items == 0) { acc_enqueue(ix, w, nullptr, 0, sms, launches); return; }
    if (n_items >= (1ull << 31)) throw std::length_error("too many weight increments in one batch: use smaller batches");
    AccSlot* slot = &w->acc[w->acc_i ^= 1];
    if (slot->timed && cudaEventQuery(slot->ev_t1) == cudaSuccess) {   // the accumulate that used this slot two batches ago: report its time now
        float ms = 0;
        if (cudaEventElapsedTime(&ms, slot->ev_t0, slot->ev_t1) == cudaSuccess) { std::lock_guard<std::mutex> lock(ix->acc_mu); ix->acc_ms_pending += ms; }
        slot->timed = false;
    }
    w->pkeys.need(4 * n_items); w->pvals.need(8 * n_items);
    CK(cudaStreamWaitEvent(st, slot->ev_done, 0));                     // ... and it has to be done with the slot before the slot is rewritten
    if (4 * n_items > slot->keys.cap || 8 * n_items > slot->vals.cap) CK(cudaEventSynchronize(slot->ev_done));   // growing frees the old arrays
    slot->keys.need(4 * n_items); slot->vals.need(8 * n_items);
    pa.keys = w->pkeys.as<uint32_t>(); pa.vals = w->pvals.as<double>();
    const int eblocks = std::max(1, std::min<int>((H + 255) / 256, sms * 8));
    kbegin(6); project_expand_kernel<<<eblocks, 256, 0, st>>>(ix->d, pa); launches++; kend();
    int end_bit = 1;
    while ((1ull << end_bit) < ix->h.nodes.size()) end_bit++;
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, w->pkeys.as<uint32_t>(), slot->keys.as<uint32_t>(), w->pvals.as<double>(), slot->vals.as<double>(),
                                    static_cast<int>(n_items), 0, end_bit, st);
    w->cub_tmp2.need(sort_tmp + 16);
    cub::DeviceRadixSort::SortPairs(w->cub_tmp2.p, sort_tmp, w->pkeys.as<uint32_t>(), slot->keys.as<uint32_t>(), w->pvals.as<double>(), slot->vals.as<double>(),
                                    static_cast<int>(n_items), 0, end_bit, st);
    CK(cudaGetLastError());
    CK(cudaEventRecord(slot->ev_sorted, st));
    acc_enqueue(ix, w, slot, static_cast<uint32_t>(n_items), sms, launches);
}

// The batch pipeline on the device. d_seq / d_off already resident.
void run_batch(grootgpu_index* ix, Workspace* w, const uint8_t* d_seq, const uint32_t* d_off, uint32_t n, uint32_t min_len, uint32_t max_len,
               const grootgpu_align_params* prm, cudaStream_t st, grootgpu_batch_result* out) {
    const bool copy_back = out && !prm->results_on_device;
    const FlatIndex& h = ix->h;
    const uint32_t S = h.p.S, k = h.p.k;
    if (max_len > 60000) throw std::runtime_error("read longer than 60000 bases (documented limit)");
    prepare_params(ix, std::max(min_len, 1u), max_len, prm->containment_threshold);
    const int sms = g_num_sms(ix->device);
    uint32_t launches = 0;
    w->kev_n = 0;
    auto kbegin = [&](int cat) { if (w->kev_n < 32) { w->kev_cat[w->kev_n] = cat; cudaEventRecord(w->kev[2 * w->kev_n], st); } };
    auto kend = [&]() { if (w->kev_n < 32) { cudaEventRecord(w->kev[2 * w->kev_n + 1], st); w->kev_n++; } };
    cudaStream_t st2 = w->st_side;
    auto kbegin2 = [&](int cat) { if (w->kev_n < 32) { w->kev_cat[w->kev_n] = cat; cudaEventRecord(w->kev[2 * w->kev_n], st2); } };
    auto kend2 = [&]() { if (w->kev_n < 32) { cudaEventRecord(w->kev[2 * w->kev_n + 1], st2); w->kev_n++; } };
    const bool project = prm->project_on_device != 0;
    const bool compact = prm->compact_records != 0;
    const uint32_t recw = ix->rec_width;
    CK(cudaStreamSynchronize(st2));   // idle unless a previous batch failed between fork and join
    if (project) push_weights_to_device(ix);   // no-op once the weights live on the device

    w->n_hits.need(4ull * n); w->hit_off.need(4ull * (n + 1)); w->stage.need(4ull * HSTAGE * n);
    w->scalars.need(64); w->tile_counter.need(16); w->error.need(16);
    // scalars: [0]=n_segs (u32), counters as u64 at +8: [0]=mapped,[1]=multimapped,[2]=records
    w->qcount.need(64);
    zero_words(st, {{w->scalars.p, 16}, {w->tile_counter.p, 4}, {w->error.p, 4}, {w->qcount.p, 16}});
    uint32_t* d_nsegs = w->scalars.as<uint32_t>();
    unsigned long long* d_counters = reinterpret_cast<unsigned long long*>(w->scalars.as<uint8_t>() + 8);
    uint64_t* d_sk = nullptr;
    if (prm->keep_sketches) { w->sketches.need(8ull * S * n); d_sk = w->sketches.as<uint64_t>(); }

    // ---- K1+K2: sketch + probe + verify ----
    SeedArgs sa{};
    sa.seq = d_seq; sa.off = d_off; sa.n_reads = n; sa.max_len = max_len; sa.len_params = ix->len_params.as<LenParam>();
    sa.n_hits = w->n_hits.as<uint32_t>(); sa.stage = w->stage.as<uint32_t>(); sa.sketches_out = d_sk;
    sa.tile_counter = w->tile_counter.as<uint32_t>(); sa.error = w->error.as<int>();
    uint32_t tile_bytes = ((static_cast<uint32_t>(kTileReads) * max_len + 31u) & ~15u) + 16u;   // per warp, per buffer
    if (tile_bytes > 20 * 1024) tile_bytes = 0;  // very long reads: no staging, threads read global memory
    sa.tile_bytes = tile_bytes;
    const size_t seed_smem = sizeof(SeedTabs) + 64 + 2ull * tile_bytes * (kSeedThreads / 32);
    // the register-resident kernels exist for maxK == 4 and the sketch sizes of GROOT_S_LIST; anything else `groot index`
    // accepts (cmd/index.go:48-49) takes the run-time-S kernels
    int occ = h.p.max_k == 4 ? seed_occupancy_dispatch(S, seed_smem) : 0;
    const bool generic = occ <= 0;
    if (generic) occ = 8;
    const uint32_t n_tiles = (n + kTileReads - 1) / kTileReads;
    int seed_blocks = static_cast<int>(std::min<uint64_t>((n_tiles + kSeedThreads / 32 - 1) / (kSeedThreads / 32), static_cast<uint64_t>(sms) * occ));
    // two passes when the optimiser probes a single band for every read length of the batch (see SEED_PRESCREEN)
    bool two_pass = !generic && !prm->keep_sketches && getenv("GROOTGPU_SEED_ONEPASS") == nullptr;
    {
        std::lock_guard<std::mutex> lock(ix->params_mu);
        for (uint32_t len = std::max(min_len, k); len <= max_len && two_pass; len++) {
            const LenParam lp = ix->h_len_params[len];
            if (lp.eq_min <= S && lp.K != 0 && lp.L != 1) two_pass = false;
        }
    }
    // 2-bit copies of the seeded reads (both orientations) for the packed walk; reads longer than 1 024 bases go byte-wise
    uint32_t nw32 = 0;   // words of 16 bases per orientation: 8 (<= 128 bases) .. 64 (<= 1024)
    if (!prm->no_align && max_len <= 1024) { nw32 = 8; while (nw32 * 16u < max_len) nw32 *= 2; }
    if (nw32) { w->reads2.need(8ull * nw32 * n + 64); w->read_ok2.need(n); w->read_oh.need(16ull * n); }
    bool packed_by_seed = false;
    CK(cudaEventRecord(w->ev[0], st));
    if (two_pass) {
        w->seed_q.need(4ull * n);
        sa.queue = w->seed_q.as<uint32_t>(); sa.n_queue = w->qcount.as<uint32_t>() + 4;
        int pocc = 0;
        CK(cudaFuncSetAttribute(seed_kernel<4, 4, SEED_PRESCREEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(seed_smem)));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pocc, seed_kernel<4, 4, SEED_PRESCREEN>, kSeedThreads, seed_smem);
        const int pblocks = static_cast<int>(std::min<uint64_t>((n_tiles + kSeedThreads / 32 - 1) / (kSeedThreads / 32), static_cast<uint64_t>(sms) * std::max(pocc, 1)));
        kbegin(0); seed_kernel<4, 4, SEED_PRESCREEN><<<std::max(pblocks, 1), kSeedThreads, seed_smem, st>>>(ix->d, sa, make_mult(k)); launches++; kend();
        CK(cudaGetLastError());
        SeedArgs sq = sa;
        sq.tile_counter = w->tile_counter.as<uint32_t>() + 1;
        const uint32_t qstride = (((max_len + 7u) >> 2) | 1u) * 4u;                 // bytes per read slot: odd number of words
        sq.tile_bytes = qstride * kTileReads <= 20 * 1024 ? qstride * kTileReads : 0u;   // per warp; very long reads: straight from global
        const size_t qsmem = sizeof(SeedTabs) + 64 + 2ull * sq.tile_bytes * (kSeedThreads / 32);
        if (nw32 && sq.tile_bytes) {   // the queued pass holds every read it hashes in shared memory: it packs the seeded ones on the way
            sq.reads2 = w->reads2.as<uint32_t>(); sq.read_ok2 = w->read_ok2.as<uint8_t>(); sq.read_oh = w->read_oh.as<uint4>(); sq.nw32 = nw32;
            packed_by_seed = true;
        }
        kbegin(0); seed_queued_dispatch(S, ix->d, sq, k, qsmem, std::max(seed_blocks, 1), st); launches++; kend();
    } else if (generic) {
        const int gblocks = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n) + kSeedThreads - 1) / kSeedThreads, static_cast<uint64_t>(sms) * 8));
        kbegin(0); seed_generic_kernel<<<std::max(gblocks, 1), kSeedThreads, 0, st>>>(ix->d, sa); launches++; kend();
    } else {
        kbegin(0); seed_dispatch(S, ix->d, sa, k, seed_smem, std::max(seed_blocks, 1), st); launches++; kend();
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(w->ev[1], st));

    // ---- exclusive scan of per-read hit counts ----
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, w->n_hits.as<uint32_t>(), w->hit_off.as<uint32_t>(), static_cast<int>(n), st);
    w->cub_tmp.need(tmp_bytes + 16);
    cub::DeviceScan::ExclusiveSum(w->cub_tmp.p, tmp_bytes, w->n_hits.as<uint32_t>(), w->hit_off.as<uint32_t>(), static_cast<int>(n), st);
    uint32_t H = 0;
    {
        const uint32_t* pk = peek(w, st, {w->hit_off.as<uint32_t>() + (n - 1), w->n_hits.as<uint32_t>() + (n - 1),
                                           w->error.as<uint32_t>(), w->error.as<uint32_t>() + 1});
        const int err0 = static_cast<int>(pk[2]), err1 = static_cast<int>(pk[3]);
        if (err0 == GROOTGPU_ERR_SHORT_READ) throw std::invalid_argument("read " + std::to_string(err1) + " is shorter than k (the reference panics at boss.go:164-166)");
        if (err0 != 0) throw std::length_error("read " + std::to_string(err1) + " exceeds the declared maximum length");
        H = pk[0] + pk[1];
    }
    poke(st, {{w->hit_off.as<uint32_t>() + n, H}});

    uint32_t n_segs = 0;
    uint64_t R = 0;
    if (H > 0) {
        w->hits.need(4ull * H); w->hit_read.need(4ull * H); w->seg_flag.need(H); w->seg_begin.need(4ull * H);
        FillArgs fa{};
        fa.seq = d_seq; fa.off = d_off; fa.n_reads = n; fa.len_params = ix->len_params.as<LenParam>(); fa.n_hits = w->n_hits.as<uint32_t>();
        fa.hit_off = w->hit_off.as<uint32_t>(); fa.stage = w->stage.as<uint32_t>(); fa.hits = w->hits.as<uint32_t>();
        fa.hit_read = w->hit_read.as<uint32_t>(); fa.seg_flag = w->seg_flag.as<uint8_t>(); fa.counters = d_counters;
        fa.n_overflow = w->qcount.as<uint32_t>() + 5;   // zeroed with the other scalars at the start of the batch
        w->seed_q.need(4ull * n);                       // the prescreen's queue is free again: reused for the reads with more than HSTAGE hits
        fa.overflow_q = w->seed_q.as<uint32_t>();
        fa.reads2 = w->reads2.as<uint32_t>(); fa.read_ok2 = w->read_ok2.as<uint8_t>(); fa.read_oh = w->read_oh.as<uint4>(); fa.nw32 = nw32;
        int fill_blocks = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n) + kSeedThreads - 1) / kSeedThreads, static_cast<uint64_t>(sms) * 8));
        kbegin(1);
        if (generic) {
            fill_kernel<8, 4, false><<<fill_blocks, kSeedThreads, 0, st>>>(ix->d, fa, make_mult(k));   // the staged hits: no sketch involved
            fill_refill_generic_kernel<<<fill_blocks, kSeedThreads, 0, st>>>(ix->d, fa);
        } else fill_dispatch(S, ix->d, fa, k, fill_blocks, st);
        launches += 2;
        if (nw32 && !packed_by_seed) { pack_reads_kernel<<<std::max(1, std::min<int>((n + 31) / 32, sms * 8)), 256, 0, st>>>(fa); launches++; }
        kend();
        CK(cudaGetLastError());
        // ---- (read, graph) segment starts ----
        thrust::counting_iterator<uint32_t> counting(0);
        size_t sel_bytes = 0;
        cub::DeviceSelect::Flagged(nullptr, sel_bytes, counting, w->seg_flag.as<uint8_t>(), w->seg_begin.as<uint32_t>(), d_nsegs, static_cast<int>(H), st);
        w->cub_tmp.need(sel_bytes + 16);
        cub::DeviceSelect::Flagged(w->cub_tmp.p, sel_bytes, counting, w->seg_flag.as<uint8_t>(), w->seg_begin.as<uint32_t>(), d_nsegs, static_cast<int>(H), st);
        n_segs = peek(w, st, {d_nsegs})[0];

        // ---- K3: align (thread per pair) -> scan -> emit ----
        w->pairs.need(sizeof(PairOut) * static_cast<size_t>(n_segs)); w->seg_nrec.need(4ull * n_segs); w->seg_locus.need(8ull * n_segs);
        w->rec_off.need(4ull * (n_segs + 1)); w->seg_ntrav.need(4ull * n_segs);
        w->seg_mask.need(4ull * kTravWords * n_segs);
        const int vthreads = 128;
        int verify_blocks = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n_segs) + vthreads - 1) / vthreads, static_cast<uint64_t>(sms) * 8));
        verify_blocks = std::max(verify_blocks, 1);
        w->stack_ws.need(static_cast<size_t>(verify_blocks) * vthreads * (max_len + 2) * sizeof(DfsFrame));
        w->mask_ws.need(static_cast<size_t>(verify_blocks) * vthreads * (max_len + 2) * kMaskWordsInline * 4);
        AlignArgs aa{};
        aa.seq = d_seq; aa.off = d_off; aa.hits = w->hits.as<uint32_t>(); aa.hit_read = w->hit_read.as<uint32_t>();
        aa.seg_begin = w->seg_begin.as<uint32_t>(); aa.n_segs_ptr = d_nsegs; aa.n_hits_ptr = w->hit_off.as<uint32_t>() + n;
        aa.pairs = w->pairs.as<PairOut>(); aa.seg_nrec = w->seg_nrec.as<uint32_t>(); aa.seg_locus = w->seg_locus.as<uint2>();
        aa.seg_mask = w->seg_mask.as<uint32_t>(); aa.seg_ntrav = w->seg_ntrav.as<uint32_t>();
        aa.stack_ws = w->stack_ws.as<DfsFrame>(); aa.mask_ws = w->mask_ws.as<uint32_t>();
        aa.max_len = max_len; aa.no_align = prm->no_align; aa.error = w->error.as<int>();
        aa.counters = d_counters;
        aa.reads2 = w->reads2.as<uint32_t>(); aa.read_ok2 = w->read_ok2.as<uint8_t>(); aa.read_oh = w->read_oh.as<uint4>(); aa.nw32 = nw32;
        if (!prm->no_align) w->sdesc.need(sizeof(ScreenDesc) * static_cast<size_t>(n_segs));
        aa.sdesc = w->sdesc.as<ScreenDesc>();
        // screen/walk rounds over a shrinking, compacted queue; the queue counts stay on the device
        w->cursor.need(8ull * n_segs); w->cand.need(8ull * n_segs); w->queue_a.need(4ull * n_segs); w->queue_b.need(4ull * n_segs);
        uint32_t* qc = w->qcount.as<uint32_t>();
        CK(cudaEventRecord(w->ev[2], st));
        w->qkey.need(4ull * n_segs); w->qkey2.need(4ull * n_segs); w->order.need(4ull * n_segs); w->slow_q.need(4ull * n_segs);
        kbegin(2); align_init_kernel<<<std::max(1, std::min<int>((n_segs + 255) / 256, sms * 8)), 256, 0, st>>>(ix->d, aa, w->cursor.as<PairCursor>(), w->queue_b.as<uint32_t>(), w->qkey.as<uint32_t>(), qc); launches++; kend();
        CK(cudaGetLastError());
        {   // pair order by window id: the 32 pairs a warp walks together sit on the same graph region (same nodes, same
            // branch pattern), instead of 32 unrelated walks of very different lengths idling on each other. The order
            // is kept for the emit and weighting kernels.
            int wbits = 1;
            while ((1ull << wbits) < ix->h.wins.size()) wbits++;
            size_t qs = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, qs, w->qkey.as<uint32_t>(), w->qkey2.as<uint32_t>(), w->queue_b.as<uint32_t>(), w->order.as<uint32_t>(),
                                            static_cast<int>(n_segs), 0, wbits, st);
            w->cub_tmp.need(qs + 16);
            cub::DeviceRadixSort::SortPairs(w->cub_tmp.p, qs, w->qkey.as<uint32_t>(), w->qkey2.as<uint32_t>(), w->queue_b.as<uint32_t>(), w->order.as<uint32_t>(),
                                            static_cast<int>(n_segs), 0, wbits, st);
        }
        int screen_blocks = static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n_segs) + 7) / 8, static_cast<uint64_t>(sms) * 8));
        screen_blocks = std::max(screen_blocks, 1);
        const int kRounds = 2;   // later rounds hold a handful of stragglers: align_finish_kernel takes them in one launch
        RoundArgs ra{};
        ra.a = aa; ra.cursor = w->cursor.as<PairCursor>(); ra.cand = w->cand.as<uint2>();
        ra.slow_queue = w->slow_q.as<uint32_t>(); ra.n_slow = qc + 3;
        for (int round = 0; round < kRounds; round++) {
            uint32_t* qa = round == 0 ? w->order.as<uint32_t>() : (round & 1) ? w->queue_b.as<uint32_t>() : w->queue_a.as<uint32_t>();
            uint32_t* qb = (round & 1) ? w->queue_a.as<uint32_t>() : w->queue_b.as<uint32_t>();
            ra.queue = qa; ra.queue_next = qb; ra.n_queue = qc + (round & 1); ra.n_queue_next = qc + ((round + 1) & 1);
            kbegin(2); align_screen_kernel<<<screen_blocks, 256, 0, st>>>(ix->d, ra); launches++; kend();
            kbegin(3); align_walk_kernel<<<verify_blocks, vthreads, 0, st>>>(ix->d, ra); launches++; kend();
            CK(cudaGetLastError());
            poke(st, {{qc + (round & 1), 0u}});   // this round's count becomes the next round's "next"
        }
        {
            uint32_t* qa = (kRounds & 1) ? w->queue_b.as<uint32_t>() : w->queue_a.as<uint32_t>();
            ra.queue = qa; ra.queue_next = nullptr; ra.n_queue = qc + (kRounds & 1); ra.n_queue_next = nullptr;
            kbegin(4); align_finish_kernel<<<verify_blocks, vthreads, 0, st>>>(ix->d, ra); launches++; kend();
            ra.queue = w->slow_q.as<uint32_t>(); ra.n_queue = qc + 3;        // pairs the packed walk could not take
            kbegin(4); align_finish_kernel<<<verify_blocks, vthreads, 0, st>>>(ix->d, ra); launches++; kend();
            CK(cudaGetLastError());
        }
        CK(cudaEventRecord(w->ev[3], st));
        // ---- a10: the ordered graph weighting starts on the side stream; it needs the pairs, not their records ----
        if (project) {
            CK(cudaEventRecord(w->ev_fork, st));
            CK(cudaStreamWaitEvent(st2, w->ev_fork, 0));
            project_begin(ix, w, d_off, n, n_segs, H, sms, st2, kbegin2, kend2, launches);
        }
        // ---- scan record counts, emit ----
        size_t scan2 = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan2, w->seg_nrec.as<uint32_t>(), w->rec_off.as<uint32_t>(), static_cast<int>(n_segs), st);
        w->cub_tmp.need(scan2 + 16);
        cub::DeviceScan::ExclusiveSum(w->cub_tmp.p, scan2, w->seg_nrec.as<uint32_t>(), w->rec_off.as<uint32_t>(), static_cast<int>(n_segs), st);
        uint32_t lo = 0, lc = 0;
        {
            const uint32_t* pk = peek(w, st, {w->rec_off.as<uint32_t>() + (n_segs - 1), w->seg_nrec.as<uint32_t>() + (n_segs - 1),
                                               w->error.as<uint32_t>(), w->error.as<uint32_t>() + 1});
            lo = pk[0]; lc = pk[1];
            if (static_cast<int>(pk[2]) == GROOTGPU_ERR_BAD_BASE) throw std::domain_error("read " + std::to_string(static_cast<int>(pk[3])) + " holds a base > 'T' and had to be reverse complemented (the reference panics at seqio.go:122)");
        }
        R = static_cast<uint64_t>(lo) + lc;
        if (compact) { w->rec_c.need(static_cast<size_t>(recw) * std::max<uint64_t>(R, 1)); w->cpairs.need(sizeof(CPairOut) * static_cast<size_t>(n_segs)); }
        else { w->rec_path.need(4ull * std::max<uint64_t>(R, 1)); w->rec_pos.need(4ull * std::max<uint64_t>(R, 1)); }
        EmitArgs ea{};
        ea.seq = d_seq; ea.off = d_off; ea.n_segs_ptr = d_nsegs; ea.pairs = w->pairs.as<PairOut>(); ea.rec_off = w->rec_off.as<uint32_t>();
        ea.seg_locus = w->seg_locus.as<uint2>(); ea.seg_mask = w->seg_mask.as<uint32_t>(); ea.seg_ntrav = w->seg_ntrav.as<uint32_t>();
        ea.rec_path = w->rec_path.as<uint32_t>(); ea.rec_pos = w->rec_pos.as<int32_t>();
        ea.rec_c = w->rec_c.p; ea.cpairs = w->cpairs.as<CPairOut>();
        ea.stack_ws = w->stack_ws.as<DfsFrame>(); ea.mask_ws = w->mask_ws.as<uint32_t>(); ea.max_len = max_len;
        ea.reads2 = w->reads2.as<uint32_t>(); ea.read_ok2 = w->read_ok2.as<uint8_t>(); ea.nw32 = nw32;
        ea.multi_queue = w->queue_a.as<uint32_t>(); ea.n_multi = qc + 2;      // the align queues are free by now
        ea.order = prm->no_align ? nullptr : w->order.as<uint32_t>();
        {
            poke(st, {{qc + 2, 0u}});
            const int emit_blocks = std::max(1, static_cast<int>(std::min<uint64_t>((static_cast<uint64_t>(n_segs) + 31) / 32, static_cast<uint64_t>(sms) * 8)));   // 32 groups of 8 lanes per block
            const int recfmt = compact ? static_cast<int>(recw) : 0;   // record format of the emit kernels (align_kernels.cuh, RECW)
            kbegin(5);
            if (recfmt == 0) align_emit_kernel<0><<<emit_blocks, 256, 0, st>>>(ix->d, ea);
            else if (recfmt == 1) align_emit_kernel<1><<<emit_blocks, 256, 0, st>>>(ix->d, ea);
            else align_emit_kernel<2><<<emit_blocks, 256, 0, st>>>(ix->d, ea);
            launches++; kend();
            kbegin(5); align_emit_classify_kernel<<<std::max(1, std::min<int>((n_segs + 255) / 256, sms * 8)), 256, 0, st>>>(ea); launches++; kend();
            // thread stacks: stack_ws / mask_ws hold verify_blocks * vthreads of them
            kbegin(5);
            if (recfmt == 0) align_emit_multi_kernel<0><<<verify_blocks, vthreads, 0, st>>>(ix->d, ea);
            else if (recfmt == 1) align_emit_multi_kernel<1><<<verify_blocks, vthreads, 0, st>>>(ix->d, ea);
            else align_emit_multi_kernel<2><<<verify_blocks, vthreads, 0, st>>>(ix->d, ea);
            launches++; kend();
        }
        CK(cudaGetLastError());
        if (project) {
            project_finish(ix, w, d_off, n, H, sms, st2, kbegin2, kend2, launches);
            CK(cudaEventRecord(w->ev_join, st2));
            CK(cudaStreamWaitEvent(st, w->ev_join, 0));
        }
    } else {
        CK(cudaEventRecord(w->ev[2], st));
        CK(cudaEventRecord(w->ev[3], st));
        if (project) acc_enqueue(ix, w, nullptr, 0, sms, launches);   // nothing to weight, but the batch still takes its turn
    }
    CK(cudaEventRecord(w->ev[4], st));

    // ---- results ----
    unsigned long long counters[4] = {0, 0, 0, 0};
    if (copy_back && compact) {
        ix->r_cpairs.need(sizeof(CPairOut) * std::max<size_t>(n_segs, 1)); ix->r_rec_c.need(static_cast<size_t>(recw) * std::max<uint64_t>(R, 1));
        if (n_segs) CK(cudaMemcpyAsync(ix->r_cpairs.p, w->cpairs.p, sizeof(CPairOut) * static_cast<size_t>(n_segs), cudaMemcpyDeviceToHost, st));
        if (R) CK(cudaMemcpyAsync(ix->r_rec_c.p, w->rec_c.p, static_cast<size_t>(recw) * R, cudaMemcpyDeviceToHost, st));
    } else if (copy_back) {
        ix->r_hit_off.need(4ull * (n + 1)); ix->r_hits.need(4ull * std::max<uint32_t>(H, 1)); ix->r_pairs.need(sizeof(PairOut) * std::max<size_t>(n_segs, 1));
        ix->r_rec_path.need(4ull * std::max<uint64_t>(R, 1)); ix->r_rec_pos.need(4ull * std::max<uint64_t>(R, 1));
        CK(cudaMemcpyAsync(ix->r_hit_off.p, w->hit_off.p, 4ull * (n + 1), cudaMemcpyDeviceToHost, st));
        if (H) CK(cudaMemcpyAsync(ix->r_hits.p, w->hits.p, 4ull * H, cudaMemcpyDeviceToHost, st));
        if (n_segs) CK(cudaMemcpyAsync(ix->r_pairs.p, w->pairs.p, sizeof(PairOut) * static_cast<size_t>(n_segs), cudaMemcpyDeviceToHost, st));
        if (R) {
            CK(cudaMemcpyAsync(ix->r_rec_path.p, w->rec_path.p, 4ull * R, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(ix->r_rec_pos.p, w->rec_pos.p, 4ull * R, cudaMemcpyDeviceToHost, st));
        }
        if (prm->keep_sketches) { ix->r_sketches.need(8ull * S * n); CK(cudaMemcpyAsync(ix->r_sketches.p, w->sketches.p, 8ull * S * n, cudaMemcpyDeviceToHost, st)); }
    }
    CK(cudaEventRecord(w->ev[5], st));
    {
        const uint32_t* c32 = reinterpret_cast<const uint32_t*>(d_counters);
        const uint32_t* pk = peek(w, st, {c32, c32 + 1, c32 + 2, c32 + 3, c32 + 4, c32 + 5, c32 + 6, c32 + 7});   // synchronises the stream
        for (int i = 0; i < 4; i++) counters[i] = static_cast<unsigned long long>(pk[2 * i]) | (static_cast<unsigned long long>(pk[2 * i + 1]) << 32);
    }
    if (out) {
        memset(out, 0, sizeof *out);
        out->n_reads = n; out->n_hits = H; out->n_pairs = n_segs; out->n_records = R;
        if (copy_back && compact) {
            out->cpairs = reinterpret_cast<const grootgpu_cpair*>(ix->r_cpairs.p); out->rec_path_c = ix->r_rec_c.p;
        } else if (copy_back) {
            out->hit_off = ix->r_hit_off.as<uint32_t>(); out->hits = ix->r_hits.as<uint32_t>();
            out->pairs = reinterpret_cast<const grootgpu_pair*>(ix->r_pairs.p);
            out->rec_path = ix->r_rec_path.as<uint32_t>(); out->rec_pos = ix->r_rec_pos.as<int32_t>();
        }
        if (copy_back) out->sketches = prm->keep_sketches ? ix->r_sketches.as<uint64_t>() : nullptr;
        out->rec_path_bytes = compact ? recw : 0u;
        out->result_set = static_cast<uint32_t>(ix->call_parity);
        out->received = n; out->mapped = counters[0]; out->multimapped = counters[1]; out->alignments = R;
        float seed_ms = 0, align_ms = 0, dev_ms = 0, all_ms = 0;
        cudaEventElapsedTime(&seed_ms, w->ev[0], w->ev[1]);
        cudaEventElapsedTime(&align_ms, w->ev[2], w->ev[3]);
        cudaEventElapsedTime(&dev_ms, w->ev[0], w->ev[4]);
        cudaEventElapsedTime(&all_ms, w->ev[0], w->ev[5]);
        out->ms[0] = all_ms; out->ms[1] = seed_ms; out->ms[2] = align_ms; out->ms[3] = dev_ms - seed_ms - align_ms;
        out->kernel_launches = launches;
        for (int i = 0; i < w->kev_n; i++) { float ms = 0; cudaEventElapsedTime(&ms, w->kev[2 * i], w->kev[2 * i + 1]); out->kernel_ms[w->kev_cat[i]] += ms; }
        out->slow_path_pairs = counters[3];
        out->d_hit_off = w->hit_off.as<uint32_t>(); out->d_hits = w->hits.as<uint32_t>();
        out->d_pairs = reinterpret_cast<const grootgpu_pair*>(w->pairs.p);
        if (compact) { out->d_cpairs = reinterpret_cast<const grootgpu_cpair*>(w->cpairs.p); out->d_rec_path_c = w->rec_c.p; }
        else { out->d_rec_path = w->rec_path.as<uint32_t>(); out->d_rec_pos = w->rec_pos.as<int32_t>(); }
        { std::lock_guard<std::mutex> lock(ix->acc_mu); out->kernel_ms[7] = ix->acc_ms_pending; ix->acc_ms_pending = 0.f; }
    }
}


// ---- chunked host path --------------------------------------------------------------------------------------
// u64 host offsets of one chunk -> u32 offsets relative to the chunk's first base, plus min / max read length
__global__ void __launch_bounds__(256) chunk_offsets_kernel(const uint64_t* __restrict__ off64, uint32_t n, uint32_t* __restrict__ off32,
                                                            uint32_t* __restrict__ minmax) {
    const uint64_t base = off64[0];
    uint32_t mn = 0xffffffffu, mx = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        const uint64_t o = off64[i];
        off32[i] = static_cast<uint32_t>(o - base);
        if (i < n) {
            const uint64_t l = off64[i + 1] - o;
            const uint32_t l32 = l > 0xffffffffull ? 0xffffffffu : static_cast<uint32_t>(l);
            mn = min(mn, l32); mx = max(mx, l32);
        }
    }
    mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
    if ((threadIdx.x & 31) == 0) { atomicMin(&minmax[0], mn); atomicMax(&minmax[1], mx); }
}

__global__ void __launch_bounds__(256) chunk_offsets_fixed_kernel(uint32_t* __restrict__ off32, uint32_t n, uint32_t len) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) off32[i] = i * len;
}

// chunk-local indices -> batch-global ones, before the chunk's arrays are copied to their place in the host result
__global__ void __launch_bounds__(256) chunk_rebase_kernel(PairOut* __restrict__ pairs, uint32_t n_pairs, uint32_t* __restrict__ hit_off, uint32_t n_off,
                                                           uint32_t read_base, uint32_t hit_base, uint32_t rec_base) {
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t i = i0; i < n_pairs; i += stride) {
        pairs[i].read += read_base; pairs[i].hit_begin += hit_base; pairs[i].rec_begin += rec_base;
    }
    for (uint32_t i = i0; i < n_off; i += stride) hit_off[i] += hit_base;
}

// grow a pinned result buffer to `need` bytes keeping its first `used` bytes (pending copies into it are drained first)
void grow_keep(HBuf& b, size_t need, size_t used, cudaStream_t drain) {
    if (need <= b.cap) return;
    CK(cudaStreamSynchronize(drain));
    HBuf nb;
    nb.need(need);
    if (used) memcpy(nb.p, b.p, used);
    std::swap(b.p, nb.p); std::swap(b.cap, nb.cap);
}

void grow_keep(DBuf& b, size_t need, size_t used, cudaStream_t drain) {   // the same for a batch-wide array in device memory
    if (need <= b.cap) return;
    CK(cudaStreamSynchronize(drain));
    DBuf nb;
    nb.need(need);
    if (used) CK(cudaMemcpy(nb.p, b.p, used, cudaMemcpyDeviceToDevice));
    std::swap(b.p, nb.p); std::swap(b.cap, nb.cap);
}

__global__ void __launch_bounds__(256) chunk_rebase_compact_kernel(CPairOut* __restrict__ cpairs, uint32_t n_pairs, uint32_t read_base) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) cpairs[i].read += read_base;
}

uint32_t chunk_reads_setting() {   // reads per pipeline chunk; GROOTGPU_CHUNK_READS overrides (tests force many small chunks)
    const char* e = getenv("GROOTGPU_CHUNK_READS");
    const long x = e ? atol(e) : 0;
    return static_cast<uint32_t>(x > 0 ? x : 2000000l);
}

// Host buffers in, host results out, as a pipeline over chunks of reads on kLanes LANES (workspaces with their own
// streams, each driven by its own host thread): chunk c runs on lane c % kLanes while
//   * the bases and offsets of the next chunks are copied in (st_in, up to two chunks ahead),
//   * the result arrays of finished chunks are copied out (st_out) straight to their final place in the batch-wide
//     host arrays, after a small kernel has rebased the chunk-local indices,
//   * the other lanes run the neighbouring chunks: their kernels fill the GPU during this chunk's latency-bound tails
//     (stragglers of the walk, the ordered f64 chains) and during the host round trips that size its buffers.
// What stays ordered: chunk c's graph weighting chain runs after chunk c-1's (an event between the lanes' side
// streams), and the batch-wide offsets of chunk c are fixed once chunk c-1 has published its totals — so every
// output word and every f64 weight equals the single-batch result.
struct ChunkShared {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<uint64_t> hit_end, pair_end, rec_end;   // inclusive prefix sums, valid for chunks < published
    uint32_t published = 0, inputs_issued = 0;
    uint32_t acc_recorded = 0;   // chunks whose "weighting done" event has been recorded (or skipped), in chunk order
    bool failed = false;
    std::exception_ptr error;
    grootgpu_batch_result total{};
};

void run_batch_chunked(grootgpu_index* ix, const uint8_t* seq, const uint64_t* seq_off, uint32_t n, const grootgpu_align_params* prm,
                       grootgpu_batch_result* out) {
    cudaStream_t st_in = ix->st_in, st_out = ix->st_out;
    const uint32_t S = ix->h.p.S;
    // reads of one length, back to back (params->fixed_read_len): offsets are arithmetic, none travel to the device
    const uint32_t fixed_len = prm->fixed_read_len;
    if (!seq_off && fixed_len == 0) throw std::runtime_error("seq_off is NULL and params->fixed_read_len is 0");
    if (fixed_len != 0 && fixed_len < ix->h.p.k) throw std::invalid_argument("a read is shorter than k (the reference panics at boss.go:164-166)");
    if (fixed_len != 0 && seq_off)
        for (uint32_t i : {0u, n / 2, n - 1}) if (seq_off[i + 1] - seq_off[i] != fixed_len || seq_off[i] - seq_off[0] != static_cast<uint64_t>(i) * fixed_len)
            throw std::runtime_error("params->fixed_read_len does not match seq_off");
    const uint64_t off_base = fixed_len != 0 && seq_off ? seq_off[0] : 0;
    auto off_at = [&](uint32_t i) -> uint64_t { return fixed_len != 0 ? off_base + static_cast<uint64_t>(i) * fixed_len : seq_off[i]; };
    // chunk boundaries (always < 4 GiB of bases per chunk). The copy-in link is only ~1.25x faster than the kernels, so
    // the schedule starts with a small chunk (the first copy-in is a transfer nothing overlaps) and grows up to
    // chunk_reads_setting() — the next chunk's bases then arrive before the current one is done. The end depends on what
    // is copied out: the full result arrays are ~90 bytes per read, so the batch ends with a small chunk (the last
    // copy-out is the other exposed transfer); the compact output and device-resident results have no copy-out to speak
    // of, so the last TWO chunks — one per lane, running side by side — are made equal and the lanes finish together.
    std::vector<uint32_t> cb{0};
    {
        const bool light_out = prm->compact_records != 0 || prm->results_on_device != 0;
        const uint32_t target = chunk_reads_setting(), edge = std::max(1u, target / 4);
        const uint64_t max_bytes = (1ull << 32) - 4096;
        uint32_t want = std::max(1u, light_out ? target / 4 : target / 2);
        while (cb.back() < n) {
            const uint32_t r0 = cb.back(), left = n - r0;
            uint32_t take = std::min(want, left);
            if (light_out) { if (left > want && left <= 2ull * want) take = (left + 1) / 2; }              // two last chunks of equal size
            else if (left > edge && left - take < edge) take = left - edge;                             // leave a last chunk of `edge` reads
            uint32_t r1 = r0 + std::max(1u, take);
            while (r1 > r0 + 1 && off_at(r1) - off_at(r0) >= max_bytes) r1 = r0 + (r1 - r0) / 2;
            if (off_at(r1) - off_at(r0) >= max_bytes) throw std::length_error("a read of 4 GiB or more");
            cb.push_back(r1);
            want = std::min<uint64_t>(target, static_cast<uint64_t>(want) + (light_out ? want / 2 : want / 4) + 1);
        }
    }
    const uint32_t C = static_cast<uint32_t>(cb.size()) - 1;
    struct Drain {   // nothing may still be copying from / into caller or handle memory when we leave, also on errors
        grootgpu_index* ix;
        ~Drain() {
            cudaStreamSynchronize(ix->st_in);
            for (Workspace& w : ix->ws) { cudaStreamSynchronize(w.stream); cudaStreamSynchronize(w.st_side); }
            cudaStreamSynchronize(ix->st_out);
        }
    } drain{ix};
    const bool on_device = prm->results_on_device != 0, compact = prm->compact_records != 0;
    const int parity = ix->call_parity;
    grootgpu_index::DevResult& dres = ix->bres[parity];
    if (on_device) { if (!compact) dres.hit_off.need(4ull * (static_cast<size_t>(n) + 1)); }
    else if (!compact) ix->r_hit_off.need(4ull * (static_cast<size_t>(n) + 1));
    if (on_device && prm->keep_sketches) throw std::runtime_error("keep_sketches needs host results");
    if (prm->keep_sketches) ix->r_sketches.need(8ull * S * n);
    if (prm->project_on_device) push_weights_to_device(ix);
    if (on_device && ix->comm) CK(cudaStreamWaitEvent(st_out, ix->comm->ev_sent[parity], 0));   // the gather of two calls ago has read this result set
    grootgpu_align_params cprm = *prm;
    cprm.results_on_device = 1;
    ChunkShared sh;
    sh.hit_end.assign(C, 0); sh.pair_end.assign(C, 0); sh.rec_end.assign(C, 0);
    const bool trace = getenv("GROOTGPU_TRACE") != nullptr;   // host-side timeline of the chunk pipeline on stderr
    const auto t_begin = std::chrono::steady_clock::now();
    auto now_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    const int device = ix->device;

    // copy-in of chunk c into input buffer c % kInBufs; issued in chunk order, at most kLanes chunks ahead of the chunk that
    // is starting (that buffer was last read by chunk c - 2 * kLanes, which completed before chunk c - kLanes — same lane — could start)
    auto issue_inputs_upto = [&](uint32_t last) {   // sh.mu held
        for (; sh.inputs_issued <= last && sh.inputs_issued < C; sh.inputs_issued++) {
            const uint32_t c = sh.inputs_issued, r0 = cb[c], nc = cb[c + 1] - r0, b = c % grootgpu_index::kInBufs;
            const uint64_t bytes = off_at(cb[c + 1]) - off_at(r0);
            ix->in_seq[b].need(bytes + 64); ix->in_off32[b].need(4ull * (nc + 1));
            CK(cudaMemcpyAsync(ix->in_seq[b].p, seq + off_at(r0), bytes, cudaMemcpyHostToDevice, st_in));
            CK(cudaMemsetAsync(ix->in_seq[b].as<uint8_t>() + bytes, 0, 64, st_in));
            if (fixed_len == 0) {
                ix->in_off64[b].need(8ull * (nc + 1));
                CK(cudaMemcpyAsync(ix->in_off64[b].p, seq_off + r0, 8ull * (nc + 1), cudaMemcpyHostToDevice, st_in));
            }
            CK(cudaEventRecord(ix->ev_in[b], st_in));
        }
    };

    auto lane_main = [&](uint32_t lane) {
        Workspace* w = &ix->ws[lane];
        cudaStream_t st = w->stream;
        try {
            pick_device(device);
            for (uint32_t c = lane; c < C; c += grootgpu_index::kLanes) {
                const uint32_t r0 = cb[c], nc = cb[c + 1] - r0, b = c % grootgpu_index::kInBufs;
                const double t_c0 = now_ms();
                {
                    std::unique_lock<std::mutex> lk(sh.mu);
                    if (sh.failed) return;
                    issue_inputs_upto(c + grootgpu_index::kLanes);
                }
                CK(cudaStreamWaitEvent(st, ix->ev_in[b], 0));
                uint32_t mm[2] = {fixed_len, fixed_len};
                if (fixed_len != 0) {   // offsets generated in place: no copy, no round trip for the length range
                    chunk_offsets_fixed_kernel<<<std::max(1u, std::min<uint32_t>((nc + 256) / 256, 1184u)), 256, 0, st>>>(ix->in_off32[b].as<uint32_t>(), nc, fixed_len);
                } else {
                    poke(st, {{w->len_minmax.as<uint32_t>(), 0xffffffffu}, {w->len_minmax.as<uint32_t>() + 1, 0u}});
                    chunk_offsets_kernel<<<std::max(1u, std::min<uint32_t>((nc + 256) / 256, 1184u)), 256, 0, st>>>(ix->in_off64[b].as<uint64_t>(), nc, ix->in_off32[b].as<uint32_t>(),
                                                                                                                 w->len_minmax.as<uint32_t>());
                    const uint32_t* pk = peek(w, st, {w->len_minmax.as<uint32_t>(), w->len_minmax.as<uint32_t>() + 1});
                    mm[0] = pk[0]; mm[1] = pk[1];
                }
                if (mm[0] < ix->h.p.k) throw std::invalid_argument("a read is shorter than k (the reference panics at boss.go:164-166)");
                w->swap_result_sets();                                   // write the result set the lane's previous chunk is NOT being copied out of
                CK(cudaStreamWaitEvent(st, w->ev_out[w->rset], 0));      // ... whose own copy-out (two chunks of this lane ago) has completed
                // chunk c's f64 chains go after chunk c-1's: both lanes enqueue them on the one accumulate stream, in chunk
                // order — the host thread of chunk c waits until chunk c-1's have been enqueued (or skipped)
                bool acc_passed = false;
                w->ring_first = c == 0; w->ring_last = c + 1 == C;
                if (cprm.project_on_device) {
                    w->acc_before = [&, c]() {
                        std::unique_lock<std::mutex> lk(sh.mu);
                        sh.cv.wait(lk, [&] { return sh.failed || sh.acc_recorded >= c; });
                        if (sh.failed) throw std::runtime_error("the other lane of the batch failed");
                    };
                    w->acc_after = [&, c]() {
                        { std::unique_lock<std::mutex> lk(sh.mu); sh.acc_recorded = c + 1; }
                        sh.cv.notify_all();
                        acc_passed = true;
                    };
                }
                grootgpu_batch_result r{};
                const double t_c1 = now_ms();
                run_batch(ix, w, ix->in_seq[b].as<uint8_t>(), ix->in_off32[b].as<uint32_t>(), nc, mm[0], mm[1], &cprm, st, &r);
                const double t_c2 = now_ms();
                if (cprm.project_on_device && !acc_passed) { uint32_t l = 0; acc_enqueue(ix, w, nullptr, 0, 148, l); }   // defensive: run_batch always takes the chunk's turn
                w->acc_before = nullptr; w->acc_after = nullptr;
                // batch-wide offsets: chunk c-1 has to have published its totals
                uint64_t hit_base = 0, pair_base = 0, rec_base = 0;
                {
                    std::unique_lock<std::mutex> lk(sh.mu);
                    sh.cv.wait(lk, [&] { return sh.failed || sh.published >= c; });
                    if (sh.failed) return;
                    if (c > 0) { hit_base = sh.hit_end[c - 1]; pair_base = sh.pair_end[c - 1]; rec_base = sh.rec_end[c - 1]; }
                    if (hit_base + r.n_hits >= (1ull << 32) || rec_base + r.n_records >= (1ull << 32))
                        throw std::length_error("more than 2^32 hits or records in one batch: use smaller batches");
                    sh.hit_end[c] = hit_base + r.n_hits; sh.pair_end[c] = pair_base + r.n_pairs; sh.rec_end[c] = rec_base + r.n_records;
                    // the batch-wide arrays (pinned host memory, or device memory when the results stay there) grow by
                    // extrapolating the yield so far; copies into them are drained first (grow_keep)
                    const double scale = 1.15 * static_cast<double>(n) / static_cast<double>(cb[c + 1]);
                    auto want = [&](uint64_t used_after, size_t elem) { return static_cast<size_t>(static_cast<double>(used_after) * scale) * elem + 4096; };
                    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
                    auto place = [&](auto& host_buf, auto& dev_buf, uint64_t base, uint64_t count, size_t elem, const void* src) {   // chunk array -> its place in the batch-wide array
                        const uint64_t end_elems = base + count;
                        if (on_device) { if (end_elems * elem > dev_buf.cap) grow_keep(dev_buf, want(end_elems, elem), base * elem, st_out); }
                        else if (end_elems * elem > host_buf.cap) grow_keep(host_buf, want(end_elems, elem), base * elem, st_out);
                        uint8_t* dst = (on_device ? static_cast<uint8_t*>(dev_buf.p) : static_cast<uint8_t*>(host_buf.p)) + base * elem;
                        if (count) CK(cudaMemcpyAsync(dst, src, count * elem, kind, st_out));
                    };
                    const uint32_t n_off = nc + (c + 1 == C ? 1u : 0u);
                    const uint32_t np = static_cast<uint32_t>(r.n_pairs);
                    if (compact) chunk_rebase_compact_kernel<<<std::max(1u, std::min<uint32_t>((np + 255) / 256, 1184u)), 256, 0, st>>>(w->cpairs.as<CPairOut>(), np, r0);
                    else chunk_rebase_kernel<<<std::max(1u, std::min<uint32_t>((std::max(np, n_off) + 255) / 256, 1184u)), 256, 0, st>>>(
                        w->pairs.as<PairOut>(), np, w->hit_off.as<uint32_t>(), n_off, r0, static_cast<uint32_t>(hit_base), static_cast<uint32_t>(rec_base));
                    CK(cudaGetLastError());
                    CK(cudaEventRecord(w->ev_done, st));
                    // copy-out, straight to the final place (st_out is shared: enqueue under the lock)
                    CK(cudaStreamWaitEvent(st_out, w->ev_done, 0));
                    if (compact) {
                        place(ix->r_cpairs, dres.cpairs, pair_base, r.n_pairs, sizeof(CPairOut), w->cpairs.p);
                        place(ix->r_rec_c, dres.rec_c, rec_base, r.n_records, ix->rec_width, w->rec_c.p);
                    } else {
                        place(ix->r_hit_off, dres.hit_off, r0, n_off, 4, w->hit_off.p);
                        place(ix->r_hits, dres.hits, hit_base, r.n_hits, 4, w->hits.p);
                        place(ix->r_pairs, dres.pairs, pair_base, r.n_pairs, sizeof(PairOut), w->pairs.p);
                        place(ix->r_rec_path, dres.rec_path, rec_base, r.n_records, 4, w->rec_path.p);
                        place(ix->r_rec_pos, dres.rec_pos, rec_base, r.n_records, 4, w->rec_pos.p);
                    }
                    if (prm->keep_sketches) CK(cudaMemcpyAsync(ix->r_sketches.as<uint64_t>() + static_cast<size_t>(r0) * S, w->sketches.p, 8ull * S * nc, cudaMemcpyDeviceToHost, st_out));
                    CK(cudaEventRecord(w->ev_out[w->rset], st_out));
                    sh.total.mapped += r.mapped; sh.total.multimapped += r.multimapped; sh.total.slow_path_pairs += r.slow_path_pairs;
                    sh.total.kernel_launches += r.kernel_launches + 2;
                    for (int i = 1; i < 4; i++) sh.total.ms[i] += r.ms[i];
                    for (int i = 0; i < 8; i++) sh.total.kernel_ms[i] += r.kernel_ms[i];
                    sh.published = c + 1;
                    if (trace) fprintf(stderr, "[grootgpu] chunk %u (lane %u): %u reads  start %.2f ms  launched %.2f  kernels done %.2f (device %.2f ms)  copy-out issued %.2f\n",
                                       c, lane, nc, t_c0, t_c1, t_c2, r.ms[1] + r.ms[2] + r.ms[3], now_ms());
                }
                sh.cv.notify_all();
            }
        } catch (...) {
            std::unique_lock<std::mutex> lk(sh.mu);
            if (!sh.failed) { sh.failed = true; sh.error = std::current_exception(); }
            lk.unlock();
            sh.cv.notify_all();
        }
    };

    CK(cudaEventRecord(ix->ev_t0, st_in));
    for (Workspace& w : ix->ws) { w.len_minmax.need(16); w.acc_before = nullptr; w.acc_after = nullptr; }
    {
        std::vector<std::thread> helpers;
        for (uint32_t l = 1; l < grootgpu_index::kLanes && l < C; l++) helpers.emplace_back(lane_main, l);
        lane_main(0u);
        for (auto& t : helpers) t.join();
    }
    for (Workspace& w : ix->ws) { w.acc_before = nullptr; w.acc_after = nullptr; }
    if (sh.failed) std::rethrow_exception(sh.error);
    CK(cudaEventRecord(ix->ev_t1, st_out));
    CK(cudaStreamSynchronize(st_out));
    if (trace) fprintf(stderr, "[grootgpu] batch of %u reads in %u chunks done at %.2f ms\n", n, C, now_ms());
    memset(out, 0, sizeof *out);
    *out = sh.total;
    out->n_reads = n; out->n_hits = sh.hit_end[C - 1]; out->n_pairs = sh.pair_end[C - 1]; out->n_records = sh.rec_end[C - 1];
    out->rec_path_bytes = compact ? ix->rec_width : 0u;
    out->result_set = static_cast<uint32_t>(parity);
    if (on_device) {
        if (compact) { out->d_cpairs = reinterpret_cast<const grootgpu_cpair*>(dres.cpairs.p); out->d_rec_path_c = dres.rec_c.p; }
        else {
            out->d_hit_off = dres.hit_off.as<uint32_t>(); out->d_hits = dres.hits.as<uint32_t>(); out->d_pairs = reinterpret_cast<const grootgpu_pair*>(dres.pairs.p);
            out->d_rec_path = dres.rec_path.as<uint32_t>(); out->d_rec_pos = dres.rec_pos.as<int32_t>();
        }
    } else if (compact) {
        out->cpairs = reinterpret_cast<const grootgpu_cpair*>(ix->r_cpairs.p); out->rec_path_c = ix->r_rec_c.p;
    } else {
        out->hit_off = ix->r_hit_off.as<uint32_t>(); out->hits = ix->r_hits.as<uint32_t>();
        out->pairs = reinterpret_cast<const grootgpu_pair*>(ix->r_pairs.p);
        out->rec_path = ix->r_rec_path.as<uint32_t>(); out->rec_pos = ix->r_rec_pos.as<int32_t>();
    }
    out->sketches = prm->keep_sketches ? ix->r_sketches.as<uint64_t>() : nullptr;
    out->received = n; out->alignments = out->n_records;
    cudaEventElapsedTime(&out->ms[0], ix->ev_t0, ix->ev_t1);
}

// ---- multi-GPU gather ------------------------------------------------------------------------------------------
// One exchange per batch: the per-rank result arrays travel to rank 0 over NVLink (grouped ncclSend / ncclRecv straight
// into their place in the merged arrays — contiguous shards in rank order, so merging is concatenation), a small kernel
// per rank turns shard-local indices into batch-wide ones. Everything runs on the communicator's own stream: the
// transfer of batch b overlaps the mapping of batch b + 1, whose result arrays are the other set.
void gather_results(grootgpu_comm* c, const grootgpu_batch_result* loc, int to_host, grootgpu_batch_result* merged) {
    grootgpu_index* ix = c->ix;
    const NcclApi& N = nccl();
    cudaStream_t sg = c->st_gather;
    const bool compact = loc->rec_path_bytes != 0;
    const int W = c->world, parity = static_cast<int>(loc->result_set & 1u);
    const uint32_t recw = ix->rec_width;
    if (loc->n_pairs && !(compact ? static_cast<const void*>(loc->d_cpairs) : static_cast<const void*>(loc->d_pairs)))
        throw std::runtime_error("grootgpu_gather needs a result with device pointers (results_on_device = 1)");
    const bool trace = getenv("GROOTGPU_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto now_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    // 1. everybody's sizes (one small all-gather; the host needs them to lay the merged arrays out)
    uint64_t* hc = c->h_counts;
    hc[0] = loc->n_reads; hc[1] = loc->n_hits; hc[2] = loc->n_pairs; hc[3] = loc->n_records;
    hc[4] = loc->mapped; hc[5] = loc->multimapped; hc[6] = loc->slow_path_pairs; hc[7] = compact ? recw : 0u;
    uint64_t* dc = c->d_counts.as<uint64_t>();
    CK(cudaMemcpyAsync(dc, hc, 8ull * kCountWords, cudaMemcpyHostToDevice, sg));
    NK(N.AllGather(dc, dc + kCountWords, kCountWords, ncclUint64, c->gath, sg));
    CK(cudaMemcpyAsync(hc + kCountWords, dc + kCountWords, 8ull * kCountWords * W, cudaMemcpyDeviceToHost, sg));
    CK(cudaStreamSynchronize(sg));   // also: the previous gather (it had the whole mapping of this batch to finish) is done
    const double t_counts = now_ms();
    const uint64_t* all = hc + kCountWords;
    std::vector<uint64_t> rb(W + 1, 0), hb(W + 1, 0), pb(W + 1, 0), cb(W + 1, 0);
    uint64_t mapped = 0, multi = 0, slow = 0;
    for (int r = 0; r < W; r++) {
        const uint64_t* a = all + static_cast<size_t>(r) * kCountWords;
        if (a[7] != hc[7]) throw std::runtime_error("grootgpu_gather: the ranks use different output formats");
        rb[r + 1] = rb[r] + a[0]; hb[r + 1] = hb[r] + a[1]; pb[r + 1] = pb[r] + a[2]; cb[r + 1] = cb[r] + a[3];
        mapped += a[4]; multi += a[5]; slow += a[6];
    }
    if (rb[W] >= (1ull << 32) || hb[W] >= (1ull << 32) || cb[W] >= (1ull << 32) || pb[W] >= (1ull << 32))
        throw std::length_error("more than 2^32 reads, hits or records in one merged batch: use smaller batches");
    auto bytes_of = [&](int sec, int r) -> size_t {   // sections: full = hit_off, hits, pairs, rec_path, rec_pos; compact = cpairs, rec_c
        const uint64_t* a = all + static_cast<size_t>(r) * kCountWords;
        if (compact) return sec == 0 ? a[2] * sizeof(CPairOut) : a[3] * recw;
        switch (sec) { case 0: return a[0] * 4; case 1: return a[1] * 4; case 2: return a[2] * sizeof(PairOut); default: return a[3] * 4; }
    };
    const int n_sec = compact ? 2 : 5;
    const void* src[5];
    if (compact) { src[0] = loc->d_cpairs; src[1] = loc->d_rec_path_c; }
    else { src[0] = loc->d_hit_off; src[1] = loc->d_hits; src[2] = loc->d_pairs; src[3] = loc->d_rec_path; src[4] = loc->d_rec_pos; }
    if (c->rank != 0) {
        NK(N.GroupStart());
        for (int sec = 0; sec < n_sec; sec++) { const size_t b = bytes_of(sec, c->rank); if (b) NK(N.Send(src[sec], b, ncclUint8, 0, c->gath, sg)); }
        NK(N.GroupEnd());
        CK(cudaEventRecord(c->ev_sent[parity], sg));
        if (merged) memset(merged, 0, sizeof *merged);
        if (trace) fprintf(stderr, "[grootgpu] gather rank %d: sizes known at %.2f ms, sends enqueued at %.2f ms\n", c->rank, t_counts, now_ms());
        return;
    }
    // 2. rank 0: receive every shard at its place
    DBuf* mb[5];
    HBuf* hbuf[5];
    size_t elem[5], base_of[5][2];   // element size; per section: which prefix array gives the base
    (void)base_of;
    grootgpu_comm::HostSet& hs = c->hs[c->hs_i ^= 1];
    if (compact) {
        mb[0] = &c->m_cpairs; mb[1] = &c->m_rec_c; hbuf[0] = &hs.cpairs; hbuf[1] = &hs.rec_c; elem[0] = sizeof(CPairOut); elem[1] = recw;
        c->m_cpairs.need(std::max<size_t>(16, pb[W] * sizeof(CPairOut))); c->m_rec_c.need(std::max<size_t>(16, cb[W] * recw));
    } else {
        mb[0] = &c->m_hit_off; mb[1] = &c->m_hits; mb[2] = &c->m_pairs; mb[3] = &c->m_rec_path; mb[4] = &c->m_rec_pos;
        hbuf[0] = &hs.hit_off; hbuf[1] = &hs.hits; hbuf[2] = &hs.pairs; hbuf[3] = &hs.rec_path; hbuf[4] = &hs.rec_pos;
        elem[0] = 4; elem[1] = 4; elem[2] = sizeof(PairOut); elem[3] = 4; elem[4] = 4;
        c->m_hit_off.need(4 * (rb[W] + 1)); c->m_hits.need(std::max<size_t>(16, 4 * hb[W])); c->m_pairs.need(std::max<size_t>(32, pb[W] * sizeof(PairOut)));
        c->m_rec_path.need(std::max<size_t>(16, 4 * cb[W])); c->m_rec_pos.need(std::max<size_t>(16, 4 * cb[W]));
    }
    auto sec_base = [&](int sec, int r) -> uint64_t {   // element offset of rank r's shard inside merged section `sec`
        if (compact) return sec == 0 ? pb[r] : cb[r];
        switch (sec) { case 0: return rb[r]; case 1: return hb[r]; case 2: return pb[r]; default: return cb[r]; }
    };
    NK(N.GroupStart());
    for (int r = 1; r < W; r++)
        for (int sec = 0; sec < n_sec; sec++) {
            const size_t b = bytes_of(sec, r);
            if (b) NK(N.Recv(mb[sec]->as<uint8_t>() + sec_base(sec, r) * elem[sec], b, ncclUint8, r, c->gath, sg));
        }
    NK(N.GroupEnd());
    for (int sec = 0; sec < n_sec; sec++) { const size_t b = bytes_of(sec, 0); if (b) CK(cudaMemcpyAsync(mb[sec]->p, src[sec], b, cudaMemcpyDeviceToDevice, sg)); }
    CK(cudaEventRecord(c->ev_sent[parity], sg));
    // 3. shard-local indices -> batch-wide ones (rank 0's own shard starts at 0 everywhere)
    for (int r = 1; r < W; r++) {
        const uint64_t* a = all + static_cast<size_t>(r) * kCountWords;
        const uint32_t np = static_cast<uint32_t>(a[2]), nr = static_cast<uint32_t>(a[0]);
        if (compact) { if (np) chunk_rebase_compact_kernel<<<std::max(1u, std::min<uint32_t>((np + 255) / 256, 1184u)), 256, 0, sg>>>(c->m_cpairs.as<CPairOut>() + pb[r], np, static_cast<uint32_t>(rb[r])); }
        else if (np || nr)
            chunk_rebase_kernel<<<std::max(1u, std::min<uint32_t>((std::max(np, nr) + 255) / 256, 1184u)), 256, 0, sg>>>(
                c->m_pairs.as<PairOut>() + pb[r], np, c->m_hit_off.as<uint32_t>() + rb[r], nr, static_cast<uint32_t>(rb[r]), static_cast<uint32_t>(hb[r]), static_cast<uint32_t>(cb[r]));
    }
    if (!compact) poke(sg, {{c->m_hit_off.as<uint32_t>() + rb[W], static_cast<uint32_t>(hb[W])}});
    CK(cudaGetLastError());
    memset(merged, 0, sizeof *merged);
    merged->n_reads = static_cast<uint32_t>(rb[W]); merged->n_hits = hb[W]; merged->n_pairs = pb[W]; merged->n_records = cb[W];
    merged->received = rb[W]; merged->mapped = mapped; merged->multimapped = multi; merged->alignments = cb[W]; merged->slow_path_pairs = slow;
    merged->rec_path_bytes = compact ? recw : 0u;
    if (compact) { merged->d_cpairs = reinterpret_cast<const grootgpu_cpair*>(c->m_cpairs.p); merged->d_rec_path_c = c->m_rec_c.p; }
    else {
        merged->d_hit_off = c->m_hit_off.as<uint32_t>(); merged->d_hits = c->m_hits.as<uint32_t>(); merged->d_pairs = reinterpret_cast<const grootgpu_pair*>(c->m_pairs.p);
        merged->d_rec_path = c->m_rec_path.as<uint32_t>(); merged->d_rec_pos = c->m_rec_pos.as<int32_t>();
    }
    const double t_enq = now_ms();
    if (to_host) {
        for (int sec = 0; sec < n_sec; sec++) {
            const uint64_t total_elems = compact ? (sec == 0 ? pb[W] : cb[W]) : (sec == 0 ? rb[W] + 1 : sec == 1 ? hb[W] : sec == 2 ? pb[W] : cb[W]);
            hbuf[sec]->need(std::max<size_t>(16, total_elems * elem[sec]));
            if (total_elems) CK(cudaMemcpyAsync(hbuf[sec]->p, mb[sec]->p, total_elems * elem[sec], cudaMemcpyDeviceToHost, sg));
        }
        if (to_host == 1) CK(cudaStreamSynchronize(sg));     // 2: asynchronous, complete after the next grootgpu_gather / grootgpu_comm_sync
        if (compact) { merged->cpairs = reinterpret_cast<const grootgpu_cpair*>(hs.cpairs.p); merged->rec_path_c = hs.rec_c.p; }
        else {
            merged->hit_off = hs.hit_off.as<uint32_t>(); merged->hits = hs.hits.as<uint32_t>(); merged->pairs = reinterpret_cast<const grootgpu_pair*>(hs.pairs.p);
            merged->rec_path = hs.rec_path.as<uint32_t>(); merged->rec_pos = hs.rec_pos.as<int32_t>();
        }
    }
    if (trace) fprintf(stderr, "[grootgpu] gather rank 0: sizes known at %.2f ms, receives + merge enqueued at %.2f ms, host copies %s at %.2f ms\n", t_counts, t_enq,
                       to_host == 1 ? "done" : to_host ? "enqueued" : "not asked for", now_ms());
}

void fnv_sink(void* ctx, const char* d, size_t n) { uint64_t& hsh = *static_cast<uint64_t*>(ctx); for (size_t i = 0; i < n; i++) { hsh ^= static_cast<uint8_t>(d[i]); hsh *= 1099511628211ULL; } }
void file_sink(void* ctx, const char* d, size_t n) { fwrite(d, 1, n, static_cast<FILE*>(ctx)); }

template <class F>
int guarded(F f) {
    try { f(); return GROOTGPU_OK; }
    catch (CudaError& e) { return fail(GROOTGPU_ERR_CUDA, e.what()); }
    catch (CommError& e) { return fail(GROOTGPU_ERR_COMM, e.what()); }
    catch (std::invalid_argument& e) { return fail(GROOTGPU_ERR_SHORT_READ, e.what()); }
    catch (std::domain_error& e) { return fail(GROOTGPU_ERR_BAD_BASE, e.what()); }
    catch (std::length_error& e) { return fail(GROOTGPU_ERR_CAPACITY, e.what()); }
    catch (std::bad_alloc&) { return fail(GROOTGPU_ERR_CAPACITY, "out of host memory"); }
    catch (std::ios_base::failure& e) { return fail(GROOTGPU_ERR_IO, e.what()); }
    catch (std::exception& e) { return fail(GROOTGPU_ERR_FORMAT, e.what()); }
}

}  // namespace

// =================================================================================================
extern "C" {

const char* grootgpu_last_error(void) { return g_err.c_str(); }
const char* grootgpu_version(void) { return GROOTGPU_VERSION " (groot " GROOTGPU_REFERENCE_VERSION " align path, sm_100a)"; }
int grootgpu_device_count(int* n) {
    if (!n) return fail(GROOTGPU_ERR_ARG, "null argument");
    *n = 0;
    if (cudaGetDeviceCount(n) != cudaSuccess) { *n = 0; return fail(GROOTGPU_ERR_CUDA, "cudaGetDeviceCount failed (no driver / no device)"); }
    return GROOTGPU_OK;
}
int grootgpu_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(GROOTGPU_ERR_ARG, "null argument");
    // GROOTGPU_HOST_ALLOC_WC=1: write-combined pinned memory (input buffers the host only writes: the DMA reads skip the CPU caches)
    const char* wc = getenv("GROOTGPU_HOST_ALLOC_WC");
    if (cudaHostAlloc(ptr, bytes, wc && *wc == '1' ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) return fail(GROOTGPU_ERR_CUDA, "cudaHostAlloc failed");
    return GROOTGPU_OK;
}
int grootgpu_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? GROOTGPU_OK : fail(GROOTGPU_ERR_CUDA, "cudaFreeHost failed"); }

int grootgpu_index_build(const char* const* msa_paths, uint32_t n_msa, const grootgpu_index_params* params, int device, grootgpu_index** out) {
    if (!msa_paths || !params || !out || n_msa == 0) return fail(GROOTGPU_ERR_ARG, "bad argument");
    *out = nullptr;
    grootgpu_index* ix = new grootgpu_index();
    int rc = guarded([&] {
        pick_device(device);
        ix->device = device;
        ix->h.p.k = params->kmer_size; ix->h.p.S = params->sketch_size; ix->h.p.w = params->window_size;
        ix->h.p.num_part = params->num_part; ix->h.p.max_k = params->max_k;
        if (ix->h.p.k < 1 || ix->h.p.w < ix->h.p.k) throw std::runtime_error("need 1 <= k <= window size");
        for (uint32_t i = 0; i < n_msa; i++) {
            std::string text;
            try { text = slurp(msa_paths[i]); } catch (std::exception& e) { throw std::ios_base::failure(e.what()); }
            append_graph_from_msa(ix->h, text);
        }
        build_windows(ix->h, sketch_cb, nullptr);
        index_to_device(ix);
    });
    if (rc != GROOTGPU_OK) { delete ix; return rc; }
    *out = ix;
    return GROOTGPU_OK;
}

int grootgpu_index_build_dir(const char* msa_dir, const grootgpu_index_params* params, int device, grootgpu_index** out) {
    if (!msa_dir) return fail(GROOTGPU_ERR_ARG, "bad argument");
    std::vector<std::string> files;
    DIR* d = opendir(msa_dir);
    if (!d) return fail(GROOTGPU_ERR_IO, std::string("cannot open MSA directory ") + msa_dir);
    while (dirent* e = readdir(d)) {
        std::string nme = e->d_name;
        if (nme.rfind("cluster", 0) == 0 && nme.size() > 4 && nme.compare(nme.size() - 4, 4, ".msa") == 0) files.push_back(std::string(msa_dir) + "/" + nme);
    }
    closedir(d);
    if (files.empty()) return fail(GROOTGPU_ERR_EMPTY, "no MSA files in the supplied directory (must be named cluster-DD.msa)");
    std::sort(files.begin(), files.end());
    std::vector<const char*> p;
    for (auto& f : files) p.push_back(f.c_str());
    return grootgpu_index_build(p.data(), static_cast<uint32_t>(p.size()), params, device, out);
}

// Host-only: MSA -> graphs (no windows, no device). Dumps the graph section (G/P/N lines) of the canonical
// text form; lets the CPU-only test-suite check the graph builder against the oracle without a GPU.
int grootgpu_graphs_dump(const char* const* msa_paths, uint32_t n_msa, const grootgpu_index_params* params, const char* dump_path, uint64_t* hash) {
    if (!msa_paths || !params || n_msa == 0) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] {
        FlatIndex h;
        h.p.k = params->kmer_size; h.p.S = params->sketch_size; h.p.w = params->window_size; h.p.num_part = params->num_part; h.p.max_k = params->max_k;
        for (uint32_t i = 0; i < n_msa; i++) {
            std::string text;
            try { text = slurp(msa_paths[i]); } catch (std::exception& e) { throw std::ios_base::failure(e.what()); }
            append_graph_from_msa(h, text);
        }
        if (hash) { *hash = 1469598103934665603ULL; dump_index(h, fnv_sink, hash); }
        if (dump_path) {
            FILE* f = fopen(dump_path, "wb");
            if (!f) throw std::ios_base::failure(std::string("cannot create ") + dump_path);
            dump_index(h, file_sink, f);
            fclose(f);
        }
    });
}

int grootgpu_index_save(const grootgpu_index* idx, const char* path) {
    if (!idx || !path) return fail(GROOTGPU_ERR_ARG, "bad argument");
    { grootgpu_index* mi = const_cast<grootgpu_index*>(idx); int rc = guarded([&] { if (mi->weights_on_device) { pick_device(mi->device); sync_weights_to_host(mi); } }); if (rc) return rc; }
    try { save_index(idx->h, path); } catch (std::exception& e) { return fail(GROOTGPU_ERR_IO, e.what()); }
    return GROOTGPU_OK;
}
int grootgpu_index_load(const char* path, int device, grootgpu_index** out) {
    if (!path || !out) return fail(GROOTGPU_ERR_ARG, "bad argument");
    *out = nullptr;
    grootgpu_index* ix = new grootgpu_index();
    int rc = guarded([&] { load_index(ix->h, path); pick_device(device); ix->device = device; index_to_device(ix); });   // the file is parsed and validated before a device is needed
    if (rc != GROOTGPU_OK) { delete ix; return rc; }
    *out = ix;
    return GROOTGPU_OK;
}
int grootgpu_index_load_gob(const char* gg_path, const char* lshe_path, int device, grootgpu_index** out) {
    if (!gg_path || !lshe_path || !out) return fail(GROOTGPU_ERR_ARG, "bad argument");
    *out = nullptr;
    grootgpu_index* ix = new grootgpu_index();
    int rc = guarded([&] { load_index_gob(ix->h, gg_path, lshe_path); validate_index(ix->h); pick_device(device); ix->device = device; index_to_device(ix); });
    if (rc != GROOTGPU_OK) { delete ix; return rc; }
    *out = ix;
    return GROOTGPU_OK;
}
int grootgpu_gob_dump(const char* gg_path, const char* lshe_path, const char* dump_path, uint64_t* hash) {
    if (!gg_path || !lshe_path) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] {
        FlatIndex h;
        load_index_gob(h, gg_path, lshe_path);
        validate_index(h);
        if (hash) { *hash = 1469598103934665603ULL; dump_index(h, fnv_sink, hash); }
        if (dump_path) {
            FILE* f = fopen(dump_path, "wb");
            if (!f) throw std::ios_base::failure(std::string("cannot create ") + dump_path);
            dump_index(h, file_sink, f);
            fclose(f);
        }
    });
}
int grootgpu_index_save_gob(const grootgpu_index* idx, const char* gg_path, const char* lshe_path) {
    if (!idx || !gg_path || !lshe_path) return fail(GROOTGPU_ERR_ARG, "bad argument");
    { grootgpu_index* mi = const_cast<grootgpu_index*>(idx); int rc = guarded([&] { if (mi->weights_on_device) { pick_device(mi->device); sync_weights_to_host(mi); } }); if (rc) return rc; }
    try { save_index_gob(idx->h, gg_path, lshe_path, GROOTGPU_REFERENCE_VERSION); } catch (std::exception& e) { return fail(GROOTGPU_ERR_IO, e.what()); }
    return GROOTGPU_OK;
}
int grootgpu_flat_to_gob(const char* flat_path, const char* gg_path, const char* lshe_path) {
    if (!flat_path || !gg_path || !lshe_path) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] { FlatIndex h; load_index(h, flat_path); save_index_gob(h, gg_path, lshe_path, GROOTGPU_REFERENCE_VERSION); });
}
int grootgpu_gob_to_flat(const char* gg_path, const char* lshe_path, const char* flat_path) {
    if (!flat_path || !gg_path || !lshe_path) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] { FlatIndex h; load_index_gob(h, gg_path, lshe_path); validate_index(h); save_index(h, flat_path); });
}
void grootgpu_index_destroy(grootgpu_index* idx) {
    if (!idx) return;
    cudaSetDevice(idx->device);
    if (idx->comm) {   // a communicator that outlives its index: detach it (grootgpu_comm_destroy then only releases NCCL)
        if (idx->st_acc) cudaStreamSynchronize(idx->st_acc);
        if (idx->comm->st_gather) cudaStreamSynchronize(idx->comm->st_gather);
        idx->comm->ix = nullptr;
    }
    delete idx;
}

int grootgpu_index_get_info(const grootgpu_index* idx, grootgpu_index_info* o) {
    if (!idx || !o) return fail(GROOTGPU_ERR_ARG, "bad argument");
    const FlatIndex& h = idx->h;
    memset(o, 0, sizeof *o);
    o->params.kmer_size = h.p.k; o->params.sketch_size = h.p.S; o->params.window_size = h.p.w; o->params.num_part = h.p.num_part; o->params.max_k = h.p.max_k;
    o->n_graphs = h.n_graphs; o->n_paths = static_cast<uint32_t>(h.path_name.size()); o->n_nodes = static_cast<uint32_t>(h.nodes.size());
    o->n_windows = static_cast<uint32_t>(h.wins.size());
    for (uint32_t g = 0; g < h.n_graphs; g++) {
        o->n_masked_graphs += h.graph_masked[g]; o->n_raw_windows += h.graph_raw_windows[g];
        o->max_paths_per_graph = std::max(o->max_paths_per_graph, h.n_paths_of(g));
        for (uint32_t p = h.graph_path_base[g]; p < h.graph_path_base[g + 1]; p++) o->n_path_bases += static_cast<uint64_t>(h.path_len[p]);
    }
    for (auto& w : h.wins) o->max_merge_span = std::max(o->max_merge_span, w.merge_span);
    return GROOTGPU_OK;
}

int grootgpu_index_dump_hash(const grootgpu_index* idx, uint64_t* hash) {
    if (!idx || !hash) return fail(GROOTGPU_ERR_ARG, "bad argument");
    *hash = 1469598103934665603ULL;
    dump_index(idx->h, fnv_sink, hash);
    return GROOTGPU_OK;
}
int grootgpu_index_dump_file(const grootgpu_index* idx, const char* path) {
    if (!idx || !path) return fail(GROOTGPU_ERR_ARG, "bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(GROOTGPU_ERR_IO, std::string("cannot create ") + path);
    dump_index(idx->h, file_sink, f);
    fclose(f);
    return GROOTGPU_OK;
}
int grootgpu_index_ref(const grootgpu_index* idx, uint32_t g, uint32_t p, const char** name, int32_t* length) {
    if (!idx || g >= idx->h.n_graphs || p >= idx->h.n_paths_of(g)) return fail(GROOTGPU_ERR_ARG, "graph / path id out of range");
    uint32_t gp = idx->h.graph_path_base[g] + p;
    if (name) *name = idx->h.path_name[gp].c_str();
    if (length) *length = idx->h.path_len[gp];
    return GROOTGPU_OK;
}
int grootgpu_index_query_params(grootgpu_index* idx, uint32_t q, double t, uint32_t* K, uint32_t* L, uint32_t* eq_min) {
    if (!idx || q == 0) return fail(GROOTGPU_ERR_ARG, "bad argument");
    LenParam lp = param_for(idx, q, t);
    if (K) *K = lp.K; if (L) *L = lp.L; if (eq_min) *eq_min = lp.eq_min;
    return GROOTGPU_OK;
}

int grootgpu_query_params_host(const grootgpu_index_params* p, uint32_t q, double t, uint32_t* K, uint32_t* L, uint32_t* eq_min) {
    if (!p || q == 0 || p->max_k == 0 || p->sketch_size < p->max_k || p->window_size < p->kmer_size) return fail(GROOTGPU_ERR_ARG, "bad argument");
    int k = 0, l = 0;
    const int x = static_cast<int>(p->window_size - p->kmer_size + 1);
    optimal_kl(static_cast<int>(p->max_k), static_cast<int>(p->sketch_size / p->max_k), x, static_cast<int>(q), t, &k, &l);
    if (K) *K = static_cast<uint32_t>(k);
    if (L) *L = static_cast<uint32_t>(l);
    if (eq_min) *eq_min = static_cast<uint32_t>(eq_min_for(static_cast<int>(p->sketch_size), static_cast<int>(q), x, t));
    return GROOTGPU_OK;
}

int grootgpu_align_batch_device(grootgpu_index* idx, const uint8_t* d_seq, const uint32_t* d_seq_off, uint32_t n_reads, uint32_t min_len,
                                uint32_t max_len, const grootgpu_align_params* params, void* stream, grootgpu_batch_result* out) {
    if (!idx || !d_seq || !d_seq_off || !params || n_reads == 0 || max_len < min_len) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] {
        pick_device(idx->device);
        Workspace* w = &idx->ws[0];
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : w->stream;
        // consecutive calls write alternating result sets: the arrays of a call stay valid during the next one (a
        // gather of batch b overlaps the mapping of batch b + 1) and are reused by the call after that
        idx->call_parity ^= 1;
        if (w->rset != idx->call_parity) w->swap_result_sets();
        if (idx->comm) CK(cudaStreamWaitEvent(st, idx->comm->ev_sent[idx->call_parity], 0));
        w->ring_first = w->ring_last = true; w->acc_before = nullptr; w->acc_after = nullptr;
        run_batch(idx, w, d_seq, d_seq_off, n_reads, min_len, max_len, params, st, out);
    });
}

int grootgpu_align_batch(grootgpu_index* idx, const uint8_t* seq, const uint64_t* seq_off, uint32_t n_reads, const grootgpu_align_params* params,
                         grootgpu_batch_result* out) {
    if (!idx || !seq || !params || !out || n_reads == 0 || (!seq_off && params->fixed_read_len == 0)) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] {
        pick_device(idx->device);
        idx->call_parity ^= 1;
        run_batch_chunked(idx, seq, seq_off, n_reads, params, out);
    });
}

int grootgpu_index_node_paths(const grootgpu_index* idx, uint32_t node, uint32_t* graph, const uint32_t** path_ids, const int32_t** positions, uint32_t* n_paths) {
    if (!idx || node >= idx->h.nodes.size()) return fail(GROOTGPU_ERR_ARG, "node index out of range");
    const NodeRec& nr = idx->h.nodes[node];
    if (graph) *graph = idx->h.graph_of_node(node);
    if (path_ids) *path_ids = idx->h.node_path_id.data() + nr.path_off;
    if (positions) *positions = idx->h.node_path_pos.data() + nr.path_off;
    if (n_paths) *n_paths = nr.path_cnt;
    return GROOTGPU_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------------
int grootgpu_comm_id(uint8_t id[GROOTGPU_COMM_ID_BYTES]) {
    if (!id) return fail(GROOTGPU_ERR_ARG, "bad argument");
    static_assert(GROOTGPU_COMM_ID_BYTES == 2 * NCCL_UNIQUE_ID_BYTES, "one id per communicator");
    return guarded([&] {
        ncclUniqueId a, b;
        NK(nccl().GetUniqueId(&a)); NK(nccl().GetUniqueId(&b));
        memcpy(id, &a, NCCL_UNIQUE_ID_BYTES); memcpy(id + NCCL_UNIQUE_ID_BYTES, &b, NCCL_UNIQUE_ID_BYTES);
    });
}

int grootgpu_comm_create(grootgpu_index* idx, const uint8_t id[GROOTGPU_COMM_ID_BYTES], int rank, int world_size, grootgpu_comm** out) {
    if (!idx || !id || !out || world_size < 1 || rank < 0 || rank >= world_size) return fail(GROOTGPU_ERR_ARG, "bad argument");
    if (idx->comm) return fail(GROOTGPU_ERR_ARG, "the index already has a communicator");
    *out = nullptr;
    grootgpu_comm* c = new grootgpu_comm();
    int rc = guarded([&] {
        pick_device(idx->device);
        c->ix = idx; c->rank = rank; c->world = world_size;
        ncclUniqueId a, b;
        memcpy(&a, id, NCCL_UNIQUE_ID_BYTES); memcpy(&b, id + NCCL_UNIQUE_ID_BYTES, NCCL_UNIQUE_ID_BYTES);
        NK(nccl().CommInitRank(&c->ring, world_size, a, rank));
        NK(nccl().CommInitRank(&c->gath, world_size, b, rank));
        {
            int prio_lo = 0, prio_hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CK(cudaStreamCreateWithPriority(&c->st_gather, cudaStreamNonBlocking, prio_lo));   // lowest: see Workspace::create
        }
        for (cudaEvent_t* e : {&c->ev_sent[0], &c->ev_sent[1]}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        c->d_counts.need(8ull * kCountWords * (world_size + 2));
        CK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_counts), 8ull * kCountWords * (world_size + 2), cudaHostAllocDefault));
        if (world_size > 1) {
            // NCCL connects two peers the first time they exchange something, and that handshake blocks the HOST until
            // both sides have issued a matching operation. In steady state the last rank's weight vector is sent one
            // batch before rank 0 asks for it (and rank 0 sits in the gather meanwhile): every connection the run will
            // use is therefore opened here, where all ranks issue their side at the same time.
            uint64_t* scratch = c->d_counts.as<uint64_t>();
            NK(nccl().GroupStart());
            NK(nccl().Send(scratch, 1, ncclUint64, (rank + 1) % world_size, c->ring, c->st_gather));
            NK(nccl().Recv(scratch + 1, 1, ncclUint64, (rank + world_size - 1) % world_size, c->ring, c->st_gather));
            NK(nccl().GroupEnd());
            NK(nccl().GroupStart());
            if (rank == 0) { for (int r = 1; r < world_size; r++) NK(nccl().Recv(scratch + 2 + r, 1, ncclUint64, r, c->gath, c->st_gather)); }
            else NK(nccl().Send(scratch, 1, ncclUint64, 0, c->gath, c->st_gather));
            NK(nccl().GroupEnd());
            NK(nccl().AllGather(scratch, scratch + kCountWords, 1, ncclUint64, c->gath, c->st_gather));
            CK(cudaStreamSynchronize(c->st_gather));
        }
        // every rank starts from its own weights; only rank 0's enter the ring (the others' are replaced by what arrives)
        if (rank != 0) { sync_weights_to_host(idx); idx->weights_on_device = false; std::fill(idx->h.kmer_freq.begin(), idx->h.kmer_freq.end(), 0.0); std::fill(idx->h.kmer_total.begin(), idx->h.kmer_total.end(), 0); }
    });
    if (rc != GROOTGPU_OK) { delete c; return rc; }
    idx->comm = c;
    *out = c;
    return GROOTGPU_OK;
}

void grootgpu_comm_destroy(grootgpu_comm* c) {
    if (!c) return;
    if (c->ix) { cudaSetDevice(c->ix->device); cudaStreamSynchronize(c->ix->st_acc); if (c->st_gather) cudaStreamSynchronize(c->st_gather); c->ix->comm = nullptr; }
    delete c;
}

int grootgpu_gather(grootgpu_comm* c, const grootgpu_batch_result* local, int to_host, grootgpu_batch_result* merged) {
    if (!c || !local) return fail(GROOTGPU_ERR_ARG, "bad argument");
    if (c->rank == 0 && !merged) return fail(GROOTGPU_ERR_ARG, "rank 0 needs a place for the merged result");
    return guarded([&] {
        grootgpu_index* ix = c->ix;
        pick_device(ix->device);
        gather_results(c, local, to_host, merged);
    });
}

int grootgpu_comm_sync(grootgpu_comm* c) {
    if (!c) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] {
        grootgpu_index* ix = c->ix;
        pick_device(ix->device);
        CK(cudaStreamSynchronize(c->st_gather));
        if (c->world > 1) {
            push_weights_to_device(ix);
            cudaStream_t sa = ix->st_acc;
            std::lock_guard<std::mutex> lock(ix->acc_mu);
            // the weight vector: whatever the last rank sent after the last batch is rank 0's result
            if (c->rank == 0 && c->ring_pending) { NK(nccl().Recv(ix->d_kmer_freq, ix->h.nodes.size(), ncclDouble, c->world - 1, c->ring, sa)); c->ring_pending = false; }
            // KmerTotal is an integer sum (graph.go:449): every rank counted its own reads
            NK(nccl().Reduce(ix->d_kmer_total, ix->d_kmer_total, ix->h.n_graphs, ncclUint64, ncclSum, 0, c->ring, sa));
            if (c->rank != 0) {
                CK(cudaMemsetAsync(ix->d_kmer_freq, 0, ix->h.nodes.size() * sizeof(double), sa));
                CK(cudaMemsetAsync(ix->d_kmer_total, 0, ix->h.n_graphs * sizeof(unsigned long long), sa));
            }
        }
        CK(cudaStreamSynchronize(ix->st_acc));
    });
}

int grootgpu_project_batch(grootgpu_index* idx, const grootgpu_batch_result* res, const uint64_t* seq_off) {
    if (!idx || !res || !seq_off) return fail(GROOTGPU_ERR_ARG, "bad argument");
    if (res->n_pairs && (!res->pairs || !res->hits)) return fail(GROOTGPU_ERR_ARG, "result holds no host arrays");
    return guarded([&] {
        pick_device(idx->device);
        sync_weights_to_host(idx);
        FlatIndex& h = idx->h;
        // pairs are ordered by (read, graph): bucketing by graph keeps read order inside every graph,
        // which is the only order the f64 accumulation of a graph depends on (one minion per graph).
        std::vector<uint32_t> cnt(h.n_graphs + 1, 0);
        for (uint64_t i = 0; i < res->n_pairs; i++) cnt[res->pairs[i].graph + 1]++;
        for (uint32_t g = 0; g < h.n_graphs; g++) cnt[g + 1] += cnt[g];
        std::vector<uint32_t> order(res->n_pairs), cur(cnt.begin(), cnt.end() - 1);
        for (uint64_t i = 0; i < res->n_pairs; i++) order[cur[res->pairs[i].graph]++] = static_cast<uint32_t>(i);
        std::vector<uint32_t> active;
        for (uint32_t g = 0; g < h.n_graphs; g++) if (cnt[g + 1] > cnt[g]) active.push_back(g);
        std::atomic<size_t> next{0};
        auto work = [&] {
            while (true) {
                size_t a = next.fetch_add(1);
                if (a >= active.size()) break;
                uint32_t g = active[a];
                for (uint32_t o = cnt[g]; o < cnt[g + 1]; o++) {
                    const grootgpu_pair& p = res->pairs[order[o]];
                    double kmers = static_cast<double>(static_cast<int64_t>(seq_off[p.read + 1] - seq_off[p.read]) - static_cast<int64_t>(h.p.k)) + 1.0;  // graphminion.go:60
                    for (uint32_t m = 0; m < p.n_incremented; m++) increment_sub_path(h, res->hits[p.hit_begin + m], kmers);
                }
            }
        };
        unsigned T = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 16u));
        if (res->n_pairs < 4096 || T == 1) work();
        else { std::vector<std::thread> th; for (unsigned t = 0; t < T; t++) th.emplace_back(work); for (auto& x : th) x.join(); }
    });
}

int grootgpu_weights(const grootgpu_index* idx_c, double* kmer_freq, uint64_t* kmer_total) {
    if (!idx_c) return fail(GROOTGPU_ERR_ARG, "bad argument");
    grootgpu_index* idx = const_cast<grootgpu_index*>(idx_c);
    int rc = guarded([&] { pick_device(idx->device); sync_weights_to_host(idx); });
    if (rc) return rc;
    if (kmer_freq) memcpy(kmer_freq, idx->h.kmer_freq.data(), idx->h.kmer_freq.size() * sizeof(double));
    if (kmer_total) memcpy(kmer_total, idx->h.kmer_total.data(), idx->h.kmer_total.size() * sizeof(uint64_t));
    return GROOTGPU_OK;
}
int grootgpu_reset_weights(grootgpu_index* idx) {
    if (!idx) return fail(GROOTGPU_ERR_ARG, "bad argument");
    idx->weights_on_device = false;   // the host copy (zeroed below) becomes authoritative
    std::fill(idx->h.kmer_freq.begin(), idx->h.kmer_freq.end(), 0.0);
    std::fill(idx->h.kmer_total.begin(), idx->h.kmer_total.end(), 0);
    return GROOTGPU_OK;
}

int grootgpu_int_issue_peak(int device, double* warp_inst_per_s) {
    if (!warp_inst_per_s) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] {
        pick_device(device);
        DBuf out; out.need(64);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int blocks = g_num_sms(device) * 8;
        const uint32_t iters = 2048;
        for (int mode = 0; mode < 3; mode++) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {      // first repetition warms up; best of the rest
                CK(cudaEventRecord(e0, 0));
                if (mode == 0) int_peak_kernel<0><<<blocks, 256>>>(out.as<uint32_t>(), iters, 0x9E3779B1u, 0x85EBCA77u, 7u);
                else if (mode == 1) int_peak_kernel<1><<<blocks, 256>>>(out.as<uint32_t>(), iters, 0x9E3779B1u, 0x85EBCA77u, 7u);
                else int_peak_kernel<2><<<blocks, 256>>>(out.as<uint32_t>(), iters, 0x9E3779B1u, 0x85EBCA77u, 7u);
                CK(cudaEventRecord(e1, 0));
                CK(cudaEventSynchronize(e1));
                float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0) best = std::min(best, ms);
            }
            const double warps = static_cast<double>(blocks) * 256.0 / 32.0;
            warp_inst_per_s[mode] = warps * iters * kIntPeakInstPerIter[mode] / (best * 1e-3);
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    });
}

int grootgpu_sketch_batch(int device, const uint8_t* seq, const uint64_t* seq_off, uint32_t n, uint32_t k, uint32_t S, uint64_t* out) {
    if (!seq || !seq_off || !out || n == 0 || k == 0) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] {
        pick_device(device);
        std::vector<uint32_t> lens(n);
        for (uint32_t i = 0; i < n; i++) {
            lens[i] = static_cast<uint32_t>(seq_off[i + 1] - seq_off[i]);
            if (lens[i] < k) throw std::invalid_argument("sequence " + std::to_string(i) + " is shorter than k (khf.go:38-41)");
        }
        sketch_host(seq, seq_off[n], seq_off, lens.data(), 0, n, k, S, out);
    });
}

int grootgpu_prune(grootgpu_index* idx, double min_cov, uint8_t* kept) {
    if (!idx) return fail(GROOTGPU_ERR_ARG, "bad argument");
    return guarded([&] { pick_device(idx->device); sync_weights_to_host(idx); for (uint32_t g = 0; g < idx->h.n_graphs; g++) { bool k = prune_graph(idx->h, g, min_cov); if (kept) kept[g] = k ? 1 : 0; } });
}
int grootgpu_graph_save_gfa(const grootgpu_index* idx, uint32_t g, const char* path, int64_t total_kmers, int* written) {
    if (!idx || !path || g >= idx->h.n_graphs) return fail(GROOTGPU_ERR_ARG, "bad argument");
    { grootgpu_index* mi = const_cast<grootgpu_index*>(idx); int rc = guarded([&] { pick_device(mi->device); sync_weights_to_host(mi); }); if (rc) return rc; }
    std::string s = graph_to_gfa(idx->h, g, total_kmers);
    if (written) *written = s.empty() ? 0 : 1;
    if (s.empty()) return GROOTGPU_OK;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(GROOTGPU_ERR_IO, std::string("cannot create ") + path);
    fwrite(s.data(), 1, s.size(), f);
    fclose(f);
    return GROOTGPU_OK;
}

}  // extern "C"
