"""ctypes binding of libgrootgpu.so (include/grootgpu.h) — the host-side mirror used by the tests, the
bench and the Python driver. There is no fallback: if the CUDA library is missing or no device is
usable every call raises.

Names follow the reference's pipeline (src/pipeline): an `Index` is what `groot index` writes and
`groot align` loads (Info + ContainmentIndex); `Index.map_reads` is theBoss.mapReads for one batch of
FASTQ reads; `Index.project` is the graph weighting the graph minions do.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GROOTGPU_LIB") or os.path.join(_HERE, "libgrootgpu.so")   # GROOTGPU_LIB: tuning variants only

ERR_NAMES = {0: "OK", -1: "ERR_ARG", -2: "ERR_CUDA", -3: "ERR_IO", -4: "ERR_FORMAT", -5: "ERR_SHORT_READ",
             -6: "ERR_BAD_BASE", -7: "ERR_CAPACITY", -8: "ERR_EMPTY", -9: "ERR_COMM"}


class GrootGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), msg))
        self.code = code


class IndexParams(C.Structure):
    _fields_ = [("kmer_size", C.c_uint32), ("sketch_size", C.c_uint32), ("window_size", C.c_uint32),
                ("num_part", C.c_uint32), ("max_k", C.c_uint32)]


class IndexInfo(C.Structure):
    _fields_ = [("params", IndexParams), ("n_graphs", C.c_uint32), ("n_masked_graphs", C.c_uint32), ("n_paths", C.c_uint32),
                ("n_nodes", C.c_uint32), ("n_path_bases", C.c_uint64), ("n_raw_windows", C.c_uint64), ("n_windows", C.c_uint32),
                ("max_merge_span", C.c_uint32), ("max_paths_per_graph", C.c_uint32)]


class AlignParams(C.Structure):
    _fields_ = [("containment_threshold", C.c_double), ("no_align", C.c_int32), ("keep_sketches", C.c_int32),
                ("project_on_device", C.c_int32), ("results_on_device", C.c_int32), ("compact_records", C.c_int32),
                ("fixed_read_len", C.c_uint32)]


class Pair(C.Structure):
    _fields_ = [("read", C.c_uint32), ("graph", C.c_uint32), ("hit_begin", C.c_uint32), ("hit_count", C.c_uint32),
                ("n_incremented", C.c_uint32), ("rec_begin", C.c_uint32), ("rec_count", C.c_uint32),
                ("reverse", C.c_uint8), ("clip_start", C.c_uint8), ("clip_end", C.c_uint8), ("stage", C.c_uint8)]


PAIR_DTYPE = np.dtype([("read", "<u4"), ("graph", "<u4"), ("hit_begin", "<u4"), ("hit_count", "<u4"), ("n_incremented", "<u4"),
                       ("rec_begin", "<u4"), ("rec_count", "<u4"), ("reverse", "u1"), ("clip_start", "u1"), ("clip_end", "u1"),
                       ("stage", "u1")])
assert PAIR_DTYPE.itemsize == C.sizeof(Pair) == 32


class CPair(C.Structure):
    _fields_ = [("read", C.c_uint32), ("node", C.c_uint32), ("offset_flags", C.c_uint32), ("rec_count", C.c_uint32)]


CPAIR_DTYPE = np.dtype([("read", "<u4"), ("node", "<u4"), ("offset_flags", "<u4"), ("rec_count", "<u4")])
CPAIR_OFFSET_MASK, CPAIR_REVERSE, CPAIR_CLIP_START, CPAIR_CLIP_END = 0x0fffffff, 0x10000000, 0x20000000, 0x40000000
COMM_ID_BYTES = 256


class BatchResultC(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_hits", C.c_uint64), ("n_pairs", C.c_uint64), ("n_records", C.c_uint64),
                ("hit_off", C.POINTER(C.c_uint32)), ("hits", C.POINTER(C.c_uint32)), ("pairs", C.POINTER(Pair)),
                ("rec_path", C.POINTER(C.c_uint32)), ("rec_pos", C.POINTER(C.c_int32)), ("sketches", C.POINTER(C.c_uint64)),
                ("received", C.c_uint64), ("mapped", C.c_uint64), ("multimapped", C.c_uint64), ("alignments", C.c_uint64),
                ("ms", C.c_float * 4), ("kernel_ms", C.c_float * 8), ("kernel_launches", C.c_uint32), ("slow_path_pairs", C.c_uint64),
                ("d_hit_off", C.c_void_p), ("d_hits", C.c_void_p), ("d_pairs", C.c_void_p), ("d_rec_path", C.c_void_p),
                ("d_rec_pos", C.c_void_p),
                ("cpairs", C.POINTER(CPair)), ("rec_path_c", C.c_void_p), ("rec_path_bytes", C.c_uint32),
                ("d_cpairs", C.c_void_p), ("d_rec_path_c", C.c_void_p), ("result_set", C.c_uint32)]


KERNEL_FAMILIES = ("seed", "fill", "align_screen", "align_walk", "align_finish", "align_emit", "project", "project_accumulate")

_lib = None


def lib():
    """Loads libgrootgpu.so; raises if it has not been built (python -m groot_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GrootGpuError(-2, "libgrootgpu.so is not built (run `python -m groot_b200.build`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.grootgpu_last_error.restype = C.c_char_p
        L.grootgpu_version.restype = C.c_char_p
        L.grootgpu_device_count.argtypes = [C.POINTER(C.c_int)]
        L.grootgpu_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        L.grootgpu_host_free.argtypes = [vp]
        L.grootgpu_index_build.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(IndexParams), C.c_int, C.POINTER(vp)]
        L.grootgpu_index_build_dir.argtypes = [C.c_char_p, C.POINTER(IndexParams), C.c_int, C.POINTER(vp)]
        L.grootgpu_graphs_dump.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(IndexParams), C.c_char_p, C.POINTER(C.c_uint64)]
        L.grootgpu_index_save.argtypes = [vp, C.c_char_p]
        L.grootgpu_index_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
        L.grootgpu_index_load_gob.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
        L.grootgpu_gob_dump.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint64)]
        L.grootgpu_index_save_gob.argtypes = [vp, C.c_char_p, C.c_char_p]
        L.grootgpu_flat_to_gob.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.grootgpu_gob_to_flat.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.grootgpu_index_destroy.argtypes = [vp]
        L.grootgpu_index_destroy.restype = None
        L.grootgpu_index_get_info.argtypes = [vp, C.POINTER(IndexInfo)]
        L.grootgpu_index_dump_file.argtypes = [vp, C.c_char_p]
        L.grootgpu_index_dump_hash.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.grootgpu_index_ref.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]
        L.grootgpu_index_query_params.argtypes = [vp, C.c_uint32, C.c_double, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.grootgpu_query_params_host.argtypes = [C.POINTER(IndexParams), C.c_uint32, C.c_double, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.grootgpu_align_batch.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(AlignParams), C.POINTER(BatchResultC)]
        L.grootgpu_align_batch_device.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(AlignParams), vp, C.POINTER(BatchResultC)]
        L.grootgpu_project_batch.argtypes = [vp, C.POINTER(BatchResultC), vp]
        L.grootgpu_index_node_paths.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_int32)),
                                                C.POINTER(C.c_uint32)]
        L.grootgpu_comm_id.argtypes = [C.c_char_p]
        L.grootgpu_comm_create.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
        L.grootgpu_comm_destroy.argtypes = [vp]
        L.grootgpu_comm_destroy.restype = None
        L.grootgpu_comm_sync.argtypes = [vp]
        L.grootgpu_gather.argtypes = [vp, C.POINTER(BatchResultC), C.c_int, C.POINTER(BatchResultC)]
        L.grootgpu_weights.argtypes = [vp, vp, vp]
        L.grootgpu_reset_weights.argtypes = [vp]
        L.grootgpu_sketch_batch.argtypes = [C.c_int, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.grootgpu_int_issue_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.grootgpu_prune.argtypes = [vp, C.c_double, vp]
        L.grootgpu_graph_save_gfa.argtypes = [vp, C.c_uint32, C.c_char_p, C.c_int64, C.POINTER(C.c_int)]
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "grootgpu_index_build", "grootgpu_index_build_dir", "grootgpu_index_save", "grootgpu_index_load", "grootgpu_index_destroy",
    "grootgpu_index_get_info", "grootgpu_graphs_dump", "grootgpu_index_dump_file", "grootgpu_index_dump_hash", "grootgpu_index_ref",
    "grootgpu_index_query_params", "grootgpu_query_params_host", "grootgpu_align_batch", "grootgpu_align_batch_device", "grootgpu_project_batch",
    "grootgpu_weights", "grootgpu_reset_weights", "grootgpu_sketch_batch", "grootgpu_prune", "grootgpu_graph_save_gfa",
    "grootgpu_host_alloc", "grootgpu_host_free", "grootgpu_device_count", "grootgpu_last_error", "grootgpu_version",
    "grootgpu_int_issue_peak", "grootgpu_index_node_paths", "grootgpu_comm_id", "grootgpu_comm_create", "grootgpu_comm_destroy", "grootgpu_comm_sync",
    "grootgpu_gather", "grootgpu_index_load_gob", "grootgpu_gob_dump", "grootgpu_index_save_gob", "grootgpu_flat_to_gob", "grootgpu_gob_to_flat",
]


def _check(rc):
    if rc != 0:
        raise GrootGpuError(rc, lib().grootgpu_last_error().decode(errors="replace"))


def device_count():
    n = C.c_int()
    rc = lib().grootgpu_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def int_issue_peak(device=0):
    """Measured integer issue rate (warp instructions / s) for IMAD only, shift/xor only, and their 1:1 mix."""
    out = (C.c_double * 3)()
    _check(lib().grootgpu_int_issue_peak(device, out))
    return dict(imad=out[0], alu=out[1], mixed=out[2])


class PinnedBuffer:
    """Pinned host memory from grootgpu_host_alloc, viewed as a numpy uint8 array (`.array`); `.ptr` for the raw-pointer calls."""

    def __init__(self, nbytes):
        p = C.c_void_p()
        _check(lib().grootgpu_host_alloc(C.byref(p), max(1, int(nbytes))))
        self.ptr, self.nbytes = p.value, int(nbytes)
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(1, int(nbytes)),))[: int(nbytes)]

    def free(self):
        if self.ptr:
            lib().grootgpu_host_free(C.c_void_p(self.ptr))
            self.ptr = None


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).view(dtype).copy()


class BatchResult:
    """Host copy of one batch's result (see grootgpu_batch_result). `raw` stays valid until the next
    map_reads on the same index and is what `Index.project` consumes."""

    def __init__(self, raw: BatchResultC, S: int, copied: bool):
        self.raw = raw
        self.n_reads = raw.n_reads
        self.n_hits, self.n_pairs, self.n_records = int(raw.n_hits), int(raw.n_pairs), int(raw.n_records)
        self.counts = dict(received=int(raw.received), mapped=int(raw.mapped), multimapped=int(raw.multimapped),
                           alignments=int(raw.alignments))
        self.ms = dict(total=raw.ms[0], seed=raw.ms[1], align=raw.ms[2], other=raw.ms[3])
        self.kernel_launches = raw.kernel_launches
        self.kernel_ms = dict(zip(KERNEL_FAMILIES, list(raw.kernel_ms)[:8]))
        self.slow_path_pairs = int(raw.slow_path_pairs)
        self.compact = raw.rec_path_bytes != 0
        if copied and self.compact:
            self.cpairs = (np.ctypeslib.as_array(C.cast(raw.cpairs, C.POINTER(C.c_uint8)), shape=(self.n_pairs * 16,)).view(CPAIR_DTYPE).copy()
                           if self.n_pairs else np.zeros(0, dtype=CPAIR_DTYPE))
            dt = np.uint8 if raw.rec_path_bytes == 1 else np.uint16
            self.rec_path_c = (np.ctypeslib.as_array(C.cast(raw.rec_path_c, C.POINTER(C.c_uint8)), shape=(self.n_records * raw.rec_path_bytes,)).view(dt).copy()
                               if self.n_records else np.zeros(0, dtype=dt))
            self.sketches = (_np(raw.sketches, raw.n_reads * S, np.uint64).reshape(raw.n_reads, S) if raw.sketches else None)
        elif copied:
            self.hit_off = _np(raw.hit_off, raw.n_reads + 1, np.uint32)
            self.hits = _np(raw.hits, raw.n_hits, np.uint32)
            self.pairs = (np.ctypeslib.as_array(C.cast(raw.pairs, C.POINTER(C.c_uint8)), shape=(self.n_pairs * 32,)).view(PAIR_DTYPE).copy()
                          if self.n_pairs else np.zeros(0, dtype=PAIR_DTYPE))
            self.rec_path = _np(raw.rec_path, raw.n_records, np.uint32)
            self.rec_pos = _np(raw.rec_pos, raw.n_records, np.int32)
            self.sketches = (_np(raw.sketches, raw.n_reads * S, np.uint64).reshape(raw.n_reads, S) if raw.sketches else None)

    def decode_compact(self, index):
        """Compact output -> the (read, graph, path, pos, flags, startClip, endClip) rows records_table() gives for the full
        output: what a BAM writer does with cpairs / rec_path_c and the nodes' path tables (grootgpu_index_node_paths)."""
        cp = self.cpairs
        rc = cp["rec_count"].astype(np.int64)
        al = cp[rc > 0]
        nodes, inv = np.unique(al["node"], return_inverse=True)
        width = 256 if self.rec_path_c.dtype == np.uint8 else 65536
        table = np.zeros((len(nodes), width), dtype=np.int64)
        graph_of = np.zeros(len(nodes), dtype=np.int64)
        for i, nd in enumerate(nodes):
            g, ids, pos = index.node_paths(int(nd))
            graph_of[i] = g
            table[i, ids] = pos
        rca = rc[rc > 0]
        row = np.repeat(inv, rca)
        out = np.zeros((self.n_records, 7), dtype=np.int64)
        path = self.rec_path_c.astype(np.int64)
        out[:, 0] = np.repeat(al["read"], rca)
        out[:, 1] = graph_of[row]
        out[:, 2] = path
        out[:, 3] = table[row, path] + np.repeat((al["offset_flags"] & CPAIR_OFFSET_MASK).astype(np.int64), rca)
        first = np.zeros(self.n_records, dtype=bool)
        first[(np.cumsum(rca) - rca)] = True
        out[:, 4] = np.where(first, 0, 0x100) | np.repeat(np.where(al["offset_flags"] & CPAIR_REVERSE, 0x10, 0), rca)
        out[:, 5] = np.repeat((al["offset_flags"] & CPAIR_CLIP_START) != 0, rca)
        out[:, 6] = np.repeat((al["offset_flags"] & CPAIR_CLIP_END) != 0, rca)
        return out

    def records_table(self):
        """(read, graph, path, pos, flags, startClip, endClip, seqLength-less) rows in (read, graph, emission) order,
        the same layout the oracle reports (flags: 0x100 on all but the first record of a pair, 0x10 when reverse)."""
        out = np.zeros((self.n_records, 7), dtype=np.int64)
        o = 0
        for p in self.pairs:
            n = int(p["rec_count"])
            if n == 0:
                continue
            b = int(p["rec_begin"])
            flags = np.full(n, 0x10 if p["reverse"] else 0, dtype=np.int64)
            if n > 1:
                flags[1:] |= 0x100
            out[o:o + n, 0] = p["read"]
            out[o:o + n, 1] = p["graph"]
            out[o:o + n, 2] = self.rec_path[b:b + n]
            out[o:o + n, 3] = self.rec_pos[b:b + n]
            out[o:o + n, 4] = flags
            out[o:o + n, 5] = p["clip_start"]
            out[o:o + n, 6] = p["clip_end"]
            o += n
        return out


class Index:
    """The graph store + containment index on one GPU (grootgpu_index)."""

    def __init__(self, handle, device):
        self.h = C.c_void_p(handle)
        self.device = device
        self._info = None

    # -- bring-up ---------------------------------------------------------------------------------
    @classmethod
    def build(cls, msa_dir=None, msa_files=None, k=31, S=21, w=100, num_part=8, max_k=4, device=0):
        """`groot index -m <msa_dir> -k -s -w -x -y` (cmd/index.go:45-51; src/pipeline/index.go:37-211)."""
        p = IndexParams(k, S, w, num_part, max_k)
        h = C.c_void_p()
        if msa_files is not None:
            arr = (C.c_char_p * len(msa_files))(*[f.encode() for f in msa_files])
            _check(lib().grootgpu_index_build(arr, len(msa_files), C.byref(p), device, C.byref(h)))
        else:
            _check(lib().grootgpu_index_build_dir(msa_dir.encode(), C.byref(p), device, C.byref(h)))
        return cls(h.value, device)

    @classmethod
    def load(cls, path, device=0):
        h = C.c_void_p()
        _check(lib().grootgpu_index_load(path.encode(), device, C.byref(h)))
        return cls(h.value, device)

    @classmethod
    def load_gob(cls, gg_path, lshe_path, device=0):
        """The reference's own index files (groot.gg + groot.lshe, Go gob): cmd/align.go:94-107."""
        h = C.c_void_p()
        _check(lib().grootgpu_index_load_gob(gg_path.encode(), lshe_path.encode(), device, C.byref(h)))
        return cls(h.value, device)

    def save(self, path):
        _check(lib().grootgpu_index_save(self.h, path.encode()))

    def save_gob(self, gg_path, lshe_path):
        """The index as the reference's own files (Info.Dump + ContainmentIndex.Dump, runtime.go:64-73, lshe.go:72-92)."""
        _check(lib().grootgpu_index_save_gob(self.h, gg_path.encode(), lshe_path.encode()))

    def close(self):
        if self.h:
            lib().grootgpu_index_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        if self._info is None:
            i = IndexInfo()
            _check(lib().grootgpu_index_get_info(self.h, C.byref(i)))
            self._info = dict(k=i.params.kmer_size, S=i.params.sketch_size, w=i.params.window_size, num_part=i.params.num_part,
                              max_k=i.params.max_k, graphs=i.n_graphs, masked=i.n_masked_graphs, paths=i.n_paths, nodes=i.n_nodes,
                              path_bases=i.n_path_bases, raw_windows=i.n_raw_windows, windows=i.n_windows,
                              max_merge_span=i.max_merge_span, max_paths_per_graph=i.max_paths_per_graph)
        return self._info

    def dump_hash(self):
        h = C.c_uint64()
        _check(lib().grootgpu_index_dump_hash(self.h, C.byref(h)))
        return h.value

    def dump_file(self, path):
        _check(lib().grootgpu_index_dump_file(self.h, path.encode()))

    def ref(self, graph, path):
        name, ln = C.c_char_p(), C.c_int32()
        _check(lib().grootgpu_index_ref(self.h, graph, path, C.byref(name), C.byref(ln)))
        return name.value.decode(), ln.value

    def query_params(self, query_kmers, threshold):
        K, L, e = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().grootgpu_index_query_params(self.h, query_kmers, threshold, C.byref(K), C.byref(L), C.byref(e)))
        return K.value, L.value, e.value

    # -- the hot path -----------------------------------------------------------------------------
    def map_reads(self, seqs, off, threshold=0.99, no_align=False, keep_sketches=False, project=False, project_on_device=False, compact=False,
                  fixed_read_len=0):
        """theBoss.mapReads for one batch (src/pipeline/boss.go:108-242): host buffers in, host result out. With
        fixed_read_len the offsets are not handed to the library (reads of one length, back to back)."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        prm = AlignParams(threshold, int(no_align), int(keep_sketches), int(project_on_device), 0, int(compact), int(fixed_read_len))
        raw = BatchResultC()
        _check(lib().grootgpu_align_batch(self.h, seqs.ctypes.data, None if fixed_read_len else off.ctypes.data, len(off) - 1, C.byref(prm), C.byref(raw)))
        res = BatchResult(raw, self.info()["S"], True)
        if project:
            self.project(res, off)
        return res

    def map_reads_raw(self, seq_ptr, off_ptr, n_reads, threshold=0.99, no_align=False, project_on_device=False, compact=False, results_on_device=False,
                      fixed_read_len=0):
        """Same call on raw host pointers (pinned buffers), returning only the C struct: what bench.py times. With
        fixed_read_len the offsets pointer may be None (reads of one length, back to back)."""
        prm = AlignParams(threshold, int(no_align), 0, int(project_on_device), int(results_on_device), int(compact), int(fixed_read_len))
        raw = BatchResultC()
        _check(lib().grootgpu_align_batch(self.h, seq_ptr, off_ptr, n_reads, C.byref(prm), C.byref(raw)))
        return raw

    def map_reads_device(self, d_seq_ptr, d_off_ptr, n_reads, min_len, max_len, threshold=0.99, no_align=False, stream=None, copy_back=False,
                         project_on_device=False, compact=False):
        """Reads already resident in HBM (device pointers as ints)."""
        prm = AlignParams(threshold, int(no_align), 0, int(project_on_device), 0 if copy_back else 1, int(compact), 0)
        raw = BatchResultC()
        _check(lib().grootgpu_align_batch_device(self.h, d_seq_ptr, d_off_ptr, n_reads, min_len, max_len, C.byref(prm), stream,
                                                 C.byref(raw)))
        return BatchResult(raw, self.info()["S"], True) if copy_back else raw

    def node_paths(self, node):
        """(GraphID, PathIDs, Position[pathID]) of a node (src/graph/node.go:13-22): turns a compact record into (Ref, Pos)."""
        g, n = C.c_uint32(), C.c_uint32()
        ids, pos = C.POINTER(C.c_uint32)(), C.POINTER(C.c_int32)()
        _check(lib().grootgpu_index_node_paths(self.h, node, C.byref(g), C.byref(ids), C.byref(pos), C.byref(n)))
        return g.value, np.ctypeslib.as_array(ids, shape=(n.value,)).copy(), np.ctypeslib.as_array(pos, shape=(n.value,)).copy()

    def project(self, res, off):
        """Ordered replay of GrootGraph.IncrementSubPath (src/graph/graph.go:401-451)."""
        off = np.ascontiguousarray(off, dtype=np.uint64)
        raw = res.raw if isinstance(res, BatchResult) else res
        _check(lib().grootgpu_project_batch(self.h, C.byref(raw), off.ctypes.data))

    def weights(self):
        i = self.info()
        kf = np.zeros(i["nodes"], dtype=np.float64)
        kt = np.zeros(i["graphs"], dtype=np.uint64)
        _check(lib().grootgpu_weights(self.h, kf.ctypes.data, kt.ctypes.data))
        return kf, kt

    def reset_weights(self):
        _check(lib().grootgpu_reset_weights(self.h))

    def prune(self, min_kmer_coverage):
        kept = np.zeros(self.info()["graphs"], dtype=np.uint8)
        _check(lib().grootgpu_prune(self.h, min_kmer_coverage, kept.ctypes.data))
        return kept

    def save_gfa(self, graph, path, total_kmers):
        w = C.c_int()
        _check(lib().grootgpu_graph_save_gfa(self.h, graph, path.encode(), total_kmers, C.byref(w)))
        return bool(w.value)


class Comm:
    """One rank of a multi-GPU run (grootgpu_comm): the weight ring and the gather of results to rank 0."""

    def __init__(self, index, comm_id, rank, world_size):
        self.index, self.rank, self.world = index, rank, world_size
        h = C.c_void_p()
        _check(lib().grootgpu_comm_create(index.h, bytes(comm_id), rank, world_size, C.byref(h)))
        self.h = h

    @staticmethod
    def new_id():
        buf = C.create_string_buffer(COMM_ID_BYTES)
        _check(lib().grootgpu_comm_id(buf))
        return buf.raw

    def gather(self, raw, to_host=0):
        """raw: BatchResultC of this rank's last align call (results_on_device). Rank 0 gets the merged batch: with
        to_host = 1 a BatchResult (numpy copies of the host arrays); with 0 (device pointers only) or 2 (host arrays
        copied asynchronously: complete after the next gather / sync) the raw C struct."""
        merged = BatchResultC()
        _check(lib().grootgpu_gather(self.h, C.byref(raw), int(to_host), C.byref(merged)))
        if self.rank != 0:
            return None
        return BatchResult(merged, self.index.info()["S"], True) if int(to_host) == 1 else merged

    def sync(self):
        _check(lib().grootgpu_comm_sync(self.h))

    def close(self):
        if self.h:
            lib().grootgpu_comm_destroy(self.h)
            self.h = None


def graphs_dump(msa_files, dump_path=None, k=31, S=21, w=100, num_part=8, max_k=4):
    """Host-only MSA -> graph step (no GPU needed); returns the FNV-1a-64 hash of the dump."""
    p = IndexParams(k, S, w, num_part, max_k)
    arr = (C.c_char_p * len(msa_files))(*[f.encode() for f in msa_files])
    h = C.c_uint64()
    _check(lib().grootgpu_graphs_dump(arr, len(msa_files), C.byref(p), dump_path.encode() if dump_path else None, C.byref(h)))
    return h.value


def gob_dump(gg_path, lshe_path, dump_path=None):
    """Host-only: decode groot.gg + groot.lshe, return the FNV-1a-64 hash of the canonical dump (and write it)."""
    h = C.c_uint64()
    _check(lib().grootgpu_gob_dump(gg_path.encode(), lshe_path.encode(), dump_path.encode() if dump_path else None, C.byref(h)))
    return h.value


def flat_to_gob(flat_path, gg_path, lshe_path):
    """Host-only: the library's flat index file -> groot.gg + groot.lshe."""
    _check(lib().grootgpu_flat_to_gob(flat_path.encode(), gg_path.encode(), lshe_path.encode()))


def gob_to_flat(gg_path, lshe_path, flat_path):
    """Host-only: groot.gg + groot.lshe -> the library's flat index file."""
    _check(lib().grootgpu_gob_to_flat(gg_path.encode(), lshe_path.encode(), flat_path.encode()))


def query_params_host(query_kmers, threshold, k=31, S=21, w=100, num_part=8, max_k=4):
    """(K, L, eq_min) the LSH Ensemble optimiser / containment threshold give for a query of that many k-mers."""
    p = IndexParams(k, S, w, num_part, max_k)
    K, L, e = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(lib().grootgpu_query_params_host(C.byref(p), query_kmers, threshold, C.byref(K), C.byref(L), C.byref(e)))
    return K.value, L.value, e.value


def sketch_batch(seqs, off, k, S, device=0):
    """Sequence.RunMinHash(k, S, false, nil) for a batch (src/seqio/seqio.go:40-68)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    n = len(off) - 1
    out = np.zeros((n, S), dtype=np.uint64)
    _check(lib().grootgpu_sketch_batch(device, seqs.ctypes.data, off.ctypes.data, n, k, S, out.ctypes.data))
    return out
