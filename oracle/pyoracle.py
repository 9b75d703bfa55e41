"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/liboracle.so, the CPU restatement of the reference's `groot align`
path (see the headers of oracle/*.hpp for the reference file:line each function follows).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under groot_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".hpp", ".cpp"))]
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        u8p, u32p, u64p, i32p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_double))
        L.oracle_ntf64.restype = C.c_uint64
        L.oracle_ntf64.argtypes = [C.c_char_p, C.c_uint]
        L.oracle_ntr64.restype = C.c_uint64
        L.oracle_ntr64.argtypes = [C.c_char_p, C.c_uint]
        L.oracle_sketch.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, u64p]
        L.oracle_kmer_hashes.restype = C.c_int64
        L.oracle_kmer_hashes.argtypes = [C.c_char_p, C.c_uint64, C.c_int, u64p]
        L.oracle_multi_hash.argtypes = [C.c_uint64, C.c_int, C.c_int, u64p]
        L.oracle_revcomp.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
        L.oracle_basecheck.argtypes = [C.c_char_p, C.c_uint64]
        L.oracle_qualtrim.restype = C.c_uint64
        L.oracle_qualtrim.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_int]
        L.oracle_optimal_kl.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_eq_min.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
        L.oracle_containment.restype = C.c_double
        L.oracle_containment.argtypes = [u64p, u64p, C.c_int, C.c_int, C.c_int]
        L.oracle_msa2gfa_text.restype = C.c_int64
        L.oracle_msa2gfa_text.argtypes = [C.c_char_p, C.c_char_p, C.c_int64]
        L.oracle_gfa_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                       u64p, C.c_int, i32p, C.c_int, C.c_char_p, C.c_int]
        L.oracle_index_build_files.restype = C.c_void_p
        L.oracle_index_build_files.argtypes = [C.POINTER(C.c_char_p), C.c_int] + [C.c_int] * 5 + [C.c_char_p, C.c_int]
        L.oracle_index_build_dir.restype = C.c_void_p
        L.oracle_index_build_dir.argtypes = [C.c_char_p] + [C.c_int] * 5 + [C.c_char_p, C.c_int]
        L.oracle_index_free.argtypes = [C.c_void_p]
        L.oracle_index_stats.argtypes = [C.c_void_p, u64p]
        L.oracle_index_dump_hash.restype = C.c_uint64
        L.oracle_index_dump_hash.argtypes = [C.c_void_p]
        L.oracle_index_dump_file.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_index_window_sketches.argtypes = [C.c_void_p, u64p]
        L.oracle_index_params.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_index_num_nodes.restype = C.c_uint64
        L.oracle_index_num_nodes.argtypes = [C.c_void_p]
        L.oracle_index_weights.argtypes = [C.c_void_p, f64p, u64p]
        L.oracle_index_reset_weights.argtypes = [C.c_void_p]
        L.oracle_prune_paths.restype = C.c_int64
        L.oracle_prune_paths.argtypes = [C.c_void_p, C.c_double, C.c_char_p, C.c_int64]
        L.oracle_gfa_text.restype = C.c_int64
        L.oracle_gfa_text.argtypes = [C.c_void_p, C.c_uint32, C.c_long, C.c_char_p, C.c_int64]
        L.oracle_map_reads.restype = C.c_void_p
        L.oracle_map_reads.argtypes = [C.c_void_p, u8p, u64p, C.c_uint32, C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.oracle_result_free.argtypes = [C.c_void_p]
        L.oracle_result_counts.argtypes = [C.c_void_p, u64p]
        L.oracle_result_sizes.argtypes = [C.c_void_p, u64p]
        L.oracle_result_hits.argtypes = [C.c_void_p, u64p, u32p]
        L.oracle_result_sketches.argtypes = [C.c_void_p, C.c_int, u64p]
        L.oracle_result_pairs.argtypes = [C.c_void_p, u32p]
        L.oracle_result_records.argtypes = [C.c_void_p, i32p]
        L.oracle_ref_name.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def ntf64(s: bytes, k: int) -> int:
    return lib().oracle_ntf64(s, k)


def ntr64(s: bytes, k: int) -> int:
    return lib().oracle_ntr64(s, k)


def sketch(seq: bytes, k: int, S: int):
    """KHF sketch (src/minhash/khf.go:35-56). Raises ValueError when len(seq) < k (khf.go:38-41)."""
    out = np.zeros(S, dtype=np.uint64)
    if lib().oracle_sketch(seq, len(seq), k, S, _p(out, C.c_uint64)) != 0:
        raise ValueError("sequence shorter than k")
    return out


def kmer_hashes(seq: bytes, k: int):
    out = np.zeros(max(1, len(seq)), dtype=np.uint64)
    n = lib().oracle_kmer_hashes(seq, len(seq), k, _p(out, C.c_uint64))
    if n < 0:
        raise ValueError("sequence shorter than k")
    return out[:n]


def multi_hash(h: int, k: int, S: int):
    out = np.zeros(S, dtype=np.uint64)
    lib().oracle_multi_hash(h, k, S, _p(out, C.c_uint64))
    return out


def revcomp(seq: bytes, qual: bytes):
    s, q = C.create_string_buffer(seq, len(seq)), C.create_string_buffer(qual, len(qual))
    if lib().oracle_revcomp(s, q, len(seq)) != 0:
        raise ValueError("base > 'T'")
    return s.raw, q.raw


def basecheck(seq: bytes) -> bytes:
    s = C.create_string_buffer(seq, len(seq))
    lib().oracle_basecheck(s, len(seq))
    return s.raw


def qualtrim(seq: bytes, qual: bytes, min_qual: int):
    s, q = C.create_string_buffer(seq, len(seq)), C.create_string_buffer(qual, len(qual))
    n = lib().oracle_qualtrim(s, q, len(seq), min_qual)
    return s.raw[:n], q.raw[:n]


def optimal_kl(max_k, max_l, x, q, t):
    K, L = C.c_int(), C.c_int()
    lib().oracle_optimal_kl(max_k, max_l, x, q, t, C.byref(K), C.byref(L))
    return K.value, L.value


def eq_min(S, q_size, x_size, t):
    return lib().oracle_eq_min(S, q_size, x_size, t)


def containment(q, x, q_size, x_size):
    q = np.ascontiguousarray(q, dtype=np.uint64)
    x = np.ascontiguousarray(x, dtype=np.uint64)
    return lib().oracle_containment(_p(q, C.c_uint64), _p(x, C.c_uint64), len(q), q_size, x_size)


def msa2gfa_text(msa_path: str) -> str:
    cap = 64 << 20
    buf = C.create_string_buffer(cap)
    n = lib().oracle_msa2gfa_text(msa_path.encode(), buf, cap)
    if n <= 0:
        raise RuntimeError("msa2gfa failed")
    return buf.raw[:n].decode()


def gfa_align(gfa_path, graph_id, read_seq: bytes, node, offset, merge_span=0, window_size=0, contained_nodes=()):
    """AlignRead on a GFA-loaded graph (fixtures of src/graph/alignment_test.go). Returns (records, path names)."""
    cn = np.asarray(list(contained_nodes), dtype=np.uint64)
    out = np.zeros((4096, 6), dtype=np.int32)
    names = C.create_string_buffer(1 << 20)
    n = lib().oracle_gfa_align(gfa_path.encode(), graph_id, read_seq, node, offset, merge_span, window_size,
                               _p(cn, C.c_uint64), len(cn), _p(out, C.c_int32), 4096, names, 1 << 20)
    if n < 0:
        raise RuntimeError("gfa_align failed")
    return out[:n].copy(), names.value.decode().split("\n")[:-1]


class MapResult:
    """Result of Index.map_reads: hits, (read, graph) pairs, alignment records, counters."""

    def __init__(self, h, n_reads, S):
        L = lib()
        self.n_reads = n_reads
        c = np.zeros(4, dtype=np.uint64)
        L.oracle_result_counts(h, _p(c, C.c_uint64))
        self.counts = dict(received=int(c[0]), mapped=int(c[1]), multimapped=int(c[2]), alignments=int(c[3]))
        sz = np.zeros(3, dtype=np.uint64)
        L.oracle_result_sizes(h, _p(sz, C.c_uint64))
        self.hit_off = np.zeros(n_reads + 1, dtype=np.uint64)
        self.hits = np.zeros(max(1, int(sz[0])), dtype=np.uint32)
        L.oracle_result_hits(h, _p(self.hit_off, C.c_uint64), _p(self.hits, C.c_uint32))
        self.hits = self.hits[: int(sz[0])]
        pairs = np.zeros((max(1, int(sz[1])), 4), dtype=np.uint32)
        L.oracle_result_pairs(h, _p(pairs, C.c_uint32))
        self.pairs = pairs[: int(sz[1])]          # read, graph, numIncremented, numRecords
        recs = np.zeros((max(1, int(sz[2])), 8), dtype=np.int32)
        L.oracle_result_records(h, _p(recs, C.c_int32))
        self.records = recs[: int(sz[2])]         # read, graph, path, pos, flags, startClip, endClip, seqLength
        self.sketches = None
        if S:
            self.sketches = np.zeros((n_reads, S), dtype=np.uint64)
            L.oracle_result_sketches(h, S, _p(self.sketches, C.c_uint64))
        L.oracle_result_free(h)


class Index:
    """CPU restatement of `groot index` output + `groot align` mapping (oracle/pipeline.hpp)."""

    def __init__(self, msa_dir=None, msa_files=None, k=31, S=21, w=100, num_part=8, max_k=4):
        err = C.create_string_buffer(1024)
        if msa_files is not None:
            arr = (C.c_char_p * len(msa_files))(*[f.encode() for f in msa_files])
            self.h = lib().oracle_index_build_files(arr, len(msa_files), k, S, w, num_part, max_k, err, 1024)
        else:
            self.h = lib().oracle_index_build_dir(msa_dir.encode(), k, S, w, num_part, max_k, err, 1024)
        if not self.h:
            raise RuntimeError("oracle index build failed: " + err.value.decode())
        self.k, self.S, self.w, self.num_part, self.max_k = k, S, w, num_part, max_k

    def close(self):
        if self.h:
            lib().oracle_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        o = np.zeros(8, dtype=np.uint64)
        lib().oracle_index_stats(self.h, _p(o, C.c_uint64))
        keys = ["graphs", "masked", "paths", "nodes", "path_bases", "windows", "raw_windows", "max_merge_span"]
        return {k: int(v) for k, v in zip(keys, o)}

    def dump_hash(self) -> int:
        return lib().oracle_index_dump_hash(self.h)

    def dump_file(self, path: str):
        if lib().oracle_index_dump_file(self.h, path.encode()) != 0:
            raise IOError(path)

    def window_sketches(self):
        n = self.stats()["windows"]
        out = np.zeros((n, self.S), dtype=np.uint64)
        lib().oracle_index_window_sketches(self.h, _p(out, C.c_uint64))
        return out

    def params(self, x, q, t):
        K, L = C.c_int(), C.c_int()
        lib().oracle_index_params(self.h, x, q, t, C.byref(K), C.byref(L))
        return K.value, L.value

    def weights(self):
        n = lib().oracle_index_num_nodes(self.h)
        g = self.stats()["graphs"]
        kf = np.zeros(n, dtype=np.float64)
        kt = np.zeros(g, dtype=np.uint64)
        lib().oracle_index_weights(self.h, _p(kf, C.c_double), _p(kt, C.c_uint64))
        return kf, kt

    def reset_weights(self):
        lib().oracle_index_reset_weights(self.h)

    def prune_paths(self, min_kmer_cov: float):
        cap = 16 << 20
        buf = C.create_string_buffer(cap)
        n = lib().oracle_prune_paths(self.h, min_kmer_cov, buf, cap)
        if n < 0:
            raise RuntimeError("buffer too small")
        return buf.raw[:n].decode().split("\n")[:-1]

    def gfa_text(self, graph_id: int, total_kmers: int) -> str:
        cap = 16 << 20
        buf = C.create_string_buffer(cap)
        n = lib().oracle_gfa_text(self.h, graph_id, total_kmers, buf, cap)
        if n < 0:
            raise RuntimeError("buffer too small")
        return buf.raw[:n].decode()

    def ref_name(self, graph_id, path_id):
        buf = C.create_string_buffer(4096)
        ln = C.c_int()
        if lib().oracle_ref_name(self.h, graph_id, path_id, buf, 4096, C.byref(ln)) != 0:
            raise KeyError((graph_id, path_id))
        return buf.value.decode(), ln.value

    def map_reads(self, seqs: np.ndarray, off: np.ndarray, threshold=0.99, no_align=False, threads=1, keep_sketches=False):
        """seqs: uint8 blob, off: uint64[n+1]. Mirrors theBoss.mapReads at -p 1 (boss.go:108-242)."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = len(off) - 1
        err = C.create_string_buffer(1024)
        h = lib().oracle_map_reads(self.h, _p(seqs, C.c_uint8), _p(off, C.c_uint64), n, threshold, int(no_align), threads,
                                   int(keep_sketches), err, 1024)
        if not h:
            raise RuntimeError("oracle map_reads failed: " + err.value.decode())
        return MapResult(h, n, self.S if keep_sketches else 0)
