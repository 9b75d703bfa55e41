// Package grootgpu is the cgo binding of libgrootgpu.so (include/grootgpu.h) for the Go host of will-rowe/groot.
//
// STATUS: NOT COMPILED OR TESTED IN THIS REPOSITORY'S ENVIRONMENT — the build image has no Go toolchain
// (`go version`: not found). The same call sequence is built and tested from C++ (groot_b200/csrc/host/pipeline.cpp)
// and from Python/ctypes (groot_b200/api.py); this file is the binding a maintainer drops into the reference tree
// (e.g. src/grootgpu/) so that theBoss.mapReads (src/pipeline/boss.go:108-242) can call the GPU path; see
// INTEGRATION.md §2 for the replacement body of mapReads.
//
// Build: CGO_CFLAGS="-I<repo>/include" CGO_LDFLAGS="-L<repo>/groot_b200 -lgrootgpu -Wl,-rpath,<repo>/groot_b200" go build ./...
package grootgpu

/*
#cgo LDFLAGS: -lgrootgpu
#include <stdlib.h>
#include "grootgpu.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"
)

// Index wraps one grootgpu_index handle: graph store + containment index + device copies + workspaces of ONE GPU.
// Calls on an Index must be serialised by the caller (one goroutine per GPU); different Index values are independent.
type Index struct{ h *C.grootgpu_index }

// Params mirrors the index parameters `groot index` stores in groot.gg (src/pipeline/runtime.go:15-27).
type Params struct{ KmerSize, SketchSize, WindowSize, NumPart, MaxK uint32 }

// The library keeps its error text per OS thread (thread_local). A goroutine may migrate between the failing call and
// grootgpu_last_error(), so every call below runs between pin() and unpin() (runtime.LockOSThread), and lastErr is read
// before unpinning.
func pin()   { runtime.LockOSThread() }
func unpin() { runtime.UnlockOSThread() }

func lastErr(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return errors.New(C.GoString(C.grootgpu_last_error()))
}

// Build replaces MSAconverter -> GraphSketcher -> SketchIndexer (src/pipeline/index.go:37-211); graph i == msaPaths[i].
func Build(msaPaths []string, p Params, device int) (*Index, error) {
	if len(msaPaths) == 0 {
		return nil, errors.New("no MSA files")
	}
	pin()
	defer unpin()
	cs := make([]*C.char, len(msaPaths))
	for i, s := range msaPaths {
		cs[i] = C.CString(s)
		defer C.free(unsafe.Pointer(cs[i]))
	}
	prm := C.grootgpu_index_params{kmer_size: C.uint32_t(p.KmerSize), sketch_size: C.uint32_t(p.SketchSize),
		window_size: C.uint32_t(p.WindowSize), num_part: C.uint32_t(p.NumPart), max_k: C.uint32_t(p.MaxK)}
	var h *C.grootgpu_index
	if err := lastErr(C.grootgpu_index_build(&cs[0], C.uint32_t(len(cs)), &prm, C.int(device), &h)); err != nil {
		return nil, err
	}
	return &Index{h}, nil
}

// Load replaces Info.Load + ContainmentIndex.Load (cmd/align.go:94-107) for the library's own index file.
func Load(path string, device int) (*Index, error) {
	pin()
	defer unpin()
	cp := C.CString(path)
	defer C.free(unsafe.Pointer(cp))
	var h *C.grootgpu_index
	if err := lastErr(C.grootgpu_index_load(cp, C.int(device), &h)); err != nil {
		return nil, err
	}
	return &Index{h}, nil
}

func (ix *Index) Save(path string) error {
	pin()
	defer unpin()
	cp := C.CString(path)
	defer C.free(unsafe.Pointer(cp))
	return lastErr(C.grootgpu_index_save(ix.h, cp))
}

// SaveGob writes the index as the reference's own groot.gg + groot.lshe (Info.Dump + ContainmentIndex.Dump,
// src/pipeline/runtime.go:64-73, src/lshe/lshe.go:72-92).
func (ix *Index) SaveGob(ggPath, lshePath string) error {
	pin()
	defer unpin()
	cg, cl := C.CString(ggPath), C.CString(lshePath)
	defer C.free(unsafe.Pointer(cg))
	defer C.free(unsafe.Pointer(cl))
	return lastErr(C.grootgpu_index_save_gob(ix.h, cg, cl))
}

func (ix *Index) Close() { C.grootgpu_index_destroy(ix.h); ix.h = nil }

// Pair is one (read, graph) unit == one graphMinionPair (src/pipeline/graphminion.go:14-17).
type Pair struct {
	Read, Graph, HitBegin, HitCount, NIncremented, RecBegin, RecCount uint32
	Reverse, ClipStart, ClipEnd, Stage                                uint8
}

// Result views the arrays of a grootgpu_batch_result. They are owned by the Index and stay valid until its next
// AlignBatch call: materialise the sam.Records (src/graph/alignment.go:114-156) before calling again.
type Result struct {
	Received, Mapped, Multimapped, Alignments uint64
	Hits                                      []uint32 // window ids; Pair.HitBegin / HitCount index into it
	Pairs                                     []Pair
	RecPath                                   []uint32
	RecPos                                    []int32
}

// AlignBatch replaces the per-read loop of theBoss.mapReads fused with the graph minions (src/pipeline/boss.go:134-203,
// src/pipeline/graphminion.go:46-102). seq = read bases back to back, off[i]..off[i+1] = read i (len(off) == reads+1).
// With projectOnDevice the ordered graph weighting (GrootGraph.IncrementSubPath, src/graph/graph.go:401-451) runs on
// the device as part of the call. Any batch size: the library streams the batch through the GPU in chunks.
func (ix *Index) AlignBatch(seq []byte, off []uint64, threshold float64, noAlign, projectOnDevice bool) (*Result, error) {
	if len(off) < 2 || len(seq) == 0 {
		return &Result{}, nil
	}
	pin()
	defer unpin()
	prm := C.grootgpu_align_params{containment_threshold: C.double(threshold)}
	if noAlign {
		prm.no_align = 1
	}
	if projectOnDevice {
		prm.project_on_device = 1
	}
	var res C.grootgpu_batch_result
	rc := C.grootgpu_align_batch(ix.h, (*C.uint8_t)(unsafe.Pointer(&seq[0])), (*C.uint64_t)(unsafe.Pointer(&off[0])),
		C.uint32_t(len(off)-1), &prm, &res)
	if err := lastErr(rc); err != nil {
		return nil, err // GROOTGPU_ERR_SHORT_READ / _BAD_BASE are the reference's panics (boss.go:164-166, seqio.go:122)
	}
	out := &Result{Received: uint64(res.received), Mapped: uint64(res.mapped), Multimapped: uint64(res.multimapped), Alignments: uint64(res.alignments)}
	if res.n_hits > 0 {
		out.Hits = unsafe.Slice((*uint32)(unsafe.Pointer(res.hits)), int(res.n_hits))
	}
	if res.n_pairs > 0 {
		out.Pairs = unsafe.Slice((*Pair)(unsafe.Pointer(res.pairs)), int(res.n_pairs)) // same 32-byte layout as grootgpu_pair
	}
	if res.n_records > 0 {
		out.RecPath = unsafe.Slice((*uint32)(unsafe.Pointer(res.rec_path)), int(res.n_records))
		out.RecPos = unsafe.Slice((*int32)(unsafe.Pointer(res.rec_pos)), int(res.n_records))
	}
	return out, nil
}

// Weights returns KmerFreq of every node (graphs ascending, SortedNodes order) and KmerTotal per graph
// (src/graph/node.go:21, src/graph/graph.go:26) so that GraphPruner and SaveGraphAsGFA run unchanged on the Go side.
func (ix *Index) Weights() ([]float64, []uint64, error) {
	pin()
	defer unpin()
	var gi C.grootgpu_index_info
	if err := lastErr(C.grootgpu_index_get_info(ix.h, &gi)); err != nil {
		return nil, nil, err
	}
	kf := make([]float64, int(gi.n_nodes))
	kt := make([]uint64, int(gi.n_graphs))
	if len(kf) == 0 || len(kt) == 0 {
		return kf, kt, nil
	}
	err := lastErr(C.grootgpu_weights(ix.h, (*C.double)(unsafe.Pointer(&kf[0])), (*C.uint64_t)(unsafe.Pointer(&kt[0]))))
	return kf, kt, err
}

// Ref returns the @SQ name and length of path `path` of graph `graph` (Store.GetSAMrefs, src/graph/graphio.go:141-154).
func (ix *Index) Ref(graph, path uint32) (string, int, bool) {
	var name *C.char
	var length C.int32_t
	if C.grootgpu_index_ref(ix.h, C.uint32_t(graph), C.uint32_t(path), &name, &length) != 0 {
		return "", 0, false
	}
	return C.GoString(name), int(length), true
}

// CPair is the compact, BAM-oriented form of a pair (grootgpu_cpair): start locus + flags + record count; the path ids of
// its records are the next RecCount entries of CompactResult.RecPath. Pos of a record = Position[path] of node Node
// (NodePaths) + Offset() (src/graph/alignment.go:296).
type CPair struct{ Read, Node, OffsetFlags, RecCount uint32 }

func (p CPair) Offset() int32   { return int32(p.OffsetFlags & C.GROOTGPU_CPAIR_OFFSET_MASK) }
func (p CPair) Reverse() bool   { return p.OffsetFlags&C.GROOTGPU_CPAIR_REVERSE != 0 }
func (p CPair) ClipStart() bool { return p.OffsetFlags&C.GROOTGPU_CPAIR_CLIP_START != 0 }
func (p CPair) ClipEnd() bool   { return p.OffsetFlags&C.GROOTGPU_CPAIR_CLIP_END != 0 }

// CompactResult views a batch result produced with compact_records = 1 (about a fifth of the bytes of Result).
type CompactResult struct {
	Received, Mapped, Multimapped, Alignments uint64
	Pairs                                     []CPair
	RecPath8                                  []uint8  // when no graph has more than 256 paths
	RecPath16                                 []uint16 // otherwise
}

// AlignBatchCompact is AlignBatch with the ordered graph weighting on the device and the compact output.
// fixedReadLen > 0 declares that every read has that many bases (then the offsets do not travel to the device).
func (ix *Index) AlignBatchCompact(seq []byte, off []uint64, threshold float64, noAlign bool, fixedReadLen uint32) (*CompactResult, error) {
	if len(off) < 2 || len(seq) == 0 {
		return &CompactResult{}, nil
	}
	pin()
	defer unpin()
	prm := C.grootgpu_align_params{containment_threshold: C.double(threshold), project_on_device: 1, compact_records: 1,
		fixed_read_len: C.uint32_t(fixedReadLen)}
	if noAlign {
		prm.no_align = 1
	}
	var res C.grootgpu_batch_result
	rc := C.grootgpu_align_batch(ix.h, (*C.uint8_t)(unsafe.Pointer(&seq[0])), (*C.uint64_t)(unsafe.Pointer(&off[0])), C.uint32_t(len(off)-1), &prm, &res)
	if err := lastErr(rc); err != nil {
		return nil, err
	}
	out := &CompactResult{Received: uint64(res.received), Mapped: uint64(res.mapped), Multimapped: uint64(res.multimapped), Alignments: uint64(res.alignments)}
	if res.n_pairs > 0 {
		out.Pairs = unsafe.Slice((*CPair)(unsafe.Pointer(res.cpairs)), int(res.n_pairs))
	}
	if res.n_records > 0 {
		if res.rec_path_bytes == 1 {
			out.RecPath8 = unsafe.Slice((*uint8)(res.rec_path_c), int(res.n_records))
		} else {
			out.RecPath16 = unsafe.Slice((*uint16)(res.rec_path_c), int(res.n_records))
		}
	}
	return out, nil
}

// NodePaths returns GraphID, PathIDs (ascending) and Position[pathID] of a node (src/graph/node.go:13-22).
func (ix *Index) NodePaths(node uint32) (graph uint32, ids []uint32, pos []int32, err error) {
	pin()
	defer unpin()
	var g, n C.uint32_t
	var pi *C.uint32_t
	var pp *C.int32_t
	if err = lastErr(C.grootgpu_index_node_paths(ix.h, C.uint32_t(node), &g, &pi, &pp, &n)); err != nil {
		return
	}
	return uint32(g), unsafe.Slice((*uint32)(unsafe.Pointer(pi)), int(n)), unsafe.Slice((*int32)(unsafe.Pointer(pp)), int(n)), nil
}

// Comm is one rank of a multi-GPU run (grootgpu_comm): N Index values on N GPUs, one goroutine each (locked to its OS
// thread for the lifetime of the run); rank 0 receives the merged batches and the graph weights.
type Comm struct{ h *C.grootgpu_comm }

func NewCommID() ([]byte, error) {
	pin()
	defer unpin()
	id := make([]byte, C.GROOTGPU_COMM_ID_BYTES)
	err := lastErr(C.grootgpu_comm_id((*C.uint8_t)(unsafe.Pointer(&id[0]))))
	return id, err
}

func NewComm(ix *Index, id []byte, rank, world int) (*Comm, error) {
	pin()
	defer unpin()
	var h *C.grootgpu_comm
	if err := lastErr(C.grootgpu_comm_create(ix.h, (*C.uint8_t)(unsafe.Pointer(&id[0])), C.int(rank), C.int(world), &h)); err != nil {
		return nil, err
	}
	return &Comm{h}, nil
}

func (c *Comm) Sync() error { pin(); defer unpin(); return lastErr(C.grootgpu_comm_sync(c.h)) }
func (c *Comm) Close()      { C.grootgpu_comm_destroy(c.h); c.h = nil }
