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
	"unsafe"
)

// Index wraps one grootgpu_index handle: graph store + containment index + device copies + workspaces of ONE GPU.
// Calls on an Index must be serialised by the caller (one goroutine per GPU); different Index values are independent.
type Index struct{ h *C.grootgpu_index }

// Params mirrors the index parameters `groot index` stores in groot.gg (src/pipeline/runtime.go:15-27).
type Params struct{ KmerSize, SketchSize, WindowSize, NumPart, MaxK uint32 }

func lastErr(rc C.int) error {
	if rc == 0 {
		return nil
	}
	return errors.New(C.GoString(C.grootgpu_last_error()))
}

// Build replaces MSAconverter -> GraphSketcher -> SketchIndexer (src/pipeline/index.go:37-211); graph i == msaPaths[i].
func Build(msaPaths []string, p Params, device int) (*Index, error) {
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
	cp := C.CString(path)
	defer C.free(unsafe.Pointer(cp))
	var h *C.grootgpu_index
	if err := lastErr(C.grootgpu_index_load(cp, C.int(device), &h)); err != nil {
		return nil, err
	}
	return &Index{h}, nil
}

func (ix *Index) Save(path string) error {
	cp := C.CString(path)
	defer C.free(unsafe.Pointer(cp))
	return lastErr(C.grootgpu_index_save(ix.h, cp))
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
	Pairs                                     []Pair
	RecPath                                   []uint32
	RecPos                                    []int32
}

// AlignBatch replaces the per-read loop of theBoss.mapReads fused with the graph minions (src/pipeline/boss.go:134-203,
// src/pipeline/graphminion.go:46-102). seq = read bases back to back, off[i]..off[i+1] = read i (len(off) == reads+1).
// With projectOnDevice the ordered graph weighting (GrootGraph.IncrementSubPath, src/graph/graph.go:401-451) runs on
// the device as part of the call. Any batch size: the library streams the batch through the GPU in chunks.
func (ix *Index) AlignBatch(seq []byte, off []uint64, threshold float64, noAlign, projectOnDevice bool) (*Result, error) {
	if len(off) < 2 {
		return &Result{}, nil
	}
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
