"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; skipped on a
one-GPU box). One index handle + one host thread per GPU in ONE process, as the `groot-b200 align --devices` driver runs
it; bench.py --gpus N drives the same entry points from one process per GPU.

The bar: the merged result of N ranks — hits, pairs, records, counters and the
ORDER-DEPENDENT f64 graph weights — equals the one-GPU result on the whole batch bit for bit."""
import os
import threading

import numpy as np
import pytest

from groot_b200 import api, synth
from groot_b200 import distributed as gd

pytestmark = pytest.mark.gpu


def _run_ranks(world, fn):
    """fn(rank) on one thread per rank; re-raises the first failure."""
    errs = [None] * world
    out = [None] * world

    def body(r):
        try:
            out[r] = fn(r)
        except BaseException as e:     # noqa: BLE001
            errs[r] = e
    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    assert not any(t.is_alive() for t in th), "a rank hung"
    for e in errs:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_n_ranks_equal_one_rank(db_dirs, world, monkeypatch):
    import torch
    if api.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    monkeypatch.setenv("GROOTGPU_CHUNK_READS", "17000")      # several chunks per shard: the ring opens with a call's first chunk and closes with its last
    d = db_dirs["arg-annot.90"]
    seqs = synth.db_sequences(d)
    L = 100
    batches = [synth.synth_reads(120_001, L, seqs, seed=42), synth.synth_reads(90_000, L, seqs, seed=43), synth.synth_reads(64, L, seqs, seed=44)]
    # ---- one GPU, the whole batches, one after the other (the weights carry over)
    one = api.Index.build(msa_dir=d, k=31, S=21, w=100, device=0)
    ref = [one.map_reads(blob, off, 0.99, project_on_device=True) for blob, off in batches]
    ref_w = one.weights()
    assert ref[0].counts["mapped"] > 50_000 and (ref_w[0] > 0).sum() > 1000
    one.close()
    # ---- N ranks
    idx = [api.Index.build(msa_dir=d, k=31, S=21, w=100, device=r) for r in range(world)]
    cid = api.Comm.new_id()
    merged = {}

    def rank_main(r):
        comm = api.Comm(idx[r], cid, r, world)
        dev = torch.device("cuda", r)
        for bi, (blob, off) in enumerate(batches):
            lo, hi = gd.shard_bounds(len(off) - 1, world, r)
            sblob, soff = blob[lo * L:hi * L], off[lo:hi + 1] - off[lo]
            for fmt in ("full", "compact"):
                if bi == 1:       # reads resident in HBM, one shot
                    d_seq = torch.zeros(len(sblob) + 64, dtype=torch.uint8, device=dev)
                    d_seq[: len(sblob)].copy_(torch.from_numpy(sblob))
                    d_off = torch.from_numpy(soff.astype(np.uint32).view(np.int32)).to(dev)
                    torch.cuda.synchronize(dev)
                    raw = idx[r].map_reads_device(d_seq.data_ptr(), d_off.data_ptr(), hi - lo, L, L, 0.99, project_on_device=(fmt == "full"), compact=(fmt == "compact"))
                else:             # host buffers through the chunked pipeline, results kept on the device
                    sb, so = np.ascontiguousarray(sblob), np.ascontiguousarray(soff)
                    raw = idx[r].map_reads_raw(sb.ctypes.data, so.ctypes.data, hi - lo, 0.99, project_on_device=(fmt == "full"), compact=(fmt == "compact"),
                                               results_on_device=True)
                m = comm.gather(raw, to_host=True)
                if r == 0:
                    merged[(bi, fmt)] = m
        comm.sync()
        w = idx[r].weights()
        comm.close()
        return w
    ws = _run_ranks(world, rank_main)
    for bi in range(len(batches)):
        full, comp, want = merged[(bi, "full")], merged[(bi, "compact")], ref[bi]
        assert full.counts == want.counts and comp.counts == want.counts
        for k in ("hit_off", "hits", "pairs", "rec_path", "rec_pos"):
            assert np.array_equal(getattr(full, k), getattr(want, k)), (bi, k)
        assert np.array_equal(comp.decode_compact(idx[0]), want.records_table()), bi
    # the graph weights: rank 0 holds the all-rank result, bit-identical to the one-GPU run; the other ranks hold nothing
    assert np.array_equal(ws[0][0], ref_w[0]) and np.array_equal(ws[0][1], ref_w[1])
    for r in range(1, world):
        assert not ws[r][0].any() and not ws[r][1].any()
    for ix in idx:
        ix.close()
