"""world_size-2 gloo test of the multi-GPU host logic (groot_b200/distributed.py): the communicator id reaches every
rank, reads are sharded into contiguous slices, every rank maps its slice independently (here: with the CPU oracle
standing in for the per-rank GPU result), the per-rank result arrays are brought to rank 0, and the merge — the host
restatement of what grootgpu_gather does on the device — must reproduce what a single rank computes on the whole
read set. (The NCCL path itself runs on the GPU box: tests/test_gpu_multi.py, bench.py --gpus N.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_result_arrays(idx, blob, off):
    from groot_b200.api import PAIR_DTYPE
    res = idx.map_reads(blob, off, 0.99)
    pairs = np.zeros(len(res.pairs), dtype=PAIR_DTYPE)
    pairs["read"], pairs["graph"] = res.pairs[:, 0], res.pairs[:, 1]
    pairs["n_incremented"], pairs["rec_count"] = res.pairs[:, 2], res.pairs[:, 3]
    pairs["rec_begin"] = np.concatenate([[0], np.cumsum(res.pairs[:, 3])[:-1]]) if len(pairs) else 0
    # hit_begin / hit_count from the per-read hit lists
    hit_begin = np.zeros(len(pairs), dtype=np.uint32)
    hit_count = np.zeros(len(pairs), dtype=np.uint32)
    by_read = {}
    for i, r in enumerate(res.pairs[:, 0]):
        by_read.setdefault(int(r), []).append(i)
    for r, plist in by_read.items():
        b = int(res.hit_off[r]); e = int(res.hit_off[r + 1])
        # single-graph reads get their exact slice; multi-graph reads keep 0 (the merge arithmetic is what is under test)
        if len(plist) == 1:
            hit_begin[plist[0]], hit_count[plist[0]] = b, e - b
    pairs["hit_begin"], pairs["hit_count"] = hit_begin, hit_count
    recs = res.records
    pairs["reverse"] = 0
    return {"hit_off": res.hit_off.astype(np.uint32), "hits": res.hits.astype(np.uint32), "pairs": pairs,
            "rec_path": recs[:, 2].astype(np.uint32), "rec_pos": recs[:, 3].astype(np.int32)}


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    from groot_b200 import distributed as gd
    from oracle import pyoracle as po
    from tests.util import load_fastq, pack_reads
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = po.Index(msa_files=[os.path.join(ROOT, "data", "graph", "test-genes.msa")], k=51, S=30, w=100)
        names, seqs, quals = load_fastq(os.path.join(ROOT, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq"))
        seqs = seqs[:401]                                            # odd count: uneven shards
        lo, hi = gd.shard_bounds(len(seqs), world, rank)
        blob, off = pack_reads(seqs[lo:hi])
        local = _oracle_result_arrays(idx, blob, off)
        # the 256-byte communicator id travels from rank 0 to everybody
        cid = gd.broadcast_comm_id(lambda: bytes(range(256)))
        assert cid == bytes(range(256))
        tens = {k: torch.from_numpy(np.ascontiguousarray(v).view(np.uint8).reshape(-1).copy()) for k, v in local.items()}
        sizes = torch.tensor([tens[k].numel() for k in gd.RESULT_KEYS], dtype=torch.int64)
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes)
        gathered = None
        if rank == 0:
            gathered = [tens]
            for r in range(1, world):
                bufs = {k: torch.empty(int(all_sizes[r][i]), dtype=torch.uint8) for i, k in enumerate(gd.RESULT_KEYS)}
                for k in gd.RESULT_KEYS:
                    if bufs[k].numel():
                        dist.recv(bufs[k], src=r)
                gathered.append(bufs)
        else:
            for k in gd.RESULT_KEYS:
                if tens[k].numel():
                    dist.send(tens[k], dst=0)
        if rank == 0:
            from groot_b200.api import PAIR_DTYPE
            dt = {"hit_off": np.uint32, "hits": np.uint32, "pairs": PAIR_DTYPE, "rec_path": np.uint32, "rec_pos": np.int32}
            per_rank = [{k: g[k].numpy().view(dt[k]) for k in gd.RESULT_KEYS} for g in gathered]
            merged = gd.merge_results(per_rank, [gd.shard_bounds(len(seqs), world, r)[0] for r in range(world)])
            blob_all, off_all = pack_reads(seqs)
            whole = _oracle_result_arrays(idx, blob_all, off_all)
            ok = (np.array_equal(merged["hit_off"], whole["hit_off"].astype(np.uint64)) and np.array_equal(merged["hits"], whole["hits"])
                  and np.array_equal(merged["rec_path"], whole["rec_path"]) and np.array_equal(merged["rec_pos"], whole["rec_pos"])
                  and np.array_equal(merged["pairs"], whole["pairs"]))
            open(os.path.join(tmp, "ok"), "w").write("1" if ok else "0")
        else:
            assert gathered is None
    finally:
        dist.destroy_process_group()


def test_shard_bounds_partition():
    from groot_b200 import distributed as gd
    for n in (0, 1, 7, 100, 10_000_001):
        for w in (1, 2, 3, 8):
            b = [gd.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_gather_and_merge_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
