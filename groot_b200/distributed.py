"""Host-side plumbing of the multi-GPU `groot align` path (one process per GPU).

Reads are independent units (the reference already fans them over NumProc workers, src/pipeline/boss.go:134-203), so
each rank maps a contiguous shard against a replicated index with NO data-path collective. What crosses NVLink is done
by libgrootgpu itself over NCCL (include/grootgpu.h "multi-GPU": grootgpu_gather, the weight ring); torch.distributed
only carries the 256-byte communicator id from rank 0 to the other ranks (NCCL on the GPU box, gloo in the CPU tests).
merge_results() restates on the host what the device-side merge of grootgpu_gather does; the tests compare against it.
"""
import numpy as np
import torch
import torch.distributed as dist

RESULT_KEYS = ("hit_off", "hits", "pairs", "rec_path", "rec_pos")


def shard_bounds(n_total, world_size, rank):
    """Contiguous 1/world_size slices of the read array."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_comm_id(make_id, group=None, device=None, nbytes=256):
    """Rank 0 calls make_id() (api.Comm.new_id); every rank returns the same bytes."""
    rank = dist.get_rank(group)
    buf = torch.zeros(nbytes, dtype=torch.uint8)
    if rank == 0:
        raw = make_id()
        assert len(raw) == nbytes
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def merge_results(per_rank, read_base):
    """Merge in global read order: per_rank[r] = dict of numpy arrays (hit_off u32 [n_r + 1], hits u32, pairs structured,
    rec_path, rec_pos) of the shard starting at global read read_base[r]. Returns one dict with global read indices and
    rebased hit / record offsets — the layout a single-GPU run over the whole batch produces."""
    hit_off, hits, pairs, rec_path, rec_pos = [np.zeros(1, dtype=np.uint64)], [], [], [], []
    h_base = r_base = 0
    for r, res in enumerate(per_rank):
        ho = res["hit_off"].astype(np.uint64)
        hit_off.append(ho[1:] + np.uint64(h_base))
        hits.append(res["hits"])
        p = res["pairs"].copy()
        p["read"] += np.uint32(read_base[r])
        p["hit_begin"] += np.uint32(h_base)
        p["rec_begin"] += np.uint32(r_base)
        pairs.append(p)
        rec_path.append(res["rec_path"])
        rec_pos.append(res["rec_pos"])
        h_base += len(res["hits"])
        r_base += len(res["rec_path"])
    return {"hit_off": np.concatenate(hit_off), "hits": np.concatenate(hits), "pairs": np.concatenate(pairs),
            "rec_path": np.concatenate(rec_path), "rec_pos": np.concatenate(rec_pos)}
