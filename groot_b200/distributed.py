"""Multi-GPU sharding of `groot align`: reads are independent units (the reference already fans them over
NumProc workers, src/pipeline/boss.go:134-203), so each rank maps its own shard against a replicated index
with NO data-path collective; the only exchange is ONE gather of the per-rank result arrays (pairs, hits,
records) to rank 0, which merges them in global read order, replays the graph weighting and writes the BAM.

torch.distributed is plumbing here (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

RESULT_KEYS = ("hit_off", "hits", "pairs", "rec_path", "rec_pos")


class _DevPtr:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_bytes(ptr, nbytes, device):
    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)


def result_tensors_from_raw(raw, device):
    """Byte views of the device-resident result arrays of a grootgpu_batch_result (results_on_device mode)."""
    n = raw.n_reads
    return {
        "hit_off": device_bytes(raw.d_hit_off, 4 * (n + 1), device),
        "hits": device_bytes(raw.d_hits, 4 * raw.n_hits, device),
        "pairs": device_bytes(raw.d_pairs, 32 * raw.n_pairs, device),
        "rec_path": device_bytes(raw.d_rec_path, 4 * raw.n_records, device),
        "rec_pos": device_bytes(raw.d_rec_pos, 4 * raw.n_records, device),
    }


def shard_bounds(n_total, world_size, rank):
    """Contiguous 1/world_size slices of the read array."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local, dst=0, group=None):
    """local: dict RESULT_KEYS -> 1-D uint8 tensor (same device on every rank). One size exchange
    (all_gather of 5 int64) followed by one batched send/recv of the payloads to `dst`.
    Returns on dst a list (per rank) of dicts of uint8 tensors; None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = local["hits"].device
    sizes = torch.tensor([local[k].numel() for k in RESULT_KEYS], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    if world == 1:
        return [local]
    ops, out = [], None
    if rank == dst:
        out = []
        for r in range(world):
            if r == dst:
                out.append(local)
                continue
            bufs = {k: torch.empty(int(all_sizes[r][i]), dtype=torch.uint8, device=dev) for i, k in enumerate(RESULT_KEYS)}
            for k in RESULT_KEYS:
                if bufs[k].numel():
                    ops.append(dist.P2POp(dist.irecv, bufs[k], r, group=group))
            out.append(bufs)
    else:
        for k in RESULT_KEYS:
            if local[k].numel():
                ops.append(dist.P2POp(dist.isend, local[k], dst, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


class OverlappedGather:
    """The same one-gather-per-batch exchange, taken off the critical path: the batch's result arrays are copied
    (device to device) into staging tensors and sent from there on a side stream while the next batch is already
    being mapped; rank `dst` receives into buffers it keeps across batches. flush() waits for everything in flight.
    Works on any backend whose P2P ops are stream-ordered (NCCL); with gloo it degrades to the blocking gather."""

    def __init__(self, device, dst=0, group=None, slack=1.25):
        self.dev, self.dst, self.group, self.slack = device, dst, group, slack
        self.cuda = device.type == "cuda"
        self.side = torch.cuda.Stream(device=device) if self.cuda else None
        self.staging = {}
        self.recv = {}
        self.last = None

    def _buf(self, cache, key, n):
        t = cache.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(int(n * self.slack) + 64, dtype=torch.uint8, device=self.dev)
            cache[key] = t
        return t[:n]

    def submit(self, local):
        """local: dict RESULT_KEYS -> uint8 tensors that stay valid only until the next batch starts."""
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if not self.cuda:
            self.last = gather_results(local, self.dst, self.group)
            return
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.side)                       # the previous send has left the staging tensors
        staged = {}
        for k in RESULT_KEYS:
            staged[k] = self._buf(self.staging, k, local[k].numel())
            staged[k].copy_(local[k])
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            sizes = torch.tensor([staged[k].numel() for k in RESULT_KEYS], dtype=torch.int64, device=self.dev)
            all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
            dist.all_gather(all_sizes, sizes, group=self.group)
            ops, out = [], None
            if rank == self.dst:
                out = []
                for r in range(world):
                    if r == self.dst:
                        out.append(staged)
                        continue
                    sz = all_sizes[r].tolist()
                    bufs = {k: self._buf(self.recv, (r, k), sz[i]) for i, k in enumerate(RESULT_KEYS)}
                    ops += [dist.P2POp(dist.irecv, bufs[k], r, group=self.group) for k in RESULT_KEYS if bufs[k].numel()]
                    out.append(bufs)
            else:
                ops += [dist.P2POp(dist.isend, staged[k], self.dst, group=self.group) for k in RESULT_KEYS if staged[k].numel()]
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()                             # stream-ordered on the side stream, does not block the host
            self.last = out

    def flush(self):
        if self.cuda:
            self.side.synchronize()
        return self.last


def merge_results(per_rank, read_base):
    """Rank-0 merge in global read order: per_rank[r] = dict of numpy arrays (hit_off u32, hits u32, pairs
    structured, rec_path, rec_pos) of the shard starting at global read read_base[r]. Returns one dict with
    global read indices and rebased hit/record offsets — the same layout a single-GPU run produces."""
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
