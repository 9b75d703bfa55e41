#!/usr/bin/env python
"""bench.py — `groot align` hot path on B200: reads/s for 100 bp synthetic reads vs arg-annot.90.

A "step" is one pass of the hot path (KHF sketch -> LSH Ensemble containment query -> hierarchical exact graph
alignment -> ordered graph weighting) over one batch of synthetic reads per GPU. The headline workload is
BASELINE.json configs[2] (C3: 10 M x 100 bp, arg-annot.90 -w 100, full align path); at N = 1 the line also carries
C2 (seeding only, --noAlign) and C4 (10 M x 150 bp vs card.90 -w 150) under "other_configs".

  value   the batch already resident in HBM, result arrays (the compact BAM-oriented form: what the BAM writer consumes
          and what crosses NVLink and PCIe) left on the device, CUDA events on the launching stream. N > 1: every rank
          maps its own shard, the result arrays are gathered to rank 0 over NCCL / NVLink (grootgpu_gather) and merged
          there, the order-dependent f64 graph weights are chained rank after rank (weight ring) — all of it, and the
          final drain, inside the timed region. At N = 1 the line also gives the step with the full result arrays
          (hit lists, 32-byte pairs, 8-byte records: round 1's output format) as `full_format`.
  e2e     the call a user makes: pinned HOST buffers -> grootgpu_align_batch (H2D, kernels, compact BAM-oriented result)
          -> host. N > 1: + gather to rank 0 and rank 0's device->host copy of the merged batch.
  --impl reference   the CPU restatement of the reference (oracle/, all host threads) on a bounded sample of the
          same workload — the reference itself is Go and cannot be built in this image (DESIGN.md "Oracle").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2|C3|C4] [--extra C2,C4|none] [--reads R]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "groot_align_reads_per_sec_100bp_argannot90"
UNIT = "reads/s"
THRESHOLD = 0.99

# BASELINE.json configs[1..3] (C2, C3, C4). `workload` is the string both arms print.
CONFIGS = {
    "C2": dict(db="arg-annot.90", read_len=100, no_align=True, index=dict(k=31, S=21, w=100, num_part=8, max_k=4),
               workload="10M x 100bp synthetic reads vs arg-annot.90 (-w 100 -k 31 -s 21 -x 8 -y 4, t=0.99), seeding only "
                        "(sketch + LSH Ensemble query, every mapping weighted: --noAlign) [BASELINE.json configs[1]]"),
    "C3": dict(db="arg-annot.90", read_len=100, no_align=False, index=dict(k=31, S=21, w=100, num_part=8, max_k=4),
               workload="10M x 100bp synthetic reads vs arg-annot.90 (-w 100 -k 31 -s 21 -x 8 -y 4, t=0.99), full align path "
                        "(sketch + LSH Ensemble query + exact graph alignment) [BASELINE.json configs[2]]"),
    "C4": dict(db="card.90", read_len=150, no_align=False, index=dict(k=31, S=21, w=150, num_part=8, max_k=4),
               workload="10M x 150bp synthetic reads vs card.90 (-w 150 -k 31 -s 21 -x 8 -y 4, t=0.99), full align path "
                        "(sketch + LSH Ensemble query + exact graph alignment) [BASELINE.json configs[3]]"),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe). nvidia-smi takes
    up to a second to start, so it is launched before the warm-up; every line is stamped on arrival and only the
    samples that fall inside the timed region are used (if the region was too short to catch one, the samples taken
    under the warm-up load are reported instead, and the line says so)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def stats(lines):
            sm, mx, reasons = [], None, set()
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx = float(f[2])
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, reasons
        inside = [x for x in self.lines if self.t_begin is not None and self.t_begin <= x[0] <= (self.t_end or time.time()) + 0.03]
        sm, mx, reasons = stats(inside)
        out = {"window": "timed region"}
        if not sm:
            sm, mx, reasons = stats(self.lines)
            out = {"window": "warm-up + timed region (the timed region was shorter than one sampling interval)"}
        out.update({"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)})
        return out


def prepare_db(db="arg-annot.90"):
    from groot_b200 import synth
    cache = os.path.join(tempfile.gettempdir(), "groot_b200_db_%d" % os.getuid())
    return synth.unpack_db(os.path.join(ROOT, "data", "db", db + ".tar"), cache)


def algorithmic_bytes_per_read(L, S, hits_per_read, pairs_per_read, recs_per_read):
    """seeding L + 8 (off,len) + 32 (one band-table sector) + 8*S*c (candidate sketches) + 4 (hit count)
    + 8*h (hits); full path adds m*(L+16) (re-read read + window meta per mapping) + 24*r... with this library's 8-byte
    records: + 8*r + 32 per pair. c is taken equal to h (at eq_min == S nearly every candidate is a hit)."""
    seed = L + 8 + 32 + 8 * S * hits_per_read + 4 + 8 * hits_per_read
    full = seed + hits_per_read * (L + 16) + 32 * pairs_per_read + 8 * recs_per_read
    return seed, full


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """CPU restatement of the reference (oracle/, kind=port) on all host threads; rank 0 only."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from groot_b200 import synth
    from oracle import pyoracle as po
    cfg = CONFIGS[args.config]
    msa_dir = prepare_db(cfg["db"])
    cores = os.cpu_count() or 1
    t0 = time.time()
    idx = po.Index(msa_dir=msa_dir, **cfg["index"])
    t_index = time.time() - t0
    seqs = synth.db_sequences(msa_dir)
    L = cfg["read_len"]
    probe_n = 20000
    blob, off = synth.synth_reads(probe_n, L, seqs, seed=42)
    t0 = time.time(); idx.map_reads(blob, off, THRESHOLD, no_align=cfg["no_align"], threads=cores); probe = probe_n / (time.time() - t0)
    total_budget = 120.0   # seconds for all steps
    per_step = max(20000, min(2_000_000, int(probe * total_budget / max(1, args.steps + args.warmup))))
    blob, off = synth.synth_reads(per_step, L, seqs, seed=42)
    for _ in range(args.warmup):
        idx.map_reads(blob, off, THRESHOLD, no_align=cfg["no_align"], threads=cores)
    t0 = time.time()
    for _ in range(args.steps):
        idx.reset_weights()
        res = idx.map_reads(blob, off, THRESHOLD, no_align=cfg["no_align"], threads=cores)
    dt = time.time() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": cfg["workload"], "config_id": args.config, "read_len": L, "threshold": THRESHOLD},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d reads/step x %d steps of the workload's synthetic mix (seed 42); oracle index build %.1fs not counted" % (per_step, args.steps, t_index)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "mapped_fraction": res.counts["mapped"] / res.counts["received"],
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
class Ctx:
    """Rank plumbing: torch.distributed is used for barriers / reductions of the timing and to carry the communicator id;
    the data path (gather, weight ring) is libgrootgpu's own NCCL code."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())


def run_config(ctx, name, steps, warmup, n, sample_clocks, cpu_baseline):
    """Times one BASELINE config on this rank's GPU; returns the result dict (meaningful on rank 0)."""
    import numpy as np
    from groot_b200 import api, synth
    from groot_b200 import distributed as gd
    torch = ctx.torch
    cfg = CONFIGS[name]
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    L, no_align = cfg["read_len"], cfg["no_align"]

    msa_dir = prepare_db(cfg["db"]) if rank == 0 else None
    ctx.barrier()
    if msa_dir is None:
        msa_dir = prepare_db(cfg["db"])
    t0 = time.time()
    idx = api.Index.build(msa_dir=msa_dir, device=ctx.local, **cfg["index"])
    t_index = time.time() - t0
    info = idx.info()
    comm = None
    if world > 1:   # the data-path communicator lives in the library; torch only carries its id
        comm = api.Comm(idx, gd.broadcast_comm_id(api.Comm.new_id, device=dev), rank, world)

    seqs = synth.db_sequences(msa_dir)
    blob, off = synth.synth_reads(n, L, seqs, seed=42 + rank)      # weak scaling: rank r maps reads [r*n, (r+1)*n) of the global batch
    p_seq = api.PinnedBuffer(n * L)                                   # pinned host input buffer (grootgpu_host_alloc)
    p_seq.array[:] = blob
    d_seq = torch.zeros(n * L + 64, dtype=torch.uint8, device=dev)
    d_seq[: n * L].copy_(torch.from_numpy(blob))
    d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
    stream = torch.cuda.current_stream().cuda_stream

    # The gather call waits for every rank's sizes (one small all-gather), i.e. for the slowest rank of the batch: it is
    # issued from a helper thread while this thread already maps the next batch (include/grootgpu.h, grootgpu_gather). At
    # most one gather is outstanding: gather b has returned before align b + 2 starts.
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=1) if comm else None
    pending = [None]
    last_merged = [None]

    def gather_async(raw, to_host):
        if pending[0] is not None:
            last_merged[0] = pending[0].result()
        pending[0] = pool.submit(comm.gather, raw, to_host)

    def step_device(compact=True):
        # stream=None: the library's own compute stream (it sits between the priorities of the f64 chains and of the gather).
        # The call returns after the batch's kernels have finished, so the events recorded on torch's stream around the
        # loop still bracket all of the device work.
        raw = idx.map_reads_device(d_seq.data_ptr(), d_off.data_ptr(), n, L, L, THRESHOLD, no_align=no_align, stream=None, project_on_device=True, compact=compact)
        if comm:
            gather_async(raw, 0)   # the one collective of the path: result arrays to rank 0 over NVLink, merged there
        return raw, last_merged[0]

    def drain():
        if comm:
            if pending[0] is not None:
                last_merged[0] = pending[0].result()
                pending[0] = None
            comm.sync()            # gathers done, the weight vector has been round every rank, rank 0 holds the weights
        else:
            idx.weights()          # the f64 chains run behind the batches: wait for the last ones
        torch.cuda.synchronize()

    e2e_parts = {"align_batch_ms": 0.0, "gather_ms": 0.0}

    def step_e2e():
        t0 = time.perf_counter()
        # reads of one length: params->fixed_read_len, no offsets array handed over (it would be 8 more bytes per read on the link)
        raw = idx.map_reads_raw(p_seq.ptr, None, n, THRESHOLD, no_align=no_align, project_on_device=True, compact=True,
                                results_on_device=comm is not None, fixed_read_len=L)
        t1 = time.perf_counter()
        if comm:
            gather_async(raw, 2)   # rank 0: merged compact batch -> host, asynchronously (complete at the next gather / sync)
        e2e_parts["align_batch_ms"] += (t1 - t0) * 1e3
        e2e_parts["gather_ms"] += (time.perf_counter() - t1) * 1e3
        return raw, last_merged[0]

    # ---- value: inputs resident in HBM ----
    sampler = ClockSampler(ctx.local) if sample_clocks else None
    if sampler and rank == 0:
        sampler.start()
    for _ in range(warmup):
        raw, merged = step_device()
    drain(); ctx.barrier()
    e0, e_map, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    fam_ms = {k: [] for k in api.KERNEL_FAMILIES}
    launches = 0
    torch.cuda.synchronize(); ctx.barrier()
    if sampler:
        sampler.mark_begin()
    e0.record()
    for _ in range(steps):
        raw, merged = step_device()
        for k, v in zip(api.KERNEL_FAMILIES, list(raw.kernel_ms)[:8]):
            fam_ms[k].append(v)
        launches += raw.kernel_launches
    e_map.record()         # this rank's mapping stream is done ...
    drain()                # ... and so is everything behind it on the library's own streams: last gather, last chains of the weight ring
    e1.record()            # recorded after the drain returned: the event pair spans all of it
    torch.cuda.synchronize(); ctx.barrier()
    if sampler:
        sampler.mark_end()
    dev_ms = ctx.reduce(e0.elapsed_time(e1), "MAX")
    map_ms = ctx.reduce(e0.elapsed_time(e_map), "MAX")
    clocks = sampler.stop() if (sampler and rank == 0) else None
    total_reads = ctx.reduce(float(n), "SUM")
    value = total_reads * steps / (dev_ms / 1000.0)
    stats = dict(hits=raw.n_hits / n, pairs=raw.n_pairs / n, records=raw.n_records / n, mapped=raw.mapped / n)
    merged_check = None
    if comm and rank == 0:
        merged = last_merged[0]
        merged_check = dict(reads=int(merged.n_reads), pairs=int(merged.n_pairs), records=int(merged.n_records), mapped=int(merged.mapped))

    # ---- the same step with the full result arrays (round 1's output format), N = 1 only ----
    full_format = None
    if world == 1:
        for _ in range(3):
            step_device(compact=False)
        drain()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fsteps = max(3, steps // 2)
        f0.record()
        for _ in range(fsteps):
            step_device(compact=False)
        drain()
        f1.record()
        torch.cuda.synchronize()
        fms = f0.elapsed_time(f1) / fsteps
        full_format = {"value": n / (fms / 1000.0), "unit": UNIT, "ms_per_step": fms, "steps": fsteps,
                       "output": "hit_off / hits / 32-byte pairs / (u32 path, i32 pos) records, left on the device"}

    # ---- e2e: pinned host buffers through the C ABI, compact result to the host ----
    for _ in range(max(3, warmup // 2)):      # at least 3: both alternating result sets (device and pinned host) are allocated in the first two calls
        raw_e, merged_e = step_e2e()
    drain(); ctx.barrier()
    e2e_parts.update(align_batch_ms=0.0, gather_ms=0.0)
    t0 = time.perf_counter()
    for _ in range(steps):
        raw_e, merged_e = step_e2e()
    drain()
    e2e_s = ctx.reduce(time.perf_counter() - t0, "MAX")
    e2e_parts = {k: v / steps for k, v in e2e_parts.items()}
    if world > 1:          # every rank's time inside grootgpu_align_batch (H2D + kernels): shows how evenly the host feeds the GPUs
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = e2e_parts["align_batch_ms"]
        ctx.dist.all_reduce(t)
        e2e_parts["align_batch_ms_by_rank"] = [round(x, 2) for x in t.tolist()]
    ctx.barrier()
    e2e_value = total_reads * steps / e2e_s
    recw = raw_e.rec_path_bytes
    h2d = n * L                                                       # per rank: the bases (fixed_read_len: no offsets travel)
    d2h_one = 16 * raw_e.n_pairs + recw * raw_e.n_records             # compact: 16-byte pairs + path ids
    d2h = ctx.reduce(float(d2h_one), "SUM")                           # N > 1: all of it leaves through rank 0 after the gather

    # ---- roofline of the dominant kernel (largest share of the step), algorithmic bytes per DESIGN.md "Kernels" ----
    peak, peak_src = measured_peaks()
    S = info["S"]
    fam_avg = {k: sum(v) / len(v) for k, v in fam_ms.items()}
    fam_bytes = {   # per launch family, for n reads
        "seed": n * (L + 8 + 32 + 4) + raw.n_hits * (8 * S + 8) + (0 if no_align else raw.mapped * (L // 2 + 17)),   # L + 8 + 32 + 8*S*c + 4 + 8*h with c := h; + the 2-bit copies (both strands), flag and one-hot prefixes of the seeded reads, written by the queued pass
        "fill": 4 * n + raw.n_hits * 13,                                         # hit counts in, hits + owner + segment flag out
        "align_screen": raw.n_pairs * (L + 16 + 8) + raw.n_hits * 32,            # read + pair bookkeeping + window records
        "align_walk": raw.n_pairs * (L + 16 + 32 + 8 + 32),                      # re-read read + window meta + pair out + locus + path bitset
        "align_finish": 0,
        "align_emit": raw.n_pairs * (32 + 8 + 32 + 16) + raw.n_records * raw.rec_path_bytes,   # pair + locus + bitset in, 16-byte compact pair + path ids out
        "project": raw.n_pairs * 40,                                             # pair in; the (node, f64) items are internal traffic
        "project_accumulate": 0,
    }
    dom = max(fam_avg, key=fam_avg.get)
    achieved = fam_bytes[dom] / (fam_avg[dom] / 1000.0) / 1e9
    traffic, traffic_src, inst_per_read = None, None, None
    prof = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof)).get(dom, {})
            if tj.get("dram_bytes_per_read") is not None:
                traffic = tj["dram_bytes_per_read"] * n
                traffic_src = "ncu --set full capture %s (%d reads per launch), scaled to this launch; not re-measured in this run" % (tj.get("source"), tj.get("reads_in_profiled_launch", 0))
            inst_per_read = tj.get("warp_inst_per_read")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": dom, "kernel_ms": fam_avg[dom], "algorithmic_bytes_per_launch": fam_bytes[dom], "peak_source": peak_src,
                "all_kernels": {k: {"ms": fam_avg[k], "algorithmic_GBps": (fam_bytes[k] / (fam_avg[k] / 1000.0) / 1e9) if fam_avg[k] > 0 else 0.0}
                                for k in api.KERNEL_FAMILIES}}
    if rank == 0:
        # the resource that really bounds the sketch kernel: integer instruction issue. Peak = measured here
        # (grootgpu_int_issue_peak: 16 independent IMAD / SHF / LOP3 chains per thread, full occupancy); the kernel's
        # instruction count per read comes from the committed ncu capture (smsp__inst_executed.sum).
        try:
            ip = api.int_issue_peak(ctx.local)
            roofline["int_issue_peak_warp_inst_per_s"] = ip
            if inst_per_read and dom == "seed":
                rate = inst_per_read * n / (fam_avg["seed"] / 1000.0)
                roofline["int_issue"] = {"warp_inst_per_read": inst_per_read, "achieved_warp_inst_per_s": rate, "frac": rate / ip["mixed"],
                                         "note": "seed_kernel is integer-issue bound: this fraction, not the HBM one, says how close it is to its roofline"}
                roofline["int_issue_frac"] = rate / ip["mixed"]
        except Exception as e:   # noqa: BLE001
            roofline["int_issue_peak_error"] = str(e)

    cpu = None
    if rank == 0 and world == 1 and cpu_baseline:
        from oracle import pyoracle as po
        cores = os.cpu_count() or 1
        oidx = po.Index(msa_dir=msa_dir, **cfg["index"])
        pn = 20000
        t0 = time.time(); oidx.map_reads(blob[: pn * L], off[: pn + 1], THRESHOLD, no_align=no_align, threads=cores); rate = pn / (time.time() - t0)
        sn = int(max(pn, min(n, rate * 15.0)))
        t0 = time.time(); oidx.map_reads(blob[: sn * L], off[: sn + 1], THRESHOLD, no_align=no_align, threads=cores); dt = time.time() - t0
        cpu = {"value": sn / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d reads of the step's batch, oracle/ C++ restatement, %d threads, %.1fs" % (sn, cores, dt)}
    if pool:
        pool.shutdown()
    if comm:
        comm.close()
    idx.close()
    del d_seq, d_off
    p_seq.free()
    torch.cuda.empty_cache()
    par = "reads sharded over %d GPU(s), index replicated" % world
    if world > 1:
        par += "; per step ONE NCCL gather of the result arrays to rank 0 (merged on its device) + the weight vector sent rank to rank (f64 chains in global read order)"
    return {
        "value": value, "ms_per_step": dev_ms / steps, "mapping_stream_ms_per_step": map_ms / steps, "full_format": full_format,
        "config": {"workload": cfg["workload"], "config_id": name, "reads_per_gpu_per_step": n, "read_len": L, "threshold": THRESHOLD, "index_windows": info["windows"],
                   "output": "compact records (16-byte pairs + %d-byte path ids); graph weights on the device" % raw.rec_path_bytes,
                   "l2": "inputs (%.2f GB of reads per step) are larger than the 126 MB L2; no explicit flush" % (n * L / 1e9),
                   "parallelism": par, "per_read": stats, "index_build_s": t_index, "merged_on_rank0": merged_check},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": int(d2h),
                "h2d_GBps_per_rank": h2d / (e2e_s / steps) / 1e9, "h2d_GBps_all_ranks": h2d * world / (e2e_s / steps) / 1e9, "d2h_GBps_rank0": d2h / (e2e_s / steps) / 1e9,
                "includes": "pinned host buffer of bases (reads of one length: fixed_read_len, no offsets) -> grootgpu_align_batch (H2D in chunks on two lanes overlapped with the kernels: sketch + query + align + ordered graph "
                            "weighting) -> compact result (16-byte pairs + %d-byte path ids) to the host%s" % (recw, "; N > 1: results kept on the device, gathered to rank 0 over "
                            "NVLink, merged, copied to rank 0's host (asynchronously, drained inside the timed region)" if world > 1 else ""),
                "per_step_ms": e2e_parts},
        "gpu_launches": launches, "kernel_ms": fam_avg, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
    }


def run_ours(args):
    ctx = Ctx()
    res = run_config(ctx, args.config, args.steps, args.warmup, args.reads, True, not args.no_cpu_baseline)
    others = {}
    extra = [] if args.extra in ("none", "") else [c for c in args.extra.split(",") if c != args.config]
    if ctx.world == 1:
        for name in extra:   # the other BASELINE configs, shorter runs; same measurement
            r = run_config(ctx, name, max(3, args.steps // 4), 3, args.reads, False, not args.no_cpu_baseline)
            others[name] = {k: r[k] for k in ("value", "ms_per_step", "config", "e2e", "kernel_ms", "cpu_baseline", "gpu_launches")}
            others[name]["roofline"] = {k: r["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "kernel_ms") if k in r["roofline"]}
            if "int_issue_frac" in r["roofline"]:
                others[name]["roofline"]["int_issue_frac"] = r["roofline"]["int_issue_frac"]
    if ctx.rank == 0:
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": res["config"], "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "kernel_ms": res["kernel_ms"],
            "mapping_stream_ms_per_step": res["mapping_stream_ms_per_step"], "full_format": res["full_format"],
            "roofline": res["roofline"], "cpu_baseline": res["cpu_baseline"], "clocks": res["clocks"],
        }
        if others:
            line["other_configs"] = others
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--extra", default="C2,C4", help="further BASELINE configs measured at N = 1 (shorter runs), or 'none'")
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
