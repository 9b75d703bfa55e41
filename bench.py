#!/usr/bin/env python
"""bench.py — `groot align` hot path on B200: reads/s for 100 bp synthetic reads vs arg-annot.90.

A "step" is one pass of the hot path (KHF sketch -> LSH Ensemble containment query -> hierarchical exact graph
alignment) over one batch of synthetic reads (BASELINE.json configs[2]: 10 M x 100 bp, arg-annot.90 -w 100,
full align path). `value` is measured with the batch already resident in HBM; `e2e` goes through the
reference-facing C-ABI call with pinned HOST buffers (H2D of reads, D2H of hits/pairs/records inside the timed
region, plus the ordered host replay of the graph weighting). `--impl reference` times the CPU restatement of
the reference (oracle/, all host threads) on a bounded sample of the same workload — the reference itself is
Go and cannot be built in this image (DESIGN.md "Oracle").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R] [--read-len L]
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
INDEX_PARAMS = CONFIGS["C3"]["index"]


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
    import numpy as np
    from groot_b200 import synth
    from oracle import pyoracle as po
    msa_dir = prepare_db()
    cores = os.cpu_count() or 1
    t0 = time.time()
    idx = po.Index(msa_dir=msa_dir, k=INDEX_PARAMS["k"], S=INDEX_PARAMS["S"], w=INDEX_PARAMS["w"])
    t_index = time.time() - t0
    seqs = synth.db_sequences(msa_dir)
    L = args.read_len
    probe_n = 20000
    blob, off = synth.synth_reads(probe_n, L, seqs, seed=42)
    t0 = time.time(); idx.map_reads(blob, off, THRESHOLD, threads=cores); probe = probe_n / (time.time() - t0)
    total_budget = 120.0   # seconds for all steps
    per_step = max(20000, min(2_000_000, int(probe * total_budget / max(1, args.steps + args.warmup))))
    blob, off = synth.synth_reads(per_step, L, seqs, seed=42)
    for _ in range(args.warmup):
        idx.map_reads(blob, off, THRESHOLD, threads=cores)
    t0 = time.time()
    for _ in range(args.steps):
        idx.reset_weights()
        res = idx.map_reads(blob, off, THRESHOLD, threads=cores)
    dt = time.time() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": "10M x 100bp synthetic reads vs arg-annot.90 (-w 100 -k 31 -s 21), full align path; bounded sample per step",
                   "reads_per_step": per_step, "read_len": L, "threshold": THRESHOLD},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d reads/step x %d steps of the same synthetic mix (seed 42); oracle index build %.1fs not counted" % (per_step, args.steps, t_index)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "mapped_fraction": res.counts["mapped"] / res.counts["received"],
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from groot_b200 import api, synth
    from groot_b200 import distributed as gd

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    msa_dir = prepare_db() if rank == 0 else None
    barrier()
    if msa_dir is None:
        msa_dir = prepare_db()
    t0 = time.time()
    idx = api.Index.build(msa_dir=msa_dir, device=local, **INDEX_PARAMS)
    t_index = time.time() - t0
    info = idx.info()

    n, L = args.reads, args.read_len
    seqs = synth.db_sequences(msa_dir)
    blob, off = synth.synth_reads(n, L, seqs, seed=42 + rank)      # weak scaling: every rank maps its own n reads
    h_seq = torch.from_numpy(blob).pin_memory()
    h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
    d_seq = torch.zeros(n * L + 64, dtype=torch.uint8, device=dev)
    d_seq[: n * L].copy_(h_seq)
    d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
    stream = torch.cuda.current_stream().cuda_stream

    gatherer = gd.OverlappedGather(dev, dst=0) if world > 1 else None

    def step_device():
        raw = idx.map_reads_device(d_seq.data_ptr(), d_off.data_ptr(), n, L, L, THRESHOLD, stream=stream, project_on_device=True)
        if world > 1:   # the one collective of the path: every rank's result arrays go to rank 0 over NVLink, sent from
            gatherer.submit(gd.result_tensors_from_raw(raw, dev))   # staging copies while the next batch is being mapped
        return raw

    e2e_parts = {"align_batch_ms": 0.0, "copy_in_to_copy_out_ms": 0.0}

    def step_e2e():
        t0 = time.perf_counter()
        raw = idx.map_reads_raw(h_seq.data_ptr(), h_off.data_ptr(), n, THRESHOLD, project_on_device=True)
        e2e_parts["align_batch_ms"] += (time.perf_counter() - t0) * 1e3
        e2e_parts["copy_in_to_copy_out_ms"] += raw.ms[0]   # CUDA events: first copy-in issued -> last copy-out complete
        return raw

    # ---- value: inputs resident in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        raw = step_device()
    if gatherer is not None:
        gatherer.flush()
    torch.cuda.synchronize(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fam_ms = {k: [] for k in api.KERNEL_FAMILIES}
    launches = 0
    torch.cuda.synchronize(); barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        raw = step_device()
        for k, v in zip(api.KERNEL_FAMILIES, list(raw.kernel_ms)[:8]):
            fam_ms[k].append(v)
        launches += raw.kernel_launches
    if gatherer is not None:
        gatherer.flush()            # every gather of the K steps completes inside the timed region
    e1.record()
    torch.cuda.synchronize(); barrier()
    sampler.mark_end()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    total_reads = sum_over_ranks(float(n))
    value = total_reads * args.steps / (dev_ms / 1000.0)
    stats = dict(hits=raw.n_hits / n, pairs=raw.n_pairs / n, records=raw.n_records / n, mapped=raw.mapped / n)

    # ---- e2e: pinned host buffers through the C ABI + host replay of the graph weighting ----
    for _ in range(max(1, args.warmup // 2)):
        raw = step_e2e()
    torch.cuda.synchronize(); barrier()
    e2e_parts.update(align_batch_ms=0.0, copy_in_to_copy_out_ms=0.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        raw = step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_parts = {k: v / args.steps for k, v in e2e_parts.items()}
    barrier()
    e2e_value = total_reads * args.steps / e2e_s
    h2d = n * L + 4 * (n + 1)
    d2h = 4 * (n + 1) + 4 * raw.n_hits + 32 * raw.n_pairs + 8 * raw.n_records

    # ---- roofline of the dominant kernel (largest share of the step), algorithmic bytes per DESIGN.md "Rooflines" ----
    peak, peak_src = measured_peaks()
    S = info["S"]
    fam_avg = {k: sum(v) / len(v) for k, v in fam_ms.items()}
    fam_bytes = {   # per launch family, for n reads
        "seed": n * (L + 8 + 32 + 4) + raw.n_hits * (8 * S + 8),                # L + 8 + 32 + 8*S*c + 4 + 8*h with c := h
        "fill": 4 * n + raw.n_hits * 13,                                         # hit counts in, hits + owner + segment flag out
        "align_screen": raw.n_pairs * (L + 16 + 8) + raw.n_hits * 32,            # read + pair bookkeeping + window records
        "align_walk": raw.n_pairs * (L + 16 + 32 + 8 + 32),                      # re-read read + window meta + pair out + locus + path bitset
        "align_finish": 0,
        "align_emit": raw.n_pairs * (32 + 8 + 32) + raw.n_records * 8,           # pair + locus + bitset in, 8-byte records out
        "project": raw.n_pairs * 40,                                             # pair in; the (node, f64) items are internal traffic
    }
    dom = max(fam_avg, key=fam_avg.get)
    achieved = fam_bytes[dom] / (fam_avg[dom] / 1000.0) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            traffic = tj.get(dom, {}).get("dram_bytes_per_read", None)
            if traffic is not None:
                traffic = traffic * n
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": dom, "kernel_ms": fam_avg[dom], "algorithmic_bytes_per_launch": fam_bytes[dom], "peak_source": peak_src,
                "note": "integer / pointer-chasing kernels: the HBM fraction is small by construction (DESIGN.md Rooflines); "
                        "seed_kernel is INT-ALU bound (%d integer ops per read)" % ((L - info["k"] + 1) * ((S - 1) * 12 + 20)),
                "all_kernels": {k: {"ms": fam_avg[k], "algorithmic_GBps": (fam_bytes[k] / (fam_avg[k] / 1000.0) / 1e9) if fam_avg[k] > 0 else 0.0}
                                for k in api.KERNEL_FAMILIES}}

    # ---- CPU baseline (rank 0, N == 1): the oracle port on all host threads, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as po
        cores = os.cpu_count() or 1
        oidx = po.Index(msa_dir=msa_dir, k=INDEX_PARAMS["k"], S=INDEX_PARAMS["S"], w=INDEX_PARAMS["w"])
        pn = 20000
        t0 = time.time(); oidx.map_reads(blob[: pn * L], off[: pn + 1], THRESHOLD, threads=cores); rate = pn / (time.time() - t0)
        sn = int(max(pn, min(n, rate * 15.0)))
        t0 = time.time(); oidx.map_reads(blob[: sn * L], off[: sn + 1], THRESHOLD, threads=cores); dt = time.time() - t0
        cpu = {"value": sn / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d reads of the step's batch, oracle/ C++ restatement, %d threads, %.1fs" % (sn, cores, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": "10M x 100bp synthetic reads vs arg-annot.90 (-w 100 -k 31 -s 21 -x 8 -y 4, t=0.99), full align path "
                                   "(sketch + LSH Ensemble query + exact graph alignment) [BASELINE.json configs[2]]",
                       "reads_per_gpu_per_step": n, "read_len": L, "threshold": THRESHOLD, "index_windows": info["windows"],
                       "l2": "inputs (%.2f GB of reads per step) are larger than the 126 MB L2; no explicit flush" % (n * L / 1e9),
                       "parallelism": "reads sharded over %d GPU(s), index replicated%s" % (world, ", one NCCL gather of results to rank 0 per step" if world > 1 else ""),
                       "per_read": stats, "index_build_s": t_index},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes": "pinned host buffers -> grootgpu_align_batch (H2D, sketch+query+align+ordered graph weighting kernels, D2H of hits/pairs/records); "
                                "inside the call the batch is streamed through the device in chunks on two lanes, copies overlapped with kernels",
                    "per_step_ms": e2e_parts},
            "gpu_launches": launches,
            "kernel_ms": fam_avg,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
