"""Host-only throughput of the BAM side of `groot-b200 align` (format_batch_bam: record formatting + BGZF deflate), on a
fabricated batch shaped like the C3 workload (100 bp reads, ~0.52 pairs per read, ~17 records per pair as against
arg-annot.90): tests/cpp/bam_batch.cpp timed for zlib-only and for the hint-driven block writer (host/bgzf.h).
   python tools/bam_throughput.py [n_reads] [workers]"""
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
workers = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
tmp = "/tmp/groot_bam_tp"
os.makedirs(tmp, exist_ok=True)
exe = os.path.join(tmp, "bam_batch")
host = os.path.join(ROOT, "groot_b200", "csrc", "host")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "bam_batch.cpp"), os.path.join(host, "pipeline.cpp"),
                       "-L" + os.path.join(ROOT, "groot_b200"), "-lgrootgpu", "-lz", "-pthread", "-Wl,-rpath," + os.path.join(ROOT, "groot_b200")])
rng = np.random.default_rng(1)
L = 100
n_graphs, paths_per_graph, n_nodes = 300, 40, 20000
seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), (n, L))
qual = np.repeat(rng.choice(np.frombuffer(b"F:,#", dtype=np.uint8), (n, L // 5), p=[0.8, 0.1, 0.07, 0.03]), 5, axis=1)     # binned qualities
idx = np.arange(n)
ids = np.empty((n, 14), dtype=np.uint8)
ids[:, :5] = np.frombuffer(b"@SYN_", dtype=np.uint8)
ids[:, 5:] = np.stack([(idx // 10 ** p) % 10 for p in range(8, -1, -1)], axis=1).astype(np.uint8) + ord("0")
has = rng.random(n) < 0.52
reads = idx[has].astype(np.uint32)
cnt = np.clip(rng.geometric(1 / 17.0, len(reads)), 1, paths_per_graph).astype(np.uint32)
node = rng.integers(0, n_nodes, len(reads)).astype(np.uint32)
flags = rng.integers(0, 60, len(reads)).astype(np.uint32) | (rng.random(len(reads)) < 0.5).astype(np.uint32) << 28
cpairs = np.stack([reads, node, flags, cnt], axis=1).astype(np.uint32)
rec_path = np.concatenate([np.sort(rng.choice(paths_per_graph, c, replace=False)) for c in cnt]).astype(np.uint8)
f = os.path.join(tmp, "batch.bin")
with open(f, "wb") as fh:
    fh.write(b"BAMT" + struct.pack("<IIQIIII", n, len(cpairs), len(rec_path), 1, n_graphs, n_graphs * paths_per_graph, n_nodes))
    for w in (14, L, L): fh.write((np.arange(n + 1, dtype=np.uint64) * w).tobytes())
    fh.write(ids.tobytes()); fh.write(seq.tobytes()); fh.write(qual.tobytes())
    fh.write(cpairs.tobytes()); fh.write(rec_path.tobytes())
    fh.write((np.arange(n_graphs + 1, dtype=np.uint32) * paths_per_graph).tobytes())
    for g in range(n_graphs):
        for p in range(paths_per_graph):
            nm = b"gene_%d_%d" % (g, p)
            fh.write(struct.pack("<I", len(nm)) + nm + struct.pack("<i", 1200))
    node_graph = rng.integers(0, n_graphs, n_nodes)
    all_ids = np.arange(paths_per_graph, dtype=np.uint32).tobytes()
    for k in range(n_nodes):
        fh.write(struct.pack("<II", int(node_graph[k]), paths_per_graph) + all_ids + rng.integers(0, 1100, paths_per_graph).astype(np.int32).tobytes())
print("batch: %d reads, %d pairs, %d records" % (n, len(cpairs), len(rec_path)))
for level, delta in ((0, 0), (1, 0), (-1, 0), (1, 1), (-1, 1)):
    out = os.path.join(tmp, "o.bam")
    r = subprocess.run([exe, f, out, str(workers), str(level), str(delta), "3"], stdout=subprocess.PIPE, check=True)
    secs, raw_bytes, bam_bytes, delta_blocks, own_code, zlib_blocks = r.stdout.decode().split()
    secs = float(secs)
    print("workers %2d level %2d delta %d: %.3f s  %.2f M reads/s  %.2f GB/s of records  BAM/raw %.3f  delta blocks %s" %
          (workers, level, delta, secs, n / secs / 1e6, int(raw_bytes) / secs / 1e9, int(bam_bytes) / int(raw_bytes), delta_blocks))
