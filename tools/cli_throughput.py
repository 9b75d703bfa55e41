"""Wall-clock throughput of the C++ host driver (FASTQ -> groot-b200 align -> BAM), for the I/O rows of SURVEY.md 8(f).
   python tools/cli_throughput.py [n_reads]      (GPU box)"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from groot_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
L = 100
CLI = os.path.join(ROOT, "groot_b200", "groot-b200")
tmp = "/tmp/groot_cli_tp"
os.makedirs(tmp, exist_ok=True)
msa_dir = synth.unpack_db(os.path.join(ROOT, "data", "db", "arg-annot.90.tar"), "/tmp/groot_b200_db_%d" % os.getuid())
blob, off = synth.synth_reads(n, L, synth.db_sequences(msa_dir), seed=42)
# fixed-width FASTQ rows assembled as one byte matrix: "@SYN_%09d\n" + seq + "\n+\n" + qual + "\n"
idx = np.arange(n)
digits = np.stack([(idx // 10 ** p) % 10 for p in range(8, -1, -1)], axis=1).astype(np.uint8) + ord("0")
row = np.empty((n, 5 + 9 + 1 + L + 3 + L + 1), dtype=np.uint8)
row[:, :5] = np.frombuffer(b"@SYN_", dtype=np.uint8)
row[:, 5:14] = digits
row[:, 14] = ord("\n")
row[:, 15:15 + L] = blob.reshape(n, L)
row[:, 15 + L:18 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
row[:, 18 + L:18 + 2 * L] = ord("I")
row[:, 18 + 2 * L] = ord("\n")
fq = os.path.join(tmp, "reads.fq")
row.tofile(fq)
print("fastq: %d reads, %.1f MB" % (n, os.path.getsize(fq) / 1e6))
t0 = time.time()
subprocess.run([CLI, "index", "-m", msa_dir, "-i", os.path.join(tmp, "idx"), "-w", "100", "-k", "31", "-s", "21"], check=True, stderr=subprocess.DEVNULL)
print("index: %.2f s" % (time.time() - t0))
for extra in (["--noAlign"], ["-p", "1"], ["-p", "16"], ["-p", "16", "--bamDelta", "0"], ["-p", "16", "--bamDelta", "0", "--bamLevel", "1"],
              ["-p", "16", "--bamLevel", "0"]):
    out = os.path.join(tmp, "out.bam")
    t0 = time.time()
    with open(out, "wb") as f:
        r = subprocess.run([CLI, "align", "-i", os.path.join(tmp, "idx"), "-f", fq, "-g", os.path.join(tmp, "graphs")] + extra,
                           stdout=f, stderr=subprocess.PIPE)
    dt = time.time() - t0
    assert r.returncode == 0, r.stderr.decode()
    print("align %-28s %.2f s  %.3f M reads/s  BAM %.1f MB" % (" ".join(extra), dt, n / dt / 1e6, os.path.getsize(out) / 1e6))
