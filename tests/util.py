"""Shared helpers for the test-suite (FASTQ loading, synthetic reads, the `groot report` restatement)."""
import gzip

import numpy as np


def load_fastq(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        lines = f.read().split(b"\n")
    names, seqs, quals = [], [], []
    for i in range(0, len(lines) - 3, 4):
        names.append(lines[i])
        seqs.append(lines[i + 1])
        quals.append(lines[i + 3])
    return names, seqs, quals


def pack_reads(seqs):
    blob = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return blob, off


_COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]


def read_msa(path):
    rows, name, buf = [], None, []
    for line in open(path):
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            if name is not None and name != "consensus":
                rows.append((name, "".join(buf)))
            name, buf = line[1:].split()[0], []
        else:
            buf.append(line)
    if name is not None and name != "consensus":
        rows.append((name, "".join(buf)))
    return rows


def report(records, ref_len, cutoff):
    """Restatement of `groot report` (src/reporting/reporting.go:33-173): records = iterable of
    (ref name, pos, aligned length); returns {gene: (count, length)} for genes whose pileup coverage
    >= cutoff. Note the reference's inclusive end (reporting.go:104-119)."""
    by_ref = {}
    for name, pos, ln in records:
        by_ref.setdefault(name, []).append((pos, ln))
    out = {}
    for name, recs in by_ref.items():
        n = ref_len[name]
        cov = np.zeros(n, dtype=np.int64)
        for pos, ln in recs:
            end = min(pos + ln, n - 1)
            cov[pos:end + 1] += 1
        if (cov > 0).sum() / n >= cutoff:
            out[name[1:] if name.startswith("*") else name] = (len(recs), n)
    return out
