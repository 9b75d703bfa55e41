"""Deterministic synthetic read sets for the parity tests and bench.py.

Composition: 50 % exact substrings of the de-gapped, upper-cased MSA sequences of the
database (uniform sequence, uniform start, half of them reverse-complemented), 25 % the same with one
random substitution, 25 % uniform random ACGT. Generator: numpy default_rng(seed) (PCG64), draws in
the order documented in synth_reads(); the data never leaves the box, so there is no file format.
"""
import glob
import os
import tarfile

import numpy as np

_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b


def unpack_db(tar_path, dest_dir):
    """Extracts a clustered-ARG-database tarball (data/db/*.tar); returns the MSA directory."""
    name = os.path.basename(tar_path)[:-4]
    out = os.path.join(dest_dir, name)
    if not os.path.isdir(out) or not glob.glob(os.path.join(out, "cluster*.msa")):
        os.makedirs(dest_dir, exist_ok=True)
        with tarfile.open(tar_path) as t:
            t.extractall(dest_dir)
    return out


def msa_files(msa_dir):
    """cluster*.msa in the order filepath.Glob returns them (lexicographic; cmd/index.go:143)."""
    return sorted(glob.glob(os.path.join(msa_dir, "cluster*.msa")))


def db_sequences(msa_dir, min_len=0):
    """De-gapped, upper-cased sequences (non-ACGT -> N) of every MSA row except `consensus`."""
    seqs = []
    for f in msa_files(msa_dir):
        name, buf = None, []
        for line in open(f, "rb"):
            line = line.strip()
            if not line:
                continue
            if line.startswith(b">"):
                if name is not None and name != b"consensus":
                    seqs.append(b"".join(buf))
                name, buf = line[1:].split()[0], []
            else:
                buf.append(line)
        if name is not None and name != b"consensus":
            seqs.append(b"".join(buf))
    out = []
    keep = np.full(256, ord("N"), dtype=np.uint8)
    for c in b"ACGTN":
        keep[c] = c
        keep[c + 32] = c
    for s in seqs:
        a = np.frombuffer(s, dtype=np.uint8)
        a = keep[a[a != ord("-")]]
        if len(a) >= max(1, min_len):
            out.append(a)
    return out


def synth_reads(n, read_len, seqs, seed=42, chunk=1 << 20, frac_exact=0.5, frac_sub=0.25):
    """Returns (blob uint8[n*read_len], off uint64[n+1]).

    Draw order per chunk: kind (uniform float) -> sequence index -> start -> strand bit -> substitution
    position -> substitution base offset (1..3) -> random bases."""
    rng = np.random.default_rng(seed)
    seqs = [s for s in seqs if len(s) >= read_len]
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    starts = np.zeros(len(seqs) + 1, dtype=np.int64)
    starts[1:] = np.cumsum(lens)
    cat = np.concatenate(seqs)
    blob = np.empty(n * read_len, dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    code = np.zeros(256, dtype=np.int64)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    ar = np.arange(read_len, dtype=np.int64)
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        kind = rng.random(m)
        si = rng.integers(0, len(seqs), size=m)
        st = (rng.random(m) * (lens[si] - read_len + 1)).astype(np.int64)
        strand = rng.integers(0, 2, size=m)
        subpos = rng.integers(0, read_len, size=m)
        subofs = rng.integers(1, 4, size=m)
        rnd = acgt[rng.integers(0, 4, size=(m, read_len))]
        reads = cat[(starts[si] + st)[:, None] + ar[None, :]]
        is_sub = (kind >= frac_exact) & (kind < frac_exact + frac_sub)
        rows = np.nonzero(is_sub)[0]
        if len(rows):
            old = reads[rows, subpos[rows]]
            reads[rows, subpos[rows]] = acgt[(code[old] + subofs[rows]) % 4]
        rc = strand == 1
        reads[rc] = _COMP[reads[rc]][:, ::-1]
        is_rnd = kind >= frac_exact + frac_sub
        reads[is_rnd] = rnd[is_rnd]
        blob[b * read_len:(b + m) * read_len] = reads.reshape(-1)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(read_len)
    return blob, off
