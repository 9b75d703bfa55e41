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


def read_bam(path):
    """Minimal BAM decoder (BGZF == concatenated gzip members). Returns (header text, [(ref name, length)], records)
    with records = list of dicts(name, ref, pos, mapq, flag, cigar, seq, qual, next_ref, next_pos, tlen)."""
    import struct
    raw = gzip.decompress(open(path, "rb").read())
    assert raw[:4] == b"BAM\x01"
    p = 4
    l_text, = struct.unpack_from("<i", raw, p); p += 4
    text = raw[p:p + l_text].decode(); p += l_text
    n_ref, = struct.unpack_from("<i", raw, p); p += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", raw, p); p += 4
        name = raw[p:p + l_name - 1].decode(); p += l_name
        l_ref, = struct.unpack_from("<i", raw, p); p += 4
        refs.append((name, l_ref))
    recs = []
    codes = "=ACMGRSVTWYHKDBN"
    while p < len(raw):
        block, = struct.unpack_from("<i", raw, p); p += 4
        ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHiiii", raw, p)
        q = p + 32
        name = raw[q:q + l_name - 1].decode(); q += l_name
        cig = ""
        for _ in range(n_cig):
            c, = struct.unpack_from("<I", raw, q); q += 4
            cig += "%d%s" % (c >> 4, "MIDNSHP=X"[c & 15])
        sb = raw[q:q + (l_seq + 1) // 2]; q += (l_seq + 1) // 2
        seq = "".join(codes[(sb[i // 2] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        qual = raw[q:q + l_seq]; q += l_seq
        recs.append(dict(name=name, ref=refs[ref_id][0], pos=pos, mapq=mapq, flag=flag, cigar=cig, seq=seq, qual=qual,
                         next_ref=nref, next_pos=npos, tlen=tlen))
        p += block
    return text, refs, recs


def canonical_records_from_oracle(idx, res, names, seqs, quals):
    """What AlignRead puts into sam.Record for every oracle record (src/graph/alignment.go:114-156), as sortable tuples
    (name, ref, pos, cigar, flag, seq, qual)."""
    out = []
    for r in res.records:
        ri, g, path, pos, flag, sclip, eclip, mlen = [int(x) for x in r]
        s, q = seqs[ri], quals[ri]
        if flag & 0x10:
            s, q = revcomp(s), q[::-1]
        cigar = ("%dH" % sclip if sclip else "") + "%dM" % mlen + ("%dH" % eclip if eclip else "")
        out.append((names[ri][1:].decode(), idx.ref_name(g, path)[0], pos, cigar, flag, s[:mlen].decode(), bytes(q[:mlen])))
    return sorted(out)


def assert_same_result_fast(g, o, check_records=True):
    """Vectorised form of the bit-exact comparison for large batches. g: groot_b200.api.BatchResult (or any object with
    the same arrays), o: oracle.pyoracle.MapResult. Hits, pairs, every record (path, pos, flags, clips) and the counters."""
    assert g.counts == o.counts, (g.counts, o.counts)
    assert np.array_equal(g.hit_off.astype(np.uint64), o.hit_off)
    assert np.array_equal(g.hits, o.hits)
    assert g.n_pairs == len(o.pairs)
    for col, name in enumerate(("read", "graph", "n_incremented", "rec_count")):
        assert np.array_equal(g.pairs[name], o.pairs[:, col]), name
    if not check_records:
        return
    rc = g.pairs["rec_count"].astype(np.int64)
    assert int(rc.sum()) == g.n_records == len(o.records)
    assert np.array_equal(g.pairs["rec_begin"][rc > 0], (np.cumsum(rc) - rc)[rc > 0].astype(np.uint32))
    rec = o.records
    assert np.array_equal(np.repeat(g.pairs["read"], rc), rec[:, 0].astype(np.uint32))
    assert np.array_equal(np.repeat(g.pairs["graph"], rc), rec[:, 1].astype(np.uint32))
    assert np.array_equal(g.rec_path, rec[:, 2].astype(np.uint32))
    assert np.array_equal(g.rec_pos, rec[:, 3])
    first = np.zeros(len(rec), dtype=bool)
    first[(np.cumsum(rc) - rc)[rc > 0]] = True
    flags = np.where(first, 0, 0x100) | np.repeat(np.where(g.pairs["reverse"] != 0, 0x10, 0), rc)
    assert np.array_equal(flags, rec[:, 4])
    assert np.array_equal(np.repeat(g.pairs["clip_start"], rc), rec[:, 5])
    assert np.array_equal(np.repeat(g.pairs["clip_end"], rc), rec[:, 6])
