"""CPU-only tests (no GPU): the C-ABI library loads and exports every symbol include/grootgpu.h declares, the
host-side logic (MSA -> graph builder, LSH parameter optimiser, synthetic reads, multi-rank gather/merge) agrees
with the oracle, and compute entry points fail loudly without a device (no CPU fallback)."""
import glob
import os
import re

import numpy as np
import pytest

from groot_b200 import api, synth
from oracle import pyoracle as po


def test_library_exports_every_header_symbol(root):
    header = open(os.path.join(root, "include", "grootgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(grootgpu_[a-z_]+)\s*\(", header)))
    assert len(declared) >= 20
    lib = api.lib()
    for sym in declared:
        assert hasattr(lib, sym), "libgrootgpu.so does not export " + sym
    assert sorted(api.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in lib.grootgpu_version()


def test_no_cpu_fallback_without_device(root):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.GrootGpuError) as e:
        api.Index.build(msa_files=[os.path.join(root, "data", "graph", "test-genes.msa")], k=51, S=30, w=100)
    assert e.value.code == -2
    with pytest.raises(api.GrootGpuError) as e:
        api.sketch_batch(np.frombuffer(b"ACGTACGTACGT", dtype=np.uint8), np.array([0, 12], dtype=np.uint64), 7, 10)
    assert e.value.code == -2


def test_index_file_is_validated_before_a_device_is_needed(tmp_path):
    """grootgpu_index_load parses and range-checks the file on the host first: a missing file is an I/O error, a file
    that is not an index (or is cut short) a format error — on a box without a GPU too."""
    with pytest.raises(api.GrootGpuError) as e:
        api.Index.load(str(tmp_path / "missing.grootb200"))
    assert e.value.code == -3
    junk = tmp_path / "junk.grootb200"
    junk.write_bytes(b"not an index at all" * 10)
    with pytest.raises(api.GrootGpuError) as e:
        api.Index.load(str(junk))
    assert e.value.code == -4
    cut = tmp_path / "cut.grootb200"
    cut.write_bytes(b"GRTB200\x01" + b"\x1f\0\0\0" * 5 + b"\x01\0\0\0" + b"\xff" * 8)      # magic, params, one graph, absurd array length
    with pytest.raises(api.GrootGpuError) as e:
        api.Index.load(str(cut))
    assert e.value.code == -4


def test_graph_builder_matches_oracle(db_dirs, root, tmp_path):
    """Product MSA->graph builder (groot_b200/csrc/host/graph_build.cpp) vs the oracle's restatement of
    gfa.MSA2GFA + graph.CreateGrootGraph on every cluster of arg-annot.90 and on the OXA test cluster."""
    for files, params in ((sorted(glob.glob(os.path.join(db_dirs["arg-annot.90"], "cluster*.msa"))), dict(k=31, S=21, w=100)),
                          ([os.path.join(root, "data", "graph", "test-genes.msa")], dict(k=51, S=30, w=100))):
        api.graphs_dump(files, str(tmp_path / "g.txt"), **params)
        o = po.Index(msa_files=files, **params)
        o.dump_file(str(tmp_path / "o.txt"))
        mine = [ln for ln in open(tmp_path / "g.txt") if ln[0] in "GPN"]
        ref = [ln for ln in open(tmp_path / "o.txt") if ln[0] in "GPN"]
        assert len(mine) > 100 and mine == ref


def test_masked_graph_and_errors(db_dirs, tmp_path):
    files = sorted(glob.glob(os.path.join(db_dirs["card.90"], "cluster*.msa")))
    api.graphs_dump(files, str(tmp_path / "g.txt"), k=31, S=21, w=150)
    masked = [ln for ln in open(tmp_path / "g.txt") if ln.startswith("G ") and "masked=1" in ln]
    assert len(masked) >= 1                       # a card.90 cluster holds a sequence < 150 bp (pipeline/index.go:59-65)
    with pytest.raises(api.GrootGpuError) as e:
        api.graphs_dump([str(tmp_path / "missing.msa")])
    assert e.value.code == -3
    bad = tmp_path / "bad.msa"
    bad.write_text(">a\nACGT\n>b\nACG\n")
    with pytest.raises(api.GrootGpuError) as e:
        api.graphs_dump([str(bad)])
    assert e.value.code == -4


def test_query_params_match_oracle():
    for q, t in [(70, .99), (60, .99), (80, .99), (70, .97), (70, .95), (20, .99), (1, .99), (50, .9), (69, .99), (71, .99)]:
        K, L, e = api.query_params_host(q, t)
        assert (K, L) == po.optimal_kl(4, 5, 70, q, t)
        assert e == po.eq_min(21, q, 70, t)
    assert api.query_params_host(120, .99, w=150)[:2] == (4, 1)
    assert api.query_params_host(50, .99, k=51, S=30, w=100) == po.optimal_kl(4, 7, 50, 50, .99) + (po.eq_min(30, 50, 50, .99),)


def test_synth_reads_deterministic_and_composed(db_dirs):
    seqs = synth.db_sequences(db_dirs["arg-annot.90"])
    assert len(seqs) == 1749
    b1, o1 = synth.synth_reads(5000, 100, seqs, seed=42)
    b2, o2 = synth.synth_reads(5000, 100, seqs, seed=42)
    assert np.array_equal(b1, b2) and np.array_equal(o1, o2)
    assert not np.array_equal(b1, synth.synth_reads(5000, 100, seqs, seed=43)[0])
    assert set(np.unique(b1)) <= set(b"ACGTN")
    res = po.Index(msa_dir=db_dirs["arg-annot.90"]).map_reads(b1[:100 * 1000], o1[:1001], 0.99, threads=4)
    assert 0.4 < res.counts["mapped"] / 1000 < 0.6         # ~50 % exact substrings seed, the rest never does


def _cigar_clean_py(s):
    """src/reporting/reporting.go:178-213, statement for statement."""
    counter, pre, cigar, dm = 1, s[0], "", {"D": 0, "M": 0}
    for i, v in enumerate(s):
        if i == 0:
            continue
        if i == len(s) - 1:
            if v == pre:
                counter += 1
                cigar += "%d%s" % (counter, v)
            else:
                cigar += "%d%s1%s" % (counter, pre, v)
            dm[v] += 1
            break
        if v == pre:
            counter += 1
        else:
            dm[pre] += 1
            cigar += "%d%s" % (counter, pre)
            pre, counter = v, 1
    return cigar, not ((dm["D"] + dm["M"]) <= 2 or (dm["D"] == 2 and dm["M"] == 1))


def _bam_bytes(refs, recs):
    """Uncompressed BAM of records (ref_id, pos, flag, cigar ops [(len, op)]) with empty names / sequences."""
    import struct
    out = b"BAM\x01" + struct.pack("<i", 0) + struct.pack("<i", len(refs))
    for name, ln in refs:
        out += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    for ref_id, pos, flag, cig in recs:
        body = struct.pack("<iiBBHHHiiii", ref_id, pos, 2, 30, 0, len(cig), flag, 0, -1, -1, 0) + b"r\0"
        body += b"".join(struct.pack("<I", (n << 4) | op) for n, op in cig)
        out += struct.pack("<i", len(body)) + body
    return out


def test_cli_report_matches_reference_logic(root, tmp_path):
    """`groot-b200 report` (host only, no device): pileup with the reference's inclusive end, coverage cutoff, asterisk
    stripping, Flags == 4 skipped, cigarClean incl. its last-element quirk, --lowCov dropping internal gaps
    (src/reporting/reporting.go:33-213, cmd/report.go:104-129)."""
    import gzip
    import subprocess
    cli = os.path.join(root, "groot_b200", "groot-b200")
    refs = [("*geneA", 300), ("geneB", 200), ("geneC", 150), ("geneD", 100)]
    recs = [(0, p, 0, [(100, 0)]) for p in (0, 90, 180, 199)]                       # geneA fully covered
    recs += [(1, 0, 16, [(1, 5), (99, 0)]), (1, 120, 256, [(79, 0), (1, 5)])]      # geneB: internal gap, hard clips do not count
    recs += [(2, 10, 0, [(100, 0)])]                                               # geneC: 5' and 3' uncovered
    recs += [(3, 0, 4, [(100, 0)])]                                                # geneD: unaligned flag only
    raw = _bam_bytes(refs, recs)
    open(tmp_path / "x.bam", "wb").write(gzip.compress(raw[:200]) + gzip.compress(raw[200:]))    # two gzip members, like BGZF blocks
    def run(*extra):
        r = subprocess.run([cli, "report", "--bamFile", str(tmp_path / "x.bam")] + list(extra), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        return [l.split("\t") for l in r.stdout.decode().splitlines()]
    def expect(name, n, positions):
        cov = ["D"] * n
        for pos, ln in positions:
            for i in range(pos, min(pos + ln, n - 1) + 1):
                cov[i] = "M"
        return cov
    a = expect("geneA", 300, [(0, 100), (90, 100), (180, 100), (199, 100)])
    b = expect("geneB", 200, [(0, 99), (120, 79)])
    c = expect("geneC", 150, [(10, 100)])
    assert a.count("M") == 300
    got = run("-c", "0.5")
    assert got == [["geneA", "4", "300", _cigar_clean_py(a)[0]], ["geneB", "2", "200", _cigar_clean_py(b)[0]], ["geneC", "1", "150", _cigar_clean_py(c)[0]]]
    assert run() == [["geneA", "4", "300", _cigar_clean_py(a)[0]]]                  # default cutoff 0.97
    assert _cigar_clean_py(b)[1] and not _cigar_clean_py(c)[1]
    assert run("-c", "0.5", "--lowCov") == [["geneA", "4", "300", _cigar_clean_py(a)[0]]]   # --lowCov forces 0.97 (cmd/report.go:118-122)
    r = subprocess.run([cli, "report", "--bamFile", str(tmp_path / "x.txt")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0


@pytest.fixture(scope="module")
def fastq_dump(root, tmp_path_factory):
    """tests/cpp/fastq_dump.cpp linked with the host driver's reader: prints every read FastqStream yields."""
    import subprocess
    exe = str(tmp_path_factory.mktemp("fastq") / "fastq_dump")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(root, "tests", "cpp", "fastq_dump.cpp"),
                           os.path.join(root, "groot_b200", "csrc", "host", "pipeline.cpp"), "-L" + os.path.join(root, "groot_b200"), "-lgrootgpu", "-lz",
                           "-pthread", "-Wl,-rpath," + os.path.join(root, "groot_b200")])
    return exe


def test_fastq_stream_edge_cases(root, tmp_path, fastq_dump):
    """The host driver's reader (block buffer + memchr) keeps the reference's line semantics (src/pipeline/sketch.go:41-77,
    175-238): CRLF, a last line without newline, an incomplete trailing record dropped, every file scanned on its own,
    gzip by content, FASTA mode ('>' entries, an empty line ends the input), '@' check -> fatal, batches of any size."""
    import gzip
    import subprocess
    exe = fastq_dump
    def run(*args):
        r = subprocess.run([exe] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        return r.returncode, r.stdout.decode().splitlines(), r.stderr.decode()
    a = tmp_path / "a.fq"; a.write_bytes(b"@r1 x y\nACGT\n+\nIIII\n@r2\r\nAC\r\n+r2\r\nI#\r\n@r3\nGG\n+\n!!")           # no newline at the end
    b = tmp_path / "b.fq.gz"; b.write_bytes(gzip.compress(b"@r4\nTTTT\n+\nJJJJ\n@r5\nAAAA\n+\n"))                      # r5 lacks its quality line
    c = tmp_path / "c.fq"; c.write_bytes(b"@r6\nT\n+\nJ\n\n")                                                          # a trailing blank line never completes a record
    rc, out, _ = run(a, b)
    assert rc == 0 and out == ["@r1 x y\tACGT\tIIII", "@r2\tAC\tI#", "@r3\tGG\t!!", "@r4\tTTTT\tJJJJ", "#4 12"]
    assert run("--batch", "1", a, b)[1] == out and run("--batch", "1000", a, b)[1] == out
    assert run(c) == (0, ["@r6\tT\tJ", "#1 1"], "")
    # blank lines never fill a record slot in the reference (an empty scanner line is a nil slice, sketch.go:49,70,216-236):
    # between records, inside a record, many at the end — every read still comes through
    d = tmp_path / "d.fq"; d.write_bytes(b"\n@r7\nACGT\n+\nIIII\n\n\n@r8\n\nGGCC\n+\n\nJJJJ\n\n\n\n\n\n")
    assert run(d) == (0, ["@r7\tACGT\tIIII", "@r8\tGGCC\tJJJJ", "#2 8"], "")
    assert run("--batch", "1", d)[1] == ["@r7\tACGT\tIIII", "@r8\tGGCC\tJJJJ", "#2 8"]
    big = tmp_path / "big.fq"                                                                                           # lines across buffer refills
    long_seq = "ACGT" * 3_000_000
    big.write_text("@L\n%s\n+\n%s\n@S\nA\n+\nI\n" % (long_seq, "I" * len(long_seq)))
    rc, out, _ = run(big)
    assert rc == 0 and out[-1] == "#2 %d" % (len(long_seq) + 1) and out[1] == "@S\tA\tI" and len(out[0]) == 4 + 2 * len(long_seq)
    bad = tmp_path / "bad.fq"; bad.write_bytes(b"@ok\nA\n+\nI\nnot a header\nA\n+\nI\n")
    rc, _, err = run(bad)
    assert rc == 2 and "does not begin with @" in err
    fa = tmp_path / "x.fa"; fa.write_bytes(b">s1 d\nACGT\nAC\n>s2\nGG\n\n>never\nTT\n")
    rc, out, _ = run("--fasta", fa)
    assert rc == 0 and out == ["@s1 d\tACGTAC\t", "@s2\tGG\t", "#2 8"]
    rc, _, err = run(tmp_path / "missing.fq")
    assert rc == 2 and "no such file" in err
    # a FIFO (process substitution) as a file argument, plain and gzipped: it cannot be rewound after a look at its first bytes
    for payload in (a.read_bytes(), gzip.compress(a.read_bytes())):
        fifo = tmp_path / "pipe.fq"
        if fifo.exists():
            fifo.unlink()
        os.mkfifo(fifo)
        p = subprocess.Popen([exe, str(fifo)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        with open(fifo, "wb") as w:
            w.write(payload)
        o, _ = p.communicate(timeout=60)
        assert p.returncode == 0 and o.decode().splitlines()[-1] == "#3 8"


def test_header_is_c99_and_matches_the_ctypes_mirror(root, tmp_path):
    """include/grootgpu.h is what cgo and any C host compile: it must be plain C, and the Python mirror
    of its structs (groot_b200/api.py) must agree with the C compiler on every field offset and size."""
    import ctypes as C
    import subprocess
    from groot_b200 import api
    mirrors = {"grootgpu_index_params": api.IndexParams, "grootgpu_index_info": api.IndexInfo, "grootgpu_align_params": api.AlignParams,
               "grootgpu_pair": api.Pair, "grootgpu_cpair": api.CPair, "grootgpu_batch_result": api.BatchResultC}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "grootgpu.h"', "int main(void) {"]
    for cname, cls in mirrors.items():
        lines.append('    printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('    printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['    printf("COMM_ID %d\\n", GROOTGPU_COMM_ID_BYTES);', "    return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines) + "\n")
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"), "-o", exe, str(src)])
    got = dict(l.split() for l in subprocess.check_output([exe]).decode().splitlines())
    for cname, cls in mirrors.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)
    assert int(got["COMM_ID"]) == api.COMM_ID_BYTES


def test_fastq_stream_differential_fuzz(root, tmp_path, fastq_dump):
    """FastqStream against a ten-line Python restatement of the reference's reader (DataStreamer: bufio.ScanLines per file,
    '\\r' stripped, every line forwarded, empty ones as nil; FastqHandler: four non-nil lines make a read, line 1 must start
    with '@', an incomplete last group is dropped — src/pipeline/sketch.go:41-77,214-238) on random messy files: blank
    lines anywhere, CRLF, no final newline, '@' and '+' starting quality lines, several files, gzip, tiny batches."""
    import gzip
    import random
    import subprocess
    exe = fastq_dump
    rng = random.Random(5)
    alphabet = "ACGTN@+I#!:F"
    for case in range(40):
        files, lines_model = [], []
        for fi in range(rng.randint(1, 3)):
            lines = []
            for _ in range(rng.randint(0, 60)):
                kind = rng.random()
                if kind < 0.2:
                    lines.append("")
                elif kind < 0.45:
                    lines.append("@" + "".join(rng.choice(alphabet + " _/1") for _ in range(rng.randint(0, 20))))
                else:
                    lines.append("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 130))))
            eol = "\r\n" if rng.random() < 0.3 else "\n"
            text = eol.join(lines) + (eol if lines and rng.random() < 0.7 else "")
            path = tmp_path / ("c%d_%d.fq%s" % (case, fi, ".gz" if rng.random() < 0.3 else ""))
            data = text.encode()
            path.write_bytes(gzip.compress(data) if str(path).endswith(".gz") else data)
            files.append(str(path))
            # bufio.ScanLines: a final empty line (text ending in a newline) is not a line; "\r" before "\n" or at the very end is dropped
            per_file = text.split("\n")
            if per_file and per_file[-1] == "":
                per_file.pop()
            lines_model += [l[:-1] if l.endswith("\r") else l for l in per_file]
        non_empty = [l for l in lines_model if l != ""]
        want, fatal = [], False
        for i in range(0, len(non_empty) - 3, 4):
            if not non_empty[i].startswith("@"):
                fatal = True
                break
            want.append("%s\t%s\t%s" % (non_empty[i], non_empty[i + 1], non_empty[i + 3]))
        r = subprocess.run([exe, "--batch", str(rng.choice([1, 2, 7, 1000]))] + files, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if fatal:
            assert r.returncode == 2 and "does not begin with @" in r.stderr.decode(), (case, files)
        else:
            assert r.returncode == 0, (case, r.stderr.decode())
            assert r.stdout.decode().splitlines()[:-1] == want, (case, files)


def test_fastq_stream_parallel_copy(root, tmp_path, fastq_dump):
    """Blocks with thousands of records take the planned path of FastqStream::next (offsets first, then the line copies
    spread over helper threads): same reads, same order, for any thread count and batch size, from a mapped plain file,
    a gzip file and a pipe; also under ThreadSanitizer."""
    import gzip
    import subprocess
    import numpy as np
    rng = np.random.default_rng(3)
    n = 60000
    recs = []
    for r in range(n):
        L = int(rng.integers(30, 150))
        recs.append(("@r%d len=%d" % (r, L), "".join(rng.choice(list("ACGTN"), L)), "".join(rng.choice(list("FI:#@+"), L))))
    text = "".join("%s\n%s\n+\n%s\n" % rec for rec in recs[:40000]) + "\n\n" + "".join("%s\r\n%s\r\n+x\r\n%s\r\n" % rec for rec in recs[40000:])
    plain = tmp_path / "big.fq"; plain.write_text(text)
    gz = tmp_path / "big.fq.gz"; gz.write_bytes(gzip.compress(text.encode(), 1))
    want = ["%s\t%s\t%s" % rec for rec in recs] + ["#%d %d" % (n, sum(len(r[1]) for r in recs))]
    src = [os.path.join(root, "tests", "cpp", "fastq_dump.cpp"), os.path.join(root, "groot_b200", "csrc", "host", "pipeline.cpp"),
           "-L" + os.path.join(root, "groot_b200"), "-lgrootgpu", "-lz", "-pthread", "-Wl,-rpath," + os.path.join(root, "groot_b200")]
    exe = fastq_dump
    for threads, batch, f in ((1, 1000000, plain), (4, 1000000, plain), (3, 7001, plain), (4, 1000000, gz), (2, 20000, gz)):
        out = subprocess.run([exe, "--threads", str(threads), "--batch", str(batch), str(f)], stdout=subprocess.PIPE, check=True).stdout.decode().splitlines()
        assert out == want, (threads, batch, str(f))
    with open(plain, "rb") as fh:                                                  # STDIN: a pipe
        out = subprocess.run([exe, "--threads", "4", "--batch", "1000000"], stdin=fh, stdout=subprocess.PIPE, check=True).stdout.decode().splitlines()
    assert out == want
    tsan = str(tmp_path / "fastq_dump_tsan")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=thread", "-std=c++17", "-o", tsan] + src)
    r = subprocess.run([tsan, "--threads", "4", "--batch", "25000", "--count", str(plain), str(gz)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if b"FATAL: ThreadSanitizer" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment")
    assert r.returncode == 0 and b"ThreadSanitizer" not in r.stderr, r.stderr.decode()[-3000:]
    assert r.stdout.decode().split()[0] == "#%d" % (2 * n)
