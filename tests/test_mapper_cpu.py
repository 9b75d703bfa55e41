"""The host driver's ReadMapper (FASTQ -> batches -> device calls -> BAM; groot_b200/csrc/host/pipeline.cpp, mirror of
theBoss.mapReads, src/pipeline/boss.go:107-242) on a machine WITHOUT a GPU: the driver is linked against
tests/cpp/mock_grootgpu.cpp, a stand-in for libgrootgpu.so whose "alignment" is a pure function of a hash of the read
(restated below). What is under test is everything around the device calls: the reader thread and its batch slots, the
BAM stage, the lifetime of the result arrays (the mock poisons them at the next call), the rank team of --devices, the
counters — under ThreadSanitizer as well. The device side itself is covered by the -m gpu tests."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

from tests.bam_model import Batch, check_blocks as _check_blocks

GRAPHS, PATHS, NODES = 11, 25, 500


def _build(root, out, extra=()):
    host = os.path.join(root, "groot_b200", "csrc", "host")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-Wall", "-pthread", *extra, "-o", out, os.path.join(root, "tests", "cpp", "mapper_mock.cpp"),
                           os.path.join(root, "tests", "cpp", "mock_grootgpu.cpp"), os.path.join(host, "pipeline.cpp"), "-lz"])
    return out


@pytest.fixture(scope="module")
def mapper(root, tmp_path_factory):
    return _build(root, str(tmp_path_factory.mktemp("mapper") / "mapper_mock"))


def _fnv1a(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _reads(rng, n, fixed_len=None):
    out = []
    for r in range(n):
        L = fixed_len or int(rng.integers(1, 260))
        out.append((b"@r%06d/%d extra" % (r, r % 7), bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), L, p=[0.24, 0.24, 0.24, 0.24, 0.04])),
                    bytes(rng.integers(33, 75, L).astype(np.uint8)).replace(b"@", b"A")))
    return out


def _expected(reads, no_align=False):
    """What the mock decides for every read (tests/cpp/mock_grootgpu.cpp: grootgpu_align_batch), as a test_bam_cpu.Batch."""
    b = Batch.__new__(Batch)
    b.path_bytes = 1
    b.graph_paths = np.full(GRAPHS, PATHS)
    b.graph_ref_base = (np.arange(GRAPHS + 1) * PATHS).astype(np.uint32)
    b.refs = [("gene_%d_%d" % (g, p), 2000) for g in range(GRAPHS) for p in range(PATHS)]
    b.nodes = [(n % GRAPHS, np.arange(PATHS, dtype=np.uint32), np.array([(n * 7 + p * 13) % 1000 for p in range(PATHS)], dtype=np.int32)) for n in range(NODES)]
    b.ids, b.seqs, b.quals = [r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads]
    b.cpairs, b.rec_path = [], []
    mapped = kmers = 0
    for r, (_, seq, _) in enumerate(reads):
        h = _fnv1a(seq)
        if h % 100 >= 52:
            continue
        mapped += 1
        if no_align:
            continue
        L = len(seq)
        cnt = 1 + (h >> 24) % min(PATHS, 20)
        flags = (h >> 44) % 50
        if (h >> 40) & 1: flags |= 0x10000000
        if L >= 3 and (h >> 52) & 1: flags |= 0x20000000
        if L >= 3 and (h >> 53) & 1: flags |= 0x40000000
        b.cpairs.append((r, (h >> 8) % NODES, flags, cnt))
        b.rec_path += list(range(cnt))
        kmers += L
    return b, mapped, kmers


def _write_fastq(path, reads, gz=False):
    data = b"".join(i + b"\n" + s + b"\n+\n" + q + b"\n" for i, s, q in reads)
    open(path, "wb").write(gzip.compress(data) if gz else data)


def _records_part(raw):
    """Decompressed BAM minus the header text (it carries the date): (references block + records)."""
    data = _check_blocks(raw)
    assert data[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", data, 4)[0]
    return data[8:8 + l_text], data[8 + l_text:]


def _run(exe, args):
    r = subprocess.run([exe] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    f = r.stdout.decode().split()
    return [int(x) for x in f[:5]], float(f[5]), r.stderr.decode()


def _check(exe, tmp_path, reads, args, files=None):
    if files is None:
        files = [str(tmp_path / "reads.fq")]
        _write_fastq(files[0], reads)
    out = str(tmp_path / "out.bam")
    stats, _, _ = _run(exe, ["--bam", out] + list(args) + files)
    want, mapped, kmers = _expected(reads)
    exp = want.expected()
    text, rest = _records_part(open(out, "rb").read())
    refs_block = 4 + sum(4 + len(n) + 1 + 4 for n, _ in want.refs)                     # n_ref + the reference entries
    assert rest == exp[want.header_bytes - refs_block:]
    assert text.startswith(b"@HD\tVN:1.5\tSO:unknown\n@SQ\tSN:gene_0_0\tLN:2000\n") and b"@PG\tID:1\tPN:groot\tCL:groot align\tVN:1.1.2\n" in text
    assert stats == [len(reads), mapped, 0, len(want.rec_path), kmers]


@pytest.mark.parametrize("args", [("-p", 1, "--batch", 100), ("-p", 4, "--batch", 257), ("-p", 3, "--batch", 100000), ("-p", 2, "--batch", 64, "--delta", 0),
                                  ("-p", 5, "--batch", 301, "--level", 1)])
def test_mapper_bam_equals_expected(mapper, tmp_path, args):
    _check(mapper, tmp_path, _reads(np.random.default_rng(7), 3000), args)


def test_mapper_fixed_length_reads_and_several_files(mapper, tmp_path):
    rng = np.random.default_rng(8)
    reads = _reads(rng, 2500, fixed_len=100)
    files = [str(tmp_path / "a.fq"), str(tmp_path / "b.fq.gz"), str(tmp_path / "c.fq")]
    _write_fastq(files[0], reads[:900]); _write_fastq(files[1], reads[900:1700], gz=True); _write_fastq(files[2], reads[1700:])
    _check(mapper, tmp_path, reads, ("-p", 4, "--batch", 333), files)


@pytest.mark.parametrize("devices", [2, 3, 8])
def test_mapper_rank_team_equals_one_device(mapper, tmp_path, devices):
    """--devices: shards per rank, one gather per batch (emulated by the mock with threads): same BAM, same counters —
    also when a batch has fewer reads than ranks (empty shards take part in the gather)."""
    _check(mapper, tmp_path, _reads(np.random.default_rng(9), 2001), ("-p", 3, "--batch", 250, "--devices", devices))
    _check(mapper, tmp_path, _reads(np.random.default_rng(10), 5), ("-p", 2, "--batch", 3, "--devices", devices))


def test_mapper_no_align_and_empty_input(mapper, tmp_path):
    reads = _reads(np.random.default_rng(11), 500)
    f = str(tmp_path / "reads.fq")
    _write_fastq(f, reads)
    stats, _, _ = _run(mapper, ["--bam", str(tmp_path / "none.bam"), "--noAlign", "--batch", 100, f])
    _, mapped, _ = _expected(reads, no_align=True)
    assert stats[:4] == [500, mapped, 0, 0] and not os.path.exists(tmp_path / "none.bam")          # boss.go:112-116: no BAM with --noAlign
    open(tmp_path / "empty.fq", "wb").close()
    r = subprocess.run([mapper, "--bam", str(tmp_path / "e.bam"), str(tmp_path / "empty.fq")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 2 and "no fastq reads received" in r.stderr.decode()                   # sketch.go:275-277


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_mapper_under_sanitizers(root, tmp_path, sanitizer):
    exe = _build(root, str(tmp_path / "mapper_san"), extra=("-fsanitize=" + sanitizer, "-fno-sanitize-recover=undefined"))
    reads = _reads(np.random.default_rng(12), 1500)
    probe = subprocess.run([exe, "--bam", str(tmp_path / "probe.bam"), "/dev/null"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if b"FATAL: ThreadSanitizer" in probe.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment")
    for args in (("-p", 4, "--batch", 200), ("-p", 3, "--batch", 128, "--devices", 4)):
        _check(exe, tmp_path, reads, args)
        out = subprocess.run([exe, "--bam", str(tmp_path / "t.bam")] + [str(a) for a in args] + [str(tmp_path / "reads.fq")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert out.returncode == 0 and b"Sanitizer" not in out.stderr, out.stderr.decode()[-3000:]
