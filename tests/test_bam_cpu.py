"""Host driver's BAM output (groot_b200/csrc/host: format_batch_bam, BamWriter, bgzf.h) on fabricated batches, no device:
the decompressed stream must equal, byte for byte, the BAM an independent Python writer produces from the same compact
results (record fields as AlignRead builds them, src/graph/alignment.go:114-156,296-315; header as setupBAM,
src/pipeline/boss.go:45-105), every BGZF block must be well-formed, and that for every worker count, deflate level and with
the hint-driven block encoder on and off."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

from tests.bam_model import Batch, check_blocks as _check_blocks


@pytest.fixture(scope="module")
def harness(root, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bam") / "bam_batch")
    host = os.path.join(root, "groot_b200", "csrc", "host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(root, "tests", "cpp", "bam_batch.cpp"), os.path.join(host, "pipeline.cpp"),
                           "-L" + os.path.join(root, "groot_b200"), "-lgrootgpu", "-lz", "-pthread", "-Wl,-rpath," + os.path.join(root, "groot_b200")])
    return exe


COVERAGE = {"fixed": 0, "own": 0, "zlib": 0}       # blocks by encoder over the whole module (checked by the last test)


def _run(harness, batch_file, out, workers, level, delta):
    r = subprocess.run([harness, batch_file, out, str(workers), str(level), str(delta)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    secs, raw_bytes, bam_bytes, delta_blocks, own_code, zlib_blocks = r.stdout.decode().split()
    COVERAGE["fixed"] += int(delta_blocks) - int(own_code); COVERAGE["own"] += int(own_code); COVERAGE["zlib"] += int(zlib_blocks)
    return int(raw_bytes), int(bam_bytes), int(delta_blocks)


@pytest.mark.parametrize("case", ["mixed", "many_paths", "single_path", "long_reads", "two_byte_paths", "binned_quals"])
def test_bam_stream_equals_independent_writer(harness, tmp_path, case):
    rng = np.random.default_rng(sum(map(ord, case)))
    b = {"mixed": lambda: Batch(rng, 3000, 1, 300, 20),
         "many_paths": lambda: Batch(rng, 2500, 90, 110, 40, zero_frac=0.02),
         "single_path": lambda: Batch(rng, 6000, 100, 100, 1),
         "long_reads": lambda: Batch(rng, 60, 20000, 70000, 4),                    # records longer than a block and than the deflate window
         "two_byte_paths": lambda: Batch(rng, 2000, 50, 150, 12, path_bytes=2),
         "binned_quals": lambda: Batch(rng, 2500, 150, 150, 25, quals="binned")}[case]()
    f = str(tmp_path / "batch.bin")
    b.write(f)
    want = b.expected()
    for workers, level, delta in ((1, -1, 1), (1, -1, 0), (3, -1, 1), (8, 1, 1), (2, 0, 1), (5, 9, 0)):
        out = str(tmp_path / ("o_%d_%d_%d.bam" % (workers, level, delta)))
        raw_bytes, bam_bytes, delta_blocks = _run(harness, f, out, workers, level, delta)
        raw = open(out, "rb").read()
        got = _check_blocks(raw)
        assert got == want, (case, workers, level, delta)
        assert gzip.decompress(raw) == want
        assert raw_bytes == len(want) - b.header_bytes
        if not delta or level == 0:
            assert delta_blocks == 0
        if case in ("many_paths", "binned_quals") and delta and level != 0:
            assert delta_blocks > 0 and bam_bytes < raw_bytes // 4            # the repeats are found: most blocks go through the hint-driven encoder


def test_bam_tiny_batch_takes_the_fixed_code(harness, tmp_path):
    """A block of a few short records: a code of its own would cost more than it saves (RFC 1951 3.2.6 block)."""
    rng = np.random.default_rng(3)
    b = Batch(rng, 1, 12, 12, 1, zero_frac=0.0)
    b.cpairs, b.rec_path = [(0, 0, 5, 4)], [int(b.nodes[0][1][0])] * 4
    f = str(tmp_path / "batch.bin")
    b.write(f)
    out = str(tmp_path / "o.bam")
    before = dict(COVERAGE)
    _run(harness, f, out, 1, -1, 1)
    assert COVERAGE["fixed"] == before["fixed"] + 1
    assert _check_blocks(open(out, "rb").read()) == b.expected()


def test_bam_empty_batch(harness, tmp_path):
    rng = np.random.default_rng(5)
    b = Batch(rng, 10, 50, 60, 3, zero_frac=1.0)                                   # pairs without records only
    f = str(tmp_path / "batch.bin")
    b.write(f)
    out = str(tmp_path / "o.bam")
    _run(harness, f, out, 4, -1, 1)
    assert _check_blocks(open(out, "rb").read()) == b.expected()


def test_report_reads_what_the_block_writer_wrote(harness, root, tmp_path):
    """`groot-b200 report` (its own BGZF reader, src/reporting/reporting.go:33-173) gives the same lines for the stream
    written through the hint-driven encoder as for the zlib-only one."""
    rng = np.random.default_rng(11)
    b = Batch(rng, 1500, 80, 120, 30, zero_frac=0.0)
    for i, (name, _) in enumerate(b.refs):
        b.refs[i] = (name, 8000)                                                   # longer than any position + read
    f = str(tmp_path / "batch.bin")
    b.write(f)
    lines = []
    for delta in (1, 0):
        out = str(tmp_path / ("d%d.bam" % delta))
        _, _, delta_blocks = _run(harness, f, out, 3, -1, delta)
        assert (delta_blocks > 0) == bool(delta)
        r = subprocess.run([os.path.join(root, "groot_b200", "groot-b200"), "report", "--bamFile", out, "-c", "0.01"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        lines.append(r.stdout.decode())
    assert lines[0] == lines[1] and lines[0].count("\n") > 10


def test_every_block_encoder_was_exercised():
    """The cases above went through all three ways a block is written: fixed code, a code of its own, zlib."""
    assert COVERAGE["fixed"] > 0 and COVERAGE["own"] > 0 and COVERAGE["zlib"] > 0, COVERAGE


def test_block_writer_fuzz(root, tmp_path):
    """bgzf.h on random streams with right, wrong and absent hints, under AddressSanitizer / UBSan: every block inflates
    (zlib) to exactly the bytes that went in — the encoder never trusts a hint."""
    exe = str(tmp_path / "bgzf_fuzz")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17", "-Wall", "-o", exe,
                           os.path.join(root, "tests", "cpp", "bgzf_fuzz.cpp"), "-lz"])
    for seed in (11, 12, 13):
        r = subprocess.run([exe, str(seed), "25"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0 and r.stdout.decode().startswith("ok "), r.stdout.decode() + r.stderr.decode()
        _, blocks, delta, own = r.stdout.decode().split()
        assert int(delta) > 100 and int(own) > 100


def test_bam_read_without_qualities_is_an_error(harness, tmp_path):
    """FASTA input reaching the BAM writer: the reference panics when it reverse-complements or slices the empty quality
    string (src/seqio/seqio.go:125-127, src/graph/alignment.go:121); the driver reports it instead of writing garbage."""
    rng = np.random.default_rng(4)
    b = Batch(rng, 20, 50, 60, 3, zero_frac=0.0)
    b.quals = [b""] * len(b.quals)
    f = str(tmp_path / "batch.bin")
    b.write(f)
    r = subprocess.run([harness, f, str(tmp_path / "o.bam"), "2", "-1", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 2 and "quality string" in r.stderr.decode()
