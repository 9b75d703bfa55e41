"""End-to-end through the C++ host driver `groot-b200` (groot_b200/csrc/host): FASTQ in, BAM + weighted GFAs out,
compared in the canonical decoded form of SURVEY.md §8(c) with what the reference semantics (oracle) produce:
the multiset of (read name, ref NAME, pos, CIGAR, flag, seq, qual) must be identical; header compared as the set of
(@SQ name, len) + @PG fields."""
import os
import shutil
import subprocess

import pytest

from groot_b200 import api
from oracle import pyoracle as po
from tests.util import canonical_records_from_oracle, load_fastq, pack_reads, read_bam, report

pytestmark = pytest.mark.gpu
CLI = os.path.join(os.path.dirname(api.LIB_PATH), "groot-b200")


def _run(*args, stdout=None):
    r = subprocess.run([CLI] + list(args), stdout=stdout, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    return r.stderr.decode()


def test_cli_oxa_cluster_bam_and_gfa(root, tmp_path):
    msa_dir = tmp_path / "msa"; msa_dir.mkdir()
    shutil.copy(os.path.join(root, "data", "graph", "test-genes.msa"), msa_dir / "cluster-0.msa")
    fq = os.path.join(root, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq")
    _run("index", "-m", str(msa_dir), "-i", str(tmp_path / "idx"), "-k", "51", "-s", "30", "-w", "100")
    log = _run("align", "-i", str(tmp_path / "idx"), "-f", fq, "-t", "0.99", "-c", "10", "-g", str(tmp_path / "graphs"),
               "--bamOut", str(tmp_path / "out.bam"), "--batchReads", "500")
    text, refs, recs = read_bam(str(tmp_path / "out.bam"))
    o = po.Index(msa_files=[str(msa_dir / "cluster-0.msa")], k=51, S=30, w=100)
    names, seqs, quals = load_fastq(fq)
    blob, off = pack_reads(seqs)
    res = o.map_reads(blob, off, 0.99)
    want = canonical_records_from_oracle(o, res, names, seqs, quals)
    got = sorted((r["name"], r["ref"], r["pos"], r["cigar"], r["flag"], r["seq"], r["qual"]) for r in recs)
    assert len(got) == res.counts["alignments"] > 1000
    assert got == want
    assert all(r["mapq"] == 30 and r["next_ref"] == -1 and r["tlen"] == 0 for r in recs)       # alignment.go:118-143
    assert sorted(refs) == sorted(o.ref_name(0, p) for p in range(o.stats()["paths"]))           # @SQ == every path (boss.go:63-67)
    assert "@PG\tID:1\tPN:groot\tCL:groot align\tVN:1.1.2" in text and "@HD\tVN:1.5" in text   # boss.go:55,74
    assert "number of reads received from input: 2062" in log
    assert "total number of mapped reads: %d" % res.counts["mapped"] in log
    # weighted GFA of the surviving graph (cmd/align.go:153-161): same S/L/P lines as the oracle's graph after Prune(10)
    kept = o.prune_paths(10.0)
    assert "argannot~~~(Bla)OXA-90~~~EU547443:1-825" in kept                                     # 3_sketch_test.go:49-58
    total_kmers = int(o.weights()[1].sum())
    gfa = open(tmp_path / "graphs" / "groot-graph-0.gfa").read()
    assert gfa == o.gfa_text(0, total_kmers)
    # -p 5: five workers format and deflate slices of every batch; the decoded stream is the serial writer's, record for record
    _run("align", "-i", str(tmp_path / "idx"), "-f", fq, "-t", "0.99", "-c", "10", "-g", str(tmp_path / "graphs5"),
         "--bamOut", str(tmp_path / "out5.bam"), "--batchReads", "700", "-p", "5", "--bamLevel", "1")
    text5, refs5, recs5 = read_bam(str(tmp_path / "out5.bam"))
    assert refs5 == refs and recs5 == recs


def test_cli_travis_blaB7(db_dirs, root, tmp_path):
    """testing/run_travis_tests.sh: index arg-annot.90 -w 150 -k 31 -s 20, align bla-b7-150bp-5x.fq, report -c 0.97
    must print exactly argannot~~~(Bla)B-7~~~AF189304:1-747."""
    fq = os.path.join(root, "data", "reads", "bla-b7-150bp-5x.fq")
    _run("index", "-m", db_dirs["arg-annot.90"], "-i", str(tmp_path / "idx"), "-w", "150", "-k", "31", "-s", "20")
    with open(tmp_path / "groot.bam", "wb") as f:
        _run("align", "-i", str(tmp_path / "idx"), "-f", fq, "-t", "0.99", "-g", str(tmp_path / "graphs"), stdout=f)
    text, refs, recs = read_bam(str(tmp_path / "groot.bam"))
    ref_len = dict(refs)
    rep = report([(r["ref"], r["pos"], int(r["cigar"].rstrip("MH").split("H")[-1])) for r in recs if r["flag"] != 4], ref_len, 0.97)
    assert list(rep) == ["argannot~~~(Bla)B-7~~~AF189304:1-747"]
    # the same through the driver's own report command, fed from STDIN like `groot align | groot report` (run_travis_tests.sh:43-56)
    with open(tmp_path / "groot.bam", "rb") as f:
        r = subprocess.run([CLI, "report", "-c", "0.97"], stdin=f, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    lines = [l.split("\t") for l in r.stdout.decode().splitlines()]
    assert [l[0] for l in lines] == ["argannot~~~(Bla)B-7~~~AF189304:1-747"] and int(lines[0][2]) == 747
    assert int(lines[0][1]) == rep["argannot~~~(Bla)B-7~~~AF189304:1-747"][0]


def test_cli_errors(tmp_path, root):
    r = subprocess.run([CLI, "align", "-i", str(tmp_path / "nope"), "-f", "x.fq"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0                                           # misc.ErrorCheck -> log.Fatal
    r = subprocess.run([CLI, "align", "-f", "x.fq"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"--indexDir" in r.stdout            # cmd/align.go:57-60


def test_cli_two_devices_equal_one(db_dirs, root, tmp_path):
    """`groot-b200 align --devices 0,1`: one index replica per GPU, every batch sharded over them, ONE NCCL gather of the
    compact records to the first GPU, the graph weights chained rank after rank. BAM records and weighted GFAs must be
    those of the one-GPU run (needs 2 GPUs: gpurun --gpus 2)."""
    import glob
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fq = os.path.join(root, "data", "reads", "full-argannot-perfect-reads-small.fq.gz")
    _run("index", "-m", db_dirs["arg-annot.90"], "-i", str(tmp_path / "idx"))
    _run("align", "-i", str(tmp_path / "idx"), "-f", fq, "-g", str(tmp_path / "g1"), "--bamOut", str(tmp_path / "one.bam"), "--batchReads", "301")
    log = _run("align", "-i", str(tmp_path / "idx"), "-f", fq, "-g", str(tmp_path / "g2"), "--bamOut", str(tmp_path / "two.bam"), "--batchReads", "301",
               "--devices", "0,1", "-p", "3")
    t1, refs1, recs1 = read_bam(str(tmp_path / "one.bam"))
    t2, refs2, recs2 = read_bam(str(tmp_path / "two.bam"))
    assert refs1 == refs2 and len(recs1) > 1000 and recs1 == recs2
    g1 = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "g1" / "*.gfa")))
    g2 = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "g2" / "*.gfa")))
    assert g1 == g2 and len(g1) > 0
    for f in g1:
        assert open(tmp_path / "g1" / f).read() == open(tmp_path / "g2" / f).read(), f
