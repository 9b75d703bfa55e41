"""Pins the CPU oracle (oracle/) against every golden vector / known-answer test the reference holds
for the align hot path:

  1. ntHash constants + rolling == from-scratch + RC invariance (src/minhash/minhash_test.go:10-14,111-157)
     + short-sequence error (minhash_test.go:86,101)
  2. seqio goldens (src/seqio/seqio_test.go:19-21,43-67)
  3. MSA -> GFA golden pair: arg-annot.90 cluster-139.msa <-> src/graph/test.gfa
  4. pipeline integration test (src/pipeline/1_pipeline_test.go:32-55, 3_sketch_test.go:49-58)
  5. Travis end-to-end test (testing/run_travis_tests.sh)
  6. accuracy self-check (testing/run_accuracy_tests.sh + testing/groot-accuracy.go)
  7. alignment fixtures (src/graph/alignment_test.go:13,27,41)
"""
import os
import random

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import load_fastq, pack_reads, read_msa, report, revcomp

M64 = (1 << 64) - 1


# ---------------------------------------------------------------- 1. ntHash / KHF
def test_nthash_known_answers():
    assert po.ntf64(b"TGCAG", 5) == 0x0BAFA6728FC6DABF
    assert po.ntr64(b"TGCAG", 5) == 0x8CF2D4072CCA480E
    assert min(po.ntf64(b"ACGTC", 5), po.ntr64(b"ACGTC", 5)) == 0x480202D54E8EBECD
    assert po.kmer_hashes(b"TGCAG", 5)[0] == 0x0BAFA6728FC6DABF
    assert po.kmer_hashes(b"ACGTC", 5)[0] == 0x480202D54E8EBECD


def test_nthash_rolling_equals_scratch_and_rc():
    rng = random.Random(7)
    for k in (5, 7, 31, 51, 64, 70):
        s = bytes(rng.choice(b"ACGTN" if k == 7 else b"ACGT") for _ in range(200))
        rolled = po.kmer_hashes(s, k)
        scratch = [min(po.ntf64(s[i:i + k], k), po.ntr64(s[i:i + k], k)) for i in range(len(s) - k + 1)]
        assert rolled.tolist() == scratch
        if b"N" not in s:
            assert sorted(po.kmer_hashes(revcomp(s), k).tolist()) == sorted(scratch)


def test_multihash_formula():
    h, k, S = 0x0BAFA6728FC6DABF, 31, 21
    mh = po.multi_hash(h, k, S)
    assert mh[0] == h
    for i in range(1, S):
        x = (h * (i ^ ((k * 0x90B45D39FB6DA1FA) & M64))) & M64
        x ^= x >> 27
        assert int(mh[i]) == x


def test_khf_rc_invariance_reference_vectors():
    # src/minhash/minhash_test.go:10-14,111-146: similarity of seqA and its reverse complement must be 1.0
    seqA = b"ACTGCGTGCGTGAAACGTGCACGTGACGTG"
    seqArc = b"CACGTCACGTGCACGTTTCACGCACGCAGT"
    assert revcomp(seqA) == seqArc
    a, b = po.sketch(seqA, 7, 10), po.sketch(seqArc, 7, 10)
    assert (a == b).all()
    # KHF definition: slot i is the min over k-mers of multihash i
    hs = po.kmer_hashes(seqA, 7)
    want = np.min(np.stack([po.multi_hash(int(h), 7, 10) for h in hs]), axis=0)
    assert (a == want).all()


def test_khf_short_sequence_errors():
    with pytest.raises(ValueError):
        po.sketch(b"A", 7, 10)            # minhash_test.go:86
    po.sketch(b"ACGTACG", 7, 10)          # len == k is fine


# ---------------------------------------------------------------- 2. seqio goldens
def test_seqio_goldens():
    l2 = b"acagcaggaaggcttactggagaaacgtatcgactataagaatcgggtgatggaacctcactctcccatcagcgcacaacatagttcgacgggtatgacc"
    l4 = b"====@==@AAD?>D@@==DACBC?@BB@C==AB==A@D>AD==?CB==@=B?=A>D?=DB=?>>D@EB===??=@C=?C>@>@B>=?C@@>=====?@>="
    upper = b"ACAGCAGGAAGGCTTACTGGAGAAACGTATCGACTATAAGAATCGGGTGATGGAACCTCACTCTCCCATCAGCGCACAACATAGTTCGACGGGTATGACC"
    trimmed = b"GAAGGCTTACTGGAGAAACGTATCGACTATAAGAATCGGGTGATGGAACCTCACTCTCCCATCAGCGCACAACATAGTTCGAC"
    rc = b"GTCGAACTATGTTGTGCGCTGATGGGAGAGTGAGGTTCCATCACCCGATTCTTATAGTCGATACGTTTCTCCAGTAAGCCTTC"
    s = po.basecheck(l2)
    assert s == upper
    s, q = po.qualtrim(s, l4, 30)
    assert s == trimmed
    s2, q2 = po.revcomp(s, q)
    assert s2 == rc and q2 == q[::-1]
    assert po.basecheck(b"acgtRYn-x") == b"ACGTNNNNN"
    with pytest.raises(ValueError):
        po.revcomp(b"ACGa", b"IIII")      # Go: index out of range on complementBases


# ---------------------------------------------------------------- 3. MSA -> GFA golden pair
def _gfa_signatures(text):
    segs, links, paths = {}, [], []
    for line in text.split("\n"):
        f = line.split("\t")
        if f[0] == "S":
            segs[f[1]] = f[2].upper()
        elif f[0] == "L":
            links.append((f[1], f[3]))
        elif f[0] == "P":
            paths.append((f[1], [s.rstrip("+") for s in f[2].split(",") if s]))
    # signature of a node = (sequence, sorted tuple of (path name, start position in that path))
    member = {s: [] for s in segs}
    spelled = {}
    for name, ps in paths:
        pos = 0
        for s in ps:
            member[s].append((name, pos))
            pos += len(segs[s])
        spelled[name] = "".join(segs[s] for s in ps)
    sig = {s: (segs[s], tuple(sorted(member[s]))) for s in segs}
    return sorted(sig.values()), sorted((sig[a], sig[b]) for a, b in links), spelled, [p[0] for p in paths]


def test_msa2gfa_matches_reference_golden_gfa(db_dirs, root):
    mine = po.msa2gfa_text(os.path.join(db_dirs["arg-annot.90"], "cluster-139.msa"))
    gold = open(os.path.join(root, "data", "graph", "test.gfa")).read()
    n1, l1, s1, o1 = _gfa_signatures(mine)
    n2, l2, s2, o2 = _gfa_signatures(gold)
    assert len(n1) == 133 and len(l1) == 176 and len(o1) == 6
    assert n1 == n2          # same segments: sequence + path membership + positions
    assert l1 == l2          # same link set
    assert o1 == o2          # same path order (== MSA row order)
    rows = dict(read_msa(os.path.join(db_dirs["arg-annot.90"], "cluster-139.msa")))
    for name, seq in s1.items():
        assert seq == rows[name].replace("-", "").upper() == s2[name]


# ---------------------------------------------------------------- LSH Ensemble parameters
def test_optimal_kl_and_containment_thresholds():
    for x, q, t in [(70, 70, .99), (120, 120, .99), (70, 60, .99), (70, 90, .99), (50, 50, .99), (70, 70, .97)]:
        assert po.optimal_kl(4, 5, x, q, t) == (4, 1)
    assert po.eq_min(21, 70, 70, .99) == 21
    assert po.eq_min(21, 60, 70, .99) == 18
    assert po.eq_min(21, 80, 70, .99) == 22      # unsatisfiable: reads longer than the window never seed
    assert po.eq_min(21, 70, 70, .97) == 20
    assert po.eq_min(21, 70, 70, .95) == 19
    a = np.arange(21, dtype=np.uint64)
    b = a.copy(); b[3] = 99
    assert po.containment(a, a, 70, 70) == 1.0
    assert abs(po.containment(a, b, 70, 70) - 2 * (20 / 21) / (1 + 20 / 21)) < 1e-15
    assert po.containment(a, a + 100, 70, 70) == 0.0


# ---------------------------------------------------------------- 4. pipeline integration test
def test_pipeline_integration_OXA90(root):
    idx = po.Index(msa_files=[os.path.join(root, "data", "graph", "test-genes.msa")], k=51, S=30, w=100, num_part=8, max_k=4)
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq"))
    blob, off = pack_reads(seqs)
    res = idx.map_reads(blob, off, threshold=0.99)
    assert res.counts["received"] == len(seqs) == 2062
    assert res.counts["mapped"] > 0
    kept = idx.prune_paths(10.0)
    assert "argannot~~~(Bla)OXA-90~~~EU547443:1-825" in kept      # 3_sketch_test.go:49-58


# ---------------------------------------------------------------- 5. Travis end-to-end
@pytest.fixture(scope="module")
def travis_index(db_dirs):
    return po.Index(msa_dir=db_dirs["arg-annot.90"], k=31, S=20, w=150)


def test_travis_e2e_reports_only_blaB7(travis_index, root):
    idx = travis_index
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "bla-b7-150bp-5x.fq"))
    blob, off = pack_reads(seqs)
    res = idx.map_reads(blob, off, threshold=0.99)
    recs, ref_len = [], {}
    for r in res.records:
        name, ln = idx.ref_name(int(r[1]), int(r[2]))
        ref_len[name] = ln
        recs.append((name, int(r[3]), int(r[7])))
    rep = report(recs, ref_len, 0.97)
    assert list(rep) == ["argannot~~~(Bla)B-7~~~AF189304:1-747"]    # run_travis_tests.sh:43-56


# ---------------------------------------------------------------- 6. accuracy self-check
def test_accuracy_selfcheck_perfect_reads(db_dirs, root):
    idx = po.Index(msa_dir=db_dirs["arg-annot.90"], k=41, S=21, w=150)
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "argannot-150bp-10000-reads.fq.gz"))
    n = 2000
    names, seqs = names[:n], seqs[:n]
    blob, off = pack_reads(seqs)
    res = idx.map_reads(blob, off, threshold=0.99, threads=4)
    by_read = {}
    for r in res.records:
        by_read.setdefault(int(r[0]), []).append((idx.ref_name(int(r[1]), int(r[2]))[0].lstrip("*"), int(r[3])))
    aligned = correct = correct_start = 0
    for i, nm in enumerate(names):
        if i not in by_read:
            continue
        aligned += 1
        f = nm[1:].decode().split("_")               # groot-accuracy.go:86-96
        true_ref = f[9].split("$")[0].split(" ")[0].replace("{", "_")   # bbmap writes '_' as '{' in read names
        true_pos = int(f[2])
        if any(ref == true_ref for ref, _ in by_read[i]):
            correct += 1
        if (true_ref, true_pos) in by_read[i]:
            correct_start += 1
    # perfect reads: nearly everything aligns (misses = the never-emitted last window group of each path)
    assert aligned / n > 0.97
    assert correct / aligned > 0.995         # the reference's own accuracy script also tolerates a few "incorrectly aligned reads"
    assert correct_start / aligned > 0.99


# ---------------------------------------------------------------- 7. alignment fixtures
def test_alignment_fixtures_on_test_gfa(root):
    gfa = os.path.join(root, "data", "graph", "test.gfa")
    golden = os.path.join(root, "tests", "golden", "alignment_fixtures.txt")
    b10 = ("ATGAAAGGATTAAAAGGGCTATTGGTTCTGGCTTTAGGCTTTACAGGACTACAGGTTTTTGGGCAACAGAACCCTGATATTAAAATTGAAAAATTAAAAGATAATTTATACGTCTATACAACCTATAATACCTTCAAAGGAACTAAATATGCGGCTAATGCGGTATATATGGTAACCGATAAAGGAGTAGTGGTTATAGACTCTCCATGGGGAGAAGATAAATTTAAAAGTTTTACAGACGAGATTTATAAAAAGCACGGAAAGAAAGTTATCATGAACATTGCAACCCACTCTCATGATGATAGAGCCGGAGGTCTTGAATATTTTGGTAAACTAGGTGCAAAAACTTATTCTACTAAAATGACAGATTCTATTTTAGCAAAAGAGAATAAGCCAAGAGCAAAGTACACTTTTGATAATAATAAATCTTTTAAAGTAGGAAAGACTGAGTTTCAGGTTTATTATCCGGGAAAAGGTCATACAGCAGATAATGTGGTTGTGTGGTTTCCTAAAGACAAAGTATTAGTAGGAGGCTGCATTGTAAAAAGTGGTGATTCGAAAGACCTTGGGTTTATTGGGGAAGCTTATGTAAACGACTGGACACAGTCCATACACAACATTCAGCAGAAATTTCCCTATGTTCAGTATGTCGTTGCAGGTCATGACGACTGGAAAGATCAAACATCAATACAACATACACTGGATTTAATCAGTGAATATCAACAAAAACAAAAGGCTTCAAATTAA")
    cases = [
        ("multimap-B7", b"ATGAAAGGATTAAAAGGG", 2, 0),                                                   # alignment_test.go:13
        ("segment-26", b"CCTGATATTAAAATTGAAAAATTAAAAGATAATTTATACGTCTATACAAC", 26, 0),                  # :27
        ("uniq-B10", b10.encode(), 2, 0),                                                             # :41
    ]
    lines = []
    for tag, seq, node, off in cases:
        recs, names = po.gfa_align(gfa, 1, seq, node, off)
        for r in recs:
            lines.append("%s\t%s\t%d\t%d\t%dH%dM%dH" % (tag, names[r[0]], r[1], r[2], r[3], r[5], r[4]))
    text = "\n".join(lines) + "\n"
    # the reference test only LOGS these records (alignment_test.go:86-92); the expectations below are
    # what its semantics imply and are frozen in tests/golden/alignment_fixtures.txt
    got = {l.split("\t")[0]: [] for l in lines}
    for l in lines:
        got[l.split("\t")[0]].append(l.split("\t")[1])
    assert "argannot~~~(Bla)B-7~~~AF189304:1-747" in got["multimap-B7"] and len(got["multimap-B7"]) > 1
    assert got["uniq-B10"] == ["*argannot~~~(Bla)B-10~~~AY348325:1-747"]
    assert len(got["segment-26"]) == 6     # segment 26 is shared by all six paths
    if not os.path.exists(golden):
        pytest.fail("golden file missing: run tests/golden/make_golden.py")
    assert text == open(golden).read()
