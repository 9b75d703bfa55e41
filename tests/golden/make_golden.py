"""Regenerates the golden fixtures under tests/golden/ from the CPU oracle.

The reference holds no golden hash values or alignment records (its tests only log them —
src/graph/alignment_test.go:86-92, src/seqio/seqio_test.go:69-86), and its Go toolchain is absent
here, so these vectors are frozen from oracle/ AFTER the oracle itself passed every known-answer
test in tests/test_oracle_kat.py. They guard the oracle (and, on the GPU box, the CUDA path)
against regressions.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests.util import load_fastq, pack_reads  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def alignment_fixtures():
    gfa = os.path.join(ROOT, "data", "graph", "test.gfa")
    from tests.test_oracle_kat import test_alignment_fixtures_on_test_gfa  # noqa: F401  (cases live there)
    import tests.test_oracle_kat as t
    src = open(t.__file__).read()
    b10 = src.split('b10 = ("')[1].split('")')[0]
    cases = [("multimap-B7", b"ATGAAAGGATTAAAAGGG", 2, 0),
             ("segment-26", b"CCTGATATTAAAATTGAAAAATTAAAAGATAATTTATACGTCTATACAAC", 26, 0),
             ("uniq-B10", b10.encode(), 2, 0)]
    lines = []
    for tag, seq, node, off in cases:
        recs, names = po.gfa_align(gfa, 1, seq, node, off)
        for r in recs:
            lines.append("%s\t%s\t%d\t%d\t%dH%dM%dH" % (tag, names[r[0]], r[1], r[2], r[3], r[5], r[4]))
    open(os.path.join(HERE, "alignment_fixtures.txt"), "w").write("\n".join(lines) + "\n")


def sketch_vectors():
    seqs = {
        "seqA_k7_s10": ("ACTGCGTGCGTGAAACGTGCACGTGACGTG", 7, 10),
        "seqio_l2_k7_s10": ("ACAGCAGGAAGGCTTACTGGAGAAACGTATCGACTATAAGAATCGGGTGATGGAACCTCACTCTCCCATCAGCGCACAACATAGTTCGACGGGTATGACC", 7, 10),
        "seqio_l2_k31_s21": ("ACAGCAGGAAGGCTTACTGGAGAAACGTATCGACTATAAGAATCGGGTGATGGAACCTCACTCTCCCATCAGCGCACAACATAGTTCGACGGGTATGACC", 31, 21),
        "withN_k31_s21": ("ACAGCAGGAAGGCTTACTGGAGAAACGTATCGACTNTAAGAATCGGGTGATGGAACCTCACTCTCCCATCAGCGCACAACATAGTTCGACGGGTATGACC", 31, 21),
        "lower_k31_s21": ("acagcaggaaggcttactggagaaacgtatcgactataagaatcgggtgatggaacctcactctcccatcagcgcacaacatagttcgacgggtatgacc", 31, 21),
    }
    out = {}
    for name, (s, k, S) in seqs.items():
        out[name] = {"seq": s, "k": k, "S": S, "sketch": ["%016x" % int(v) for v in po.sketch(s.encode(), k, S)]}
    json.dump(out, open(os.path.join(HERE, "sketch_vectors.json"), "w"), indent=1)


def small_align_records():
    """OXA test cluster (src/pipeline/test-data/test-genes.msa) at the reference's integration-test
    parameters; first 300 reads of the reference's read set."""
    idx = po.Index(msa_files=[os.path.join(ROOT, "data", "graph", "test-genes.msa")], k=51, S=30, w=100)
    names, seqs, quals = load_fastq(os.path.join(ROOT, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq"))
    blob, off = pack_reads(seqs[:300])
    res = idx.map_reads(blob, off, 0.99)
    np.savez_compressed(os.path.join(HERE, "oxa_small_align.npz"), hit_off=res.hit_off, hits=res.hits, pairs=res.pairs,
                        records=res.records, counts=np.array([res.counts[k] for k in ("received", "mapped", "multimapped", "alignments")]),
                        index_hash=np.array([idx.dump_hash()], dtype=np.uint64), weights=idx.weights()[0])


if __name__ == "__main__":
    alignment_fixtures()
    sketch_vectors()
    small_align_records()
    print("golden fixtures written to", HERE)
