"""GPU parity tests (run on the B200 box: pytest -m gpu). Every test drives the CUDA path through the C ABI
(libgrootgpu.so via groot_b200.api) and compares it, bit for bit, with the CPU oracle on the same inputs,
with the committed golden fixtures, or through size-independent properties at full size.

Parity bar: integer / byte / index work -> bit-exact. The f64 graph weights are compared bit-exactly too,
because the host replay performs the additions in the reference's (read) order.
"""
import json
import os

import numpy as np
import pytest

from groot_b200 import api, synth
from oracle import pyoracle as po
from tests.util import assert_same_result_fast, load_fastq, pack_reads, revcomp

pytestmark = pytest.mark.gpu


def _have_gpu():
    try:
        return api.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if not _have_gpu():
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box (there is no CPU fallback)")


def oracle_records_table(res):
    return res.records[:, :7].astype(np.int64)


def assert_same_result(g, o, check_records=True):
    """g: api.BatchResult, o: pyoracle.MapResult"""
    assert g.counts == o.counts
    assert np.array_equal(g.hit_off.astype(np.uint64), o.hit_off)
    assert np.array_equal(g.hits, o.hits)
    assert g.n_pairs == len(o.pairs)
    assert np.array_equal(g.pairs["read"], o.pairs[:, 0])
    assert np.array_equal(g.pairs["graph"], o.pairs[:, 1])
    assert np.array_equal(g.pairs["n_incremented"], o.pairs[:, 2])
    assert np.array_equal(g.pairs["rec_count"], o.pairs[:, 3])
    if check_records:
        assert np.array_equal(g.records_table(), oracle_records_table(o))


# ------------------------------------------------------------------------------------------------- sketches
def test_sketch_golden_vectors(root):
    vec = json.load(open(os.path.join(root, "tests", "golden", "sketch_vectors.json")))
    for name, v in vec.items():
        blob, off = pack_reads([v["seq"].encode()])
        got = api.sketch_batch(blob, off, v["k"], v["S"])[0]
        assert ["%016x" % int(x) for x in got] == v["sketch"], name


def test_sketch_matches_oracle_random():
    rng = np.random.default_rng(5)
    for k, S in [(7, 10), (31, 21), (51, 30), (31, 20), (41, 21), (31, 32), (15, 8), (31, 16), (31, 24)]:
        seqs = []
        for i in range(300):
            ln = int(rng.integers(k, 260))
            alphabet = b"ACGT" if i % 3 else b"ACGTNacgtRY"
            seqs.append(bytes(rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=ln)))
        seqs.append(seqs[0][:k])     # len == k: a single k-mer
        blob, off = pack_reads(seqs)
        got = api.sketch_batch(blob, off, k, S)
        for i, s in enumerate(seqs):
            assert np.array_equal(got[i], po.sketch(s, k, S)), (k, S, i)
        # RC invariance on the ACGT-only ones (src/minhash/minhash_test.go:111-157)
        acgt = [s for s in seqs if set(s) <= set(b"ACGT")]
        b2, o2 = pack_reads([revcomp(s) for s in acgt])
        b1, o1 = pack_reads(acgt)
        assert np.array_equal(api.sketch_batch(b1, o1, k, S), api.sketch_batch(b2, o2, k, S))


def test_sketch_tie_path_is_exact(monkeypatch, argannot, db_dirs):
    """The sketch kernel orders k-mer products by their top 27 bits and only finishes / compares in full on a tie
    (seed_kernels.cuh, variant 6) — about once in 10^5 reads. GROOTGPU_KHF_KEYBITS lowers the number of deciding bits so
    that nearly every comparison ties and takes the exact path: sketches, hits and records must not change."""
    rng = np.random.default_rng(11)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(rng.integers(31, 300)))) for _ in range(400)]
    seqs.append(b"A" * 120)                       # one k-mer repeated: every comparison is an exact tie
    seqs.append(b"ACGTN" * 30)
    blob, off = pack_reads(seqs)
    want = {(k, S): np.stack([po.sketch(s, k, S) for s in seqs]) for k, S in ((31, 21), (21, 32), (31, 8))}
    g, o = argannot
    rb, ro = _c1_reads(db_dirs["arg-annot.90"], 3000, 100, seed=21)
    orr = o.map_reads(rb, ro, 0.99, threads=8)
    for bits in ("27", "12", "3", "1"):
        monkeypatch.setenv("GROOTGPU_KHF_KEYBITS", bits)
        for (k, S), w in want.items():
            assert np.array_equal(api.sketch_batch(blob, off, k, S), w), (bits, k, S)
        for keep in (False, True):               # two-pass (prescreen + queued) and one-pass seed kernels
            gr = g.map_reads(rb, ro, 0.99, keep_sketches=keep)
            assert_same_result(gr, orr)
    monkeypatch.delenv("GROOTGPU_KHF_KEYBITS")


def test_sketch_short_sequence_is_an_error():
    blob, off = pack_reads([b"ACGTACGTAC", b"ACG"])
    with pytest.raises(api.GrootGpuError) as e:
        api.sketch_batch(blob, off, 7, 10)
    assert e.value.code == -5       # minhash_test.go:86: AddSequence must fail when len < k


# ------------------------------------------------------------------------------------------------- index
@pytest.fixture(scope="module")
def oxa(root):
    f = [os.path.join(root, "data", "graph", "test-genes.msa")]
    return api.Index.build(msa_files=f, k=51, S=30, w=100), po.Index(msa_files=f, k=51, S=30, w=100)


def test_index_parity_oxa_cluster(oxa, tmp_path):
    g, o = oxa
    g.dump_file(str(tmp_path / "g.txt"))
    o.dump_file(str(tmp_path / "o.txt"))
    assert open(tmp_path / "g.txt").read() == open(tmp_path / "o.txt").read()
    assert g.dump_hash() == o.dump_hash()
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "oxa_small_align.npz"))
    assert g.dump_hash() == int(golden["index_hash"][0])


@pytest.fixture(scope="module")
def argannot(db_dirs):
    d = db_dirs["arg-annot.90"]
    return api.Index.build(msa_dir=d, k=31, S=21, w=100), po.Index(msa_dir=d, k=31, S=21, w=100)


def test_index_parity_argannot(argannot):
    g, o = argannot
    gi, oi = g.info(), o.stats()
    for key in ("graphs", "masked", "paths", "nodes", "path_bases", "windows", "raw_windows", "max_merge_span"):
        assert gi[key] == oi[key], key
    assert gi["graphs"] == 583 and gi["paths"] == 1749 and gi["raw_windows"] == 1356246   # SURVEY.md §8 sizes
    assert g.dump_hash() == o.dump_hash()
    assert g.query_params(70, 0.99) == (4, 1, 21)
    assert g.query_params(60, 0.99) == (4, 1, 18)
    assert g.query_params(80, 0.99)[2] == 22


def test_index_save_load_roundtrip(oxa, tmp_path):
    g, _ = oxa
    p = str(tmp_path / "oxa.grootb200")
    g.save(p)
    g2 = api.Index.load(p)
    assert g2.dump_hash() == g.dump_hash()
    with pytest.raises(api.GrootGpuError):
        api.Index.load(str(tmp_path / "missing.grootb200"))
    # a cut or patched file must be rejected by the loader's range checks, not crash a kernel later
    raw = open(p, "rb").read()
    open(tmp_path / "cut.grootb200", "wb").write(raw[: len(raw) // 2])
    with pytest.raises(api.GrootGpuError) as e:
        api.Index.load(str(tmp_path / "cut.grootb200"))
    assert e.value.code == -4
    patched = bytearray(raw)
    patched[8 + 4:8 + 8] = (0).to_bytes(4, "little")          # sketch size 0
    open(tmp_path / "s0.grootb200", "wb").write(bytes(patched))
    with pytest.raises(api.GrootGpuError) as e:
        api.Index.load(str(tmp_path / "s0.grootb200"))
    assert e.value.code == -4


# ------------------------------------------------------------------------------------------------- align
def test_align_golden_oxa(oxa, root):
    g, o = oxa
    golden = np.load(os.path.join(root, "tests", "golden", "oxa_small_align.npz"))
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq"))
    blob, off = pack_reads(seqs[:300])
    g.reset_weights()
    res = g.map_reads(blob, off, 0.99, keep_sketches=True, project=True)
    assert np.array_equal(res.hit_off.astype(np.uint64), golden["hit_off"])
    assert np.array_equal(res.hits, golden["hits"])
    assert np.array_equal(res.records_table(), golden["records"][:, :7].astype(np.int64))
    assert [res.counts[k] for k in ("received", "mapped", "multimapped", "alignments")] == golden["counts"].tolist()
    assert np.array_equal(g.weights()[0], golden["weights"])          # bit-exact f64
    for i in (0, 17, 299):
        assert np.array_equal(res.sketches[i], po.sketch(seqs[i], 51, 30))


def test_align_parity_oxa_full_and_prune(oxa, root):
    g, o = oxa
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq"))
    blob, off = pack_reads(seqs)
    g.reset_weights(); o.reset_weights()
    gr = g.map_reads(blob, off, 0.99, project=True)
    orr = o.map_reads(blob, off, 0.99)
    assert_same_result(gr, orr)
    gw, gt = g.weights(); ow, ot = o.weights()
    assert np.array_equal(gw, ow) and np.array_equal(gt, ot)
    # the reference's own assertion (src/pipeline/3_sketch_test.go:49-58): OXA-90 survives pruning at 10
    kept = g.prune(10.0)
    assert kept[0] == 1
    assert "argannot~~~(Bla)OXA-90~~~EU547443:1-825" in [g.ref(0, p)[0] for p in range(g.info()["paths"])]


def _c1_reads(db_dir, n, L, seed=42):
    return synth.synth_reads(n, L, synth.db_sequences(db_dir), seed=seed)


def test_align_parity_argannot_c1(argannot, db_dirs, root):
    """Config C1: 1k x 100 bp vs arg-annot.90 — the reference's shipped perfect reads and the synthetic mix."""
    g, o = argannot
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "full-argannot-perfect-reads-small.fq.gz"))
    for blob, off in (pack_reads(seqs), _c1_reads(db_dirs["arg-annot.90"], 4000, 100)):
        g.reset_weights(); o.reset_weights()
        gr = g.map_reads(blob, off, 0.99, project=True)
        orr = o.map_reads(blob, off, 0.99, threads=8)
        assert orr.counts["mapped"] > 0.4 * orr.counts["received"]
        assert_same_result(gr, orr)
        assert np.array_equal(g.weights()[0], o.weights()[0])
        assert np.array_equal(g.weights()[1], o.weights()[1])


def test_align_parity_variable_length_and_thresholds(argannot, root):
    g, o = argannot
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "full-argannot-perfect-reads-small-variable-rl.fq.gz"))
    seqs = [s for s in seqs if len(s) >= 31]
    blob, off = pack_reads(seqs)
    for t in (0.99, 0.97, 0.9):
        g.reset_weights(); o.reset_weights()
        gr = g.map_reads(blob, off, t, project=True)
        orr = o.map_reads(blob, off, t, threads=8)
        assert_same_result(gr, orr)
        assert np.array_equal(g.weights()[0], o.weights()[0])


def test_device_projection_is_bit_exact(argannot, db_dirs):
    """The on-device ordered graph weighting (project_*_kernel + stable radix sort) must equal the oracle's sequential
    f64 accumulation bit for bit, across several batches (KmerFreq carries over) and mixed with the host replay."""
    g, o = argannot
    g.reset_weights(); o.reset_weights()
    for seed in (1, 2, 3):
        blob, off = _c1_reads(db_dirs["arg-annot.90"], 20000, 100, seed=seed)
        if seed == 2:
            g.map_reads(blob, off, 0.99, project=True)                  # host replay in the middle
        else:
            g.map_reads(blob, off, 0.99, project_on_device=True)
        o.map_reads(blob, off, 0.99, threads=8)
        gw, gt = g.weights(); ow, ot = o.weights()
        assert np.array_equal(gw, ow) and np.array_equal(gt, ot)
    assert (g.weights()[0] > 0).sum() > 1000


def test_align_chunked_pipeline_matches_single_shot(argannot, db_dirs, root, monkeypatch):
    """grootgpu_align_batch streams the batch through the device in chunks (copy-in / kernels / copy-out overlapped,
    results rebased to batch-wide indices). Forcing many ragged chunks must not change a single output word, nor the
    order-dependent f64 graph weights, with the weighting done on the device or replayed on the host."""
    g, o = argannot
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "full-argannot-perfect-reads-small-variable-rl.fq.gz"))
    blob_v, off_v = pack_reads([s for s in seqs if len(s) >= 31])
    for (blob, off), chunk in ((_c1_reads(db_dirs["arg-annot.90"], 5003, 100), "257"), ((blob_v, off_v), "64"), (_c1_reads(db_dirs["arg-annot.90"], 300, 100), "1")):
        monkeypatch.setenv("GROOTGPU_CHUNK_READS", chunk)
        for on_device in (True, False):
            g.reset_weights(); o.reset_weights()
            gr = g.map_reads(blob, off, 0.99, project=not on_device, project_on_device=on_device, keep_sketches=True)
            orr = o.map_reads(blob, off, 0.99, threads=8, keep_sketches=True)
            assert_same_result(gr, orr)
            assert np.array_equal(gr.sketches, orr.sketches)
            assert np.array_equal(g.weights()[0], o.weights()[0])
            assert np.array_equal(g.weights()[1], o.weights()[1])
    monkeypatch.delenv("GROOTGPU_CHUNK_READS")


def test_align_no_align_mode(argannot, db_dirs):
    g, o = argannot
    blob, off = _c1_reads(db_dirs["arg-annot.90"], 2000, 100, seed=7)
    g.reset_weights(); o.reset_weights()
    gr = g.map_reads(blob, off, 0.99, no_align=True, project=True)
    orr = o.map_reads(blob, off, 0.99, no_align=True)
    assert gr.n_records == 0
    assert_same_result(gr, orr)
    assert np.array_equal(g.weights()[0], o.weights()[0])       # graphminion.go:70-72: every mapping is weighted


def test_align_edge_cases(argannot):
    g, o = argannot
    k = 31
    # a read shorter than k: the reference panics (boss.go:164-166) -> error code, nothing computed
    blob, off = pack_reads([b"ACGT" * 25, b"ACGTACGT"])
    with pytest.raises(api.GrootGpuError) as e:
        g.map_reads(blob, off)
    assert e.value.code == -5
    # empty batch
    with pytest.raises(api.GrootGpuError):
        g.map_reads(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    # reads that cannot seed: longer than the window (eq_min unsatisfiable), exactly k long, all-N, single read
    rng = np.random.default_rng(3)
    long_read = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=180))
    for seqs in ([long_read], [long_read[:k]], [b"N" * 100], [long_read[:100]] * 3):
        blob, off = pack_reads(seqs)
        gr = g.map_reads(blob, off)
        orr = o.map_reads(blob, off)
        assert_same_result(gr, orr)


def test_align_walk_variants(argannot, db_dirs, root):
    """The three walks must agree with the oracle: the packed 2-bit walk (reads of upper-case ACGT), the byte-wise walk
    (a read holding an 'N' near its end still seeds at t = 0.9 and is aligned through the slow queue, forward and
    reverse) and the packed walk at the wider read copies (250 bp reads against -w 250: 16 words per orientation)."""
    g, o = argannot
    seqs = synth.db_sequences(db_dirs["arg-annot.90"])
    rng = np.random.default_rng(5)
    reads = []
    for i in range(600):
        s = seqs[int(rng.integers(len(seqs)))]
        if len(s) < 120:
            continue
        a = int(rng.integers(0, len(s) - 100))
        r = bytearray(bytes(s[a:a + 100]))
        if i % 2:
            r = bytearray(revcomp(bytes(r)))
        if i % 3 == 0:
            r[int(rng.choice([0, 1, 98, 99]))] = ord("N")
        reads.append(bytes(r))
    blob, off = pack_reads(reads)
    for t in (0.9, 0.99):
        g.reset_weights(); o.reset_weights()
        gr = g.map_reads(blob, off, t, project=True)
        orr = o.map_reads(blob, off, t, threads=8)
        assert_same_result(gr, orr)
        assert np.array_equal(g.weights()[0], o.weights()[0])
    assert orr.counts["mapped"] > 100
    msa = [os.path.join(root, "data", "graph", "test-genes.msa")]
    g2 = api.Index.build(msa_files=msa, k=31, S=21, w=250)
    o2 = po.Index(msa_files=msa, k=31, S=21, w=250)
    assert g2.dump_hash() == o2.dump_hash()
    import re
    txt = open(msa[0]).read()
    genes = [re.sub(r"[^ACGTacgt]", "", "".join(b.split("\n")[1:])).upper().encode() for b in txt.split(">")[1:]]
    genes = [x for x in genes if len(x) >= 300]
    reads = []
    for i in range(400):
        s = genes[int(rng.integers(len(genes)))]
        a = int(rng.integers(0, len(s) - 250))
        r = s[a:a + 250]
        reads.append(revcomp(r) if i % 2 else r)
    blob, off = pack_reads(reads)
    gr = g2.map_reads(blob, off, 0.99, project=True)
    orr = o2.map_reads(blob, off, 0.99, threads=8)
    assert orr.counts["mapped"] > 100
    assert_same_result(gr, orr)
    assert np.array_equal(g2.weights()[0], o2.weights()[0])


def test_align_lowercase_read_needing_revcomp_is_an_error(argannot, db_dirs):
    """seqio.go:17-23,122: complementBases is indexed by the base; a byte > 'T' panics. Lower-case reads hash
    like upper-case ones (nthash seed table), so a lower-case copy of a reverse-strand read seeds, fails the
    forward alignment and reaches RevComplement."""
    g, o = argannot
    seqs = synth.db_sequences(db_dirs["arg-annot.90"])
    s = bytes(seqs[10][40:140])
    blob, off = pack_reads([revcomp(s).lower()])
    with pytest.raises(api.GrootGpuError) as e:
        g.map_reads(blob, off)
    assert e.value.code == -6
    with pytest.raises(RuntimeError):
        o.map_reads(blob, off)


def test_align_parity_card_150bp(db_dirs):
    """Config C4 (reduced): 150 bp reads vs card.90 -w 150 (masked graph, larger merge spans)."""
    d = db_dirs["card.90"]
    g = api.Index.build(msa_dir=d, k=31, S=21, w=150)
    o = po.Index(msa_dir=d, k=31, S=21, w=150)
    assert g.info()["masked"] == o.stats()["masked"] >= 1
    assert g.dump_hash() == o.dump_hash()
    blob, off = synth.synth_reads(3000, 150, synth.db_sequences(d), seed=11)
    gr = g.map_reads(blob, off, 0.99, project=True)
    orr = o.map_reads(blob, off, 0.99, threads=8)
    assert_same_result(gr, orr)
    assert np.array_equal(g.weights()[0], o.weights()[0])


# ------------------------------------------------------------------------------------------------- full size
def test_full_size_properties(argannot, db_dirs):
    """BASELINE config sizes through size-independent properties: (1) batch-split invariance — a 2M-read batch
    equals the concatenation of its halves; (2) strand symmetry — reverse-complementing every read leaves the
    sketches' hits unchanged and flips the reverse flag of every aligned pair; (3) idempotence; (4) a sampled
    slice equals the oracle."""
    g, o = argannot
    n, L = 2_000_000, 100
    blob, off = synth.synth_reads(n, L, synth.db_sequences(db_dirs["arg-annot.90"]), seed=42)
    full = g.map_reads(blob, off, 0.99)
    assert full.counts["received"] == n
    assert 0.45 * n < full.counts["mapped"] < 0.55 * n          # 50 % exact substrings minus the W2 quirk
    again = g.map_reads(blob, off, 0.99)
    assert np.array_equal(full.hits, again.hits) and np.array_equal(full.rec_pos, again.rec_pos)
    assert np.array_equal(full.pairs, again.pairs)
    h = n // 2
    a = g.map_reads(blob[:h * L], off[:h + 1], 0.99)
    b = g.map_reads(blob[h * L:], off[h:] - off[h], 0.99)
    assert np.array_equal(np.concatenate([a.hits, b.hits]), full.hits)
    assert np.array_equal(np.concatenate([a.rec_path, b.rec_path]), full.rec_path)
    assert np.array_equal(np.concatenate([a.rec_pos, b.rec_pos]), full.rec_pos)
    assert a.counts["mapped"] + b.counts["mapped"] == full.counts["mapped"]
    # strand symmetry on a 200k slice
    m = 200_000
    comp = synth._COMP
    rc = comp[blob[:m * L].reshape(m, L)][:, ::-1].reshape(-1).copy()
    fwd = g.map_reads(blob[:m * L], off[:m + 1], 0.99)
    rev = g.map_reads(rc, off[:m + 1], 0.99)
    assert np.array_equal(fwd.hit_off, rev.hit_off) and np.array_equal(fwd.hits, rev.hits)
    both = (fwd.pairs["rec_count"] > 0) & (rev.pairs["rec_count"] > 0)
    assert both.sum() > 0.9 * len(fwd.pairs)
    # a palindromic placement can align on both strands; everything else must flip
    flipped = fwd.pairs["reverse"][both] != rev.pairs["reverse"][both]
    assert flipped.mean() > 0.99
    # sampled slice vs oracle
    s0, s1 = 1_234_000, 1_240_000
    sub = g.map_reads(blob[s0 * L:s1 * L], off[s0:s1 + 1] - off[s0], 0.99)
    orr = o.map_reads(blob[s0 * L:s1 * L], off[s0:s1 + 1] - off[s0], 0.99, threads=8)
    assert_same_result(sub, orr)


def _stage_forcing_reads(seqs, n, L=100, seed=9):
    """Reads that cannot be placed by the first stage of AlignRead (alignment.go:35-103): first / last base substituted
    (1-base start / end clip: stages 3 and 4), reads from the tail of a sequence (the never-emitted final window group of
    a path, graph.go:285-338: they seed through merged neighbours or not at all), reads overhanging a sequence end by one
    base on either side, reads from the head of a sequence; every second group reverse-complemented."""
    rng = np.random.default_rng(seed)
    seqs = [bytes(s) for s in seqs if len(s) >= 2 * L]

    def mut(b):
        return b"ACGT"[(b"ACGT".index(bytes([b])) + int(rng.integers(1, 4))) % 4] if bytes([b]) in b"ACGT" else ord("A")
    reads = []
    for i in range(n):
        s = seqs[int(rng.integers(len(seqs)))]
        kind = i % 6
        if kind == 0:
            a = int(rng.integers(0, len(s) - L)); r = bytearray(s[a:a + L]); r[0] = mut(r[0])
        elif kind == 1:
            a = int(rng.integers(0, len(s) - L)); r = bytearray(s[a:a + L]); r[L - 1] = mut(r[L - 1])
        elif kind == 2:
            a = len(s) - L - int(rng.integers(0, 25)); r = bytearray(s[a:a + L])
        elif kind == 3:
            r = bytearray(s[len(s) - (L - 1):] + b"ACGT"[int(rng.integers(4)):][:1])
        elif kind == 4:
            r = bytearray(b"ACGT"[int(rng.integers(4)):][:1] + s[:L - 1])
        else:
            a = int(rng.integers(0, 12)); r = bytearray(s[a:a + L])
        r = bytes(r)
        reads.append(revcomp(r) if (i // 6) % 2 else r)
    return pack_reads(reads)


def test_align_stages_2_3_4_at_scale(argannot, db_dirs):
    """240 k reads built so that the hierarchy has to go past stage 1: every stage must occur, with both clips, on both
    strands, and every hit / pair / record / f64 weight must equal the oracle's."""
    g, o = argannot
    blob, off = _stage_forcing_reads(synth.db_sequences(db_dirs["arg-annot.90"]), 240_000)
    g.reset_weights(); o.reset_weights()
    gr = g.map_reads(blob, off, 0.99, project_on_device=True)
    orr = o.map_reads(blob, off, 0.99, threads=os.cpu_count() or 8)
    assert_same_result_fast(gr, orr)
    assert np.array_equal(g.weights()[0], o.weights()[0]) and np.array_equal(g.weights()[1], o.weights()[1])
    aligned = gr.pairs[gr.pairs["rec_count"] > 0]
    stages = np.bincount(aligned["stage"], minlength=5)
    assert stages[1] > 10_000 and stages[3] > 1_000 and stages[4] > 1_000, stages
    assert stages[2] > 0, stages
    assert (aligned["clip_start"] == (aligned["stage"] == 3)).all() and (aligned["clip_end"] == (aligned["stage"] == 4)).all()
    for st in (3, 4):
        assert set(np.unique(aligned["reverse"][aligned["stage"] == st])) == {0, 1}


def test_c3_full_size_vs_oracle(argannot, db_dirs):
    """BASELINE config C3 at full size (10 M x 100 bp, seed 42, the bench workload): the batch runs (a) in ONE piece with
    the reads resident in HBM — grootgpu_align_batch_device, the call bench.py's `value` times — and (b) through the
    chunked two-lane host path — grootgpu_align_batch, the call `e2e` times. Both must give the same arrays, and those
    must equal the oracle's: hits, pairs and every record of a 1 M-read subsample, hits / pairs /
    counters on all 10 M, and the order-dependent f64 graph weights after all 10 M."""
    import torch
    g, o = argannot
    n, L = 10_000_000, 100
    blob, off = synth.synth_reads(n, L, synth.db_sequences(db_dirs["arg-annot.90"]), seed=42)
    dev = torch.device("cuda", 0)
    d_seq = torch.zeros(n * L + 64, dtype=torch.uint8, device=dev)
    d_seq[: n * L].copy_(torch.from_numpy(blob))
    d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
    torch.cuda.synchronize()
    g.reset_weights()
    one = g.map_reads_device(d_seq.data_ptr(), d_off.data_ptr(), n, L, L, 0.99, copy_back=True, project_on_device=True)
    w_one = g.weights()
    del d_seq, d_off
    g.reset_weights()
    chunked = g.map_reads(blob, off, 0.99, project_on_device=True)
    w_chunked = g.weights()
    assert one.counts == chunked.counts and one.counts["received"] == n
    for k in ("hit_off", "hits", "pairs", "rec_path", "rec_pos"):
        assert np.array_equal(getattr(one, k), getattr(chunked, k)), k
    assert np.array_equal(w_one[0], w_chunked[0]) and np.array_equal(w_one[1], w_chunked[1])
    # the oracle, 1 M reads at a time (reads are independent; the weights carry over from slice to slice)
    o.reset_weights()
    step = 1_000_000
    tot = dict(received=0, mapped=0, multimapped=0, alignments=0)
    threads = os.cpu_count() or 8
    for s0 in range(0, n, step):
        s1 = s0 + step
        orr = o.map_reads(blob[s0 * L:s1 * L], off[s0:s1 + 1] - off[s0], 0.99, threads=threads)
        for k in tot:
            tot[k] += orr.counts[k]
        h0, h1 = int(one.hit_off[s0]), int(one.hit_off[s1])
        assert np.array_equal(one.hit_off[s0:s1 + 1].astype(np.uint64) - np.uint64(h0), orr.hit_off)
        assert np.array_equal(one.hits[h0:h1], orr.hits)
        p0, p1 = np.searchsorted(one.pairs["read"], [s0, s1])
        pr = one.pairs[p0:p1]
        assert len(pr) == len(orr.pairs)
        assert np.array_equal(pr["read"] - np.uint32(s0), orr.pairs[:, 0]) and np.array_equal(pr["graph"], orr.pairs[:, 1])
        assert np.array_equal(pr["n_incremented"], orr.pairs[:, 2]) and np.array_equal(pr["rec_count"], orr.pairs[:, 3])
        if s0 == 3 * step:                              # the 1 M-read subsample: every record
            class _Slice:
                pass
            sl = _Slice()
            sl.counts = orr.counts
            sl.hit_off = one.hit_off[s0:s1 + 1] - np.uint32(h0)
            sl.hits = one.hits[h0:h1]
            sl.n_pairs = len(pr)
            sl.pairs = pr.copy()
            r0 = int(pr["rec_begin"][0]) if len(pr) else 0
            sl.pairs["read"] -= np.uint32(s0)
            sl.pairs["rec_begin"] -= np.uint32(r0)
            sl.n_records = int(pr["rec_count"].sum())
            sl.rec_path = one.rec_path[r0:r0 + sl.n_records]
            sl.rec_pos = one.rec_pos[r0:r0 + sl.n_records]
            assert_same_result_fast(sl, orr)
    assert tot == one.counts
    ow = o.weights()
    assert np.array_equal(w_one[0], ow[0]) and np.array_equal(w_one[1], ow[1])


def test_compact_output_decodes_to_the_full_records(argannot, db_dirs, monkeypatch):
    """params->compact_records: cpairs + 1-byte path ids instead of hit_off / hits / pairs / rec_path / rec_pos. Decoded
    with the nodes' path tables (what a BAM writer does) it must give exactly the records of the full output and of the
    oracle — single shot on the device, chunked through the host path (ragged chunks), stage 2/3/4 reads included."""
    import torch
    g, o = argannot
    seqs = synth.db_sequences(db_dirs["arg-annot.90"])
    b1, o1 = _c1_reads(db_dirs["arg-annot.90"], 30_011, 100, seed=5)
    b2, o2 = _stage_forcing_reads(seqs, 24_000, seed=3)
    for blob, off in ((b1, o1), (b2, o2)):
        full = g.map_reads(blob, off, 0.99)
        orr = o.map_reads(blob, off, 0.99, threads=8)
        assert np.array_equal(full.records_table(), oracle_records_table(orr))
        for chunk in (None, "997"):
            if chunk:
                monkeypatch.setenv("GROOTGPU_CHUNK_READS", chunk)
            c = g.map_reads(blob, off, 0.99, compact=True, project_on_device=True)
            if chunk:
                monkeypatch.delenv("GROOTGPU_CHUNK_READS")
            assert c.compact and c.rec_path_c.dtype == np.uint8 and c.counts == full.counts
            assert c.n_pairs == full.n_pairs and c.n_records == full.n_records
            assert np.array_equal(c.cpairs["read"], full.pairs["read"]) and np.array_equal(c.cpairs["rec_count"], full.pairs["rec_count"])
            assert np.array_equal(c.decode_compact(g), full.records_table())
        n = len(off) - 1
        dev = torch.device("cuda", 0)
        d_seq = torch.zeros(len(blob) + 64, dtype=torch.uint8, device=dev)
        d_seq[: len(blob)].copy_(torch.from_numpy(blob))
        d_off = torch.from_numpy(off.astype(np.uint32).view(np.int32)).to(dev)
        torch.cuda.synchronize()
        c = g.map_reads_device(d_seq.data_ptr(), d_off.data_ptr(), n, 100, 100, 0.99, copy_back=True, compact=True)
        assert np.array_equal(c.decode_compact(g), full.records_table())
    # the weights are the same whichever output format the batches used
    g.reset_weights(); o.reset_weights()
    g.map_reads(b1, o1, 0.99, compact=True, project_on_device=True)
    o.map_reads(b1, o1, 0.99, threads=8)
    assert np.array_equal(g.weights()[0], o.weights()[0]) and np.array_equal(g.weights()[1], o.weights()[1])


def test_any_sketch_size_and_maxk(root, db_dirs):
    """`groot index` accepts any -s / -y (cmd/index.go:48-49). Combinations outside the compiled register-resident kernels
    (maxK == 4, eight sketch sizes) run the run-time-S kernels: index dump, sketches, hits, records and weights must equal
    the oracle's — maxK 2 with S = 25 (12 bands of 2), maxK 6 (band prefixes longer than a table slot's four words),
    maxK 1, and a compiled S with maxK 3."""
    msa = [os.path.join(root, "data", "graph", "test-genes.msa")]
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq"))
    blob, off = pack_reads(seqs[:600])
    rng = np.random.default_rng(1)
    rs = [bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=int(rng.integers(31, 200)))) for _ in range(50)]
    rblob, roff = pack_reads(rs)
    for S, max_k, k, t in ((25, 2, 31, 0.99), (25, 2, 31, 0.9), (30, 6, 31, 0.95), (13, 1, 21, 0.9), (21, 3, 31, 0.97), (40, 4, 31, 0.99)):
        g = api.Index.build(msa_files=msa, k=k, S=S, w=100, max_k=max_k)
        o = po.Index(msa_files=msa, k=k, S=S, w=100, max_k=max_k)
        assert g.dump_hash() == o.dump_hash(), (S, max_k)
        assert np.array_equal(api.sketch_batch(rblob, roff, k, S), np.stack([po.sketch(r, k, S) for r in rs]))
        gr = g.map_reads(blob, off, t, project=True, keep_sketches=True)
        orr = o.map_reads(blob, off, t, threads=8, keep_sketches=True)
        assert orr.counts["mapped"] > 50, (S, max_k, orr.counts)
        assert_same_result(gr, orr)
        assert np.array_equal(gr.sketches, orr.sketches)
        assert np.array_equal(g.weights()[0], o.weights()[0])
        gr2 = g.map_reads(blob, off, t)           # without keep_sketches (the two-pass switch must not matter here)
        assert np.array_equal(gr2.hits, orr.hits) and np.array_equal(gr2.records_table(), oracle_records_table(orr))
        g.close()
    # arg-annot.90 at -y 2 -s 25: reads with many hits (the refill kernel), several graphs per read
    d = db_dirs["arg-annot.90"]
    g = api.Index.build(msa_dir=d, k=31, S=25, w=100, max_k=2)
    o = po.Index(msa_dir=d, k=31, S=25, w=100, max_k=2)
    assert g.dump_hash() == o.dump_hash()
    blob, off = _c1_reads(d, 3000, 100, seed=8)
    for t in (0.99, 0.9):
        g.reset_weights(); o.reset_weights()
        gr = g.map_reads(blob, off, t, project_on_device=True)
        orr = o.map_reads(blob, off, t, threads=8)
        assert_same_result(gr, orr)
        assert np.array_equal(g.weights()[0], o.weights()[0])


def test_index_from_reference_gob_files(oxa, root, tmp_path):
    """grootgpu_index_load_gob: groot.gg + groot.lshe as `groot index` writes them (here: written by tests/gob_writer.py
    from the oracle's index, maps in random order) give the same index and the same alignments as the natively built one;
    the driver picks them up when the index directory holds no groot.grootb200."""
    import subprocess
    from tests import gob_writer as gw
    g, o = oxa
    dump = str(tmp_path / "o.txt")
    o.dump_file(dump)
    d = tmp_path / "idx"; d.mkdir()
    gw.write_reference_index(dump, str(d / "groot.gg"), str(d / "groot.lshe"), seed=7)
    gg = api.Index.load_gob(str(d / "groot.gg"), str(d / "groot.lshe"))
    assert gg.dump_hash() == o.dump_hash()          # (the shared GPU index `g` has been pruned by an earlier test: its path lengths differ)
    names, seqs, quals = load_fastq(os.path.join(root, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq"))
    blob, off = pack_reads(seqs[:500])
    o.reset_weights()
    orr = o.map_reads(blob, off, 0.99)
    gr = gg.map_reads(blob, off, 0.99, project_on_device=True)
    assert_same_result(gr, orr)
    assert np.array_equal(gg.weights()[0], o.weights()[0])
    cli = os.path.join(os.path.dirname(api.LIB_PATH), "groot-b200")
    fq = os.path.join(root, "data", "reads", "test-reads-OXA90-OXA106-100bp-with-errors.fastq")
    r = subprocess.run([cli, "align", "-i", str(d), "-f", fq, "-g", str(tmp_path / "graphs"), "--bamOut", str(tmp_path / "out.bam")], stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    assert b"number of reads received from input: 2062" in r.stderr


def test_fixed_read_length_path(argannot, db_dirs, monkeypatch):
    """params->fixed_read_len: reads of one length, back to back, seq_off == NULL — the offsets are generated on the device
    instead of travelling there (8 of 108 bytes per read). Same result as with offsets, in ragged chunks too."""
    g, o = argannot
    blob, off = _c1_reads(db_dirs["arg-annot.90"], 20_003, 100, seed=12)
    want = g.map_reads(blob, off, 0.99)
    for chunk in (None, "3001"):
        if chunk:
            monkeypatch.setenv("GROOTGPU_CHUNK_READS", chunk)
        got = g.map_reads(blob, off, 0.99, fixed_read_len=100)
        assert got.counts == want.counts
        for k in ("hit_off", "hits", "pairs", "rec_path", "rec_pos"):
            assert np.array_equal(getattr(got, k), getattr(want, k)), k
        c = g.map_reads(blob, off, 0.99, fixed_read_len=100, compact=True, project_on_device=True)
        assert np.array_equal(c.decode_compact(g), want.records_table())
        if chunk:
            monkeypatch.delenv("GROOTGPU_CHUNK_READS")
    with pytest.raises(api.GrootGpuError) as e:          # shorter than k: the reference panics (boss.go:164-166)
        g.map_reads(blob, off, 0.99, fixed_read_len=20)
    assert e.value.code == -5
