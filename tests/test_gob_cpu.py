"""The library's reader for the reference's own index files (groot.gg / groot.lshe, Go encoding/gob:
src/pipeline/runtime.go:64-91, src/lshe/lshe.go:72-146). No Go toolchain exists here, so the streams come from
tests/gob_writer.py, a restatement of the gob ENCODER that is itself pinned to the byte sequence the encoding/gob
documentation gives for Point{22, 33}. CPU only: grootgpu_gob_dump decodes, validates and dumps without a device."""
import glob
import os

import pytest

from groot_b200 import api
from oracle import pyoracle as po
from tests import gob_writer as gw


def test_gob_writer_reproduces_the_documented_stream():
    want = bytes.fromhex("1fff8103010105506f696e7401ff82000102010158010400010159010400000007ff82012c014200")
    got = gw.Encoder().encode(gw.Struct("Point", [("X", gw.INT), ("Y", gw.INT)]), {"X": 22, "Y": 33})
    assert got == want
    assert gw.uvarint(7) == b"\x07" and gw.uvarint(256) == b"\xfe\x01\x00" and gw.varint(-129) == b"\xfe\x01\x01"     # doc: 256 -> FE 01 00, -129 -> FE 01 01
    assert gw.gfloat(17.0) == b"\xfe\x31\x40"                                                                          # doc: float64 17 -> FE 31 40


def test_gob_index_roundtrip(root, db_dirs, tmp_path):
    """An index written the way `groot index` writes it (maps in random order) must come back as the same index: graphs
    in Store order, nodes in SortedNodes order, out-edges in the stored order, windows sorted (graph, Node, OffSet, the
    counter of the lookup string)."""
    cases = [([os.path.join(root, "data", "graph", "test-genes.msa")], dict(k=51, S=30, w=100)),
             (sorted(glob.glob(os.path.join(db_dirs["arg-annot.90"], "cluster*.msa")))[:25], dict(k=31, S=21, w=100)),
             (sorted(glob.glob(os.path.join(db_dirs["card.90"], "cluster*.msa")))[:12], dict(k=31, S=20, w=150, max_k=2))]
    for i, (files, prm) in enumerate(cases):
        o = po.Index(msa_files=files, **prm)
        dump = str(tmp_path / ("o%d.txt" % i))
        o.dump_file(dump)
        gg, lshe = str(tmp_path / "groot.gg"), str(tmp_path / "groot.lshe")
        for seed in (1, 2):
            gw.write_reference_index(dump, gg, lshe, seed=seed)
            out = str(tmp_path / "g.txt")
            assert api.gob_dump(gg, lshe, out) == o.dump_hash()
            assert open(out).read() == open(dump).read()


def test_gob_errors(root, tmp_path):
    o = po.Index(msa_files=[os.path.join(root, "data", "graph", "test-genes.msa")], k=51, S=30, w=100)
    dump = str(tmp_path / "o.txt")
    o.dump_file(dump)
    gg, lshe = str(tmp_path / "groot.gg"), str(tmp_path / "groot.lshe")
    gw.write_reference_index(dump, gg, lshe)
    with pytest.raises(api.GrootGpuError) as e:
        api.gob_dump(str(tmp_path / "missing.gg"), lshe)
    assert e.value.code == -3
    raw = open(gg, "rb").read()
    open(tmp_path / "cut.gg", "wb").write(raw[: len(raw) // 2])
    with pytest.raises(api.GrootGpuError) as e:
        api.gob_dump(str(tmp_path / "cut.gg"), lshe)
    assert e.value.code == -4
    open(tmp_path / "empty.gg", "wb").write(b"")
    with pytest.raises(api.GrootGpuError) as e:                     # runtime.go:86-88 "groot graph store appears empty"
        api.gob_dump(str(tmp_path / "empty.gg"), lshe)
    assert e.value.code == -4
    with pytest.raises(api.GrootGpuError) as e:                     # the two files swapped: wrong top-level type
        api.gob_dump(lshe, gg)
    assert e.value.code == -4
    # a groot.lshe of another index (different sketch size) must be refused
    o2 = po.Index(msa_files=[os.path.join(root, "data", "graph", "test-genes.msa")], k=51, S=21, w=100)
    o2.dump_file(str(tmp_path / "o2.txt"))
    gw.write_reference_index(str(tmp_path / "o2.txt"), str(tmp_path / "g2.gg"), str(tmp_path / "g2.lshe"))
    with pytest.raises(api.GrootGpuError) as e:
        api.gob_dump(gg, str(tmp_path / "g2.lshe"))
    assert e.value.code == -4


def test_damaged_index_files_are_refused_not_crashed_on(root, tmp_path):
    """Fuzz (tests/cpp/gob_fuzz.cpp, ASan + UBSan): a groot.gg / groot.lshe pair — and the library's own flat index file —
    with flipped, inserted, duplicated, zeroed or cut-off bytes either loads and validates or raises; it never crashes
    (a damaged type table can make a gob type contain itself: the decoder bounds its depth)."""
    import subprocess
    host = os.path.join(root, "groot_b200", "csrc", "host")
    exe = str(tmp_path / "gob_fuzz")
    subprocess.check_call(["g++", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17", "-o", exe,
                           os.path.join(root, "tests", "cpp", "gob_fuzz.cpp")] + [os.path.join(host, f) for f in ("gob_reader.cpp", "index_io.cpp")])
    o = po.Index(msa_files=[os.path.join(root, "data", "graph", "test-genes.msa")], k=51, S=30, w=100)
    o.dump_file(str(tmp_path / "o.txt"))
    gg, lshe = str(tmp_path / "groot.gg"), str(tmp_path / "groot.lshe")
    gw.write_reference_index(str(tmp_path / "o.txt"), gg, lshe)
    env = dict(os.environ, ASAN_OPTIONS="allocator_may_return_null=1")
    for args in ((gg, lshe, str(tmp_path), "6", "30"), (gg, lshe, str(tmp_path), "3", "30", "flat")):
        r = subprocess.run([exe] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
        assert r.returncode == 0 and r.stdout.decode().startswith("ok "), r.stderr.decode()[-2000:]
        assert int(r.stdout.decode().split()[2]) > 15                               # most damaged files are refused


def test_gob_type_that_contains_itself(root, tmp_path):
    """What the fuzz found: a type table in which a struct has a field of its own type lets every input byte open another
    level — the decoder must refuse at a fixed depth instead of running out of stack."""
    u, i = gw.uvarint, gw.varint
    tdef = i(-65) + u(3) + u(1) + gw.Encoder._common("T", 65) + u(1) + u(1) + (u(1) + u(1) + b"X" + u(1) + i(65) + b"\x00") + b"\x00" + b"\x00"
    val = i(65) + u(1) * 300000
    stream = u(len(tdef)) + tdef + u(len(val)) + val
    open(tmp_path / "self.gg", "wb").write(stream)
    o = po.Index(msa_files=[os.path.join(root, "data", "graph", "test-genes.msa")], k=51, S=30, w=100)
    o.dump_file(str(tmp_path / "o.txt"))
    gw.write_reference_index(str(tmp_path / "o.txt"), str(tmp_path / "groot.gg"), str(tmp_path / "groot.lshe"))
    with pytest.raises(api.GrootGpuError) as e:
        api.gob_dump(str(tmp_path / "self.gg"), str(tmp_path / "groot.lshe"))
    assert e.value.code == -4 and "nested too deeply" in str(e.value)


@pytest.mark.parametrize("params", [dict(k=51, S=30, w=100), dict(k=7, S=21, w=20)])
def test_gob_writer_cpp_equals_the_python_encoder(root, tmp_path, params):
    """host/gob_writer.cpp against the independent Python encoder, byte for byte: an index goes Python gob (shuffled maps, as Go
    writes them) -> flat file -> C++ gob, which must equal what the Python encoder writes with its maps in key order; and
    the C++ files load back to the same canonical dump (oracle hash)."""
    o = po.Index(msa_files=[os.path.join(root, "data", "graph", "test-genes.msa")], **params)
    dump = str(tmp_path / "o.txt")
    o.dump_file(dump)
    import numpy as np
    n_nodes = sum(len(g["nodes"]) for g in gw.parse_dump(dump)[1])
    kf = np.random.default_rng(0).random(n_nodes) * 3                       # graph weights travel too (KmerFreq, zero ones omitted)
    kf[::3] = 0
    gw.write_reference_index(dump, str(tmp_path / "py.gg"), str(tmp_path / "py.lshe"), seed=5, kmer_freq=kf)
    api.gob_to_flat(str(tmp_path / "py.gg"), str(tmp_path / "py.lshe"), str(tmp_path / "x.grootb200"))
    api.flat_to_gob(str(tmp_path / "x.grootb200"), str(tmp_path / "cpp.gg"), str(tmp_path / "cpp.lshe"))
    gw.write_reference_index(dump, str(tmp_path / "want.gg"), str(tmp_path / "want.lshe"), seed=None, kmer_freq=kf)
    assert open(tmp_path / "cpp.gg", "rb").read() == open(tmp_path / "want.gg", "rb").read()
    assert open(tmp_path / "cpp.lshe", "rb").read() == open(tmp_path / "want.lshe", "rb").read()
    assert api.gob_dump(str(tmp_path / "cpp.gg"), str(tmp_path / "cpp.lshe")) == o.dump_hash()
    with pytest.raises(api.GrootGpuError) as e:
        api.flat_to_gob(str(tmp_path / "missing.grootb200"), str(tmp_path / "a.gg"), str(tmp_path / "a.lshe"))
    assert e.value.code == -3
    with pytest.raises(api.GrootGpuError) as e:
        api.flat_to_gob(str(tmp_path / "x.grootb200"), str(tmp_path / "no_such_dir" / "a.gg"), str(tmp_path / "a.lshe"))
    assert e.value.code == -3
