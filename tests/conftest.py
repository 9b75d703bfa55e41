import os
import sys
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_build():
    """libgrootgpu.so, the groot-b200 driver and the oracle are built in-tree by __graft_entry__.build(). The tests only
    make sure they EXIST (building what is missing; nvcc cross-compiles without a GPU) — whether a copied tree's
    binaries are stale cannot be told from file times, and rebuilding is build()'s job."""
    from groot_b200 import build as gb
    if not os.path.exists(gb.OUT):
        gb.build(force=True)
    if not os.path.exists(gb.CLI):
        gb.build_cli()
    from oracle import pyoracle
    pyoracle.build()


@pytest.fixture(scope="session")
def root():
    return ROOT


@pytest.fixture(scope="session")
def db_dirs(tmp_path_factory):
    """Unpack the clustered ARG databases shipped under data/db (copied from the reference's
    db/clustered-ARG-databases/1.1/) and return {name: msa_dir}."""
    base = tmp_path_factory.mktemp("db")
    out = {}
    for name in ("arg-annot.90", "card.90"):
        tar = os.path.join(ROOT, "data", "db", name + ".tar")
        with tarfile.open(tar) as t:
            t.extractall(base)
        out[name] = str(base / name)
    return out
