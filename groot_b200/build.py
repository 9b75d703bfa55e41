"""Builds libgrootgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m groot_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgrootgpu.so")
CLI = os.path.join(HERE, "groot-b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CU = ["capi.cu"]
CPP = ["host/graph_build.cpp", "host/index_io.cpp", "host/lshe_params.cpp", "host/replay.cpp", "host/prefix_table.cpp", "host/gob_reader.cpp", "host/gob_writer.cpp"]


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".cpp", ".h"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "grootgpu.h"))
    return out


def build(force=False, verbose=False, out=OUT, defines=()):
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(s) for s in sources()):
        return out
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
           "-Xcompiler", "-fPIC,-O2,-Wall,-pthread", "--expt-extended-lambda", "-o", out] + ["-D" + d for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, f) for f in CU + CPP]
    subprocess.check_call(cmd)
    if out == OUT:
        build_cli()
    return out


def build_cli():
    """groot-b200: the C++ host driver (pipeline mirror + CLI) linked against libgrootgpu.so."""
    host = os.path.join(CSRC, "host")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", CLI, os.path.join(host, "main.cpp"), os.path.join(host, "pipeline.cpp"),
           "-L" + HERE, "-lgrootgpu", "-lz", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return CLI


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else OUT, defines=defs))
