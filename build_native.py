"""Build libollamamq_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python build_native.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(ROOT, "ollamamq_b200")
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libollamamq_b200.so")
INCLUDE = os.path.join(ROOT, "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
                  "-I", INCLUDE, "-I", CSRC]
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-I", INCLUDE, "-I", CSRC, "-I", "/usr/local/cuda/include"]


def _sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(".cu") or f.endswith(".cpp"):
            out.append(os.path.join(CSRC, f))
    return out


def _headers_mtime():
    m = 0.0
    for root in (CSRC, INCLUDE):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh", ".hpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, force, hdr_m):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
            and os.path.getmtime(obj) > hdr_m):
        return obj, False
    if src.endswith(".cu"):
        cmd = [NVCC] + CUFLAGS + ["-c", src, "-o", obj]
    else:
        cmd = ["g++"] + CXXFLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return obj, True


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    hdr_m = _headers_mtime()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_m), srcs))
    objs = [o for o, _ in res]
    rebuilt = any(ch for _, ch in res)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xlinker", "-z,defs", "-lpthread", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            print("built", LIB)
    elif verbose:
        print("up to date", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
