"""Builds parallel_amg_b200/libpamg.so (host setup + sm_100a kernels + C ABI) in-tree.

nvcc cross-compiles for sm_100a without a GPU; the .so travels to the GPU box with the repo
snapshot.  Usage: python -m parallel_amg_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpamg.so")
STAMP = os.path.join(HERE, ".libpamg.stamp")

CXX_SOURCES = ["host_setup.cpp", "hierarchy_io.cpp", "capi.cpp"]
CU_SOURCES = ["engine.cu", "setup_gpu.cu"]
DEPS = ["host.hpp", "engine.hpp", "formats.hpp", "kernels.cuh", os.path.join("..", "..", "include", "pamg.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off", "-Xptxas", "-v"]
NVCC_FLAGS += os.environ.get("PAMG_EXTRA_NVCC_FLAGS", "").split()  # experiments only (A/B builds under _ab/)
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wall"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for f in CXX_SOURCES + CU_SOURCES + DEPS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    objs = []
    nvcc = _nvcc()
    log = []
    for f in CXX_SOURCES:
        o = os.path.join(bdir, f + ".o")
        cmd = ["g++"] + CXX_FLAGS + ["-I", os.path.join(nvcc.rsplit("/bin/", 1)[0], "include"), "-c",
                                     os.path.join(CSRC, f), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode:
            raise RuntimeError("g++ failed:\n" + r.stderr)
        objs.append(o)
    for f in CU_SOURCES:
        o = os.path.join(bdir, f + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, f), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed:\n" + r.stderr[-8000:])
        objs.append(o)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("link failed:\n" + r.stderr)
    with open(os.path.join(bdir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
