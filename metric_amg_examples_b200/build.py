"""Build libmamg.so (host setup in C++ + sm_100a CUDA kernels + the C-ABI) in-tree.

Usage: python metric_amg_examples_b200/build.py [--force] [-v]
The shared object lands next to this file so that it travels with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# experiment variants: MAMG_DEFS="-DMAMG_SW_MINB24=4 ..." MAMG_VARIANT=name -> libmamg_<name>.so (selected with MAMG_LIB)
VARIANT = os.environ.get("MAMG_VARIANT", "")
DEFS = os.environ.get("MAMG_DEFS", "").split()
OUT = os.path.join(HERE, f"libmamg_{VARIANT}.so" if VARIANT else "libmamg.so")
OBJ = os.path.join(HERE, "build_" + VARIANT if VARIANT else "build")

HOST_SRC = ["host/assemble.cpp", "host/setup.cpp", "host/capi.cpp"]
CUDA_SRC = ["cuda/device.cu"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CXXFLAGS = ["-O3", "-std=c++17", "-fPIC", "-fopenmp", "-Wall", "-Wno-unknown-pragmas"]
NVFLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false", "-Xcompiler", "-fPIC,-fopenmp",
           "-Xptxas", "-v"]


def _sources():
    files = []
    for root, _, names in os.walk(CSRC):
        for n in names:
            files.append(os.path.join(root, n))
    files.append(os.path.join(HERE, "..", "include", "mamg.h"))
    files.append(os.path.abspath(__file__))
    return sorted(files)


def _digest():
    h = hashlib.sha256()
    for f in _sources():
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError("build failed:\n" + "\n".join(log[-1:]))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    log = []
    jobs = []
    objs = []
    for s in HOST_SRC:
        o = os.path.join(OBJ, s.replace("/", "_") + ".o")
        objs.append(o)
        jobs.append(["g++"] + CXXFLAGS + ["-c", os.path.join(CSRC, s), "-o", o])
    for s in CUDA_SRC:
        o = os.path.join(OBJ, s.replace("/", "_") + ".o")
        objs.append(o)
        flags = [f for f in NVFLAGS if f != "--use_fast_math=false"]
        jobs.append([NVCC] + ARCH + flags + DEFS + ["-c", os.path.join(CSRC, s), "-o", o])
    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(lambda c: _run(c, log), jobs))
    _run([NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-Xcompiler", "-fopenmp", "-lgomp", "-lnccl", "-ldl"], log)
    with open(os.path.join(OBJ, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(stamp, "w") as fh:
        fh.write(dig)
    if verbose:
        print("\n".join(log))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
