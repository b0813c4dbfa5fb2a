"""Build the C-ABI shared library ``alphasurf_b200/csrc/libasurf.so`` from ``csrc/src/*.cu`` for sm_100a.

In-tree build with plain ``nvcc`` (no torch headers: the C ABI carries no torch types).  The ``.so`` is git-ignored
but travels to the GPU box with the snapshot.  ``python -m alphasurf_b200.build`` rebuilds it.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "src")
INC = os.path.join(os.path.dirname(HERE), "include")
OBJ = os.path.join(HERE, "csrc", "obj")
LIB = os.path.join(HERE, "csrc", "libasurf.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + INC, "-I" + SRC]


def _nvcc():
    n = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(n):
        raise RuntimeError("nvcc not found: cannot build libasurf.so")
    return n


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cuh", ".h", ".inl"))]
    hdrs += [os.path.join(INC, f) for f in os.listdir(INC)]
    return hdrs


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in sources() + _deps())


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = max(os.path.getmtime(p) for p in _deps())

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
