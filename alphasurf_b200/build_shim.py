"""Build the compiled ``svox2.csrc`` shim (csrc/host/svox2_shim.cpp: pybind11 + torch C++ headers over the C ABI).

    python -m alphasurf_b200.build_shim                         # in-tree module alphasurf_b200/csrc/svox2_csrc_shim*.so
    python -m alphasurf_b200.build_shim --name csrc --out DIR   # the drop-in: DIR = the reference's svox2/ package directory

Plain ``g++``: the shim holds no kernels.  It links against libasurf.so (built first if needed) with an rpath to it.
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "host", "svox2_shim.cpp")


def shim_path(name="svox2_csrc_shim", out_dir=None):
    return os.path.join(out_dir or os.path.join(HERE, "csrc"), name + sysconfig.get_config_var("EXT_SUFFIX"))


def build_shim(name="svox2_csrc_shim", out_dir=None, force=False) -> str:
    from . import build
    lib = build.build_library()
    out = shim_path(name, out_dir)
    deps = [SRC, os.path.join(ROOT, "include", "asurf.h")]
    if not force and os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(p) for p in deps):
        return out
    import torch
    from torch.utils import cpp_extension
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found: cannot build the svox2.csrc shim")
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-DTORCH_EXTENSION_NAME=" + name,
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + p for p in cpp_extension.include_paths()] + ["-I" + sysconfig.get_paths()["include"], "-I" + cuda_inc,
                                                                "-I" + os.path.join(ROOT, "include")]
    cmd += [SRC, "-o", out, lib, "-L" + tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-ldl",
            "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + tlib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the svox2.csrc shim failed:\n%s\n%s" % (r.stdout[-4000:], r.stderr[-4000:]))
    return out


def load(name="svox2_csrc_shim"):
    """import the in-tree shim (building it first if needed)"""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    path = build_shim(name)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="svox2_csrc_shim")
    ap.add_argument("--out", default=None)
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    print(build_shim(a.name, a.out, a.force))
