"""GPU: the compiled svox2.csrc module (alphasurf_b200/csrc/host/svox2_shim.cpp, pybind11 + torch C++ over the C ABI) gives
the same results as the ctypes mirror most of the kernel-level tests run through -- both are argument plumbing over one
libasurf.so.  The comparison (tests/shim_gpu_check.py) runs in a process of its own.  The compiled module is what bench.py
times and what tests/test_dropin_gpu.py puts under the reference's own Python; its CPU-side behaviour (import, spec classes,
checks, call-site arity against the reference's Python) is covered by tests/test_reference_binding.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_equals_ctypes_mirror_on_gpu():
    r = subprocess.run([sys.executable, "-m", "tests.shim_gpu_check"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
