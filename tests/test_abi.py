"""CPU: the C-ABI library loads and exports every symbol include/asurf.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from alphasurf_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "asurf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asurf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(capi.EXPORTS)


def test_library_exports_every_symbol():
    lib = capi.lib()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.asurf_abi_version() == 3
    assert isinstance(lib.asurf_last_error(), bytes)


def test_accel_words_is_host_only():
    lib = capi.lib()
    n = lib.asurf_accel_words((ctypes.c_int32 * 3)(512, 512, 512))
    # three pyramid levels + the list of non-empty 16^3 blocks (count word + uint32 ids, 2 per word) + stored-vertex count
    # + the list of the 16^3 vertex blocks with a stored vertex (same form) + X*Y + 1 per-column prefix counts (uint32)
    assert n == (128 ** 3 + 32 ** 3 + 8 ** 3 + 1 + (32 ** 3 + 1) // 2 + 1) + (1 + (32 ** 3 + 1) // 2) + (512 * 512 + 2) // 2


def test_struct_layouts_match_header():
    # sizes follow from the field lists in include/asurf.h (natural alignment)
    assert ctypes.sizeof(capi.OptT) == 17 * 4
    assert ctypes.sizeof(capi.RaysT) == 24
    assert ctypes.sizeof(capi.GradsT) == 56
    assert ctypes.sizeof(capi.GridT) == 8 + 12 + 4 + 8 * 4 + 4 * 3 + 4 + 8 + 12 + 12 + 4 + 4 + 8 + 8 + 8 + 8 + 4 + 4
    assert ctypes.sizeof(capi.FusedT) == 18 * 4 + 8


def test_header_is_plain_c(tmp_path):
    """include/asurf.h is the drop-in boundary: it must compile as C99 (no C++ or torch types) and link against the
    library from a C translation unit."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "asurf.h"\n#include <stdio.h>\n'
                   'int main(void) { int32_t sz[3] = {512, 512, 512};\n'
                   '  printf("%d %lld\\n", (int)asurf_abi_version(), (long long)asurf_accel_words(sz)); return 0; }\n')
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                           "-I", os.path.join(root, "include"), str(src)])
    lib = os.path.join(root, "alphasurf_b200", "csrc", "libasurf.so")
    exe = tmp_path / "t"
    subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(root, "include"), str(src), lib, "-o", str(exe),
                           "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) >= 3 and int(out[1]) == capi.lib().asurf_accel_words((ctypes.c_int32 * 3)(512, 512, 512))


def test_ctypes_call_sites_pass_the_declared_number_of_arguments():
    """ctypes does not check arity: every `...asurf_xxx(...)` call in the package, the tests and bench.py must pass exactly
    as many arguments as include/asurf.h declares for that entry (the compiled shim gets this check from the C++ compiler)."""
    import ast
    src = open(os.path.join(ROOT, "include", "asurf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = {}
    for m in re.finditer(r"\b(asurf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = m.group(2).strip()
        declared[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    assert len(declared) > 40
    files = [os.path.join(ROOT, "bench.py")]
    for d in ("alphasurf_b200", "tests", "scratch"):
        files += [os.path.join(ROOT, d, f) for f in os.listdir(os.path.join(ROOT, d)) if f.endswith(".py")]
    checked, bad = 0, []
    for path in files:
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in declared:
                starred = sum(isinstance(a, ast.Starred) for a in node.args)
                if starred:
                    continue          # argument packs (camera tuples) are expanded at run time
                checked += 1
                if len(node.args) != declared[node.func.attr]:
                    bad.append("%s:%d %s passes %d of %d" % (os.path.basename(path), node.lineno, node.func.attr, len(node.args),
                                                            declared[node.func.attr]))
    assert checked > 50, checked
    assert not bad, bad
