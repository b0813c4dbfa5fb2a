"""CPU: the reference's own Python package (/root/reference/svox2, unmodified, imported from where it lies) accepts
alphasurf_b200.svox2_csrc as its `svox2.csrc` extension, builds its spec objects through it, and every call site of the
extension in svox2/svox2.py passes a number of positional arguments our function of that name accepts.

Reads /root/reference, which exists in the development container only: skipped elsewhere (never part of the GPU tier)."""
import ast
import inspect
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "svox2")), reason="reference checkout not present")

BACKENDS = ("cuvol", "surf_trav")          # the backends on the B200 hot path (the others raise NotImplementedError)


@pytest.fixture(scope="module", params=["ctypes mirror", "compiled shim"])
def ref_svox2(request):
    if request.param == "ctypes mirror":
        import alphasurf_b200.svox2_csrc as ours
    else:
        from alphasurf_b200 import build_shim
        ours = build_shim.load()          # csrc/host/svox2_shim.cpp: pybind11 + torch C++ over the same C ABI
    names = ("mcubes", "svox2", "svox2.csrc", "svox2.svox2", "svox2.utils", "svox2.defs", "svox2.version")
    saved = {k: sys.modules.get(k) for k in names}
    for k in names[1:]:          # a copy imported earlier in the session (e.g. without an extension, by the L0 harness) must
        sys.modules.pop(k, None)  # not be reused: svox2/utils.py looks for svox2.csrc once, at import
    sys.modules.setdefault("mcubes", types.ModuleType("mcubes"))     # module-level import of an absent package (svox2.py:16)
    sys.modules["svox2.csrc"] = ours                                 # what INTEGRATION.md's svox2/csrc.py amounts to
    sys.path.insert(0, REF)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import svox2
        yield svox2, ours
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_package_accepts_the_module(ref_svox2):
    svox2, ours = ref_svox2
    from svox2 import utils
    assert utils._get_c_extension() is ours          # svox2/utils.py:32-46 found `sample_grid` and kept the module
    grid = svox2.SparseGrid(reso=16, radius=1.0, center=0.0, basis_dim=4, use_z_order=True, device="cpu")
    spec = grid._to_cpp()                            # svox2.py:6234-6272 fills our SparseGridSpec by attribute
    assert isinstance(spec, ours.SparseGridSpec)
    assert spec.links is grid.links and spec.basis_dim == 4
    assert spec._offset.device.type == "cpu" and spec._scaling.shape == (3,)
    opt = grid.opt._to_cpp()
    assert isinstance(opt, ours.RenderOptions)
    for k in ("step_size", "sigma_thresh", "stop_thresh", "near_clip", "alpha_activation_type", "only_outward_intersect",
              "truncated_vol_render", "trunc_vol_weight_min"):
        assert hasattr(opt, k), k
    rays = svox2.Rays(torch.zeros(4, 3), torch.ones(4, 3))._to_cpp()
    assert isinstance(rays, ours.RaysSpec) and rays.origins.shape == (4, 3)
    cam = svox2.Camera(torch.eye(4), 100.0, 100.0, 32.0, 24.0, 64, 48)._to_cpp()
    assert isinstance(cam, ours.CameraSpec) and cam.width == 64 and cam.height == 48


def _accepts(fn, n_args):
    try:
        sig = inspect.signature(fn)
    except ValueError:
        # a pybind11 function: its docstring starts with "name(arg0: T, arg1: T, ...) -> R" (or "(*args, **kwargs)")
        head = (fn.__doc__ or "").split("\n")[0]
        inner = head[head.index("(") + 1:head.rindex(")")] if "(" in head else ""
        if "*args" in inner:
            return True
        depth, count = 0, (1 if inner.strip() else 0)
        for ch in inner:
            depth += ch in "[("
            depth -= ch in "])"
            count += (ch == "," and depth == 0)
        return count == n_args
    try:
        sig.bind(*([None] * n_args))
        return True
    except TypeError:
        return False


def test_every_call_site_matches_our_signatures(ref_svox2):
    svox2, ours = ref_svox2
    src = open(os.path.join(REF, "svox2", "svox2.py")).read()
    tree = ast.parse(src)
    checked, problems, seen = 0, [], {}

    def names_of(node):
        """_C.__dict__[<str or f-string>] -> candidate function names"""
        if isinstance(node, ast.Constant) and isinstance(node.value, str):
            return [node.value]
        if isinstance(node, ast.JoinedStr):
            outs = [""]
            for part in node.values:
                if isinstance(part, ast.Constant):
                    outs = [o + part.value for o in outs]
                else:
                    outs = [o + b for o in outs for b in BACKENDS]
            return outs
        return []

    for fn in ast.walk(tree):
        if not isinstance(fn, (ast.FunctionDef, ast.AsyncFunctionDef)):
            continue
        events = []
        for node in ast.walk(fn):
            if isinstance(node, ast.Assign) and isinstance(node.value, ast.Subscript):
                v = node.value.value
                if isinstance(v, ast.Attribute) and v.attr == "__dict__" and isinstance(v.value, ast.Name) and v.value.id == "_C":
                    events.append((node.lineno, "assign", names_of(node.value.slice)))
            if isinstance(node, ast.Call):
                f = node.func
                n = len(node.args)
                if any(isinstance(a, ast.Starred) for a in node.args):
                    continue
                if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "_C":
                    events.append((node.lineno, "direct", (f.attr, n)))
                elif isinstance(f, ast.Name) and f.id == "cu_fn":
                    events.append((node.lineno, "cu_fn", n))
        groups = []            # [candidate names of a `cu_fn = _C.__dict__[...]`, arities of the cu_fn(...) calls that follow]
        for _, kind, payload in sorted(events, key=lambda e: e[0]):
            if kind == "assign":
                groups.append([payload, set()])
            elif kind == "direct":
                name, n = payload
                if not hasattr(ours, name):
                    if name not in ("alpha_lap_grad_sparse", "sparse_grid_mask_renderalpha_rescale"):   # absent from svox2.cpp too
                        problems.append("missing %s" % name)
                    continue
                obj = getattr(ours, name)
                if inspect.isclass(obj):
                    continue
                checked += 1
                if not _accepts(obj, n):
                    problems.append("%s called with %d positional args" % (name, n))
            elif groups:
                groups[-1][1].add(payload)
        for cands, arities in groups:
            for c in cands:
                if hasattr(ours, c):
                    seen.setdefault(c, set()).update(arities)
    # `cu_fn` call sites sit in per-backend branches and classes (the `surface` backend passes one more argument, and which
    # branch serves which backend is not visible statically): every on-path function must be callable with at least one of
    # the arities the file uses for its name pattern
    for c, arities in sorted(seen.items()):
        if not arities:
            continue
        checked += 1
        if not any(_accepts(getattr(ours, c), n) for n in arities):
            problems.append("%s: no call site with a matching arity among %s" % (c, sorted(arities)))
    assert checked > 40, checked
    assert not problems, problems
