"""CPU-only checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/vrad_cuda.h declares; the host-side helpers behave; no compute is attempted without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="vrad_cuda.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vrad_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vrad_b200 import lib
    handle = lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for sym in declared:
        assert hasattr(handle, sym), f"libvradcuda.so does not export {sym}"
    assert sorted(lib.SYMBOLS) == declared
    assert b"sm_100a" in handle.vrad_version()
    # the BSP side of the boundary (include/vrad_bsp.h)
    from vrad_b200 import bspfile
    declared_bsp = _declared_symbols("vrad_bsp.h")
    assert len(declared_bsp) >= 20 and sorted(bspfile.SYMBOLS) == declared_bsp
    for sym in declared_bsp:
        assert hasattr(handle, sym), f"libvradcuda.so does not export {sym}"
    assert sorted(os.listdir(os.path.join(ROOT, "include"))) == ["vrad_bsp.h", "vrad_cuda.h"]


def test_struct_sizes_match_header():
    from vrad_b200 import lib, scenes
    assert lib.TRI48_DTYPE.itemsize == 48 and scenes.LIGHT_DTYPE.itemsize == 96
    assert C.sizeof(lib.VradConfig) == 16


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vrad_b200.environment import Environment, VradError
    with pytest.raises(VradError) as ei:
        Environment()
    assert ei.value.status == -2 and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under vrad_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vrad_b200")):
        if "_lib" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in text and "liboracle" not in text and "oracle.h" not in text, f


def test_row_partition_rule():
    from vrad_b200.environment import row_partition
    for n, w in ((187328, 8), (10, 3), (7, 8), (4096, 1)):
        parts = row_partition(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n and len(parts) == w
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        rpr = ((n + w - 1) // w + 3) & ~3
        assert all(hi - lo <= rpr for lo, hi in parts) and all(lo % 4 == 0 or lo == n for lo, _ in parts)


def test_splitmix64_reference_vector():
    """splitmix64 known answers (seed 1234567: first outputs of the public reference implementation)."""
    from vrad_b200.scenes import SplitMix64
    got = SplitMix64(1234567).u64(3)
    assert [int(x) for x in got] == [6457827717110365317, 3203168211198807973, 9817491932198370423]


def test_bsp_side_entry_points_reject_null_arguments():
    """Every entry point of include/vrad_bsp.h called with nothing but zeros / NULLs: a status (0 only where 'nothing to do' is a valid
    request), never a crash.  (x86-64 SysV: surplus integer arguments are ignored by the callee.)"""
    import ctypes as C
    from vrad_b200 import bspfile, lib
    handle = lib.load()
    zeros = [C.c_void_p(0)] * 14
    ok_on_empty = {"vrad_bsp_rescale_lightmap_vecs", "vrad_color_to_rgbexp32", "vrad_color_from_rgbexp32", "vrad_luxel_nearest_patch",
                   "vrad_luxel_radial_light_host"}
    for sym in bspfile.SYMBOLS:
        fn = getattr(handle, sym)
        if sym == "vrad_bspfile_close":
            fn.restype = None
            fn(None)
            continue
        fn.restype = C.c_int
        rc = fn(*zeros)
        assert (rc == 0) if sym in ok_on_empty else (rc < 0), (sym, rc)
        if rc < 0:
            assert handle.vrad_last_error()                      # and it says why


def _top_level_args(text, start):
    """Number of top-level comma-separated arguments of the call whose '(' is at text[start]."""
    depth, args, i, seen = 0, 0, start, False
    while i < len(text):
        c = text[i]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                return args + (1 if seen else 0)
        elif c == "," and depth == 1:
            args += 1
        elif not c.isspace() and depth >= 1:
            seen = True
        i += 1
    raise AssertionError("unbalanced call")


def test_cgo_sources_call_declared_entry_points_with_the_right_arity():
    """The Go shims cannot be compiled here (no Go toolchain), so at least every C.vrad_* call in integration/go names a function the
    headers declare and passes as many arguments as the declaration has parameters."""
    decl = {}
    for header in ("vrad_cuda.h", "vrad_bsp.h"):
        src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", header)).read(), flags=re.S)
        for m in re.finditer(r"\b(vrad_[a-z0-9_]+)\s*\(", src):
            n = _top_level_args(src, m.end() - 1)
            params = src[m.end():src.index(")", m.end())].strip()
            decl[m.group(1)] = 0 if params in ("", "void") else n
    calls = 0
    for dirpath, _, files in os.walk(os.path.join(ROOT, "integration", "go")):
        for f in files:
            if not f.endswith(".go"):
                continue
            text = open(os.path.join(dirpath, f)).read()
            text = re.sub(r"//[^\n]*", "", text)
            for m in re.finditer(r"\bC\.(vrad_[a-z0-9_]+)\(", text):
                name = m.group(1)
                assert name in decl, f"{f}: C.{name} is not declared in include/"
                assert _top_level_args(text, m.end() - 1) == decl[name], f"{f}: C.{name} called with the wrong number of arguments"
                calls += 1
    assert calls >= 40
