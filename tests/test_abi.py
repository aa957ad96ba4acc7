"""CPU tests: libntx.so loads without a GPU and exports exactly the C ABI declared in include/ntx.h; the ctypes binding
(nerf_texture_b200/_lib.py) covers every declared entry point; the product never imports the oracle."""
import ctypes
import os
import re

from _util import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "ntx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ntx_[a-z0-9_A-Z]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib_path = os.path.join(ROOT, "nerf_texture_b200", "lib", "libntx.so")
    assert os.path.exists(lib_path), "build with `python -m nerf_texture_b200.build` (or __graft_entry__.build())"
    lib = ctypes.CDLL(lib_path)          # no CUDA call is made by loading
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.ntx_last_error.restype = ctypes.c_char_p
    assert lib.ntx_version() >= 100
    lib.ntx_compact_rays_workspace_bytes.restype = ctypes.c_size_t
    assert lib.ntx_compact_rays_workspace_bytes(1 << 20) >= 8 + 8 * 1024


def test_binding_covers_the_abi():
    from nerf_texture_b200 import _lib
    bound = set(_lib._SIGS) | set(_lib._SIZE_FNS) | {"ntx_last_error", "ntx_version", "ntx_device_ok"}
    assert set(_declared()) <= bound, sorted(set(_declared()) - bound)


def test_argument_validation_needs_no_gpu():
    """error paths return before any CUDA call, with the reference's messages (gridencoder.cu:355, ffmlp.cu:658, ...)"""
    from nerf_texture_b200 import _lib as L
    import pytest
    with pytest.raises(RuntimeError, match="C must be 1, 2, 4, or 8"):
        L.call("ntx_grid_encode_forward", 8, 8, 8, 8, 4, 3, 3, 1, 1.0, 16, 0, None, 0, 0, 0, 1, None)
    with pytest.raises(RuntimeError, match="hidden_dim should in"):
        L.call("ntx_ffmlp_inference", 16, 16, 128, 32, 16, 48, 2, 0, 6, None, 16, None)
    with pytest.raises(RuntimeError, match="degree in"):
        L.call("ntx_sh_encode_forward", 8, 8, 4, 3, 9, 0, None, None)
    with pytest.raises(RuntimeError, match="workspace"):
        L.call("ntx_compact_rays", 4, 8, 8, 8, 8, 8, None, None)


def test_product_never_touches_the_oracle():
    """only tests/, bench.py's cpu legs and __graft_entry__.smoke() may use oracle/ (the checker is not the product)"""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "nerf_texture_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|libntx_oracle|#include.*oracle|dlopen", txt, flags=re.M):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_binding_arity_matches_the_header():
    """every ctypes signature in _lib.py has exactly as many arguments as the declaration in include/ntx.h"""
    from nerf_texture_b200 import _lib
    src = open(os.path.join(ROOT, "include", "ntx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(ntx_[a-z0-9_A-Z]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    sigs = dict(_lib._SIGS)
    sigs.update(_lib._SIZE_FNS)
    wrong = {n: (len(a), decl[n]) for n, a in sigs.items() if n in decl and len(a) != decl[n]}
    assert not wrong, wrong
    assert set(sigs) <= set(decl), sorted(set(sigs) - set(decl))
