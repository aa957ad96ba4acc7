"""Build the reference's own CUDA extensions for sm_100a into oracle/_ref/  (TEST INFRASTRUCTURE ONLY).

The reference (yihua7/NeRF-Texture) has no CPU implementation of the hot path: every native entry
point is CUDA-only (gridencoder/src/gridencoder.cu:420, ffmlp/src/ffmlp.cu:636,
shencoder/src/shencoder.cu:403, raymarching/raymarching.py:34).  So the strongest checker we can
have is the reference's *own* kernels recompiled for the B200.  This script compiles them straight
from where they lie under /root/reference (nothing is copied into the repo) with nvcc, one shared
object per extension, named `_ref_<pkg>` so they can never be confused with the product:

    oracle/_ref/_ref_gridencoder.so   <- gridencoder/src/{gridencoder.cu,bindings.cpp}
    oracle/_ref/_ref_shencoder.so     <- shencoder/src/{shencoder.cu,bindings.cpp}
    oracle/_ref/_ref_raymarching.so   <- raymarching/src/{raymarching.cu,bindings.cpp}
    oracle/_ref/_ref_ffmlp.so         <- ffmlp/src/{ffmlp.cu,bindings.cpp} (+ vendored CUTLASS 2.8 headers)

Only deviations from the reference's own setup.py flags: `-std=c++17` (torch >= 2.1 headers reject
the reference's hard-coded c++14) and an explicit `-gencode arch=compute_100a,code=sm_100a`.

`oracle/_ref/` is git-ignored but NOT gpurun-ignored: the .so files travel to the GPU box, where
/root/reference does not exist.  Only tests/, __graft_entry__.smoke() and oracle/ scripts load them.
"""
import os
import subprocess
import sys
import sysconfig
import time

REF = os.environ.get("NTX_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "gridencoder": dict(srcs=["gridencoder/src/gridencoder.cu", "gridencoder/src/bindings.cpp"], inc=[], extra=[]),
    "shencoder": dict(srcs=["shencoder/src/shencoder.cu", "shencoder/src/bindings.cpp"], inc=[], extra=[]),
    "raymarching": dict(srcs=["raymarching/src/raymarching.cu", "raymarching/src/bindings.cpp"], inc=[], extra=[]),
    "ffmlp": dict(
        srcs=["ffmlp/src/ffmlp.cu", "ffmlp/src/bindings.cpp"],
        inc=["ffmlp/dependencies/cutlass/include", "ffmlp/dependencies/cutlass/tools/util/include"],
        extra=["--expt-extended-lambda", "-Xcompiler=-mf16c",
               "-Xcompiler=-Wno-float-conversion", "-Xcompiler=-fno-strict-aliasing"],
    ),
}


def _torch_flags():
    import torch
    from torch.utils.cpp_extension import include_paths, library_paths
    inc = include_paths("cuda") if "device_type" in include_paths.__code__.co_varnames else include_paths(True)
    libs = library_paths("cuda") if "device_type" in library_paths.__code__.co_varnames else library_paths(True)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc, libs, abi


def build_one(name, force=False, verbose=True):
    spec = EXTS[name]
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, f"_ref_{name}.so")
    srcs = [os.path.join(REF, s) for s in spec["srcs"]]
    if not all(os.path.exists(s) for s in srcs):
        return None  # reference not present (GPU box): use the prebuilt file
    if os.path.exists(out) and not force and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    inc, libs, abi = _torch_flags()
    cmd = ["nvcc", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "-gencode", "arch=compute_100a,code=sm_100a",
           "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
           f"-DTORCH_EXTENSION_NAME=_ref_{name}", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-w",
           "--expt-relaxed-constexpr"]  # torch's BuildExtension always passes this (COMMON_NVCC_FLAGS)
    cmd += spec["extra"]
    cmd += ["-I" + sysconfig.get_paths()["include"]]
    cmd += ["-I" + p for p in inc]
    cmd += ["-I" + os.path.join(REF, p) for p in spec["inc"]]
    cmd += srcs
    cmd += ["-L" + p for p in libs]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    cmd += ["-Xlinker", "-rpath," + ":".join(libs)]
    cmd += ["-o", out]
    t0 = time.time()
    if verbose:
        print(f"[build_ref] {name}: nvcc ... ({len(cmd)} args)", flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:] + r.stderr[-8000:])
        raise RuntimeError(f"reference extension {name} failed to build")
    if verbose:
        print(f"[build_ref] {name}: ok in {time.time() - t0:.0f}s -> {out}", flush=True)
    return out


def build_all(force=False, names=None):
    res = {}
    for n in (names or EXTS):
        res[n] = build_one(n, force=force)
    return res


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or None
    build_all(force="--force" in sys.argv, names=names)
