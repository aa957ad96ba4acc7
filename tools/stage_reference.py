#!/usr/bin/env python
"""Stage the reference's OWN Python files, unmodified, into baseline/_ref/ (TEST / BENCH INFRASTRUCTURE ONLY).

    python tools/stage_reference.py          # needs $NTX_REFERENCE_ROOT (default /root/reference); no-op when it is absent

baseline/_ref/ is git-ignored (nothing of the reference enters the history) but travels to the GPU box with the snapshot,
like oracle/_ref/ (the reference's CUDA rebuilt for sm_100a).  Two trees are staged:

  baseline/_ref/callers/   nerf/{renderer,network_ff,utils}.py, tools/{encoding,activation,shape_tools,map,__init__}.py
                           — the code that CALLS the hot path (renderer.py:338 run_cuda, network_ff.py:55 forward; map.py:414
                           MeshProjector.project for the mesh front end; network_curvedfield.py
                           + the light-model files it imports: the product's own model);
  baseline/_ref/wrappers/  gridencoder/, ffmlp/, shencoder/, raymarching/ *.py — the reference's own operator wrappers, used only
                           by the reference arm (on top of oracle/_ref/_ref_*.so) in tests/test_gpu_reference_files.py and bench.py.

tools/run_reference_files.py imports the callers over either nerf_texture_b200/compat (product) or the wrappers (reference).
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NTX_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

CALLERS = ["nerf/renderer.py", "nerf/network_ff.py", "nerf/utils.py", "tools/__init__.py", "tools/encoding.py", "tools/activation.py", "tools/shape_tools.py",
           "tools/map.py", "nerf/network_curvedfield.py", "nerf/sg_light_model.py", "nerf/sh_light_model.py", "nerf/envmap_light_model.py"]
WRAPPERS = ["gridencoder/__init__.py", "gridencoder/grid.py", "gridencoder/grid_clustering.py", "ffmlp/__init__.py", "ffmlp/ffmlp.py",
            "shencoder/__init__.py", "shencoder/sphere_harmonics.py", "raymarching/__init__.py", "raymarching/raymarching.py"]


def stage(verbose=True):
    if not os.path.isdir(REF):
        return None
    n = 0
    for sub, files in (("callers", CALLERS), ("wrappers", WRAPPERS)):
        for f in files:
            src, dst = os.path.join(REF, f), os.path.join(DST, sub, f)
            if not os.path.exists(src):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            n += 1
    if verbose:
        print("[stage_reference] %d files -> %s" % (n, DST))
    return DST


if __name__ == "__main__":
    if stage() is None:
        print("[stage_reference] %s not present: nothing staged" % REF)
        sys.exit(0)
