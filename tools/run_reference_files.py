#!/usr/bin/env python
"""Run the reference's OWN, UNMODIFIED caller files (nerf/renderer.py, nerf/network_ff.py, tools/encoding.py, tools/activation.py,
staged by tools/stage_reference.py into baseline/_ref/callers/) on top of either operator stack:

    --backend ntx   the drop-in packages of nerf_texture_b200/compat  (gridencoder, ffmlp, shencoder, raymarching -> libntx.so)
    --backend ref   the reference's own wrappers (baseline/_ref/wrappers/) on the reference's own CUDA rebuilt for sm_100a
                    (oracle/_ref/_ref_*.so pre-seeded as `_gridencoder`, `_ffmlp`, `_shencoder`, `_raymarching`, the names the
                    wrappers try first: gridencoder/grid.py:9-12)

builds `NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True)` (network_ff.py:11), gives it the bench scene (random table,
seed-42 MLP weights — FFMLP.reset_parameters reseeds, ffmlp.py:142 —, ball occupancy), renders one frame through
`NeRFRenderer.render` -> `run_cuda` (renderer.py:665, :338) under fp16 autocast like the reference's evaluation loop
(nerf/utils.py `with torch.cuda.amp.autocast(enabled=self.fp16)`), and writes image/depth + loop statistics.

Absent third-party modules the callers import at module level but never touch on this path (turtle needs tkinter; trimesh,
tensorboardX, mcubes, torch_ema, ... are not in the image) are replaced by inert stubs — the reference files themselves are
byte-for-byte the reference's.  TEST / BENCH INFRASTRUCTURE: nothing under nerf_texture_b200/ imports this.
"""
import argparse
import importlib.abc
import importlib.machinery
import importlib.util
import json
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "baseline", "_ref")
STUBBED = ("turtle", "trimesh", "tensorboardX", "matplotlib", "mcubes", "torch_ema", "plyfile", "pytorch3d", "imageio", "pymesh", "open3d",
           "cv2", "sklearn", "PIL", "xatlas")


class _Inert:
    """anything you ask of a stub module: callable, subscriptable, attribute-able, usable as a base class"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """import hook: the STUBBED top-level packages (and any submodule of them) that are NOT importable here become inert stubs"""

    def __init__(self):
        self.missing = set()
        for name in STUBBED:
            try:
                if name == "turtle":
                    raise ImportError("turtle needs tkinter")      # do not even try (it may open a display)
                importlib.import_module(name)
            except Exception:
                self.missing.add(name)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.missing:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _load_ref_ext(name):
    path = os.path.join(ROOT, "oracle", "_ref", "_ref_%s.so" % name)
    spec = importlib.util.spec_from_file_location("_ref_%s" % name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def available(backend):
    ok = os.path.exists(os.path.join(STAGE, "callers", "nerf", "renderer.py"))
    if backend == "ref":
        ok = ok and os.path.exists(os.path.join(STAGE, "wrappers", "gridencoder", "grid.py")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "_ref_raymarching.so"))
    return ok


def import_reference_network(backend):
    """returns (NeRFNetwork class of the reference's nerf/network_ff.py, the `raymarching` module it will call)"""
    import torch  # noqa: F401  (must be loaded before the extensions)
    warnings.filterwarnings("ignore", category=FutureWarning)     # torch.cuda.amp.custom_fwd deprecation in the reference's files
    sys.meta_path.insert(0, _StubFinder())
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)                                   # nerf_texture_b200.scene (test scene) — NOT the compat packages
    if backend == "ntx":
        import nerf_texture_b200
        nerf_texture_b200.install()                                # compat/ first on sys.path: gridencoder, ffmlp, ... resolve to libntx
    else:
        for ext in ("gridencoder", "ffmlp", "shencoder", "raymarching"):
            sys.modules["_" + ext] = _load_ref_ext(ext)            # what `import _gridencoder as _backend` finds (grid.py:9)
        sys.path.insert(0, os.path.join(STAGE, "wrappers"))
    sys.path.insert(0, os.path.join(STAGE, "callers"))
    if "nerf" not in sys.modules:                                  # the reference's nerf/ has no __init__.py: a namespace package would do,
        pkg = types.ModuleType("nerf")                             # but be explicit about where it lives
        pkg.__path__ = [os.path.join(STAGE, "callers", "nerf")]
        sys.modules["nerf"] = pkg
    import raymarching
    from nerf.network_ff import NeRFNetwork
    return NeRFNetwork, raymarching


def import_reference_map():
    """the reference's tools/map.py (MeshProjector, map.py:340) over the drop-in `frnn` / `RayTracer` / `tinycudann` / `gridencoder`
    packages; the mesh libraries it imports at module level (trimesh, xatlas, open3d, pytorch3d) are inert stubs — MeshProjector.project
    and .knn only use torch, frnn and RayTracer."""
    import torch  # noqa: F401
    warnings.filterwarnings("ignore", category=FutureWarning)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import nerf_texture_b200
    nerf_texture_b200.install()
    callers = os.path.join(STAGE, "callers")
    if callers not in sys.path:
        sys.path.insert(0, callers)
    import tools.map as ref_map
    return ref_map


def import_reference_product_model():
    """(tools.map module, NeRFNetwork of the reference's nerf/network_curvedfield.py — the PRODUCT's model: mesh-anchored texture field,
    tcnn sigma / colour networks, SH direction encoding) over the drop-in packages"""
    ref_map = import_reference_map()
    if "nerf" not in sys.modules:
        pkg = types.ModuleType("nerf")
        pkg.__path__ = [os.path.join(STAGE, "callers", "nerf")]
        sys.modules["nerf"] = pkg
    from nerf.network_curvedfield import NeRFNetwork
    return ref_map, NeRFNetwork


def build_model(NeRFNetwork, device, seed=0):
    import torch
    sys.path.insert(0, ROOT)
    from nerf_texture_b200 import scene
    model = NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True)          # network_ff.py:11-49
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    with torch.no_grad():
        model.encoder.embeddings.copy_(torch.rand(model.encoder.embeddings.shape, generator=g) * 2 - 1)   # same table as bench.py's NGPField.random
    model = model.to(device).eval()
    model.density_bitfield.copy_(scene.ball_bitfield(1, 128, 1.0, device))
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["ntx", "ref"], required=True)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--out", default=None, help=".npz with image / depth")
    ap.add_argument("--time", type=int, default=0, help="timed frames (CUDA events, L2 flushed between frames)")
    ap.add_argument("--save-ckpt", default=None, help="write a checkpoint in the layout of Trainer.save_checkpoint (nerf/utils.py:1485-1521)")
    ap.add_argument("--load-ckpt", default=None, help="load such a checkpoint the way Trainer.load_checkpoint does (nerf/utils.py:1553-1574) instead of seeding the scene")
    args = ap.parse_args()
    if not available(args.backend):
        print(json.dumps({"unavailable": "reference files not staged (tools/stage_reference.py) or oracle/_ref missing"}))
        return
    import numpy as np
    import torch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    NeRFNetwork, raymarching = import_reference_network(args.backend)
    from nerf_texture_b200 import scene
    if args.load_ckpt:
        model = NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True)
        ckpt = torch.load(args.load_ckpt, map_location=dev)
        missing, unexpected = model.load_state_dict(ckpt["model"], strict=False)       # utils.py:1560
        assert not missing and not unexpected, (missing, unexpected)
        model.mean_count, model.mean_density = ckpt["mean_count"], ckpt["mean_density"]
        model = model.to(dev).eval()
    else:
        model = build_model(NeRFNetwork, dev)
    if args.save_ckpt:
        torch.save({"epoch": 1, "global_step": 1, "stats": {}, "mean_count": model.mean_count, "mean_density": model.mean_density, "model": model.state_dict()},
                   args.save_ckpt)
    rays_o, rays_d = scene.pinhole_rays(args.size, args.size, dev)
    calls = {"march_rays": 0, "samples": 0}
    orig_march = raymarching.march_rays

    def counting_march(n_alive, n_step, *a, **k):
        calls["march_rays"] += 1
        out = orig_march(n_alive, n_step, *a, **k)
        if calls.get("count"):
            calls["samples"] += int((out[2][:, 0] > 0).sum().item())
        return out

    raymarching.march_rays = counting_march

    def frame():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.half):
            return model.render(rays_o[None], rays_d[None], staged=False, bg_color=1, perturb=False, max_steps=1024)

    calls["count"] = True
    out = frame()
    torch.cuda.synchronize()
    res = {"backend": args.backend, "size": args.size, "iterations": calls["march_rays"], "samples": calls["samples"],
           "renderer": sys.modules["nerf.renderer"].__file__, "raymarching": raymarching.__file__}
    calls["count"] = False
    if args.out:
        np.savez(args.out, image=out["image"].float().cpu().numpy(), depth=out["depth"].float().cpu().numpy())
    if args.time:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(2):
            frame()
        evs = []
        for _ in range(args.time):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); frame(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        res["frame_ms"] = ts[len(ts) // 2]
        res["samples_per_s"] = res["samples"] / (res["frame_ms"] * 1e-3)
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    main()
