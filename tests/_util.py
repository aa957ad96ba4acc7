"""Shared helpers of the test-suite: loaders for the checker libraries and synthetic scene builders."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def oracle():
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from oracle import oracle as O
    O.lib()
    return O


def ref(name):
    """the reference's own CUDA extension rebuilt for sm_100a (oracle/_ref, see oracle/build_ref.py); skip if absent"""
    import torch  # noqa: F401  (libtorch must be loaded first)
    path = os.path.join(ROOT, "oracle", "_ref", "_ref_%s.so" % name)
    if not os.path.exists(path):
        pytest.skip("reference extension %s not built (oracle/_ref)" % name)
    modname = "_ref_%s" % name
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[modname] = mod
    return mod


def ntx():
    import nerf_texture_b200
    nerf_texture_b200.install()
    from nerf_texture_b200 import _lib
    _lib.lib()
    return _lib


# ------------------------------------------------------------------------------------------------ encoder configs
def cfgA():
    """network_ff's encoder: get_encoder('hashgrid', desired_resolution=2048) (tools/encoding.py:48,61-63)"""
    return dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048, align_corners=True)


def cfgB():
    """BASELINE config 1: L=4, T=2^14, F=2, constructor defaults otherwise"""
    return dict(input_dim=3, num_levels=4, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=14, align_corners=False)


def cfgT():
    """NeRF-Texture's texture grid (tools/map.py:563)"""
    return dict(input_dim=3, num_levels=8, level_dim=2, base_resolution=512, log2_hashmap_size=19, desired_resolution=1024, align_corners=True)


# ------------------------------------------------------------------------------------------------ synthetic scene
def _expand_bits(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3D_np(x, y, z):
    return _expand_bits(x) | (_expand_bits(y) << np.uint32(1)) | (_expand_bits(z) << np.uint32(2))


def ball_density_grid(cascade, H, bound, radius=0.5, center=(0.0, 0.0, 0.0)):
    """density_grid [cascade, H^3] (Morton order, like renderer.py:585-600): 1 inside the ball, 0 outside"""
    g = np.arange(H, dtype=np.uint32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    idx = morton3D_np(X.ravel(), Y.ravel(), Z.ravel()).astype(np.int64)
    grid = np.zeros((cascade, H ** 3), np.float32)
    for c in range(cascade):
        b = min(2.0 ** c, bound)
        xyz = (np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1).astype(np.float32) + 0.5) / H * 2 - 1
        xyz = xyz * b - np.asarray(center, np.float32)
        grid[c, idx] = (np.linalg.norm(xyz, axis=1) < radius).astype(np.float32)
    return grid


def pinhole_rays(H, W, fovy_deg=50.0, radius=2.5, azim_deg=30.0, elev_deg=20.0, dtype=np.float32):
    """rays of a camera on a sphere looking at the origin (OpenGL convention: camera looks down -z)"""
    az, el = np.deg2rad(azim_deg), np.deg2rad(elev_deg)
    eye = np.array([radius * np.cos(el) * np.sin(az), radius * np.sin(el), radius * np.cos(el) * np.cos(az)], np.float64)
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, np.array([0.0, 1.0, 0.0])); right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    focal = 0.5 * H / np.tan(0.5 * np.deg2rad(fovy_deg))
    i, j = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5, indexing="xy")
    d = ((i - W / 2) / focal)[..., None] * right + (-(j - H / 2) / focal)[..., None] * up + fwd
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    o = np.broadcast_to(eye, d.shape)
    return np.ascontiguousarray(o.reshape(-1, 3).astype(dtype)), np.ascontiguousarray(d.reshape(-1, 3).astype(dtype))


def ulp16(x):
    """fp16 unit in the last place at |x| (as float32)"""
    x = np.abs(np.asarray(x, np.float32))
    e = np.floor(np.log2(np.maximum(x, 2.0 ** -14)))
    return (2.0 ** (e - 10)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ synthetic meshes
def bumpy_sphere(n_lat, n_lon, bump=0.1, radius=0.7):
    """closed-ish lat/lon sphere with a smooth bump pattern: vertices [n,3] f32, faces [m,3] i32, outward vertex normals [n,3] f32"""
    th = np.linspace(0.05, np.pi - 0.05, n_lat)
    ph = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)
    T, P = np.meshgrid(th, ph, indexing="ij")
    r = radius + bump * np.sin(5 * T) * np.cos(3 * P)
    v = np.stack([r * np.sin(T) * np.cos(P), r * np.sin(T) * np.sin(P), r * np.cos(T)], -1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(n_lat - 1), np.arange(n_lon), indexing="ij")
    a, b = i * n_lon + j, i * n_lon + (j + 1) % n_lon
    c, d = (i + 1) * n_lon + j, (i + 1) * n_lon + (j + 1) % n_lon
    f = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], 2).reshape(-1, 3).astype(np.int32)
    return v, f, vertex_normals(v, f)


def vertex_normals(v, f):
    """area-weighted vertex normals (what open3d's compute_vertex_normals gives the reference, tools/map.py:369,394)"""
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    vn = np.zeros_like(v)
    for k in range(3):
        np.add.at(vn, f[:, k], fn)
    return (vn / (np.linalg.norm(vn, axis=1, keepdims=True) + 1e-12)).astype(np.float32)


def random_rays(rng, n, extent=1.0):
    o = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    return o, (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def adversarial_rays(rng, v, f):
    """rays aimed exactly at vertices, edge midpoints and centroids (ties between the triangles that share them), from inside and from
    outside; axis-parallel rays from lattice origins (rays that run IN box faces); rays that start on vertices"""
    tgt = np.concatenate([v, 0.5 * (v[f[:, 0]] + v[f[:, 1]]), (v[f[:, 0]] + v[f[:, 1]] + v[f[:, 2]]) / 3]).astype(np.float32)
    d = (tgt / np.linalg.norm(tgt, axis=1, keepdims=True)).astype(np.float32)
    ax = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 4000)] * rng.choice([-1.0, 1.0], (4000, 1)).astype(np.float32)
    oa = rng.uniform(-1, 1, (4000, 3)).astype(np.float32)
    oa[:2000] = np.round(oa[:2000] * 4) / 4
    ov = v[rng.integers(0, len(v), 4000)]
    dv = rng.normal(size=(4000, 3)).astype(np.float32)
    o = np.concatenate([np.zeros_like(tgt), 2 * tgt, oa, ov]).astype(np.float32)
    return o, np.concatenate([d, -d, ax, dv]).astype(np.float32)


def build_mesh_host_check():
    """g++ build of tests/native/mesh_host_check.cpp (the product's tree builder + traversals compiled for the host); returns a CDLL"""
    import ctypes
    import subprocess
    src = os.path.join(ROOT, "tests", "native", "mesh_host_check.cpp")
    out_dir = os.path.join(ROOT, "tests", "native", "_build")
    so = os.path.join(out_dir, "libmesh_host_check.so")
    deps = [src] + [os.path.join(ROOT, "nerf_texture_b200", "csrc", f) for f in ("mesh_bvh.cuh", "mesh_build.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        os.makedirs(out_dir, exist_ok=True)
        cmd = ["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-o", so, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("mesh_host_check build failed:\n" + r.stderr[-4000:])
    return ctypes.CDLL(so)
