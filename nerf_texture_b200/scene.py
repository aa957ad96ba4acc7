"""Synthetic inputs of the benchmark configs (BASELINE.md / SURVEY.md 8d): camera rays and an analytic occupancy bit-field.
Everything is generated on the device with torch + libntx ops; nothing is read from disk."""
import math

import torch

from . import _lib as L


def pinhole_rays(H, W, device, fovy_deg=50.0, radius=2.5, azim_deg=30.0, elev_deg=20.0):
    """rays_o, rays_d [H*W, 3] of a camera on a sphere of `radius` looking at the origin (what nerf/utils.py::get_rays produces
    for such a pose: unit directions, pixel centres)."""
    az, el = math.radians(azim_deg), math.radians(elev_deg)
    eye = torch.tensor([radius * math.cos(el) * math.sin(az), radius * math.sin(el), radius * math.cos(el) * math.cos(az)], dtype=torch.float64)
    fwd = -eye / eye.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64))
    right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    focal = 0.5 * H / math.tan(0.5 * math.radians(fovy_deg))
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float64) + 0.5, torch.arange(W, dtype=torch.float64) + 0.5, indexing="ij")
    d = ((i - W / 2) / focal)[..., None] * right + (-(j - H / 2) / focal)[..., None] * up + fwd
    d = d / d.norm(dim=-1, keepdim=True)
    o = eye.expand_as(d)
    return o.reshape(-1, 3).float().contiguous().to(device), d.reshape(-1, 3).float().contiguous().to(device)


def ball_bitfield(cascade, grid_size, bound, device, radius=0.5, thresh=0.5):
    """density_bitfield [cascade*H^3/8] uint8 of the ball |x| < radius, built like NeRFRenderer.update_extra_state lays the
    density grid out (Morton order per cascade, renderer.py:585-600) and packed with the packbits kernel."""
    H = grid_size
    idx = torch.arange(H ** 3, dtype=torch.int32, device=device)
    coords = torch.empty(H ** 3, 3, dtype=torch.int32, device=device)
    L.call("ntx_morton3D_invert", L.ptr(idx), H ** 3, L.ptr(coords), L.stream())
    grid = torch.empty(cascade, H ** 3, dtype=torch.float32, device=device)
    for c in range(cascade):
        b = min(2.0 ** c, bound)
        xyz = ((coords.float() + 0.5) / H * 2 - 1) * b
        grid[c] = (xyz.norm(dim=-1) < radius).float()
    bits = torch.empty(cascade * H ** 3 // 8, dtype=torch.uint8, device=device)
    L.call("ntx_packbits", L.ptr(grid), cascade * H ** 3 // 8, float(thresh), L.ptr(bits), L.stream())
    return bits
