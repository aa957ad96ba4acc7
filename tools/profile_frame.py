#!/usr/bin/env python
"""Render frames of the bench workload for ncu:  ncu --profile-from-start off ... python tools/profile_frame.py [cfg2]
Two warm-up frames run before cudaProfilerStart so that only one steady-state frame (or one cfg2 launch set) is captured."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from nerf_texture_b200 import _lib as L  # noqa: E402
from nerf_texture_b200 import render  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
field, rays_o, rays_d, bits = bench.build_scene(dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "frame"
if mode == "frame":
    for _ in range(2):
        render.render_rays(field, rays_o, rays_d, bits, 1, 128)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    out = render.render_rays(field, rays_o, rays_d, bits, 1, 128)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("iterations", out["iterations"])
else:
    B = 1 << 20
    g = torch.Generator().manual_seed(0)
    xs = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev)
    ds = torch.randn(B, 3, generator=g)
    ds = (ds / ds.norm(dim=1, keepdim=True)).to(dev)
    sig = torch.empty(B, device=dev); rgb = torch.empty(B, 3, device=dev)
    x01 = ((xs + 1) / 2).contiguous()
    feat = torch.empty(B, 32, dtype=torch.half, device=dev)
    h = torch.empty(B, 16, dtype=torch.half, device=dev)

    def once():
        field(xs, ds, out_sigmas=sig, out_rgbs=rgb)
        L.call("ntx_grid_encode_forward", L.ptr(x01), L.ptr(field.table), L.ptr(field.offsets), L.ptr(feat), B, 3, 2, 16, field.S, field.H, 0, None, 0, 1,
               L.F16, L.LAYOUT_BLC, L.stream())
        L.call("ntx_ffmlp_inference", L.ptr(feat), L.ptr(field.w_sigma), B, 32, 16, 64, 2, 0, 6, None, L.ptr(h), L.stream())
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    once()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
