#!/usr/bin/env python
"""libntx vs the REFERENCE'S OWN CUDA kernels on the same B200.

The reference extensions are the unmodified sources of /root/reference rebuilt for sm_100a by oracle/build_ref.py
(oracle/_ref/_ref_*.so — test infrastructure, never on the product path).  They are driven here the way the reference's Python
drives them: GridEncoder.forward (grid.py:143-152: [L,B,C] output + permute), FFMLP.forward, SHEncoder, network_ff.forward's torch
glue under fp16 (network_ff.py:85-101) and NeRFRenderer.run_cuda's inference loop with its per-iteration .item() (renderer.py:436-489).

    python tools/compare_ref.py            # prints one JSON object
"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from nerf_texture_b200 import _lib as L  # noqa: E402
from nerf_texture_b200 import render  # noqa: E402


def ref(name):
    path = os.path.join(ROOT, "oracle", "_ref", "_ref_%s.so" % name)
    spec = importlib.util.spec_from_file_location("_ref_%s" % name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def time_it(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main(only_frame=False):
    dev = torch.device("cuda", torch.cuda.current_device())
    G, F, S, R = ref("gridencoder"), ref("ffmlp"), ref("shencoder"), ref("raymarching")
    F.allocate_splitk(4)
    field, rays_o, rays_d, bits = bench.build_scene(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nlev, bound = field.num_levels, field.bound
    dummy_h = torch.empty(1, dtype=torch.half, device=dev)
    dummy_f = torch.empty(1, device=dev)

    def ref_field(xyz, dirs):
        """network_ff.forward (fp16 autocast) on the reference extensions"""
        M = xyz.shape[0]
        x01 = ((xyz + bound) / (2 * bound)).contiguous()
        out = torch.empty(nlev, M, 2, dtype=torch.half, device=dev)
        G.grid_encode_forward(x01, field.table, field.offsets, out, M, 3, 2, nlev, field.S, field.H, False, dummy_h, 0, field.align_corners)
        feat = out.permute(1, 0, 2).reshape(M, nlev * 2).contiguous()
        h = torch.empty(M, 16, dtype=torch.half, device=dev)
        buf = torch.empty(M, 64, dtype=torch.half, device=dev)
        F.ffmlp_inference(feat, field.w_sigma, M, nlev * 2, 16, 64, 2, 0, 6, buf, h)
        sigma = torch.exp(h[:, 0].float())
        sh = torch.empty(M, 16, device=dev)
        S.sh_encode_forward(dirs.contiguous(), sh, M, 3, 4, False, dummy_f)
        cin = torch.cat([sh, h[:, 1:].float(), torch.zeros(M, 1, device=dev)], dim=-1).half().contiguous()
        hc = torch.empty(M, 16, dtype=torch.half, device=dev)
        F.ffmlp_inference(cin, field.w_color, M, 32, 16, 64, 3, 0, 6, buf, hc)
        return sigma, torch.sigmoid(hc[:, :3]).float()

    res = {}
    if not only_frame:
        res.update(_config2(dev, field, flush, ref_field, G, F))
    res.update(_config3(dev, field, rays_o, rays_d, bits, flush, ref_field, R))
    res["note"] = ("reference_cuda = /root/reference's gridencoder/ffmlp/shencoder/raymarching .cu rebuilt unmodified for sm_100a (oracle/_ref), driven like its "
                   "Python drives them; median of 5-10 timed runs, L2 flushed before each; same B200, same inputs")
    return res


def _config2(dev, field, flush, ref_field, G, F):
    nlev = field.num_levels
    dummy_h = torch.empty(1, dtype=torch.half, device=dev)
    res = {}
    # ---------------- config 2: 2^20 samples through the field --------------------------------------------------------
    B = 1 << 20
    g = torch.Generator().manual_seed(0)
    xs = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev)
    ds = torch.randn(B, 3, generator=g)
    ds = (ds / ds.norm(dim=1, keepdim=True)).to(dev)
    sig, rgb = torch.empty(B, device=dev), torch.empty(B, 3, device=dev)
    t_ref = time_it(lambda: ref_field(xs, ds), flush=flush)
    t_ntx = time_it(lambda: field(xs, ds, out_sigmas=sig, out_rgbs=rgb), flush=flush)
    rs, rr = ref_field(xs, ds)
    field(xs, ds, out_sigmas=sig, out_rgbs=rgb)
    torch.cuda.synchronize()
    res["cfg2_field_2^20_random"] = {"reference_cuda_ms": t_ref, "ntx_ms": t_ntx, "speedup": t_ref / t_ntx,
                                     "max_rel_sigma_diff": float(((sig - rs).abs() / rs.clamp_min(1e-30)).max()), "max_abs_rgb_diff": float((rgb - rr).abs().max())}
    x01 = ((xs + 1) / 2).contiguous()
    out_ref = torch.empty(nlev, B, 2, dtype=torch.half, device=dev)
    out_ntx = torch.empty(B, nlev * 2, dtype=torch.half, device=dev)

    def ref_grid():
        G.grid_encode_forward(x01, field.table, field.offsets, out_ref, B, 3, 2, nlev, field.S, field.H, False, dummy_h, 0, field.align_corners)
        return out_ref.permute(1, 0, 2).reshape(B, nlev * 2).contiguous()

    t_ref = time_it(ref_grid, flush=flush)
    t_ntx = time_it(lambda: L.call("ntx_grid_encode_forward", L.ptr(x01), L.ptr(field.table), L.ptr(field.offsets), L.ptr(out_ntx), B, 3, 2, nlev, field.S, field.H, 0,
                                   None, 0, int(field.align_corners), L.F16, L.LAYOUT_BLC, L.stream()), flush=flush)
    res["cfg2_grid_encode_2^20_random"] = {"reference_cuda_ms": t_ref, "ntx_ms": t_ntx, "speedup": t_ref / t_ntx, "bit_identical": bool(torch.equal(ref_grid(), out_ntx))}
    feat = torch.randn(B, 32, device=dev).half()
    h_ref, h_ntx, buf = torch.empty(B, 16, dtype=torch.half, device=dev), torch.empty(B, 16, dtype=torch.half, device=dev), torch.empty(B, 64, dtype=torch.half, device=dev)
    t_ref = time_it(lambda: F.ffmlp_inference(feat, field.w_sigma, B, 32, 16, 64, 2, 0, 6, buf, h_ref), flush=flush)
    t_ntx = time_it(lambda: L.call("ntx_ffmlp_inference", L.ptr(feat), L.ptr(field.w_sigma), B, 32, 16, 64, 2, 0, 6, None, L.ptr(h_ntx), L.stream()), flush=flush)
    res["cfg2_ffmlp_32_64_64_16_2^20"] = {"reference_cuda_ms": t_ref, "ntx_ms": t_ntx, "speedup": t_ref / t_ntx,
                                          "max_abs_diff": float((h_ref.float() - h_ntx.float()).abs().max())}

    return res


def _config3(dev, field, rays_o, rays_d, bits, flush, ref_field, R):
    bound = field.bound
    res = {}
    # ---------------- config 3: the 1024x1024 frame ------------------------------------------------------------------
    N = rays_o.shape[0]
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=dev)

    def ref_frame():
        """NeRFRenderer.run_cuda, inference branch (renderer.py:436-489), on the reference's raymarching extension"""
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        R.near_far_from_aabb(rays_o, rays_d, aabb, N, 0.2, nears, fars)
        weights_sum, depth, image = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
        n_alive = N
        alive_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        rays_alive = torch.zeros(2, n_alive, dtype=torch.int32, device=dev)
        rays_t = torch.zeros(2, n_alive, dtype=torch.float32, device=dev)
        step, i = 0, 0
        while step < 1024:
            if step == 0:
                rays_alive[0] = torch.arange(n_alive, dtype=torch.int32, device=dev)
                rays_t[0] = nears
            else:
                alive_counter.zero_()
                R.compact_rays(n_alive, rays_alive[i % 2], rays_alive[(i + 1) % 2], rays_t[i % 2], rays_t[(i + 1) % 2], alive_counter)
                n_alive = alive_counter.item()
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            M = n_alive * n_step
            M += 128 - (M % 128)
            xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
            R.march_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], rays_o, rays_d, bound, 0.0, 1024, 1, 128, bits, nears, fars, xyzs, dirs, deltas, 0)
            sigmas, rgbs = ref_field(xyzs, dirs)
            R.composite_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], sigmas, rgbs, deltas, weights_sum, depth, image)
            step += n_step
            i += 1
        return image + (1 - weights_sum).unsqueeze(-1), depth, i

    t_ref = time_it(ref_frame, iters=5, warm=2, flush=flush)
    t_ntx = time_it(lambda: render.render_rays(field, rays_o, rays_d, bits, 1, 128), iters=10, warm=3, flush=flush)
    img_r, dep_r, it_r = ref_frame()
    o = render.render_rays(field, rays_o, rays_d, bits, 1, 128, count_samples=True)
    torch.cuda.synchronize()
    res["cfg3_frame_1024x1024"] = {"reference_cuda_ms": t_ref, "ntx_ms": t_ntx, "speedup": t_ref / t_ntx, "samples": o["n_samples"],
                                   "reference_gsamples_per_s": o["n_samples"] / t_ref / 1e6, "ntx_gsamples_per_s": o["n_samples"] / t_ntx / 1e6,
                                   "iterations": [it_r, o["iterations"]], "max_abs_image_diff": float((img_r - o["image"]).abs().max()),
                                   "max_abs_depth_diff": float((dep_r - o["depth"]).abs().max())}
    return res


if __name__ == "__main__":
    torch.cuda.set_device(0)
    print(json.dumps(main()))
