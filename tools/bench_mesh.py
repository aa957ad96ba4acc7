#!/usr/bin/env python
"""Mesh front end of the texture field — what `MeshProjector.project` (tools/map.py:414-433) costs per batch of samples:

    python tools/bench_mesh.py [--log2-samples 20] [--lat 340] [--cpu-samples 4096]

  * `trace`, `knn`: one nearest-hit cast / one K = 8 neighbour search per sample (C ABI ntx_mesh_trace / ntx_mesh_knn),
  * `project_fused`: the whole projection as ONE kernel (ntx_mesh_project),
  * `project_reference_chain`: the reference's UNMODIFIED MeshProjector.project (staged tools/map.py) on the drop-in `frnn` and
    `RayTracer` packages — the same searches, but ~30 torch kernels over [N,K,3] temporaries in between,
  * `texture_field_forward_*`: the reference's unmodified `MeshFeatureField.forward` (the product's field, cfgT grids) on the drop-ins, with
    its own projection chain and with the fused projection,
  * `cpu`: the oracle's exhaustive-scan restatement on the host cores, on a bounded sample of the same batch.

Samples: 8192 camera-like rays x 128 steps through the shell 0.45 <= |x| <= 0.95 around a 230 K-triangle bumpy sphere, in ray order
(consecutive samples are neighbours in space, as they are when the renderer feeds the field).  CUDA events, median of 10, a 256 MB
write between iterations flushes L2.  `*_ray_order_ms`: the same call without visiting the batch in Morton order.  Prints `RESULT {json}`; bench.py embeds it as `mesh`."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-samples", type=int, default=20)
    ap.add_argument("--lat", type=int, default=340)
    ap.add_argument("--cpu-samples", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--profile", action="store_true", help="one call of each kernel and nothing else (for ncu)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import _util as U                                   # synthetic mesh builders shared with the tests
    import nerf_texture_b200
    nerf_texture_b200.install()
    from nerf_texture_b200 import _lib as L
    from nerf_texture_b200 import mesh as M
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)

    v, f, vn = U.bumpy_sphere(args.lat, args.lat)
    t0 = time.time()
    mesh = M.Mesh(v, f)
    build_s = time.time() - t0
    N = 1 << args.log2_samples
    steps = 128
    rays = N // steps
    g = torch.Generator(device="cpu").manual_seed(0)
    o = torch.nn.functional.normalize(torch.randn(rays, 3, generator=g), dim=-1) * 1.5
    tgt = torch.nn.functional.normalize(torch.randn(rays, 3, generator=g), dim=-1) * 0.6
    d = torch.nn.functional.normalize(tgt - o, dim=-1)
    ts = torch.linspace(0.4, 2.4, steps)
    x = (o[:, None] + ts[None, :, None] * d[:, None]).reshape(-1, 3)
    rad = x.norm(dim=-1)
    x = torch.where(((rad < 0.45) | (rad > 0.95))[:, None], x / rad[:, None] * rad.clamp(0.45, 0.95)[:, None], x).contiguous().to(dev)
    dirs = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).to(dev)
    vn_d = torch.from_numpy(vn).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        if args.profile:
            fn()
            torch.cuda.synchronize()
            return float("nan")
        ms = []
        for i in range(args.iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ms.append(e0.elapsed_time(e1))
        ms.sort()
        return ms[len(ms) // 2]

    res = {"samples": N, "triangles": int(len(f)), "vertices": int(len(v)), "tree_build_s": round(build_s, 3), "info": mesh.info()}
    res["trace_ms"] = timed(lambda: mesh.trace(x, dirs))
    res["knn_ms"] = timed(lambda: mesh.knn(x, K=8))                       # default: the batch is visited in Morton order (mesh.spatial_order)
    res["knn_ray_order_ms"] = timed(lambda: mesh.knn(x, K=8, sort=False))
    res["spatial_order_ms"] = timed(lambda: M.spatial_order(x))
    res["project_fused_ms"] = timed(lambda: mesh.project(x, vn_d, K=8))
    res["project_fused_ray_order_ms"] = timed(lambda: mesh.project(x, vn_d, K=8, sort=False))
    res["project_fused_msamples_per_s"] = N / res["project_fused_ms"] / 1e3

    import run_reference_files as R
    if args.profile:
        print("RESULT " + json.dumps(res))
        return
    if os.path.exists(os.path.join(R.STAGE, "callers", "tools", "map.py")):
        try:
            ref_map = R.import_reference_map()
            import frnn
            from RayTracer import RayTracer
            mp = ref_map.MeshProjector.__new__(ref_map.MeshProjector)
            mp.mesh_vertices, mp.vertex_normals = torch.from_numpy(v).to(dev), vn_d
            _, _, _, mp.grid = frnn.frnn_grid_points(mp.mesh_vertices.unsqueeze(0), mp.mesh_vertices.unsqueeze(0), None, None, K=8, r=100., grid=None)
            mp.radius, mp.distance_method, mp.max_K = 100., "frnn", len(v)
            mp.raytracer, mp.depth_threshold = RayTracer(v, f), 9.5
            mp.tbn = torch.zeros(len(f), 3, 3, device=dev)
            mp.faces = torch.from_numpy(f.astype(np.int64)).to(dev)
            res["project_reference_chain_ms"] = timed(lambda: mp.project(x, K=8, h_threshold=0.1))            # the reference's code on the drop-ins
            res["project_fused_same_outputs_ms"] = timed(lambda: M.project(mp, x, K=8, h_threshold=0.1))     # one kernel + tbn gather + h_mask
            res["fused_vs_chain"] = res["project_reference_chain_ms"] / res["project_fused_same_outputs_ms"]
            # the product's field: the reference's unmodified MeshFeatureField.forward (map.py:621: projection -> two cfgT hash-grid encodes ->
            # frequency encoding of the sdf -> factorised normal net) on the drop-in packages, then with only the projection swapped
            original = ref_map.MeshProjector
            ref_map.MeshProjector = lambda *a, **k: mp
            try:
                torch.manual_seed(0)
                field = ref_map.MeshFeatureField(mesh_path=None, h_threshold=0.1, K=8, bound=1).to(dev)
            finally:
                ref_map.MeshProjector = original

            def field_forward():
                with torch.no_grad():
                    field(x, no_noise=True)
            res["texture_field_forward_reference_chain_ms"] = timed(field_forward)
            mp.project = lambda xyz, K=8, h_threshold=None, requires_grad_xyz=False, use_dir_vec=True: M.project(mp, xyz, K=K, h_threshold=h_threshold)
            res["texture_field_forward_fused_projection_ms"] = timed(field_forward)
            del mp.project
            # the product's MODEL (nerf/network_curvedfield.py, unmodified) rendered by the unmodified NeRFRenderer on the drop-ins: a 512^2 frame
            try:
                from nerf_texture_b200 import scene
                _, NeRFNetwork = R.import_reference_product_model()
                ref_map.MeshProjector = lambda *a, **k: mp
                try:
                    torch.manual_seed(0)
                    model = NeRFNetwork(surface_mesh_path=None, light_model="None", bound=1, cuda_ray=True).to(dev).eval()
                finally:
                    ref_map.MeshProjector = original
                rays_o, rays_d = scene.pinhole_rays(512, 512, dev)

                def frame():
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.half):
                        model.render(rays_o[None], rays_d[None], staged=False, bg_color=1, perturb=False, max_steps=1024)
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.half):
                    t0 = time.time()
                    model.update_extra_state()
                    torch.cuda.synchronize()
                    res["product_model_update_extra_state_s"] = round(time.time() - t0, 3)
                res["product_model_frame_512_reference_chain_ms"] = timed(frame)
                mp.project = lambda xyz, K=8, h_threshold=None, requires_grad_xyz=False, use_dir_vec=True: M.project(mp, xyz, K=K, h_threshold=h_threshold)
                res["product_model_frame_512_fused_projection_ms"] = timed(frame)
                del mp.project
            except Exception as e:
                res["product_model_frame_512_reference_chain_ms"] = {"unavailable": repr(e)[:300]}
        except Exception as e:
            res["project_reference_chain_ms"] = {"unavailable": repr(e)[:300]}

    if args.cpu_samples > 0:
        from oracle import oracle as O
        n = min(args.cpu_samples, N)
        sel = np.linspace(0, N - 1, n).astype(np.int64)
        xs = x[torch.from_numpy(sel).to(dev)].cpu().numpy()
        O.mesh_project(v, vn, f, xs[:64])                # warm the thread pool
        t0 = time.time()
        want = O.mesh_project(v, vn, f, xs)
        dt = time.time() - t0
        got = mesh.project(torch.from_numpy(xs).to(dev), vn_d, K=8)
        same = (got[3].cpu().numpy() == want[3])
        res["cpu"] = {"kind": "port", "cores": O.num_threads(), "sample": "%d of the %d samples (every %d-th), exhaustive scan" % (n, N, N // n),
                      "seconds": round(dt, 2), "msamples_per_s": n / dt / 1e6, "faces_equal": float(same.mean()),
                      "max_abs_sdf_diff_same_face": float(np.abs(got[1].cpu().numpy().reshape(-1) - want[1])[same].max())}
        res["gpu_over_cpu"] = res["project_fused_msamples_per_s"] / res["cpu"]["msamples_per_s"]
    res["gpu_launches"] = L.launches
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    main()
