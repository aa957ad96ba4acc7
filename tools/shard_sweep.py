import sys; sys.path.insert(0, ".")
import torch, bench
from nerf_texture_b200 import render
dev = torch.device("cuda", 0)
field, rays_o, rays_d, bits = bench.build_scene(dev)
N = rays_o.shape[0]
for world in (8, 4, 2, 1):
    idx = render.shard_indices(N, world, 0).to(dev)
    o_, d_ = rays_o[idx].contiguous(), rays_d[idx].contiguous()
    for _ in range(3): o = render.render_rays(field, o_, d_, bits, 1, 128, count_samples=True, time_kernels=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); render.render_rays(field, o_, d_, bits, 1, 128); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("shard 1/%d" % world, "iters", o["iterations"], "ms %.2f" % min(ts), "march %.2f field %.2f" % (o["march_ms"], o["field_ms"]))
