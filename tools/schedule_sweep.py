#!/usr/bin/env python
"""Frame time of a 1/world shard (one GPU) as a function of the sample schedule (budget multiple, max samples per ray and iteration):
    python tools/schedule_sweep.py [world ...]"""
import sys
sys.path.insert(0, ".")
import torch, bench
from nerf_texture_b200 import render
dev = torch.device("cuda", 0)
field, rays_o, rays_d, bits = bench.build_scene(dev)
N = rays_o.shape[0]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for world in [int(a) for a in sys.argv[1:]] or [8, 1]:
    idx = render.shard_indices(N, world, 0).to(dev)
    o_, d_ = rays_o[idx].contiguous(), rays_d[idx].contiguous()
    for sched in [(8, 64), (8, 128), (16, 64), (16, 128), (32, 128), (32, 256), (4, 32)]:
        try:
            for _ in range(3):
                o = render.render_rays(field, o_, d_, bits, 1, 128, schedule=sched, time_kernels=True)
            torch.cuda.synchronize()
            ts = []
            for _ in range(6):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); render.render_rays(field, o_, d_, bits, 1, 128, schedule=sched); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            print("shard 1/%d schedule %-10s iters %2d  ms %.2f (min %.2f)  march %.2f field %.2f" % (world, sched, o["iterations"], sorted(ts)[len(ts) // 2], min(ts), o["march_ms"], o["field_ms"]))
        except Exception as e:
            print("shard 1/%d schedule %s FAILED %s" % (world, sched, repr(e)[:200]))
        sys.stdout.flush()
