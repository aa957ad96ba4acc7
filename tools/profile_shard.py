import sys; sys.path.insert(0, ".")
import torch, bench
from nerf_texture_b200 import render
dev = torch.device("cuda", 0)
field, rays_o, rays_d, bits = bench.build_scene(dev)
N = rays_o.shape[0]
idx = render.shard_indices(N, 8, 0).to(dev)
o_, d_ = rays_o[idx].contiguous(), rays_d[idx].contiguous()
for _ in range(3): render.render_rays(field, o_, d_, bits, 1, 128)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
render.render_rays(field, o_, d_, bits, 1, 128)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
