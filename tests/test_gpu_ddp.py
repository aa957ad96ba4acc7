"""Data-parallel training step on real NCCL — needs two GPUs on the box (`gpurun --gpus 2`); skipped otherwise.
Two ranks run the hash-grid + sigma-MLP forward/backward (the drop-in modules, fp16 autocast) on their half of a batch and exchange
gradients with nerf_texture_b200.parallel.allreduce_gradients (one packed fp16 all-reduce over NVLink); the averaged gradients must equal
the single-process gradient of the whole batch up to the fp16 rounding of the exchange and the atomics' summation order.  The same step
under torch's DistributedDataParallel (what the reference does, nerf/utils.py:439-441) must agree as well."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import nerf_texture_b200
nerf_texture_b200.install()
from nerf_texture_b200 import parallel
from gridencoder import GridEncoder
from ffmlp import FFMLP
rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=2, device_id=dev)

class Field(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=17, desired_resolution=1024, align_corners=True)
        self.sigma_net = FFMLP(input_dim=32, output_dim=16, hidden_dim=64, num_layers=2)
    def forward(self, x):
        return self.sigma_net(self.encoder(x, bound=1))

def build():
    torch.manual_seed(0)
    m = Field()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        m.encoder.embeddings.copy_(torch.rand(m.encoder.embeddings.shape, generator=g) * 2 - 1)
    return m.to(dev).train()

B = 1 << 15
g = torch.Generator().manual_seed(5)
x = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev)
gy = torch.randn(B, 16, generator=g).half().to(dev)

def step(model, xs, gs):
    for p in model.parameters(): p.grad = None
    with torch.autocast("cuda", dtype=torch.half):
        y = model(xs)
    y.backward(gs)

# whole batch on one process (each rank computes it: the reference result); loss = sum over the batch
full = build(); step(full, x, gy)
want = [p.grad.clone() for p in full.parameters()]
# data parallel: each rank its half, gradients SUMMED over the ranks == whole-batch gradient
half = slice(rank * B // 2, (rank + 1) * B // 2)
dp = build(); step(dp, x[half], gy[half])
parallel.allreduce_gradients(dp.parameters(), average=False)
ok = True
for (n, p), w in zip(dp.named_parameters(), want):
    err = (p.grad - w).abs().max().item(); scale = w.abs().max().item()
    ok = ok and err <= 2e-2 * scale + 1e-3
    print("rank", rank, n, "err", err, "scale", scale)
# torch DDP on the drop-in modules (averages): 2 * averaged == whole-batch gradient
ddp = parallel.ddp(build(), rank); step(ddp, x[half], gy[half])
for p, w in zip(ddp.module.parameters(), want):
    err = (2 * p.grad - w).abs().max().item(); scale = w.abs().max().item()
    ok = ok and err <= 2e-2 * scale + 1e-3
print("RANK%%d %%s" %% (rank, "OK" if ok else "MISMATCH"))
dist.destroy_process_group()
""" % ROOT


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_data_parallel_step_nccl():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, "-c", WORKER], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for r, (so, se) in enumerate(outs):
        assert "RANK%d OK" % r in so, (so[-3000:], se[-3000:])


FRAME_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from nerf_texture_b200 import render, scene
rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=2, device_id=dev)
field = render.NGPField.random(dev, seed=0)
H = 96
rays_o, rays_d = scene.pinhole_rays(H, H, dev)
bits = scene.ball_bitfield(1, 128, 1.0, dev)
N = rays_o.shape[0]
whole = render.render_rays(field, rays_o, rays_d, bits, 1, 128, bg_color=0.5)
ok = True
nccl = render.render_image_sharded(field, rays_o, rays_d, bits, 1, 128, tile=256, bg_color=0.5)
ok = ok and all(torch.equal(nccl[k], whole[k]) for k in ("image", "depth", "weights_sum"))
ex = render.PeerFrameExchange.create(N, tile=256, device=dev)
flag = torch.tensor([1 if ex is not None else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if int(flag.item()) == 1:
    for frame in range(4):                                    # four frames: both halves of the double buffer, twice
        peer = render.render_image_sharded(field, rays_o, rays_d, bits, 1, 128, tile=256, bg_color=0.5, exchange=ex)
        ok = ok and all(torch.equal(peer[k], whole[k]) for k in ("image", "depth", "weights_sum"))
    print("RANK%%d PEER" %% rank)
else:
    print("RANK%%d NOPEER" %% rank)
torch.cuda.synchronize()
print("RANK%%d %%s" %% (rank, "OK" if ok else "MISMATCH"))
dist.destroy_process_group()
""" % ROOT


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_sharded_frame_nccl_and_peer_exchange():
    """BASELINE config 4 on two real GPUs: the ray-sharded frame assembled (a) after ONE NCCL all-gather and (b) by the kernel that reads
    the peers' blocks over NVLink (symmetric memory) must both equal the whole-frame render bit for bit, frame after frame."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, "-c", FRAME_WORKER], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for r, (so, se) in enumerate(outs):
        assert "RANK%d OK" % r in so, (so[-3000:], se[-3000:])
    print(outs[0][0][-200:])
