#!/usr/bin/env python
"""Sweep libntx's run-time tunables (environment variables) on the B200:  python tools/tune.py [spec ...]
Each spec is a comma-separated list of KEY=VALUE; one subprocess per spec runs the cfg2 micro-benchmarks and one frame."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, sys, os
sys.path.insert(0, %r)
import torch, bench
from nerf_texture_b200 import _lib as L, render
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
field, rays_o, rays_d, bits = bench.build_scene(dev)
peaks, _ = bench.measured_peaks()
out = bench.bench_cfg2(torch, L, field, dev, peaks)
res = {k: {kk: round(vv, 1) for kk, vv in v.items() if kk.endswith('us')} for k, v in out.items() if isinstance(v, dict)}
for _ in range(2):
    render.render_rays(field, rays_o, rays_d, bits, 1, 128)
torch.cuda.synchronize()
evs = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); render.render_rays(field, rays_o, rays_d, bits, 1, 128); e1.record(); evs.append((e0, e1))
torch.cuda.synchronize()
res['frame_ms'] = round(min(a.elapsed_time(b) for a, b in evs), 2)
print('RESULT ' + json.dumps(res))
""" % ROOT


def main():
    specs = sys.argv[1:] or [""]
    for spec in specs:
        env = dict(os.environ)
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            env[k] = v
        r = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(spec or "(default)", "FAILED", r.stderr[-800:])
            continue
        d = json.loads(line[0][7:])
        print("%-44s frame %6.2f ms | rand: field %7.1f grid %6.1f | coh: field %7.1f grid %6.1f | mlp %6.1f us" % (
            spec or "(default)", d["frame_ms"], d["random"]["fused_field_us"], d["random"]["grid_encode_us"], d["coherent"]["fused_field_us"],
            d["coherent"]["grid_encode_us"], d["ffmlp_32_64_64_16"]["us"]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
