#!/usr/bin/env python
"""Which half of the warp-specialised field kernel is the bottleneck?  (development aid, needs a probe build:
    NTX_NVCC_EXTRA=-DNTX_DEV_PROBES python -m nerf_texture_b200.build ; python tools/field_probe.py )
Times ntx_ngp_field_forward on 2^20 coherent / random samples with NTX_FIELD_DEBUG = 0 (normal), 1 (no gather: consumer
chain alone), 2 (no MLP: producers alone), 3 (neither: fixed overhead)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r"""
import json, sys
sys.path.insert(0, %r)
import torch, bench
from nerf_texture_b200 import _lib as L
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
field, rays_o, rays_d, bits = bench.build_scene(dev)
peaks, _ = bench.measured_peaks()
out = bench.bench_cfg2(torch, L, field, dev, peaks)
res = {k: v.get('fused_field_us') for k, v in out.items() if isinstance(v, dict) and 'fused_field_us' in v}
# cycle accounting of one frame (one consumer thread and one producer warp per CTA; cycles per tile of that CTA)
import ctypes
from nerf_texture_b200 import render
lib = L.lib()
if hasattr(lib, 'ntx_dev_probe_read'):
    buf = (ctypes.c_ulonglong * 8)()
    render.render_rays(field, rays_o, rays_d, bits, 1, 128); torch.cuda.synchronize()
    lib.ntx_dev_probe_read(buf, 1)
    o = render.render_rays(field, rays_o, rays_d, bits, 1, 128, count_samples=True); torch.cuda.synchronize()
    lib.ntx_dev_probe_read(buf, 1)
    v = list(buf)
    res['frame_probe_Mcycles'] = dict(cons_wait_full=v[0] / 1e6, cons_wait_mma=v[1] / 1e6, cons_epilogue=v[2] / 1e6, prod_wait_empty=v[4] / 1e6,
                                      prod_gather=v[5] / 1e6, prod_loads=v[6] / 1e6, ctas=v[7])
print('RESULT ' + json.dumps(res))
""" % ROOT

for dbg in sys.argv[1:] or ["0", "1", "2", "3"]:
    env = dict(os.environ, NTX_FIELD_DEBUG=dbg)
    r = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    print("NTX_FIELD_DEBUG=%s" % dbg, line[0][7:] if line else "FAILED " + r.stderr[-600:])
    sys.stdout.flush()
