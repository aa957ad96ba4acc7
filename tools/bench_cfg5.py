#!/usr/bin/env python
"""BASELINE config 5 — training step: forward + backward of the hash-grid (table gradients) and the sigma MLP (weight gradients) on
2^18 samples, fp16 autocast, through the operator API the reference's trainer uses (GridEncoder.forward -> FFMLP.forward ->
loss.backward(), nerf/network_ff.py:85-88 + nerf/utils.py train_step).

    python tools/bench_cfg5.py --backend ntx     # nerf_texture_b200/compat -> libntx.so
    python tools/bench_cfg5.py --backend ref     # the reference's own wrappers (baseline/_ref/wrappers) on its own CUDA (oracle/_ref)

Prints `RESULT {json}`: step time (CUDA events, median, L2 flushed between steps), samples/s, and — for ntx — the device time of
each C-ABI call of the step (grid fwd, mlp fwd, mlp bwd = dgrad + wgrad, grid bwd).  bench.py runs both and reports the ratio.
Inputs: SURVEY.md 8d cfg5 (x ~ U[0,1)^3 seed 0, upstream gradient ~ N(0,1) * 128 seed 2, table U(-1,1) seed 1, FFMLP seed 42).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["ntx", "ref"], required=True)
    ap.add_argument("--log2-batch", type=int, default=18)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dump", default=None, help=".npz with the gradients of one step (parity check between the arms)")
    args = ap.parse_args()
    import run_reference_files as R
    if args.backend == "ref" and not R.available("ref"):
        print("RESULT " + json.dumps({"unavailable": "reference wrappers / oracle/_ref not present"}))
        return
    import numpy as np
    import torch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    import warnings
    warnings.filterwarnings("ignore", category=FutureWarning)
    sys.meta_path.insert(0, R._StubFinder())
    if args.backend == "ntx":
        import nerf_texture_b200
        nerf_texture_b200.install()
    else:
        for ext in ("gridencoder", "ffmlp", "shencoder", "raymarching"):
            sys.modules["_" + ext] = R._load_ref_ext(ext)
        sys.path.insert(0, os.path.join(R.STAGE, "wrappers"))
    from gridencoder import GridEncoder
    from ffmlp import FFMLP

    B = 1 << args.log2_batch
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048, gridtype="hash",
                      align_corners=True).to(dev)
    mlp = FFMLP(input_dim=32, output_dim=16, hidden_dim=64, num_layers=2).to(dev)
    g = torch.Generator(device="cpu").manual_seed(1)
    with torch.no_grad():
        enc.embeddings.copy_(torch.rand(enc.embeddings.shape, generator=g) * 2 - 1)
    x = (torch.rand(B, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1).to(dev)           # in [-bound, bound], bound = 1
    gy = (torch.randn(B, 16, generator=torch.Generator().manual_seed(2)) * 128).half().to(dev)
    enc.train(); mlp.train()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        enc.embeddings.grad = None
        mlp.weights.grad = None
        with torch.autocast("cuda", dtype=torch.half):
            feat = enc(x, bound=1)
            y = mlp(feat)
        y.backward(gy)

    for _ in range(3):
        step()
    evs = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    ms = ts[len(ts) // 2]
    res = {"backend": args.backend, "batch": B, "step_ms": ms, "samples_per_s": B / (ms * 1e-3),
           "grad_table_abs_sum": float(enc.embeddings.grad.float().abs().sum()), "grad_weights_abs_sum": float(mlp.weights.grad.float().abs().sum())}
    if args.dump:
        np.savez(args.dump, gw=mlp.weights.grad.float().cpu().numpy(), gt=enc.embeddings.grad.float().cpu().numpy()[:1 << 16])

    if args.backend == "ntx":
        # device time of each C-ABI call of the step, on resident tensors
        from nerf_texture_b200 import _lib as L
        table = enc.embeddings.detach().half().contiguous()
        x01 = ((x + 1) / 2).contiguous()
        w = mlp.weights.detach().half().contiguous()
        feat = torch.empty(B, 32, dtype=torch.half, device=dev)
        y = torch.empty(B, 16, dtype=torch.half, device=dev)
        hid = torch.empty(2, B, 64, dtype=torch.half, device=dev)
        dhid = torch.empty(2, B, 64, dtype=torch.half, device=dev)
        dfeat = torch.empty(B, 32, dtype=torch.half, device=dev)
        dw = torch.empty_like(w)
        dtab = torch.zeros_like(table)
        ws = torch.empty(L.lib().ntx_ffmlp_backward_workspace_bytes(32, 16, 64, 2), dtype=torch.uint8, device=dev)
        S, H = float(np.log2(enc.per_level_scale)), int(enc.base_resolution)

        def timed(fn, iters=15):
            for _ in range(2):
                fn()
            ev = []
            for _ in range(iters):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                ev.append((a, b))
            torch.cuda.synchronize()
            t = sorted(a.elapsed_time(b) for a, b in ev)
            return t[len(t) // 2] * 1e3

        calls = {
            "grid_fwd_us": lambda: L.call("ntx_grid_encode_forward", L.ptr(x01), L.ptr(table), L.ptr(enc.offsets), L.ptr(feat), B, 3, 2, 16, S, H, 0, None, 0, 1,
                                          L.F16, L.LAYOUT_BLC, L.stream()),
            "mlp_fwd_train_us": lambda: L.call("ntx_ffmlp_forward", L.ptr(feat), L.ptr(w), B, 32, 16, 64, 2, 0, 6, L.ptr(hid), L.ptr(y), L.stream()),
            "mlp_bwd_us": lambda: L.call("ntx_ffmlp_backward", L.ptr(gy), L.ptr(feat), L.ptr(w), L.ptr(hid), B, 32, 16, 64, 2, 0, 6, 1, L.ptr(dhid), L.ptr(dfeat), L.ptr(dw),
                                        L.ptr(ws), L.stream()),
            "grid_bwd_us": lambda: L.call("ntx_grid_encode_backward", L.ptr(dfeat), L.ptr(x01), L.ptr(table), L.ptr(enc.offsets), L.ptr(dtab), B, 3, 2, 16, S, H, 0, None,
                                          None, 0, 1, L.F16, L.LAYOUT_BLC, L.stream()),
            "table_grad_zero_fill_us": lambda: dtab.zero_(),
        }
        res["kernels"] = {k: timed(f) for k, f in calls.items()}
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    main()
