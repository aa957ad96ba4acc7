"""Seeded inputs of the golden fixtures (shared by tests/golden/make_golden.py, which runs the reference CUDA on them, and by
tests/test_oracle_cpu.py, which runs the CPU oracle on them)."""
import numpy as np

from _util import ball_density_grid, cfgA, cfgB, oracle, pinhole_rays

GRID_CASES = {
    # BASELINE config 1: L=4, T=2^14, F=2, 4096 random points, fp32 table U(-1e-4, 1e-4) like GridEncoder.reset_parameters
    "cfgB_f32_init": dict(cfg=cfgB(), B=4096, dtype=np.float32, seed=0, table_range=1e-4, gridtype=0),
    # same geometry with a U(-1,1) table so that relative error is measurable, fp16 and fp32
    "cfgB_f32": dict(cfg=cfgB(), B=4096, dtype=np.float32, seed=1, table_range=1.0, gridtype=0),
    "cfgB_f16": dict(cfg=cfgB(), B=4096, dtype=np.float16, seed=2, table_range=1.0, gridtype=0),
    # the benchmark encoder (L=16, T=2^19), fp16, few points (the table is regenerated from the seed: 24 MB)
    "cfgA_f16": dict(cfg=cfgA(), B=512, dtype=np.float16, seed=3, table_range=1.0, gridtype=0),
    # tiled grid, 2-D, 4 features
    "tiled2d_f32": dict(cfg=dict(input_dim=2, num_levels=5, level_dim=4, per_level_scale=1.5, base_resolution=8, log2_hashmap_size=10, align_corners=False),
                        B=1024, dtype=np.float32, seed=4, table_range=1.0, gridtype=1),
}


def grid_case_inputs(case):
    O = oracle()
    cfg = case["cfg"]
    offsets, pls = O.grid_offsets(**{k: v for k, v in cfg.items() if k != "level_dim"})
    rng = np.random.default_rng(case["seed"])
    D = cfg["input_dim"]
    x = rng.random((case["B"], D), dtype=np.float32)
    x[:16, 0] = 1.5     # out-of-range rows
    x[16:32, D - 1] = -0.25
    x[32] = 1.0
    x[33] = 0.0
    emb = ((rng.random((int(offsets[-1]), cfg["level_dim"]), dtype=np.float32) * 2 - 1) * case["table_range"]).astype(case["dtype"])
    return x, emb, offsets, pls, cfg


MLP_CASES = {
    "sigma_net": dict(in_dim=32, hidden=64, layers=2, B=256, seed=0),
    "color_net": dict(in_dim=32, hidden=64, layers=3, B=256, seed=1),
    "w32": dict(in_dim=16, hidden=32, layers=2, B=128, seed=2),
}


def mlp_case_inputs(case):
    rng = np.random.default_rng(case["seed"])
    x = (rng.standard_normal((case["B"], case["in_dim"])) * 0.5).astype(np.float16)
    n = case["hidden"] * (case["in_dim"] + case["hidden"] * (case["layers"] - 1) + 16)
    w = ((rng.random(n, dtype=np.float32) * 2 - 1) * np.sqrt(3.0 / case["hidden"])).astype(np.float16)
    return x, w


def sh_inputs(degree):
    rng = np.random.default_rng(100 + degree)
    d = rng.standard_normal((512, 3)).astype(np.float32)
    d[:256] /= np.linalg.norm(d[:256], axis=1, keepdims=True)   # half on the sphere, half off it (the input is used as-is)
    d[256:] *= 0.6
    return d


def scene_inputs(name):
    O = oracle()
    if name == "c1":
        cascade, H, bound, dt_gamma, perturb, max_steps = 1, 128, 1.0, 0.0, 0, 256
    else:
        cascade, H, bound, dt_gamma, perturb, max_steps = 2, 64, 2.0, 1.0 / 128, 1, 128
    o, d = pinhole_rays(24, 32)
    grid = ball_density_grid(cascade, H, bound, 0.45, (0.1, -0.05, 0.0))
    bits = O.packbits(grid, 0.5)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    return dict(rays_o=o, rays_d=d, bits=bits, aabb=aabb, cascade=cascade, H=H, bound=bound, dt_gamma=dt_gamma, perturb=perturb, max_steps=max_steps)
