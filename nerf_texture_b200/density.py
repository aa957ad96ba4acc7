"""Density-grid maintenance on libntx: NeRFRenderer.update_extra_state (nerf/renderer.py:567-660) as one launch chain.

The reference queries `self.density()` (grid encoder -> sigma MLP, two extension calls + torch glue per cascade) on 128^3 cell
centres it builds with meshgrid / morton3D / index_put, then runs five more torch passes over the grid (EMA-max, clamp, mean,
`.item()`, packbits) — every 16 training steps and 50 times when a run starts.  `update_density_grid` hands the whole thing to
ntx_update_density_grid: the fused field kernel in density mode generates each cell's position from its Morton index (same fp32
arithmetic, one rounding per torch op), gathers, runs the sigma net and writes sigma * density_scale; two small kernels do the
EMA-max + mean and packbits reads the threshold from device memory.  `update_extra_state(model, ...)` is the drop-in for the
method: same arguments, same attributes read and written on the renderer object (density_grid, density_bitfield, mean_density,
iter_density, step_counter, mean_count, local_step).
"""
import numpy as np
import torch

from . import _lib as L
from .operators import _half_table

_ws_cache = {}


def _workspace(C, H, dev):
    key = (C, H, dev.index, L.stream())
    ws = _ws_cache.get(key)
    if ws is None:
        raw = torch.empty(L.lib().ntx_update_density_grid_workspace_bytes(C, H) + 256, dtype=torch.uint8, device=dev)
        ws = _ws_cache[key] = (raw, raw.data_ptr() + (-raw.data_ptr()) % 256)
    return ws[1]


def update_density_grid(density_grid, density_bitfield, bound, density_scale, density_thresh, encoder, sigma_net, decay=0.95, cells=None, noise=None,
                        force_full_grid=False):
    """density_grid [C, H^3] f32 and density_bitfield [C*H^3/8] u8 are updated IN PLACE; returns a device tensor [2] = (mean_density,
    threshold used).  encoder: GridEncoder (D=3, level_dim 2), sigma_net: FFMLP(2L, 16, 64, num_layers=2) — the network_ff.py sigma branch.
    cells: None (full update) or int32 [C, n] Morton indices; noise: None or f32 [C, n | H^3, 3] uniform in [0, 1)."""
    C, H3 = density_grid.shape
    H = int(round(H3 ** (1.0 / 3.0)))
    assert H ** 3 == H3, "density_grid must be [cascade, H^3]"
    dev = density_grid.device
    table = _half_table(encoder.embeddings)
    w = sigma_net.weights.detach().to(torch.half).contiguous()
    assert table.shape[1] == 2 and encoder.input_dim == 3, "fused density query needs a 3-D grid with level_dim == 2"
    assert w.numel() == 64 * (2 * encoder.num_levels + 64 + 16), "sigma net must be FFMLP(2L, 16, 64, num_layers=2)"
    n = 0
    if cells is not None:
        cells = cells.to(device=dev, dtype=torch.int32).contiguous()
        assert cells.dim() == 2 and cells.shape[0] == C
        n = cells.shape[1]
    if noise is not None:
        noise = noise.to(device=dev, dtype=torch.float32).contiguous()
        assert noise.shape == (C, n if cells is not None else H3, 3)
    stats = torch.empty(2, dtype=torch.float32, device=dev)
    L.call("ntx_update_density_grid", L.ptr(density_grid, torch.float32), L.ptr(density_bitfield, torch.uint8), C, H, float(bound), float(density_scale), float(decay),
           float(density_thresh), L.ptr(table), L.ptr(encoder.offsets, torch.int32), encoder.num_levels, float(np.log2(encoder.per_level_scale)),
           int(encoder.base_resolution), int(encoder.align_corners), L.ptr(w), None if cells is None else L.ptr(cells), n, None if noise is None else L.ptr(noise),
           int(bool(force_full_grid)), _workspace(C, H, dev), L.ptr(stats), L.stream())
    from .compat.raymarching import raymarching as _rm      # the bit-field changed behind torch's version counter
    _rm._mip_forget(density_bitfield)
    return stats


@torch.no_grad()
def update_extra_state(model, decay=0.95, S=128, force_full_update=False, force_full_grid=False, jitter=True):
    """Drop-in for NeRFRenderer.update_extra_state(self, decay, S, force_full_update, force_full_grid) (renderer.py:567) on a
    network_ff-type model (`encoder`, `sigma_net`).  `S` (the reference's meshgrid block size) has no meaning here: the kernel covers a
    cascade in one launch.  jitter=False queries the cell centres (deterministic; the reference always jitters)."""
    if not model.cuda_ray:
        return
    grid = model.density_grid
    C, H3 = grid.shape
    dev = grid.device
    cells = noise = None
    if model.iter_density < 16 or force_full_update:               # full update (renderer.py:578-602)
        if jitter:
            noise = torch.rand(C, H3, 3, device=dev)
    else:                                                          # partial update (renderer.py:603-625): H^3/4 random cells + H^3/4 draws
        n = H3 // 4                                                # (with replacement) from the cells that are occupied now
        rows = []
        for cas in range(C):
            idx = torch.randint(0, H3, (n,), device=dev, dtype=torch.int64)       # a uniform random cell == the Morton code of uniform random coords
            occ = torch.nonzero(grid[cas] > 0).squeeze(-1)
            if occ.shape[0] > 0:
                idx = torch.cat([idx, occ[torch.randint(0, occ.shape[0], (n,), device=dev)]])
            else:
                idx = torch.cat([idx, idx])                        # keep the list rectangular: re-query the same cells (idempotent but for jitter)
            rows.append(idx)
        cells = torch.stack(rows).to(torch.int32)
        if jitter:
            noise = torch.rand(C, cells.shape[1], 3, device=dev)
    stats = update_density_grid(grid, model.density_bitfield, model.bound, model.density_scale, model.density_thresh, model.encoder, model.sigma_net, decay=decay,
                                cells=cells, noise=noise, force_full_grid=force_full_grid)
    model.mean_density = stats[0].item()                           # the reference's `.item()` (renderer.py:642); one sync, after everything is queued
    model.iter_density += 1
    total_step = min(16, model.local_step)                         # step counter bookkeeping (renderer.py:649-653)
    if total_step > 0:
        model.mean_count = int(model.step_counter[:total_step, 0].sum().item() / total_step)
    model.local_step = 0
