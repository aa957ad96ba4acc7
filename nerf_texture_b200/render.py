"""Inference rendering on libntx: the loop of nerf/renderer.py::run_cuda (reference :446-489) with the fused field kernel.

`NGPField` packs the parameters of the network_ff topology (hash-grid + two FFMLPs) once — fp16 table, fp16 weights —
and evaluates sigma/rgb for a batch of samples with ONE kernel launch (ntx_ngp_field_forward).
`render_rays` is the reference's march -> field -> composite -> compact loop over the same C ABI the drop-in
`raymarching` package uses; buffers are allocated once per call, the only device->host traffic is the 4-byte alive
counter per iteration (the reference has the same sync, renderer.py:469).
`render_image_sharded` splits the rays of one frame over the ranks of a torch.distributed group in interleaved tiles and
all-gathers the packed result (rgb, depth, alpha = 20 B/ray) — the only collective on the path.
"""
import math

import numpy as np
import torch

from . import _lib as L


class NGPField:
    """hash-grid (D=3, C=2) -> FFMLP(2L,16,64,2) -> exp | SH4 ++ geo_feat -> FFMLP(32,3,64,3) -> sigmoid  (nerf/network_ff.py)"""

    def __init__(self, embeddings, offsets, per_level_scale, base_resolution, w_sigma, w_color, bound=1.0, align_corners=True,
                 density_scale=1.0):
        dev = embeddings.device
        self.table = embeddings.detach().to(torch.half).contiguous()           # [n_entries, 2] fp16, cast ONCE
        self.offsets = offsets.detach().to(device=dev, dtype=torch.int32).contiguous()
        self.num_levels = int(self.offsets.shape[0] - 1)
        self.S = float(np.log2(per_level_scale))
        self.H = int(base_resolution)
        self.align_corners = bool(align_corners)
        self.w_sigma = w_sigma.detach().to(device=dev, dtype=torch.half).contiguous()
        self.w_color = w_color.detach().to(device=dev, dtype=torch.half).contiguous()
        self.bound = float(bound)
        self.density_scale = float(density_scale)
        nfeat = 2 * self.num_levels
        assert self.table.shape[1] == 2, "fused field needs level_dim == 2"
        assert self.w_sigma.numel() == 64 * (nfeat + 64 + 16), "sigma net must be FFMLP(2L, 16, 64, num_layers=2)"
        assert self.w_color.numel() == 64 * (32 + 2 * 64 + 16), "colour net must be FFMLP(32, 3, 64, num_layers=3)"

    @classmethod
    def from_modules(cls, encoder, sigma_net, color_net, bound=1.0, density_scale=1.0):
        """encoder: gridencoder.GridEncoder, sigma_net / color_net: ffmlp.FFMLP (as built by nerf/network_ff.py:29-49)"""
        return cls(encoder.embeddings, encoder.offsets, encoder.per_level_scale, encoder.base_resolution, sigma_net.weights, color_net.weights,
                   bound=bound, align_corners=encoder.align_corners, density_scale=density_scale)

    @classmethod
    def random(cls, device, num_levels=16, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048, bound=1.0, seed=0,
               table_std=1.0, align_corners=True):
        """random-weight field of BASELINE config 2/3 (table U(-1,1) seed `seed`+1, MLP weights U(+-sqrt(3/64)) like ffmlp.py:141-144)"""
        pls = float(np.exp2(np.log2(desired_resolution * bound / base_resolution) / (num_levels - 1)))
        offs, off = [], 0
        for i in range(num_levels):
            res = int(np.ceil(base_resolution * pls ** i))
            n = min(2 ** log2_hashmap_size, (res if align_corners else res + 1) ** 3)
            n = int(np.ceil(n / 8) * 8)
            offs.append(off)
            off += n
        offs.append(off)
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        table = (torch.rand(off, 2, generator=g) * 2 - 1) * table_std
        g2 = torch.Generator(device="cpu").manual_seed(42)
        std = math.sqrt(3 / 64)
        ws = (torch.rand(64 * (2 * num_levels + 64 + 16), generator=g2) * 2 - 1) * std
        wc = (torch.rand(64 * (32 + 128 + 16), generator=g2) * 2 - 1) * std
        return cls(table.to(device), torch.tensor(offs, dtype=torch.int32, device=device), pls, base_resolution, ws.to(device), wc.to(device),
                   bound=bound, align_corners=align_corners)

    def __call__(self, xyzs, dirs, deltas=None, out_sigmas=None, out_rgbs=None, M=None):
        M = xyzs.shape[0] if M is None else M
        sig = out_sigmas if out_sigmas is not None else torch.empty(M, dtype=torch.float32, device=xyzs.device)
        rgb = out_rgbs if out_rgbs is not None else torch.empty(M, 3, dtype=torch.float32, device=xyzs.device)
        L.call("ntx_ngp_field_forward", L.ptr(xyzs), L.ptr(dirs), None if deltas is None else L.ptr(deltas), int(M), self.bound, L.ptr(self.table),
               L.ptr(self.offsets), self.num_levels, self.S, self.H, int(self.align_corners), L.ptr(self.w_sigma), L.ptr(self.w_color),
               self.density_scale, L.ptr(sig), L.ptr(rgb), L.stream())
        return sig, rgb


_frame_cache = {}


SCHEDULES = {
    # name: (sample budget per iteration as a multiple of N, max samples per ray per iteration)
    "reference": (1, 8),     # n_step = clamp(N // n_alive, 1, 8)   (renderer.py:464)
    "wide": (4, 32),         # same image, ~3x fewer loop iterations; evaluates (and discards) more samples behind opaque hits
}


WALK_BUDGET = 16   # empty voxels a ray may cross per march before it pauses (ntx_render_rays, walk_budget); 0 = never pause


def auto_schedule(N):
    """(32N, 256): up to 256 samples per ray and iteration, 32N sample rows of workspace (36 B each: 1.2 GB for a 1024^2 frame).
    Round-2 sweep (tools/schedule_sweep.py, profiles/r02_schedule_sweep.txt): 3 loop iterations instead of 7 with (8N, 64) and 43 with the
    reference's (N, 8); full frame 6.35 -> 6.23 ms, a 1/8 shard 1.36 -> 1.21 ms (every iteration costs a march launch whose length is
    a latency chain per ray, whatever the ray count).  The field kernel runs over the marcher's live-row list, so rows past a ray's
    end cost nothing."""
    return (32, 256)


def _frame_buffers(dev, N, max_steps, budget, mailbox_len):
    """device workspace + pinned host mailbox of ntx_render_rays, cached per (device, stream, N, budget)"""
    key = (dev.index, L.stream(), N, mailbox_len, budget)
    buf = _frame_cache.get(key)
    if buf is None:
        ws = torch.empty(L.lib().ntx_render_rays_workspace_bytes(N, budget) + 256, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        mailbox = torch.zeros(mailbox_len, dtype=torch.int32).pin_memory()
        counter = torch.zeros(1, dtype=torch.int64, device=dev)
        buf = (ws, ws.data_ptr() + off, mailbox, counter)
        # One frame size at a time is the normal case: do not hoard workspaces.  A workspace being evicted may still be in use by a
        # frame queued on ITS stream (the key holds the stream): wait for that device's work before the allocator may recycle it.
        if _frame_cache:
            torch.cuda.synchronize(dev)
            _frame_cache.clear()
        _frame_cache[key] = buf
    return buf


def _aabb_tensor(bound, dev, cache={}):
    key = (float(bound), dev.index)
    if key not in cache:
        cache[key] = torch.tensor([-bound, -bound, -bound, bound, bound, bound], dtype=torch.float32, device=dev)
    return cache[key]


def render_rays(field, rays_o, rays_d, density_bitfield, cascade, grid_size, aabb=None, min_near=0.2, dt_gamma=0.0, max_steps=1024, bg_color=1.0,
                perturb=0, count_samples=False, profile=None, use_mip=True, device_loop=True, mip=None, schedule="auto", time_kernels=False, block_rows=None, cache_mip=False, block_out=None):
    """Inference branch of NeRFRenderer.run_cuda (renderer.py:436-489).  rays_o/d [N,3] fp32 CUDA.
    Returns dict(image [N,3], depth [N], weights_sum [N], iterations, n_samples (if count_samples)).
    device_loop=True (default): the whole loop is ONE libntx call whose iteration state stays on the device (ntx_render_rays);
    device_loop=False: the reference's structure, one extension call per step and a blocking n_alive read per iteration
    (bit-identical results; kept for profiling and as the parity reference of the device-driven loop).
    profile: optional list (stepwise loop only); gets one (start_event, stop_event, live_sample_count_tensor) per field-kernel launch.
    time_kernels (device loop only): also return march_ms / field_ms, the summed CUDA-event times of the frame's march and field launches.
    mip: optional pre-built occupancy mip (ntx_build_occupancy_mip) of density_bitfield; cache_mip=True keeps the mip this call
    builds for as long as the bit-field tensor is unchanged (3 tiny kernels per frame otherwise: bench.py rebuilds it every frame).
    schedule (device loop only): "auto" (auto_schedule(N); the reference's when perturb != 0), "reference", "wide", or a
    (budget_multiple, max_n_step) pair — the image is the same either way, see include/ntx.h.
    block_rows (device loop only, used by the sharded path): return the raw planar result block [weights_sum | depth | rgb] padded to
    block_rows rays per plane (no background term yet) instead of the finished image."""
    dev = rays_o.device
    rays_o = rays_o.contiguous().view(-1, 3).float()
    rays_d = rays_d.contiguous().view(-1, 3).float()
    N = rays_o.shape[0]
    bound = field.bound
    if aabb is None:
        aabb = _aabb_tensor(bound, dev)
    st = L.stream()
    if use_mip and mip is None and grid_size >= 16 and (grid_size & (grid_size - 1)) == 0:
        if cache_mip:
            # built once per bit-field tensor and version counter (weak reference: see compat/raymarching), like the drop-in march_rays
            from .compat.raymarching.raymarching import _occupancy_mip
            hit = _occupancy_mip(density_bitfield, int(cascade), int(grid_size))
            mip = None if hit is None else hit[1]
        else:
            mip = torch.empty(L.lib().ntx_occupancy_mip_bytes(int(cascade), int(grid_size)), dtype=torch.uint8, device=dev)
            L.call("ntx_build_occupancy_mip", L.ptr(density_bitfield), int(cascade), int(grid_size), L.ptr(mip), st)
    if not use_mip:
        mip = None
    if device_loop and profile is None and N > 0:
        import ctypes
        walk = 0 if (schedule == "reference" or perturb) else WALK_BUDGET   # "reference" = the reference's exact iteration structure
        if schedule == "auto":
            schedule = "reference" if perturb else auto_schedule(N)
        mult, cap = SCHEDULES[schedule] if isinstance(schedule, str) else schedule
        budget = int(mult) * N
        ws, ws_ptr, mailbox, counter = _frame_buffers(dev, N, int(max_steps), budget, int(max_steps) + 2 * int(grid_size) * int(cascade) + 8)
        # the raw result as ONE planar block [weights_sum (n) | depth (n) | rgb (3n)]: what a rank sends in the sharded path, and what
        # the fused epilogue (ntx_unshard_frame: un-permute + background term) reads
        n_blk = N if block_rows is None else int(block_rows)
        block = torch.empty(5 * n_blk, dtype=torch.float32, device=dev) if block_out is None else block_out    # block_out: e.g. a slice of symmetric memory
        assert block.numel() == 5 * n_blk and block.dtype == torch.float32 and block.is_contiguous()
        wsum_c, depth_c, image_c = block[:N], block[n_blk:n_blk + N], block[2 * n_blk:2 * n_blk + 3 * N].view(N, 3)
        if count_samples:
            counter.zero_()
        stats = (ctypes.c_uint32 * 2)()
        kms = (ctypes.c_float * 2)()
        L.call("ntx_render_rays", L.ptr(rays_o), L.ptr(rays_d), N, L.ptr(aabb), float(min_near), float(bound), float(dt_gamma), int(max_steps), int(perturb),
               budget, int(cap), walk, int(cascade), int(grid_size), L.ptr(density_bitfield), None if mip is None else L.ptr(mip), L.ptr(field.table), L.ptr(field.offsets),
               field.num_levels, field.S, field.H, int(field.align_corners), L.ptr(field.w_sigma), L.ptr(field.w_color), float(field.density_scale),
               L.ptr(wsum_c), L.ptr(depth_c), L.ptr(image_c), ws_ptr, mailbox.data_ptr(), counter.data_ptr() if count_samples else None,
               ctypes.addressof(stats), ctypes.addressof(kms) if time_kernels else None, st)
        L.launches += int(stats[1]) - 1          # L.call counted the call as one launch
        if block_rows is not None:               # sharded frame: the caller all-gathers the block and assembles the image
            out = dict(block=block, iterations=int(stats[0]))
        elif isinstance(bg_color, (int, float)):
            image = torch.empty(N, 3, dtype=torch.float32, device=dev)
            depth = torch.empty(N, dtype=torch.float32, device=dev)
            wsum = torch.empty(N, dtype=torch.float32, device=dev)
            L.call("ntx_unshard_frame", L.ptr(block), 1, N, max(N, 1), N, float(bg_color), L.ptr(image), L.ptr(depth), L.ptr(wsum), st)   # image + (1 - weights_sum) * bg
            out = dict(image=image, depth=depth, weights_sum=wsum, iterations=int(stats[0]))
        else:                                    # per-ray background tensor (renderer.py:354-355)
            out = dict(image=image_c + (1 - wsum_c).unsqueeze(-1) * bg_color, depth=depth_c, weights_sum=wsum_c, iterations=int(stats[0]))
        if count_samples:
            out["n_samples"] = int(counter.item())
        if time_kernels:
            out["march_ms"], out["field_ms"] = float(kms[0]), float(kms[1])
        return out
    nears = torch.empty(N, dtype=torch.float32, device=dev)
    fars = torch.empty(N, dtype=torch.float32, device=dev)
    L.call("ntx_near_far_from_aabb", L.ptr(rays_o), L.ptr(rays_d), L.ptr(aabb), N, float(min_near), L.ptr(nears), L.ptr(fars), st)

    weights_sum = torch.zeros(N, dtype=torch.float32, device=dev)
    depth = torch.zeros(N, dtype=torch.float32, device=dev)
    image = torch.zeros(N, 3, dtype=torch.float32, device=dev)
    rays_alive = torch.empty(2, N, dtype=torch.int32, device=dev)
    rays_t = torch.empty(2, N, dtype=torch.float32, device=dev)
    torch.arange(N, out=rays_alive[0])
    rays_t[0].copy_(nears)
    alive_counter = torch.zeros(1, dtype=torch.int32, device=dev)
    Mmax = N + 128                      # n_alive * n_step <= N, padded to 128 like the reference (align=128)
    xyzs = torch.empty(Mmax, 3, dtype=torch.float32, device=dev)
    dirs = torch.empty(Mmax, 3, dtype=torch.float32, device=dev)
    deltas = torch.empty(Mmax, 2, dtype=torch.float32, device=dev)
    sigmas = torch.empty(Mmax, dtype=torch.float32, device=dev)
    rgbs = torch.empty(Mmax, 3, dtype=torch.float32, device=dev)
    ws = L.workspace("compact", L.lib().ntx_compact_rays_workspace_bytes(N), dev)
    n_samples = torch.zeros(1, dtype=torch.int64, device=dev) if count_samples else None

    n_alive, step, i = N, 0, 0
    while step < max_steps:
        cur, old = i % 2, (i + 1) % 2
        if step > 0:
            alive_counter.zero_()
            L.call("ntx_compact_rays", n_alive, L.ptr(rays_alive[cur]), L.ptr(rays_alive[old]), L.ptr(rays_t[cur]), L.ptr(rays_t[old]),
                   L.ptr(alive_counter), L.ptr(ws), st)
            n_alive = int(alive_counter.item())     # the loop's one D2H sync (renderer.py:469)
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        M = n_alive * n_step
        M += 128 - (M % 128)
        L.call("ntx_march_rays", n_alive, n_step, L.ptr(rays_alive[cur]), L.ptr(rays_t[cur]), L.ptr(rays_o), L.ptr(rays_d), bound, float(dt_gamma),
               int(max_steps), int(cascade), int(grid_size), L.ptr(density_bitfield), L.ptr(nears), L.ptr(fars), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas),
               int(perturb), 1, M, None if mip is None else L.ptr(mip), st)
        if profile is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            field(xyzs, dirs, deltas, out_sigmas=sigmas, out_rgbs=rgbs, M=M)
            e1.record()
            profile.append((e0, e1, (deltas[:M, 0] > 0).sum(), M))
        else:
            field(xyzs, dirs, deltas, out_sigmas=sigmas, out_rgbs=rgbs, M=M)
        if count_samples:
            n_samples += (deltas[:M, 0] > 0).sum()
        L.call("ntx_composite_rays", n_alive, n_step, L.ptr(rays_alive[cur]), L.ptr(rays_t[cur]), L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas),
               L.ptr(weights_sum), L.ptr(depth), L.ptr(image), st)
        step += n_step
        i += 1
    image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
    out = dict(image=image, depth=depth, weights_sum=weights_sum, iterations=i)
    if count_samples:
        out["n_samples"] = int(n_samples.item())
    return out


# ------------------------------------------------------------------------------------------------ multi-GPU
def shard_indices(N, world_size, rank, tile=1024):
    """Interleaved ray tiles: tile k goes to rank k % world_size.  Per-ray cost varies >100x between rays that miss and rays
    that cross the object; contiguous blocks would put the object on a few ranks only."""
    ntiles = (N + tile - 1) // tile
    mine = torch.arange(rank, ntiles, world_size)
    idx = (mine[:, None] * tile + torch.arange(tile)[None, :]).reshape(-1)
    return idx[idx < N]


_shard_cache = {}


def _shard_plan(N, world, tile, device):
    """per-rank ray indices (on `device`), rows of the padded per-rank block, and the inverse permutation of the gathered rows"""
    key = (N, world, tile, str(device))
    plan = _shard_cache.get(key)
    if plan is None:
        idxs = [shard_indices(N, world, r, tile) for r in range(world)]
        n_max = max(i.numel() for i in idxs)
        # gathered row (r * n_max + j) holds ray idxs[r][j]  ->  dest[ray] = gathered row
        src = torch.cat([r * n_max + torch.arange(i.numel()) for r, i in enumerate(idxs)])
        inv = torch.empty(N, dtype=torch.long)
        inv[torch.cat(idxs)] = src
        plan = ([i.to(device) for i in idxs], n_max, inv.to(device))
        _shard_cache[key] = plan
    return plan


class PeerFrameExchange:
    """The sharded frame's exchange step over NVLink peer memory instead of a collective ("fused op <-> collective adjacency").
    Every rank renders its shard straight into its slice of a SYMMETRIC buffer (torch.distributed._symmetric_memory: each rank's
    allocation is mapped into every other rank's address space over NVLink / NVSwitch); after one device-side cross-rank barrier
    `ntx_unshard_frame_peers` reads all ranks' blocks in place while it puts every ray at its image position and adds the background
    term: exchange + un-permute + epilogue in ONE kernel, no all-gather, no staging copy.  Blocks are double-buffered: a rank may only
    overwrite a block two frames later, after everybody has passed the next frame's barrier.
    `PeerFrameExchange.create(...)` returns None where symmetric memory is not available (then gather_frame's NCCL all-gather is used)."""

    def __init__(self, hdl, buf, n_max, world, rank, tile, N):
        self.hdl, self.buf, self.n_max, self.world, self.rank, self.tile, self.N, self.frame = hdl, buf, n_max, world, rank, tile, N, 0

    @classmethod
    def create(cls, N, group=None, tile=1024, device=None):
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = group or dist.group.WORLD
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            n_max = _shard_plan(N, world, tile, torch.device("cpu"))[1]
            if hasattr(symm_mem, "enable_symm_mem_for_group"):
                try:
                    symm_mem.enable_symm_mem_for_group(group.group_name)
                except Exception:
                    pass
            buf = symm_mem.empty(2 * 5 * n_max, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(buf, group)
            if hdl.world_size != world or not hdl.buffer_ptrs_dev:
                return None
            return cls(hdl, buf, n_max, world, rank, tile, N)
        except Exception:
            return None

    def block(self):
        """this frame's send block of this rank (render_rays(..., block_rows=n_max, block_out=exchange.block()))"""
        h = self.frame & 1
        return self.buf[h * 5 * self.n_max:(h + 1) * 5 * self.n_max]

    def assemble(self, bg_color=1.0):
        """barrier (all blocks of this frame written) + the fused exchange / un-permute / background kernel; returns image, depth, weights_sum"""
        dev = self.buf.device
        h = self.frame & 1
        self.hdl.barrier(channel=h)
        image = torch.empty(self.N, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(self.N, dtype=torch.float32, device=dev)
        wsum = torch.empty(self.N, dtype=torch.float32, device=dev)
        L.call("ntx_unshard_frame_peers", int(self.hdl.buffer_ptrs_dev), h * 5 * self.n_max, self.world, self.n_max, self.tile, self.N, float(bg_color),
               L.ptr(image), L.ptr(depth), L.ptr(wsum), L.stream())
        self.frame += 1
        return dict(image=image, depth=depth, weights_sum=wsum)


def gather_frame(out, N, group=None, tile=1024, bg_color=1.0):
    """ONE all_gather of the per-rank planar result blocks (weights_sum | depth | rgb = 20 B/ray) + one kernel that puts every ray
    back at its image position and adds the background term (ntx_unshard_frame).
    out: render_rays(..., block_rows=n_max) (raw block, background added here) or a finished per-rank result dict (packed here)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    idxs, n_max, _ = _shard_plan(N, world, tile, torch.device("cpu"))
    if "block" in out:
        block = out["block"]                                            # rendered straight into the send layout
    else:                                                               # finished result (its image already holds the background term)
        n_local = idxs[rank].numel()
        block = torch.zeros(5 * n_max, dtype=torch.float32, device=out["image"].device)
        block[:n_local] = out["weights_sum"]
        block[n_max:n_max + n_local] = out["depth"]
        block[2 * n_max:2 * n_max + 3 * n_local] = out["image"].reshape(-1)
        bg_color = 0.0
    dev = block.device
    gathered = torch.empty(world * 5 * n_max, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, block, group=group)           # the path's only collective (NCCL over NVLink)
    if dev.type == "cuda":
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        wsum = torch.empty(N, dtype=torch.float32, device=dev)
        L.call("ntx_unshard_frame", L.ptr(gathered), world, n_max, tile, N, float(bg_color), L.ptr(image), L.ptr(depth), L.ptr(wsum), L.stream())
    else:                                                               # the same index rule on the host (gloo tests)
        ray = torch.arange(N)
        k = ray // tile
        base, j = (k % world) * 5 * n_max, (k // world) * tile + ray % tile
        wsum, depth = gathered[base + j], gathered[base + n_max + j]
        image = gathered[(base + 2 * n_max + 3 * j)[:, None] + torch.arange(3)[None, :]] + ((1 - wsum) * float(bg_color))[:, None]
    res = dict(image=image, depth=depth, weights_sum=wsum, iterations=out["iterations"])
    if "n_samples" in out:
        res["n_samples"] = out["n_samples"]
    return res


def render_image_sharded(field, rays_o, rays_d, density_bitfield, cascade, grid_size, group=None, tile=1024, bg_color=1.0, exchange=None, **kw):
    """Each rank renders its interleaved tiles of the frame straight into its planar send block; then either ONE all_gather and one
    assembly kernel (NCCL), or — with `exchange=PeerFrameExchange.create(N, group)` — one cross-rank barrier and one kernel that reads
    the peers' blocks over NVLink while it assembles the image.  rays_o/rays_d: the full frame's rays on every rank ([N,3])."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return render_rays(field, rays_o, rays_d, density_bitfield, cascade, grid_size, bg_color=bg_color, **kw)
    N = rays_o.shape[0]
    idxs, n_max, _ = _shard_plan(N, dist.get_world_size(group), tile, rays_o.device)
    idx = idxs[dist.get_rank(group)]
    if exchange is not None:
        out = render_rays(field, rays_o[idx], rays_d[idx], density_bitfield, cascade, grid_size, block_rows=n_max, block_out=exchange.block(), **kw)
        res = exchange.assemble(bg_color)
        res["iterations"] = out["iterations"]
        return res
    out = render_rays(field, rays_o[idx], rays_d[idx], density_bitfield, cascade, grid_size, block_rows=n_max, **kw)
    return gather_frame(out, N, group, tile, bg_color=bg_color)
