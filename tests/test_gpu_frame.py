"""GPU parity of the device-driven frame (ntx_render_rays: loop state on the device, no blocking n_alive read) against the
stepwise loop that mirrors NeRFRenderer.run_cuda call by call (nerf/renderer.py:436-489).  Both run the same kernels on the
same data in the same order, so everything must agree to the last bit — including the iteration count and the number of
samples marched — for every way the loop can end (all rays dead, step budget exhausted) and for ragged ray counts."""
import numpy as np
import pytest
import torch

from _util import ntx

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scene(hw, cascade, grid_size, bound, seed=0):
    from nerf_texture_b200 import render, scene
    field = render.NGPField.random(torch.device(DEV), bound=bound, seed=seed)
    rays_o, rays_d = scene.pinhole_rays(hw[0], hw[1], DEV, radius=2.5 * bound)
    bits = scene.ball_bitfield(cascade, grid_size, bound, DEV, radius=0.5 * bound)
    return render, field, rays_o, rays_d, bits


@pytest.mark.parametrize("hw,cascade,grid_size,bound,kw", [
    ((64, 64), 1, 128, 1.0, dict()),
    ((57, 41), 1, 128, 1.0, dict(perturb=7)),                       # ragged ray count, perturbed start
    ((64, 64), 2, 64, 2.0, dict(dt_gamma=1.0 / 128)),               # two cascades, growing steps
    ((64, 64), 1, 128, 1.0, dict(max_steps=24)),                    # the step budget ends the loop, rays still alive
    ((64, 64), 1, 128, 1.0, dict(use_mip=False)),
    ((3, 5), 1, 128, 1.0, dict()),                                  # fewer rays than one block
], ids=["plain", "ragged_perturb", "cascades_dtgamma", "step_budget", "no_mip", "tiny"])
def test_device_driven_frame_equals_stepwise_loop(hw, cascade, grid_size, bound, kw):
    ntx()
    render, field, rays_o, rays_d, bits = _scene(hw, cascade, grid_size, bound)
    a = render.render_rays(field, rays_o, rays_d, bits, cascade, grid_size, count_samples=True, device_loop=False, **kw)
    b = render.render_rays(field, rays_o, rays_d, bits, cascade, grid_size, count_samples=True, device_loop=True, schedule="reference", **kw)
    assert a["iterations"] == b["iterations"] and a["iterations"] > 0
    assert a["n_samples"] == b["n_samples"] and a["n_samples"] > 0
    for k in ("image", "depth", "weights_sum"):
        np.testing.assert_array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), err_msg=k)


@pytest.mark.parametrize("hw,cascade,grid_size,bound,kw", [
    ((96, 96), 1, 128, 1.0, dict()),
    ((57, 41), 1, 128, 1.0, dict(min_near=0.05)),
    ((64, 64), 2, 64, 2.0, dict(dt_gamma=1.0 / 128)),
], ids=["plain", "ragged", "cascades_dtgamma"])
@pytest.mark.parametrize("schedule", ["wide", "auto", (1, 8), (2, 13), (8, 64)])
def test_image_does_not_depend_on_the_sample_schedule(hw, cascade, grid_size, bound, kw, schedule):
    """n_step = clamp(budget // n_alive, 1, cap) only decides how a ray's samples are spread over loop iterations: each ray
    still composites the same samples in the same order and stops at the same one.  (More than 8 samples per ray and
    iteration also exercises the chunked path of the marcher.)  perturb must be 0: the reference jitters t once per
    march_rays CALL (raymarching.cu:1011), so with perturb the schedule is part of the random realisation."""
    ntx()
    render, field, rays_o, rays_d, bits = _scene(hw, cascade, grid_size, bound, seed=1)
    a = render.render_rays(field, rays_o, rays_d, bits, cascade, grid_size, count_samples=True, schedule="reference", **kw)
    b = render.render_rays(field, rays_o, rays_d, bits, cascade, grid_size, count_samples=True, schedule=schedule, **kw)
    if schedule != (1, 8):                            # (1, 8) = the reference's n_step rule, only the walk budget differs
        assert b["iterations"] < a["iterations"]
    assert b["n_samples"] >= a["n_samples"]          # samples behind an opaque hit are marched (and discarded) in bigger chunks
    for k in ("image", "depth", "weights_sum"):
        np.testing.assert_array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), err_msg=k)


def test_device_driven_frame_is_repeatable_and_reuses_its_workspace():
    ntx()
    render, field, rays_o, rays_d, bits = _scene((96, 96), 1, 128, 1.0, seed=3)
    outs = [render.render_rays(field, rays_o, rays_d, bits, 1, 128, count_samples=True) for _ in range(5)]
    for o in outs[1:]:
        assert o["iterations"] == outs[0]["iterations"] and o["n_samples"] == outs[0]["n_samples"]
        assert torch.equal(o["image"], outs[0]["image"]) and torch.equal(o["depth"], outs[0]["depth"])
    # a frame with no occupied voxel at all: nothing is marched, the image is the background
    empty = torch.zeros_like(bits)
    o = render.render_rays(field, rays_o, rays_d, empty, 1, 128, count_samples=True)
    assert o["n_samples"] == 0 and o["iterations"] <= 1
    assert torch.equal(o["image"], torch.ones_like(o["image"])) and float(o["weights_sum"].abs().max()) == 0.0


def test_render_rays_rejects_missing_workspace_and_mailbox():
    L = ntx()
    import ctypes
    it = (ctypes.c_uint32 * 2)()
    d = torch.zeros(16, device=DEV)
    args = [d.data_ptr(), d.data_ptr(), 1, d.data_ptr(), 0.2, 1.0, 0.0, 8, 0, 0, 0, 0, 1, 128, d.data_ptr(), None, d.data_ptr(), d.data_ptr(), 16, 0.5, 16, 1, d.data_ptr(),
            d.data_ptr(), 1.0, d.data_ptr(), d.data_ptr(), d.data_ptr()]
    with pytest.raises(RuntimeError, match="workspace"):
        L.call("ntx_render_rays", *args, None, None, None, ctypes.addressof(it), None, L.stream())
    ws = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    with pytest.raises(RuntimeError, match="mailbox"):
        L.call("ntx_render_rays", *args, ws.data_ptr() + (-ws.data_ptr()) % 256, None, None, ctypes.addressof(it), None, L.stream())


def test_paused_rays_cross_a_long_gap_between_two_objects():
    """Two small balls on the camera axis, 1.3 units (83 voxels) apart: after the first ball every ray walks a long empty stretch.
    With a walk budget the marcher pauses such rays several iterations in a row; they must still reach the second ball with the
    reference's samples, and the loop must not run out of its step budget on the way (max_steps kept small on purpose)."""
    L = ntx()
    import math
    from _util import ball_density_grid
    from nerf_texture_b200 import render, scene
    field = render.NGPField.random(torch.device(DEV), seed=5)
    rays_o, rays_d = scene.pinhole_rays(48, 48, DEV, fovy_deg=20.0)
    az, el = math.radians(30.0), math.radians(20.0)
    fwd = -np.array([math.cos(el) * math.sin(az), math.sin(el), math.cos(el) * math.cos(az)])
    grid = np.maximum(ball_density_grid(1, 128, 1.0, 0.12, tuple(-0.65 * fwd)), ball_density_grid(1, 128, 1.0, 0.12, tuple(0.65 * fwd)))
    dens = torch.from_numpy(grid.reshape(-1)).to(DEV)
    bits = torch.zeros(128 ** 3 // 8, dtype=torch.uint8, device=DEV)
    L.call("ntx_packbits", dens.data_ptr(), dens.numel() // 8, 0.5, bits.data_ptr(), L.stream())   # N counts bit-field bytes (raymarching.py:204)
    ref = render.render_rays(field, rays_o, rays_d, bits, 1, 128, count_samples=True, schedule="reference", max_steps=256)
    assert ref["n_samples"] > 0
    for sched in [(1, 8), "auto", (2, 3)]:
        for wb in (4, 16):
            render.WALK_BUDGET, old = wb, render.WALK_BUDGET
            try:
                out = render.render_rays(field, rays_o, rays_d, bits, 1, 128, count_samples=True, schedule=sched, max_steps=256)
            finally:
                render.WALK_BUDGET = old
            assert out["n_samples"] == ref["n_samples"], (sched, wb)
            for k in ("image", "depth", "weights_sum"):
                np.testing.assert_array_equal(out[k].cpu().numpy(), ref[k].cpu().numpy(), err_msg="%s %s %d" % (k, sched, wb))
    # both balls are really seen: some rays accumulate samples from two separate stretches
    assert float(ref["weights_sum"].max()) > 0


@pytest.mark.parametrize("world,tile,hw", [(2, 1024, (96, 96)), (8, 64, (57, 41)), (4, 1024, (64, 64))], ids=["2x1024", "8x64_ragged", "4x1024_fewer_tiles_than_ranks"])
def test_sharded_frame_assembly_equals_whole_frame(world, tile, hw):
    """BASELINE config 4 on one GPU: every rank's shard (interleaved tiles, render.shard_indices) is rendered straight into its planar
    send block, the blocks are concatenated the way all_gather_into_tensor lays them out, and ntx_unshard_frame (un-permute + background
    term) must reproduce the whole-frame render bit for bit — rays are independent, so sharding cannot change a pixel."""
    L_ = ntx()
    render, field, rays_o, rays_d, bits = _scene(hw, 1, 128, 1.0)
    N = rays_o.shape[0]
    whole = render.render_rays(field, rays_o, rays_d, bits, 1, 128, bg_color=0.25)
    idxs, n_max, _ = render._shard_plan(N, world, tile, torch.device(DEV))
    blocks = []
    for r in range(world):
        idx = idxs[r]
        if idx.numel() == 0:                                  # more ranks than tiles: an empty shard sends zeros
            blocks.append(torch.zeros(5 * n_max, device=DEV))
            continue
        out = render.render_rays(field, rays_o[idx].contiguous(), rays_d[idx].contiguous(), bits, 1, 128, block_rows=n_max)
        assert out["block"].numel() == 5 * n_max
        blocks.append(out["block"])
    gathered = torch.cat(blocks).contiguous()
    image = torch.empty(N, 3, device=DEV); depth = torch.empty(N, device=DEV); wsum = torch.empty(N, device=DEV)
    L_.call("ntx_unshard_frame", L_.ptr(gathered), world, n_max, tile, N, 0.25, L_.ptr(image), L_.ptr(depth), L_.ptr(wsum), L_.stream())
    torch.cuda.synchronize()
    for got, k in ((image, "image"), (depth, "depth"), (wsum, "weights_sum")):
        np.testing.assert_array_equal(got.cpu().numpy(), whole[k].cpu().numpy(), err_msg=k)
