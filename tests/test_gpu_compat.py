"""GPU integration tests of the drop-in packages: the reference's own call sequences run on libntx.

  * the field of nerf/network_ff.py:85-101 assembled from gridencoder / ffmlp / shencoder modules under fp16 autocast equals the
    fused kernel bit for bit;
  * the inference loop of nerf/renderer.py:446-489 written against the `raymarching` functionals (same calls, same arguments)
    produces the same image as nerf_texture_b200.render.render_rays and as the CPU oracle's loop;
  * the tinycudann shim is self-consistent with the in-tree ops.
"""
import numpy as np
import pytest
import torch

from _util import ntx, oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _NetworkFF(torch.nn.Module):
    """the sigma/colour topology of nerf/network_ff.py:29-49, built from the drop-in modules exactly like the reference does"""

    def __init__(self, bound=1):
        super().__init__()
        from ffmlp import FFMLP
        from gridencoder import GridEncoder
        from shencoder import SHEncoder
        self.bound = bound
        self.encoder = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048 * bound,
                                   gridtype="hash", align_corners=True)
        self.sigma_net = FFMLP(input_dim=32, output_dim=16, hidden_dim=64, num_layers=2)
        self.encoder_dir = SHEncoder(input_dim=3, degree=4)
        self.color_net = FFMLP(input_dim=32, output_dim=3, hidden_dim=64, num_layers=3)

    def forward(self, x, d):
        x = self.encoder(x, bound=self.bound)
        h = self.sigma_net(x)
        sigma = torch.exp(h[..., 0].float())            # trunc_exp forward (tools/activation.py:6-10)
        geo_feat = h[..., 1:]
        d = self.encoder_dir(d)
        p = torch.zeros_like(geo_feat[..., :1])
        h = torch.cat([d, geo_feat, p], dim=-1)
        h = self.color_net(h)
        return sigma, torch.sigmoid(h), {}


def _model():
    torch.manual_seed(1)
    m = _NetworkFF().to(DEV).eval()
    m.encoder.embeddings.data.uniform_(-1, 1)
    g = torch.Generator().manual_seed(5)
    m.color_net.weights.data.copy_((torch.rand(m.color_net.weights.shape, generator=g) * 2 - 1) * np.sqrt(3 / 64))
    return m


def test_modular_field_equals_fused_field():
    ntx()
    from nerf_texture_b200 import render
    m = _model()
    field = render.NGPField.from_modules(m.encoder, m.sigma_net, m.color_net, bound=1.0)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(128 * 50 + 3, 3, generator=g) * 2 - 1).to(DEV)
    d = torch.randn(128 * 50 + 3, 3, generator=g)
    d = (d / d.norm(dim=1, keepdim=True)).to(DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.half):
        sigma, rgb, _ = m(x, d)
    fs, fr = field(x, d)
    torch.cuda.synchronize()
    assert sigma.dtype == torch.float32 and rgb.dtype == torch.half
    np.testing.assert_array_equal(fs.cpu().numpy(), sigma.cpu().numpy())
    np.testing.assert_array_equal(fr.cpu().numpy(), rgb.float().cpu().numpy())


def _run_cuda_inference(model, rays_o, rays_d, bitfield, cascade, grid_size, bound, min_near=0.2, dt_gamma=0, max_steps=1024, perturb=False, bg_color=1):
    """nerf/renderer.py:436-489 (inference branch of run_cuda), call for call"""
    import raymarching
    N = rays_o.shape[0]
    device = rays_o.device
    aabb = torch.tensor([-bound, -bound, -bound, bound, bound, bound], dtype=torch.float32, device=device)
    nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    dtype = torch.float32
    weights_sum = torch.zeros(N, dtype=dtype, device=device)
    depth = torch.zeros(N, dtype=dtype, device=device)
    image = torch.zeros(N, 3, dtype=dtype, device=device)
    n_alive = N
    alive_counter = torch.zeros([1], dtype=torch.int32, device=device)
    rays_alive = torch.zeros(2, n_alive, dtype=torch.int32, device=device)
    rays_t = torch.zeros(2, n_alive, dtype=dtype, device=device)
    step, i = 0, 0
    while step < max_steps:
        if step == 0:
            torch.arange(n_alive, out=rays_alive[0])
            rays_t[0] = nears
        else:
            alive_counter.zero_()
            raymarching.compact_rays(n_alive, rays_alive[i % 2], rays_alive[(i + 1) % 2], rays_t[i % 2], rays_t[(i + 1) % 2], alive_counter)
            n_alive = alive_counter.item()
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], rays_o, rays_d, bound, bitfield, cascade, grid_size, nears,
                                                    fars, 128, perturb, dt_gamma, max_steps)
        sigmas, rgbs, _ = model(xyzs, dirs)
        raymarching.composite_rays(n_alive, n_step, rays_alive[i % 2], rays_t[i % 2], sigmas, rgbs, deltas, weights_sum, depth, image)
        step += n_step
        i += 1
    image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
    return image, depth, weights_sum, i


def test_renderer_loop_on_dropin_ops_matches_fused_render_and_oracle():
    ntx()
    O = oracle()
    from nerf_texture_b200 import render, scene
    m = _model()
    rays_o, rays_d = scene.pinhole_rays(48, 48, DEV)
    bits = scene.ball_bitfield(1, 128, 1.0, DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.half):
        image, depth, ws, iters = _run_cuda_inference(m, rays_o, rays_d, bits, 1, 128, 1.0)
    field = render.NGPField.from_modules(m.encoder, m.sigma_net, m.color_net, bound=1.0)
    out = render.render_rays(field, rays_o, rays_d, bits, 1, 128, count_samples=True, schedule="reference")
    assert out["iterations"] == iters
    np.testing.assert_array_equal(out["image"].cpu().numpy(), image.cpu().numpy())
    np.testing.assert_array_equal(out["depth"].cpu().numpy(), depth.cpu().numpy())
    # CPU oracle of the whole loop
    sc = torch.empty(16, device=DEV)
    from nerf_texture_b200 import _lib as L
    L.call("ntx_grid_level_scales", field.S, field.H, 16, sc.data_ptr(), L.stream())
    img, dep, wsum, ns, it = O.render_rays(rays_o.cpu().numpy(), rays_d.cpu().numpy(), bits.cpu().numpy(), 1, 128, 1.0, field.table.cpu().numpy(),
                                           field.offsets.cpu().numpy(), float(2 ** field.S), field.H, field.w_sigma.cpu().numpy(), field.w_color.cpu().numpy(),
                                           level_scales=sc.cpu().numpy())
    assert ns == out["n_samples"] and it == iters
    assert np.abs(out["image"].cpu().numpy() - img).max() < 5e-3
    assert np.abs(out["weights_sum"].cpu().numpy() - wsum).max() < 5e-3


def test_training_functionals_through_dropin_api():
    """march_rays_train + composite_rays_train (forward + backward) through the `raymarching` package"""
    ntx()
    O = oracle()
    import raymarching
    from nerf_texture_b200 import scene
    rays_o, rays_d = scene.pinhole_rays(32, 32, DEV)
    bits = scene.ball_bitfield(1, 128, 1.0, DEV)
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=DEV)
    nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, 1.0, bits, 1, 128, nears, fars, counter, -1, False, 128, True, 0, 256)
    m = int(counter[0].item())
    assert xyzs.shape[0] % 128 == 0 and xyzs.shape[0] >= m > 1000
    wx, wd, wl, wrays, wcnt, _ = O.march_rays_train(rays_o.cpu().numpy(), rays_d.cpu().numpy(), 1.0, bits.cpu().numpy(), 1, 128, nears.cpu().numpy(), fars.cpu().numpy(),
                                                    1024 * 256, max_steps=256)
    np.testing.assert_array_equal(rays.cpu().numpy(), wrays)
    np.testing.assert_array_equal(xyzs[:m].cpu().numpy(), wx[:m])
    sig = (torch.rand(xyzs.shape[0], device=DEV) * 10).requires_grad_(True)
    rgb = torch.rand(xyzs.shape[0], 3, device=DEV).requires_grad_(True)
    ws, dep, img = raymarching.composite_rays_train(sig, rgb, deltas, rays)
    (img.sum() + ws.sum()).backward()
    gs, gc = O.composite_rays_train_backward(np.ones(1024, np.float32), np.ones((1024, 3), np.float32), sig.detach().cpu().numpy(), rgb.detach().cpu().numpy(),
                                             deltas.cpu().numpy(), wrays, ws.detach().cpu().numpy(), img.detach().cpu().numpy())
    np.testing.assert_allclose(sig.grad.cpu().numpy(), gs, rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(rgb.grad.cpu().numpy(), gc, rtol=2e-5, atol=2e-6)


def test_tinycudann_shim_self_consistency():
    ntx()
    import tinycudann as tcnn
    torch.manual_seed(0)
    net = tcnn.Network(41, 16, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 32, "n_hidden_layers": 1}).to(DEV)
    x = torch.randn(1000, 41, device=DEV) * 0.3
    y = net(x)
    assert y.shape == (1000, 16) and y.dtype == torch.half
    W = net.params.detach().half().float()
    W0, W1 = W[:32 * 48].view(32, 48), W[32 * 48:].view(16, 32)
    xp = torch.nn.functional.pad(x, (0, 7)).half().float()
    want = torch.relu(xp @ W0.T).half().float() @ W1.T
    assert (y.float() - want).abs().max() <= 4e-3 * want.abs().max() + 1e-3
    sh = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4}).to(DEV)
    d01 = torch.rand(64, 3, device=DEV)
    assert sh.n_output_dims == 16 and sh(d01).shape == (64, 16)
    hg = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": 4, "n_features_per_level": 2, "log2_hashmap_size": 12, "base_resolution": 16, "per_level_scale": 1.5}).to(DEV)
    assert hg.n_output_dims == 8 and hg(d01).shape == (64, 8)
    y.float().sum().backward()
    assert net.params.grad is not None and torch.isfinite(net.params.grad).all()


def test_tinycudann_shim_maps_bit_for_bit_onto_the_in_tree_ops():
    """tiny-cuda-nn itself is absent (un-vendored, un-pinned: SURVEY F2), so the shim cannot be pinned against it; what CAN be pinned
    is that it is nothing but the in-tree operators under tcnn's conventions: Network == FFMLP on the zero-padded input with the same
    flat weights, SphericalHarmonics == SHEncoder(2x - 1), HashGrid == GridEncoder(2x - 1, bound=1), bit for bit."""
    ntx()
    import tinycudann as tcnn
    from ffmlp import FFMLP
    from gridencoder import GridEncoder
    from shencoder import SHEncoder
    g = torch.Generator().manual_seed(4)
    x01 = torch.rand(3000, 3, generator=g).to(DEV)
    # network: 3 matmuls for n_hidden_layers = 2  <->  FFMLP(num_layers = 2)
    net = tcnn.Network(32, 16, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2}).to(DEV).eval()
    mlp = FFMLP(32, 16, 64, 2).to(DEV).eval()
    with torch.no_grad():
        mlp.weights.copy_(net.params)
    feat = (torch.rand(3000, 32, generator=g) - 0.5).to(DEV)
    with torch.no_grad():
        assert torch.equal(net(feat), mlp(feat))
    # a ragged input width is zero-padded to a multiple of 16
    net41 = tcnn.Network(41, 3, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 32, "n_hidden_layers": 1}).to(DEV).eval()
    f41 = (torch.rand(500, 41, generator=g) - 0.5).to(DEV)
    from nerf_texture_b200.operators import ffmlp_forward
    with torch.no_grad():
        want = ffmlp_forward(torch.nn.functional.pad(f41, (0, 7)), net41.params, 48, 16, 32, 1, 0, 3, True, False)[:, :3]
        assert torch.equal(net41(f41), want)
    # encodings
    sh = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4}).to(DEV)
    assert torch.equal(sh(x01), SHEncoder(3, 4).to(DEV)(x01 * 2 - 1).half())
    cfg = {"otype": "HashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 14, "base_resolution": 16, "per_level_scale": 1.5}
    hg = tcnn.Encoding(3, cfg).to(DEV)
    ge = GridEncoder(input_dim=3, num_levels=8, level_dim=2, per_level_scale=1.5, base_resolution=16, log2_hashmap_size=14, gridtype="hash", align_corners=False).to(DEV)
    with torch.no_grad():
        ge.embeddings.copy_(hg.enc.embeddings)
        assert torch.equal(hg(x01), ge(x01 * 2 - 1, bound=1).half())


def test_march_rays_train_differentiable_backward():
    """raymarching.py:232-288: the backward pads grad_xyzs and the saved sample parameters t to [N * max_steps] rows, views them as
    [N, max_steps, .] and returns grad_rays_o = sum_k g, grad_rays_d = sum_k g * t  (d xyz / d o = I, d xyz / d d = t, with t the parameter the marcher saved: the END of the sample's step).  The drop-in must
    return exactly that function of ITS forward's samples; t of every sample is recovered here from the returned positions
    (xyz = o + t d for the unclamped samples inside the ball) and the sums are formed in fp64."""
    ntx()
    import raymarching
    from nerf_texture_b200 import scene
    side, max_steps = 16, 64
    N = side * side
    rays_o, rays_d = scene.pinhole_rays(side, side, DEV)
    bits = scene.ball_bitfield(1, 128, 1.0, DEV)
    nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, torch.tensor([-1., -1, -1, 1, 1, 1], device=DEV), 0.2)
    o = rays_o.clone().requires_grad_(True)
    d = rays_d.clone().requires_grad_(True)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    xyzs, dirs, deltas, rays = raymarching.march_rays_train_differentiable(o, d, 1.0, bits, 1, 128, nears, fars, counter, -1, False, 128, True, 0.0, max_steps)
    M = int(counter[0].item())
    assert 0 < M <= xyzs.shape[0] <= N * max_steps
    w = torch.randn(xyzs.shape, generator=torch.Generator().manual_seed(9)).to(DEV)
    (w * xyzs).sum().backward()
    # t of every real sample, from its owner ray (rays = [ray id, offset, count])
    xn = xyzs.detach().cpu().numpy().astype(np.float64)
    on, dn = rays_o.cpu().numpy().astype(np.float64), rays_d.cpu().numpy().astype(np.float64)
    t = np.zeros(N * max_steps)
    dl = deltas.detach().cpu().numpy().astype(np.float64)
    for ray, off, cnt in rays.cpu().numpy():
        if cnt > 0 and off + cnt <= M:
            # the reference saves t AFTER the step (`t += dt; ...; rays_ts[0] = t`, raymarching.cu:655-658): position parameter + dt
            t[off:off + cnt] = ((xn[off:off + cnt] - on[ray]) * dn[ray]).sum(1) / (dn[ray] ** 2).sum() + dl[off:off + cnt, 0]
    g = np.zeros((N * max_steps, 3))
    g[:xn.shape[0]] = w.cpu().numpy().astype(np.float64)
    want_o = g.reshape(N, max_steps, 3).sum(1)
    want_d = (g * t[:, None]).reshape(N, max_steps, 3).sum(1)
    np.testing.assert_allclose(o.grad.cpu().numpy(), want_o, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(d.grad.cpu().numpy(), want_d, rtol=2e-3, atol=2e-3 * max(1.0, np.abs(want_d).max()))
    assert np.abs(want_d).max() > 1.0                                            # the check is not vacuous
