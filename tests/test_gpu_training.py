"""GPU parity of the training path (BASELINE config 5): fused-MLP backward (activation + weight gradients), hash-grid table
gradients through the drop-in modules' autograd, against fp64 references, the CPU oracle and the reference's own CUDA."""
import numpy as np
import pytest
import torch

from _util import ntx, oracle, ref, ulp16

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mlp_f64(x, w, din, hid, layers, g):
    """fp64 forward/backward of the bias-free ReLU MLP with the flat weight layout of ffmlp.cu:632"""
    X = x.astype(np.float64)
    mats, o = [], 0
    mats.append(w[o:o + hid * din].astype(np.float64).reshape(hid, din)); o += hid * din
    for _ in range(layers - 1):
        mats.append(w[o:o + hid * hid].astype(np.float64).reshape(hid, hid)); o += hid * hid
    mats.append(w[o:o + 16 * hid].astype(np.float64).reshape(16, hid))
    acts = [X]
    for Wm in mats[:-1]:
        acts.append(np.maximum(acts[-1] @ Wm.T, 0))
    out = acts[-1] @ mats[-1].T
    d = g.astype(np.float64)
    grads = [None] * len(mats)
    grads[-1] = d.T @ acts[-1]
    d = (d @ mats[-1]) * (acts[-1] > 0)
    for li in range(len(mats) - 2, -1, -1):
        grads[li] = d.T @ acts[li]
        if li > 0:
            d = (d @ mats[li]) * (acts[li] > 0)
    gin = d @ mats[0]
    return out, np.concatenate([gm.ravel() for gm in grads]), gin


@pytest.mark.parametrize("din,hid,layers,B", [(32, 64, 2, 4096), (32, 64, 3, 2048 + 100), (16, 32, 2, 1024), (64, 128, 2, 512), (32, 16, 3, 640)])
def test_ffmlp_backward(din, hid, layers, B):
    L_ = ntx()
    O = oracle()
    rng = np.random.default_rng(din + hid + layers)
    x = (rng.standard_normal((B, din)) * 0.5).astype(np.float16)
    w = ((rng.random(hid * (din + hid * (layers - 1) + 16), dtype=np.float32) * 2 - 1) * np.sqrt(3 / hid)).astype(np.float16)
    g = (rng.standard_normal((B, 16)) * 0.25).astype(np.float16)
    xt, wt, gt = (torch.from_numpy(a).to(DEV) for a in (x, w, g))
    out = torch.empty(B, 16, dtype=torch.half, device=DEV)
    fb = torch.empty(layers, B, hid, dtype=torch.half, device=DEV)
    L_.call("ntx_ffmlp_forward", xt.data_ptr(), wt.data_ptr(), B, din, 16, hid, layers, 0, 6, fb.data_ptr(), out.data_ptr(), L_.stream())
    bb = torch.full((layers, B, hid), float("nan"), dtype=torch.half, device=DEV)
    gi = torch.full((B, din), float("nan"), dtype=torch.half, device=DEV)
    gw = torch.full_like(wt, float("nan"))
    ws = torch.zeros(L_.lib().ntx_ffmlp_backward_workspace_bytes(din, 16, hid, layers), dtype=torch.uint8, device=DEV)
    L_.call("ntx_ffmlp_backward", gt.data_ptr(), xt.data_ptr(), wt.data_ptr(), fb.data_ptr(), B, din, 16, hid, layers, 0, 6, 1, bb.data_ptr(), gi.data_ptr(),
            gw.data_ptr(), ws.data_ptr(), L_.stream())
    torch.cuda.synchronize()
    gw_n, gi_n, bb_n = gw.cpu().numpy().astype(np.float64), gi.cpu().numpy().astype(np.float64), bb.cpu().numpy()
    assert np.isfinite(gw_n).all() and np.isfinite(gi_n).all() and np.isfinite(bb_n.astype(np.float32)).all()
    # oracle with the same rounding points, fed OUR forward buffer so that only the backward is compared
    ogw, ogi, obb = O.ffmlp_backward(g, x, w, fb.cpu().numpy(), din, 16, hid, layers, calc_grad_inputs=True)
    assert np.abs(bb_n.astype(np.float32) - obb.astype(np.float32)).max() <= 4 * ulp16(np.abs(obb.astype(np.float32)).max())
    assert np.abs(gi_n - ogi.astype(np.float64)).max() <= 4 * ulp16(np.abs(ogi.astype(np.float32)).max())
    sw = np.abs(ogw.astype(np.float64)).max()
    assert np.abs(gw_n - ogw.astype(np.float64)).max() <= 2e-3 * sw + 2 * ulp16(sw)
    # fp64 truth
    _, tgw, tgin = _mlp_f64(x, w, din, hid, layers, g)
    err_ours = np.abs(gw_n - tgw).max() / np.abs(tgw).max()
    assert err_ours < 6e-2, err_ours            # fp16 activations / ReLU masks vs a pure fp64 model; the bar that matters is the reference's own error below
    # per-sample input gradients: a ReLU whose fp16 pre-activation rounds across zero flips its mask w.r.t. the fp64 model, so only a
    # robust statistic is meaningful here (the exact check against the oracle with identical rounding points is above)
    gerr = np.abs(gi_n - tgin) / np.abs(tgin).max()
    assert np.median(gerr) < 1e-3 and np.mean(gerr > 3e-2) < 0.01, (np.median(gerr), np.mean(gerr > 3e-2))
    if B % 128 == 0 and hid >= 32:
        m = ref("ffmlp")
        m.allocate_splitk(layers + 1)
        rout = torch.empty(B, 16, dtype=torch.half, device=DEV)
        rfb = torch.empty(layers, B, hid, dtype=torch.half, device=DEV)
        m.ffmlp_forward(xt, wt, B, din, 16, hid, layers, 0, 6, rfb, rout)
        rbb = torch.zeros(layers, B, hid, dtype=torch.half, device=DEV)
        rgi = torch.zeros(B, din, dtype=torch.half, device=DEV)
        rgw = torch.zeros_like(wt)
        m.ffmlp_backward(gt, xt, wt, rfb, B, din, 16, hid, layers, 0, 6, True, rbb, rgi, rgw)
        torch.cuda.synchronize()
        err_ref = np.abs(rgw.cpu().numpy().astype(np.float64) - tgw).max() / np.abs(tgw).max()
        assert err_ours <= err_ref * 1.1 + 1e-3, (err_ours, err_ref)     # fp32 accumulation must not be worse than the reference's fp16 split-K


def test_modules_autograd_matches_fp32_torch_model():
    """network_ff's sigma branch (hash-grid -> FFMLP) in training mode under fp16 autocast: gradients of the drop-in modules vs an fp32
    torch model evaluating the same function (features gathered with the oracle-exact encoder, dense matmuls in fp32)."""
    ntx()
    from ffmlp import FFMLP
    from gridencoder import GridEncoder
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=3, num_levels=8, level_dim=2, base_resolution=16, log2_hashmap_size=15, desired_resolution=256, align_corners=True).to(DEV)
    enc.embeddings.data.uniform_(-1, 1)
    mlp = FFMLP(16, 16, 64, 2).to(DEV)
    enc.train(); mlp.train()
    B = 2 ** 14
    x = (torch.rand(B, 3, device=DEV) * 2 - 1)
    gout = torch.randn(B, 16, device=DEV)
    with torch.autocast("cuda", dtype=torch.half):
        feat = enc(x, bound=1)
        assert feat.dtype == torch.half and feat.shape == (B, 16)
        h = mlp(feat)
        loss = (h.float() * gout).sum() * 128.0            # GradScaler-like scaling
    feat.retain_grad()
    loss.backward()
    g_emb = enc.embeddings.grad.float() / 128.0
    g_w = mlp.weights.grad.float() / 128.0
    assert g_emb.shape == enc.embeddings.shape and g_w.shape == mlp.weights.shape
    assert torch.isfinite(g_emb).all() and torch.isfinite(g_w).all()

    # fp32 reference: same features (fp32 table path of the same kernel family is bit-checked elsewhere), dense MLP in fp32
    W = mlp.weights.detach().half().float()
    W0, W1, W2 = W[:64 * 16].view(64, 16), W[64 * 16:64 * 16 + 64 * 64].view(64, 64), W[64 * 16 + 64 * 64:].view(16, 64)
    f32 = enc.embeddings.detach().half().float().requires_grad_(True)
    from gridencoder import grid_encode
    feat32 = grid_encode((x + 1) / 2, f32, enc.offsets, enc.per_level_scale, enc.base_resolution, False, 0, True)
    W0r, W1r, W2r = (t.clone().requires_grad_(True) for t in (W0, W1, W2))
    h32 = torch.relu(torch.relu(feat32 @ W0r.T) @ W1r.T) @ W2r.T
    (h32 * gout).sum().backward(retain_graph=True)
    ref_w = torch.cat([W0r.grad.flatten(), W1r.grad.flatten(), W2r.grad.flatten()])
    rel_w = (g_w - ref_w).abs().max() / ref_w.abs().max()
    assert rel_w < 2e-2, rel_w
    # table gradient: isolate the encoder by pushing OUR dL/dfeat (fp16) through the fp32 instantiation of the same backward
    f32.grad = None
    feat32.backward(feat.grad.float() / 128.0)
    ref_e = f32.grad
    rel_e = (g_emb - ref_e).abs().max() / ref_e.abs().max()
    assert rel_e < 2e-2, rel_e                   # fp16 atomics (rounding after every add) vs fp32 atomics
    # untouched table rows get exactly zero gradient
    assert ((ref_e.abs().sum(1) == 0) == (g_emb.abs().sum(1) == 0)).float().mean() > 0.999
