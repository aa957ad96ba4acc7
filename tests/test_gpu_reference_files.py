"""The reference's OWN files — nerf/renderer.py, nerf/network_ff.py, tools/encoding.py, tools/activation.py, byte-for-byte as staged
by tools/stage_reference.py into baseline/_ref/callers/ — run on the drop-in packages (nerf_texture_b200/compat -> libntx) and,
as the checker, on the reference's own wrappers + its own CUDA rebuilt for sm_100a (oracle/_ref).  north_star: "nerf/renderer.py
and network_curvedfield.py call the new kernels unmodified" (network_curvedfield.py needs six absent third-party packages,
SURVEY 8c; network_ff.py is the same call pattern on the same four packages)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "run_reference_files.py")
STAGED = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "callers", "nerf", "renderer.py"))
REF_CUDA = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "_ref_raymarching.so")) and os.path.exists(
    os.path.join(ROOT, "baseline", "_ref", "wrappers", "gridencoder", "grid.py"))


def _run(backend, size, out, *extra):
    r = subprocess.run([sys.executable, TOOL, "--backend", backend, "--size", str(size), "--out", out] + list(extra), capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, "run_reference_files.py --backend %s failed:\n%s\n%s" % (backend, r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(line[0][7:]), np.load(out)


@pytest.mark.skipif(not STAGED, reason="reference files not staged (tools/stage_reference.py needs /root/reference)")
def test_reference_callers_import_over_compat_cpu():
    """no GPU: the unmodified renderer / network import and construct on top of the compat packages"""
    code = ("import sys; sys.path.insert(0, %r); import run_reference_files as R; Net, rm = R.import_reference_network('ntx'); "
            "m = Net(encoding='hashgrid', bound=1, cuda_ray=True); import nerf.renderer as rr; "
            "assert 'baseline/_ref/callers' in rr.__file__.replace(chr(92), '/'); assert 'nerf_texture_b200/compat' in rm.__file__.replace(chr(92), '/'); "
            "assert type(m.encoder).__module__ == 'nerf_texture_b200.operators'; print('OK')") % os.path.join(ROOT, "tools")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.skipif(not STAGED, reason="reference files not staged")
def test_unmodified_renderer_on_libntx_matches_reference_stack(tmp_path):
    """NeRFRenderer.render -> run_cuda (renderer.py:338-489) at 256x256: same image, depth and loop structure on both stacks"""
    info_n, img_n = _run("ntx", 256, str(tmp_path / "ntx.npz"))
    assert "baseline/_ref/callers/nerf/renderer.py" in info_n["renderer"].replace("\\", "/")
    assert "nerf_texture_b200/compat" in info_n["raymarching"].replace("\\", "/")
    assert info_n["iterations"] > 8 and info_n["samples"] > 100000
    image = img_n["image"].reshape(-1, 3)
    assert np.isfinite(image).all() and image.min() >= -1e-6 and image.max() <= 1 + 1e-4
    hit = (1 - image).max(axis=1) > 1e-3                    # white background, random-colour object
    assert 0.05 < hit.mean() < 0.6
    if not REF_CUDA:
        pytest.skip("oracle/_ref (reference CUDA) not built: product arm checked alone")
    ckpt = str(tmp_path / "reference_stack.pth")
    info_r, img_r = _run("ref", 256, str(tmp_path / "ref.npz"), "--save-ckpt", ckpt)
    assert "baseline/_ref/wrappers" in info_r["raymarching"].replace("\\", "/")
    assert info_n["iterations"] == info_r["iterations"], (info_n, info_r)          # the reference's own loop structure (n_step = N // n_alive ...)
    assert info_n["samples"] == info_r["samples"], (info_n, info_r)                # same samples marched (integer, exact)
    d_img = np.abs(img_n["image"] - img_r["image"])
    d_dep = np.abs(img_n["depth"] - img_r["depth"])
    # tolerance: north_star's 1e-4 on the composited fp32 image (fp16 MLP outputs differ by ulps between fp32 and fp16 accumulation)
    assert d_img.max() <= 1e-4, "image differs from the reference stack: max %g mean %g" % (d_img.max(), d_img.mean())
    assert d_dep.max() <= 1e-4 * max(1.0, float(np.abs(img_r["depth"]).max())), "depth differs: max %g" % d_dep.max()
    # checkpoint round trip: the .pth the REFERENCE stack's model wrote loads into the drop-in modules (strict key / shape
    # match, Trainer.load_checkpoint's code path) and renders the same frame
    info_c, img_c = _run("ntx", 256, str(tmp_path / "ntx_ckpt.npz"), "--load-ckpt", ckpt)
    assert info_c["iterations"] == info_r["iterations"] and info_c["samples"] == info_r["samples"]
    np.testing.assert_array_equal(img_c["image"], img_n["image"])          # same parameters as the seeded build, bit for bit
    assert np.abs(img_c["image"] - img_r["image"]).max() <= 1e-4
