"""numpy front-end of the CPU oracle (oracle/ntx_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this.
Array layouts follow the reference's native entry points (e.g. grid outputs are [L,B,C], gridencoder.cu:362),
the helpers at the bottom re-express them in the layouts the Python modules return.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libntx_oracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("ntx_oracle.c", "ntx_oracle_mesh.c", "grid_impl.inc", "sh_table.inc")]
    if (not force) and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    r = subprocess.run(["make", "-C", _HERE, "libntx_oracle.so"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _SO


def lib():
    global _lib
    if _lib is None:
        try:
            if not os.path.exists(_SO):
                build()
            _lib = C.CDLL(_SO)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


_SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float16): "f16", np.dtype(np.float64): "f64"}


def num_threads():
    return int(lib().orc_num_threads())


# ------------------------------------------------------------------------------------------------ grid
def grid_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, align_corners=False):
    """Level table exactly as GridEncoder.__init__ builds it (gridencoder/grid.py:98-129)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), float(per_level_scale)


def grid_level_scales(S, H, L):
    out = np.empty(L, np.float32)
    lib().orc_grid_level_scales(C.c_float(S), C.c_uint32(H), C.c_uint32(L), _p(out))
    return out


def grid_indices(inputs, offsets, level, scale, gridtype=0, align_corners=False):
    inputs = _c(inputs, np.float32)
    B, D = inputs.shape
    out = np.empty((B, 1 << D), np.uint32)
    lib().orc_grid_indices(_p(inputs), _p(_c(offsets, np.int32)), C.c_uint32(B), C.c_uint32(D), C.c_uint32(level),
                           C.c_float(scale), C.c_uint32(gridtype), C.c_int(int(align_corners)), _p(out))
    return out


def grid_encode_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                        gridtype=0, align_corners=False, level_scales=None):
    """-> outputs [L,B,C] (reference layout), dy_dx [B, L*D*C] or None.  dtype follows `embeddings`."""
    inputs = _c(inputs, np.float32)
    emb = np.ascontiguousarray(embeddings)
    offsets = _c(offsets, np.int32)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    Cc = emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    out = np.empty((L, B, Cc), emb.dtype)
    dy_dx = np.empty((B, L * D * Cc), emb.dtype) if calc_grad_inputs else None
    ls = None if level_scales is None else _c(level_scales, np.float32)
    fn = getattr(lib(), "orc_grid_encode_forward_" + _SUFFIX[emb.dtype])
    rc = fn(_p(inputs), _p(emb), _p(offsets), _p(out), C.c_uint32(B), C.c_uint32(D), C.c_uint32(Cc), C.c_uint32(L),
            C.c_float(S), C.c_uint32(base_resolution), C.c_int(int(calc_grad_inputs)), _p(dy_dx),
            C.c_uint32(gridtype), C.c_int(int(align_corners)), _p(ls))
    if rc != 0:
        raise RuntimeError("GridEncoding: C must be 1, 2, 4, or 8." if rc == -2 else "GridEncoding: D must be 2 or 3.")
    return out, dy_dx


def grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, **kw):
    """[B, L*C] like `grid_encode` returns (gridencoder/grid.py:52)."""
    out, _ = grid_encode_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, **kw)
    L, B, Cc = out.shape
    return np.ascontiguousarray(out.transpose(1, 0, 2)).reshape(B, L * Cc)


def grid_encode_backward(grad_LBC, inputs, embeddings, offsets, per_level_scale, base_resolution, dy_dx=None,
                         gridtype=0, align_corners=False, level_scales=None):
    """grad_LBC [L,B,C] -> grad_embeddings [n,C], grad_inputs [B,D] (or None)."""
    inputs = _c(inputs, np.float32)
    grad = np.ascontiguousarray(grad_LBC)
    emb = np.ascontiguousarray(embeddings).astype(grad.dtype, copy=False)
    offsets = _c(offsets, np.int32)
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    Cc = emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    ge = np.zeros_like(emb)
    gi = np.zeros((B, D), grad.dtype) if dy_dx is not None else None
    dd = None if dy_dx is None else np.ascontiguousarray(dy_dx).astype(grad.dtype, copy=False)
    ls = None if level_scales is None else _c(level_scales, np.float32)
    fn = getattr(lib(), "orc_grid_encode_backward_" + _SUFFIX[grad.dtype])
    rc = fn(_p(grad), _p(inputs), _p(emb), _p(offsets), _p(ge), C.c_uint32(B), C.c_uint32(D), C.c_uint32(Cc),
            C.c_uint32(L), C.c_float(S), C.c_uint32(base_resolution), C.c_int(int(dy_dx is not None)), _p(dd), _p(gi),
            C.c_uint32(gridtype), C.c_int(int(align_corners)), _p(ls))
    if rc != 0:
        raise RuntimeError("grid_encode_backward: unsupported D/C")
    return ge, gi


# ------------------------------------------------------------------------------------------------ mlp
ACT = {"relu": 0, "exponential": 1, "sine": 2, "sigmoid": 3, "squareplus": 4, "softplus": 5, "none": 6}


def ffmlp_forward(inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation=0, output_activation=6,
                  want_forward_buffer=False, acc_mode=0):
    x = _c(inputs, np.float16)
    w = _c(weights, np.float16)
    B = x.shape[0]
    out = np.empty((B, output_dim), np.float16)
    fb = np.empty((num_layers, B, hidden_dim), np.float16) if want_forward_buffer else None
    rc = lib().orc_ffmlp_forward(_p(x), _p(w), _p(out), _p(fb), C.c_uint32(B), C.c_uint32(input_dim), C.c_uint32(output_dim),
                                 C.c_uint32(hidden_dim), C.c_uint32(num_layers), C.c_int(activation), C.c_int(output_activation),
                                 C.c_int(acc_mode))
    if rc != 0:
        raise RuntimeError("hidden_dim should in [16, 32, 64, 128, 256]")
    return (out, fb) if want_forward_buffer else out


def ffmlp_backward(grad, inputs, weights, forward_buffer, input_dim, output_dim, hidden_dim, num_layers, activation=0,
                   calc_grad_inputs=False):
    g = _c(grad, np.float16)
    x = _c(inputs, np.float16)
    w = _c(weights, np.float16)
    fb = _c(forward_buffer, np.float16)
    B = x.shape[0]
    bb = np.zeros((num_layers, B, hidden_dim), np.float16)
    gi = np.zeros((B, input_dim), np.float16) if calc_grad_inputs else None
    gw = np.zeros_like(w)
    lib().orc_ffmlp_backward(_p(g), _p(x), _p(w), _p(fb), _p(bb), _p(gi), _p(gw), C.c_uint32(B), C.c_uint32(input_dim),
                             C.c_uint32(output_dim), C.c_uint32(hidden_dim), C.c_uint32(num_layers), C.c_int(activation))
    return gw, gi, bb


# ------------------------------------------------------------------------------------------------ sh
def sh_encode_forward(inputs, degree, calc_grad_inputs=False):
    x = _c(inputs, np.float32)
    B, D = x.shape
    out = np.empty((B, degree * degree), np.float32)
    dy_dx = np.empty((B, D * degree * degree), np.float32) if calc_grad_inputs else None
    rc = lib().orc_sh_encode_forward(_p(x), _p(out), C.c_uint32(B), C.c_uint32(D), C.c_uint32(degree),
                                     C.c_int(int(calc_grad_inputs)), _p(dy_dx))
    if rc != 0:
        raise RuntimeError("SH encoder only supports input dim == 3 and degree in [1, 8]")
    return (out, dy_dx) if calc_grad_inputs else out


def sh_encode_backward(grad, degree, dy_dx):
    g = _c(grad, np.float32)
    B = g.shape[0]
    dd = _c(dy_dx, np.float32)
    gi = np.zeros((B, 3), np.float32)
    lib().orc_sh_encode_backward(_p(g), C.c_uint32(B), C.c_uint32(3), C.c_uint32(degree), _p(dd), _p(gi))
    return gi


# ------------------------------------------------------------------------------------------------ raymarching
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    o = _c(rays_o, np.float32).reshape(-1, 3)
    d = _c(rays_d, np.float32).reshape(-1, 3)
    N = o.shape[0]
    nears = np.empty(N, np.float32)
    fars = np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(o), _p(d), _p(_c(aabb, np.float32)), C.c_uint32(N), C.c_float(min_near), _p(nears), _p(fars))
    return nears, fars


def polar_from_ray(rays_o, rays_d, radius):
    o = _c(rays_o, np.float32).reshape(-1, 3)
    d = _c(rays_d, np.float32).reshape(-1, 3)
    N = o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().orc_polar_from_ray(_p(o), _p(d), C.c_float(radius), C.c_uint32(N), _p(coords))
    return coords


def morton3D(coords):
    c = _c(coords, np.int32)
    out = np.empty(c.shape[0], np.int32)
    lib().orc_morton3D(_p(c), C.c_uint32(c.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    i = _c(indices, np.int32)
    out = np.empty((i.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_p(i), C.c_uint32(i.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    g = _c(grid, np.float32)
    N = g.size // 8
    out = np.empty(N, np.uint8)
    lib().orc_packbits(_p(g), C.c_uint32(N), C.c_float(thresh), _p(out))
    return out


def pcg32_stream(seed, advance, count):
    u = np.empty(count, np.uint32)
    f = np.empty(count, np.float32)
    lib().orc_pcg32_stream(C.c_uint64(seed), C.c_int64(advance), C.c_uint32(count), _p(u), _p(f))
    return u, f


def march_rays_train(rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, M, perturb=False, dt_gamma=0.0, max_steps=1024,
                     want_ts=False):
    o = _c(rays_o, np.float32).reshape(-1, 3)
    d = _c(rays_d, np.float32).reshape(-1, 3)
    N = o.shape[0]
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    ts = np.zeros((M, 1), np.float32) if want_ts else None
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32)
    lib().orc_march_rays_train(_p(o), _p(d), _p(_c(bitfield, np.uint8)), C.c_float(bound), C.c_float(dt_gamma), C.c_uint32(max_steps),
                               C.c_uint32(N), C.c_uint32(Cc), C.c_uint32(H), C.c_uint32(M), _p(_c(nears, np.float32)),
                               _p(_c(fars, np.float32)), _p(xyzs), _p(dirs), _p(deltas), _p(ts), _p(rays), _p(counter),
                               C.c_uint32(int(perturb)))
    return xyzs, dirs, deltas, rays, counter, ts


def composite_rays_train_forward(sigmas, rgbs, deltas, rays):
    s = _c(sigmas, np.float32)
    c = _c(rgbs, np.float32)
    dl = _c(deltas, np.float32)
    r = _c(rays, np.int32)
    M, N = s.shape[0], r.shape[0]
    ws = np.empty(N, np.float32)
    depth = np.empty(N, np.float32)
    image = np.empty((N, 3), np.float32)
    lib().orc_composite_rays_train_forward(_p(s), _p(c), _p(dl), _p(r), C.c_uint32(M), C.c_uint32(N), _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image):
    s = _c(sigmas, np.float32)
    c = _c(rgbs, np.float32)
    M, N = s.shape[0], rays.shape[0]
    gs = np.zeros_like(s)
    gc = np.zeros_like(c)
    lib().orc_composite_rays_train_backward(_p(_c(grad_ws, np.float32)), _p(_c(grad_image, np.float32)), _p(s), _p(c),
                                            _p(_c(deltas, np.float32)), _p(_c(rays, np.int32)), _p(_c(weights_sum, np.float32)),
                                            _p(_c(image, np.float32)), C.c_uint32(M), C.c_uint32(N), _p(gs), _p(gc))
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, align=-1,
               perturb=0, dt_gamma=0.0, max_steps=1024):
    o = _c(rays_o, np.float32).reshape(-1, 3)
    d = _c(rays_d, np.float32).reshape(-1, 3)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    lib().orc_march_rays(C.c_uint32(n_alive), C.c_uint32(n_step), _p(_c(rays_alive, np.int32)), _p(_c(rays_t, np.float32)), _p(o), _p(d),
                         C.c_float(bound), C.c_float(dt_gamma), C.c_uint32(max_steps), C.c_uint32(Cc), C.c_uint32(H),
                         _p(_c(bitfield, np.uint8)), _p(_c(nears, np.float32)), _p(_c(fars, np.float32)), _p(xyzs), _p(dirs), _p(deltas),
                         C.c_uint32(int(perturb)))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    """in place on rays_t, weights_sum, depth, image (all float32 contiguous numpy arrays)"""
    for a in (rays_t, weights_sum, depth, image):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().orc_composite_rays(C.c_uint32(n_alive), C.c_uint32(n_step), _p(_c(rays_alive, np.int32)), _p(rays_t), _p(_c(sigmas, np.float32)),
                             _p(_c(rgbs, np.float32)), _p(_c(deltas, np.float32)), _p(weights_sum), _p(depth), _p(image))


def compact_rays(n_alive, rays_alive_old, rays_t_old):
    ra = np.zeros_like(_c(rays_alive_old, np.int32))
    rt = np.zeros_like(_c(rays_t_old, np.float32))
    cnt = np.zeros(1, np.int32)
    lib().orc_compact_rays(C.c_uint32(n_alive), _p(ra), _p(_c(rays_alive_old, np.int32)), _p(rt), _p(_c(rays_t_old, np.float32)), _p(cnt))
    return ra, rt, int(cnt[0])


# ------------------------------------------------------------------------------------------------ composed field
def trunc_exp(x):
    return np.exp(x.astype(np.float32))


def ngp_field(xyz, dirs, bound, embeddings_f16, offsets, per_level_scale, base_resolution, w_sigma, w_color,
              align_corners=True, level_scales=None, acc_mode=0, geo_feat_dim=15, sh_degree=4):
    """network_ff.NeRFNetwork.forward (nerf/network_ff.py:85-101) under fp16 autocast, inference mode:
    hashgrid -> FFMLP(32,16,64,2) -> trunc_exp | SH(4) ++ geo_feat ++ 0 -> FFMLP(32,3,64,3) -> sigmoid."""
    x01 = ((np.asarray(xyz, np.float32) + np.float32(bound)) / np.float32(2 * bound)).astype(np.float32)
    feat = grid_encode(x01, embeddings_f16, offsets, per_level_scale, base_resolution, gridtype=0,
                       align_corners=align_corners, level_scales=level_scales)
    h = ffmlp_forward(feat, w_sigma, feat.shape[1], 16, 64, 2, acc_mode=acc_mode)
    sigma = trunc_exp(h[:, 0])
    geo = h[:, 1:1 + geo_feat_dim]
    sh = sh_encode_forward(dirs, sh_degree).astype(np.float16)
    cin = np.concatenate([sh, geo, np.zeros((geo.shape[0], 1), np.float16)], axis=1)
    hc = ffmlp_forward(cin, w_color, cin.shape[1], 16, 64, 3, acc_mode=acc_mode)[:, :3]
    rgb = (1.0 / (1.0 + np.exp(-hc.astype(np.float32)))).astype(np.float16).astype(np.float32)
    return sigma, rgb


# ------------------------------------------------------------------------------------------------ whole-frame loop
def render_rays(rays_o, rays_d, bitfield, cascade, grid_size, bound, embeddings_f16, offsets, per_level_scale, base_resolution, w_sigma,
                w_color, align_corners=True, min_near=0.2, dt_gamma=0.0, max_steps=1024, bg_color=1.0, perturb=0, density_scale=1.0,
                level_scales=None):
    """CPU restatement of the inference branch of NeRFRenderer.run_cuda (nerf/renderer.py:436-489) over the oracle ops.
    Returns image [N,3], depth [N], weights_sum [N], number of non-sentinel samples, iterations."""
    o = _c(rays_o, np.float32).reshape(-1, 3)
    d = _c(rays_d, np.float32).reshape(-1, 3)
    N = o.shape[0]
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = near_far_from_aabb(o, d, aabb, min_near)
    weights_sum, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    alive, t = np.arange(N, dtype=np.int32), nears.copy()
    n_alive, step, it, n_samples = N, 0, 0, 0
    while step < max_steps:
        if step > 0:
            alive, t, n_alive = compact_rays(n_alive, alive, t)
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = march_rays(n_alive, n_step, alive, t, o, d, bound, bitfield, cascade, grid_size, nears, fars, align=128,
                                        perturb=perturb, dt_gamma=dt_gamma, max_steps=max_steps)
        live = deltas[:, 0] > 0
        n_samples += int(live.sum())
        sig, rgb = ngp_field(xyzs, dirs, bound, embeddings_f16, offsets, per_level_scale, base_resolution, w_sigma, w_color,
                             align_corners=align_corners, level_scales=level_scales)
        sig = (np.float32(density_scale) * sig).astype(np.float32)
        tt = np.ascontiguousarray(t[:n_alive])
        composite_rays(n_alive, n_step, alive, tt, sig, rgb, deltas, weights_sum, depth, image)
        t = tt
        alive = alive[:n_alive]
        step += n_step
        it += 1
    image = image + (1 - weights_sum)[:, None] * np.float32(bg_color)
    return image, depth, weights_sum, n_samples, it


# ------------------------------------------------------------------------------------------------ mesh front end (ntx_oracle_mesh.c)
def mesh_trace(vertices, triangles, rays_o, rays_d):
    """external/RayTracer raytracer.py:31-68 `trace`: positions [N,3], face_normals [N,3], depth [N], face_idx [N] i64 (-1 = no hit)."""
    v, t = _c(vertices, np.float32), _c(triangles, np.int32)
    o, d = _c(rays_o, np.float32).reshape(-1, 3), _c(rays_d, np.float32).reshape(-1, 3)
    N = o.shape[0]
    pos, nrm, depth = np.empty((N, 3), np.float32), np.empty((N, 3), np.float32), np.empty(N, np.float32)
    face = np.full(N, -1, np.int64)
    lib().orc_mesh_trace(_p(v), _p(t), C.c_uint32(t.shape[0]), _p(o), _p(d), C.c_uint32(N), _p(pos), _p(nrm), _p(depth), _p(face))
    return pos, nrm, depth, face


def points_knn(points, queries, K, r):
    """frnn.frnn_grid_points(queries, points, K=K, r=r, return_sorted=True) for one batch element: squared dists [N,K], idxs [N,K] i64."""
    p, q = _c(points, np.float32), _c(queries, np.float32).reshape(-1, 3)
    N = q.shape[0]
    dists, idxs = np.empty((N, K), np.float32), np.empty((N, K), np.int64)
    lib().orc_points_knn(_p(p), C.c_uint32(p.shape[0]), _p(q), C.c_uint32(N), C.c_uint32(K), C.c_float(r), _p(dists), _p(idxs))
    return dists, idxs


def mesh_project(vertices, vertex_normals, triangles, xyz, K=8, r=100.0, dir_vec_wdist=0.05):
    """tools/map.py:414-433 MeshProjector.project: p_sur [N,3], sdf [N], normal [N,3], face_idx [N] i64."""
    v, vn, t = _c(vertices, np.float32), _c(vertex_normals, np.float32), _c(triangles, np.int32)
    x = _c(xyz, np.float32).reshape(-1, 3)
    N = x.shape[0]
    p_sur, sdf, normal, face = np.empty((N, 3), np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32), np.empty(N, np.int64)
    lib().orc_mesh_project(_p(v), _p(vn), C.c_uint32(v.shape[0]), _p(t), C.c_uint32(t.shape[0]), _p(x), C.c_uint32(N), C.c_uint32(K),
                           C.c_float(r), C.c_float(dir_vec_wdist), _p(p_sur), _p(sdf), _p(normal), _p(face))
    return p_sur, sdf, normal, face
