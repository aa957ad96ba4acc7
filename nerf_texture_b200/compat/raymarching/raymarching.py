"""Drop-in `raymarching.raymarching` on libntx (B200).

Same eleven functionals, positional signatures and in-place conventions as the reference's raymarching/raymarching.py
(:49 near_far_from_aabb, :80 polar_from_ray, :104 morton3D, :126 morton3D_invert, :155 packbits, :230 march_rays_train,
:290 march_rays_train_differentiable, :346 composite_rays_train, :397 march_rays, :422 composite_rays, :442 compact_rays).

Behavioural notes (all members of the reference's own outcome set):
  * ray compaction and training segment allocation are ordered scans, so `rays_alive` / `rays` come out in ascending
    order instead of the arbitrary order of the reference's global atomics;
  * `march_rays` lets the kernel zero the unused sample slots instead of three `torch.zeros` fills per call;
  * no `torch.cuda.empty_cache()` inside `march_rays_train` (raymarching.py:226) — it only stalls the allocator.
"""
import weakref

import torch
from torch.autograd import Function

from nerf_texture_b200 import _lib as L

__all__ = ["near_far_from_aabb", "polar_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
           "march_rays_train_differentiable", "composite_rays_train", "march_rays", "composite_rays", "compact_rays"]

_F32, _I32 = torch.float32, torch.int32
_fwd32 = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = torch.amp.custom_bwd(device_type="cuda")


def _f32c(t):
    """fp32, contiguous, on the GPU — what the reference gets from custom_fwd(cast_inputs=float32) + its .cuda()/.contiguous() calls"""
    t = t if t.is_cuda else t.cuda()
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _rays(rays_o, rays_d):
    return _f32c(rays_o).view(-1, 3), _f32c(rays_d).view(-1, 3)


def _new(like, *shape, dtype=None):
    return torch.empty(*shape, dtype=dtype or like.dtype, device=like.device)


# Only two of the eleven functionals have a backward (composite_rays_train, march_rays_train_differentiable); the other nine
# are plain functions here: they take and return tensors that never require grad, so an autograd node buys nothing.

# ---------------------------------------------------------------------------------------------------- geometry helpers
@torch.no_grad()
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """rays_o/d [N,3], aabb [6] (xmin,ymin,zmin,xmax,ymax,zmax) -> nears, fars [N]"""
    o, d = _rays(rays_o, rays_d)
    n = o.shape[0]
    nears, fars = _new(o, n), _new(o, n)
    L.call("ntx_near_far_from_aabb", L.ptr(o), L.ptr(d), L.ptr(_f32c(aabb)), n, float(min_near), L.ptr(nears), L.ptr(fars), L.stream())
    return nears, fars


@torch.no_grad()
def polar_from_ray(rays_o, rays_d, radius):
    """(theta, phi) in [-1,1]^2 of the far intersection with the sphere of `radius` -> coords [N,2]"""
    o, d = _rays(rays_o, rays_d)
    coords = _new(o, o.shape[0], 2)
    L.call("ntx_polar_from_ray", L.ptr(o), L.ptr(d), float(radius), o.shape[0], L.ptr(coords), L.stream())
    return coords


def _i32c(t):
    return (t if t.is_cuda else t.cuda()).int().contiguous()


@torch.no_grad()
def morton3D(coords):
    """coords [N,3] int32 in [0,1024) -> Morton codes [N] int32"""
    c = _i32c(coords)
    codes = _new(c, c.shape[0])
    L.call("ntx_morton3D", L.ptr(c), c.shape[0], L.ptr(codes), L.stream())
    return codes


@torch.no_grad()
def morton3D_invert(indices):
    """Morton codes [N] int32 -> coords [N,3] int32"""
    i = _i32c(indices)
    coords = _new(i, i.shape[0], 3)
    L.call("ntx_morton3D_invert", L.ptr(i), i.shape[0], L.ptr(coords), L.stream())
    return coords


@torch.no_grad()
def packbits(grid, thresh, bitfield=None):
    """grid [C, H^3] float -> bitfield [C*H^3/8] uint8, bit i of byte n = grid[8n+i] > thresh"""
    g = _f32c(grid)
    nbytes = g.shape[0] * g.shape[1] // 8
    if bitfield is None:
        bitfield = _new(g, nbytes, dtype=torch.uint8)
    L.call("ntx_packbits", L.ptr(g), nbytes, float(thresh), L.ptr(bitfield), L.stream())
    _mip_forget(bitfield)   # the bit-field changed behind torch's version counter
    return bitfield


# ---------------------------------------------------------------------------------------------------- training
def _march_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, mean_count, perturb, align, force_all_rays,
                 dt_gamma, max_steps, want_ts):
    """shared body of march_rays_train(_differentiable): sample budget M as the reference sizes it (raymarching.py:200-214)"""
    o, d = _rays(rays_o, rays_d)
    bits = (density_bitfield if density_bitfield.is_cuda else density_bitfield.cuda()).contiguous()
    nears, fars = _f32c(nears), _f32c(fars)
    n = o.shape[0]
    budget = n * max_steps
    if not force_all_rays and mean_count > 0:
        budget = mean_count + (align - mean_count % align if align > 0 else 0)
    xyzs, dirs, deltas = (torch.zeros(budget, k, dtype=o.dtype, device=o.device) for k in (3, 3, 2))
    rays_ts = torch.zeros(budget, 1, dtype=o.dtype, device=o.device) if want_ts else None
    rays = _new(o, n, 3, dtype=torch.int32)
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=o.device)
    ws = L.workspace("march_train", L.lib().ntx_march_rays_train_workspace_bytes(n), o.device)
    L.call("ntx_march_rays_train", L.ptr(o), L.ptr(d), L.ptr(bits), float(bound), float(dt_gamma), int(max_steps), n, int(C), int(H), budget, L.ptr(nears),
           L.ptr(fars), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.ptr(rays_ts), L.ptr(rays), L.ptr(step_counter, _I32), int(bool(perturb)), L.ptr(ws), L.stream())
    if force_all_rays or mean_count <= 0:
        used = step_counter[0].item()              # D2H copy, only in the first epochs (raymarching.py:219-224)
        if align > 0:
            used += align - used % align
        xyzs, dirs, deltas = xyzs[:used], dirs[:used], deltas[:used]
    return xyzs, dirs, deltas, rays, rays_ts, n


@torch.no_grad()
def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1,
                     force_all_rays=False, dt_gamma=0, max_steps=1024):
    """-> xyzs [M,3], dirs [M,3], deltas [M,2] (dt, real delta), rays [N,3] int32 (ray id, sample offset, sample count)"""
    return _march_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, mean_count, perturb, align, force_all_rays, dt_gamma,
                        max_steps, False)[:4]


class MarchTrainDifferentiableOp(Function):
    """march_rays_train with d xyz / d rays_o = I and d xyz / d rays_d = t (raymarching.py:232-288)"""

    @staticmethod
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1,
                force_all_rays=False, dt_gamma=0, max_steps=1024):
        xyzs, dirs, deltas, rays, rays_ts, n = _march_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, mean_count, perturb,
                                                            align, force_all_rays, dt_gamma, max_steps, True)
        ctx.n_rays, ctx.max_steps = n, max_steps
        ctx.save_for_backward(rays_ts)
        return xyzs, dirs, deltas, rays

    @staticmethod
    def backward(ctx, grad_xyzs, grad_dirs, grad_deltas, grad_rays):
        # like the reference's backward this assumes max_steps sample slots per ray (raymarching.py:275-286)
        (rays_ts,) = ctx.saved_tensors
        slots = ctx.n_rays * ctx.max_steps
        g = grad_xyzs.new_zeros(slots, 3)
        g[:grad_xyzs.shape[0]] = grad_xyzs
        t = grad_xyzs.new_zeros(slots, 1)
        t[:rays_ts.shape[0]] = rays_ts
        g, t = g.view(ctx.n_rays, -1, 3), t.view(ctx.n_rays, -1, 1)
        return (g.sum(dim=1), (g * t).sum(dim=1)) + (None,) * 13


march_rays_train_differentiable = MarchTrainDifferentiableOp.apply


class CompositeTrainOp(Function):
    """composite_rays_train (raymarching.py:297-346): sigmas [M], rgbs [M,3], deltas [M,2], rays [N,3] -> weights_sum [N], depth [N], image [N,3]"""

    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, rgbs, deltas, rays):
        sigmas, rgbs, deltas = _f32c(sigmas), _f32c(rgbs), _f32c(deltas)
        n_samples, n_rays = sigmas.shape[0], rays.shape[0]
        weights_sum, depth, image = _new(sigmas, n_rays), _new(sigmas, n_rays), _new(sigmas, n_rays, 3)
        L.call("ntx_composite_rays_train_forward", L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(rays, _I32), n_samples, n_rays, L.ptr(weights_sum), L.ptr(depth),
               L.ptr(image), L.stream())
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image)
        return weights_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # the gradient of depth is dropped, as in the reference (raymarching.py:330)
        sigmas, rgbs, deltas, rays, weights_sum, image = ctx.saved_tensors
        d_sigmas, d_rgbs = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        L.call("ntx_composite_rays_train_backward", L.ptr(grad_weights_sum.contiguous()), L.ptr(grad_image.contiguous()), L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas),
               L.ptr(rays), L.ptr(weights_sum), L.ptr(image), sigmas.shape[0], rays.shape[0], L.ptr(d_sigmas), L.ptr(d_rgbs), L.stream())
        return d_sigmas, d_rgbs, None, None


composite_rays_train = CompositeTrainOp.apply


# ---------------------------------------------------------------------------------------------------- inference
_mip_cache = {}          # id(bit-field tensor) -> (weakref to that tensor, (version, C, H), mip)
_MIP_CACHE_MAX = 8


def _mip_forget(bitfield):
    _mip_cache.pop(id(bitfield), None)


def _occupancy_mip(bitfield, C, H):
    """conservative occupancy mip of `bitfield` (see ntx_build_occupancy_mip).  Cached per TENSOR OBJECT (a weak reference pins the
    identity: an entry dies with its tensor, so a new bit-field that the caching allocator places at the same address — or a
    state_dict load into it — can never inherit another model's mip) and per version counter; our own packbits(), which writes
    through a raw pointer, drops the entry explicitly.  Tensors without a version counter (inference mode) are not cached."""
    if H < 16 or (H & (H - 1)) != 0 or (bitfield.data_ptr() & 15) != 0:
        return None
    try:
        version = bitfield._version
    except Exception:
        version = None
    hit = _mip_cache.get(id(bitfield)) if version is not None else None
    if hit is not None and (hit[0]() is not bitfield or hit[1] != (version, C, H, bitfield.data_ptr())):
        hit = None
    if hit is None:
        mip = torch.empty(L.lib().ntx_occupancy_mip_bytes(C, H), dtype=torch.uint8, device=bitfield.device)
        L.call("ntx_build_occupancy_mip", L.ptr(bitfield, torch.uint8), C, H, L.ptr(mip), L.stream())
        if version is None:
            return mip.data_ptr(), mip
        key = id(bitfield)
        if len(_mip_cache) >= _MIP_CACHE_MAX:
            for k in [k for k, v in _mip_cache.items() if v[0]() is None] or [next(iter(_mip_cache))]:
                _mip_cache.pop(k, None)
        hit = _mip_cache[key] = (weakref.ref(bitfield, lambda _r, key=key: _mip_cache.pop(key, None)), (version, C, H, bitfield.data_ptr()), mip)
    return hit[2].data_ptr(), hit[2]


@torch.no_grad()
def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1, perturb=False, dt_gamma=0,
               max_steps=1024):
    """-> xyzs, dirs [n_alive*n_step (+pad to `align`), 3], deltas [.., 2]; unused slots are zero (delta 0 = sentinel)"""
    o, d = _rays(rays_o, rays_d)
    rows = n_alive * n_step
    if align > 0:
        rows += align - (rows % align)
    xyzs, dirs, deltas = _new(o, rows, 3), _new(o, rows, 3), _new(o, rows, 2)     # the kernel zero-fills what it does not use
    mip = _occupancy_mip(density_bitfield, int(C), int(H))                          # (pointer, tensor that keeps it alive during the call)
    L.call("ntx_march_rays", int(n_alive), int(n_step), L.ptr(rays_alive, _I32), L.ptr(rays_t, _F32), L.ptr(o), L.ptr(d), float(bound), float(dt_gamma),
           int(max_steps), int(C), int(H), L.ptr(density_bitfield, torch.uint8), L.ptr(near, _F32), L.ptr(far, _F32), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas),
           int(perturb), 1, rows, None if mip is None else mip[0], L.stream())
    return xyzs, dirs, deltas


@torch.no_grad()
def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
    """continues the front-to-back accumulation IN PLACE in weights_sum / depth / image; dead rays get rays_t = -1"""
    L.call("ntx_composite_rays", int(n_alive), int(n_step), L.ptr(rays_alive, _I32), L.ptr(rays_t, _F32), L.ptr(_f32c(sigmas)), L.ptr(_f32c(rgbs)),
           L.ptr(_f32c(deltas)), L.ptr(weights_sum, _F32), L.ptr(depth, _F32), L.ptr(image, _F32), L.stream())
    return tuple()


@torch.no_grad()
def compact_rays(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter):
    """survivors (rays_t_old >= 0) -> rays_alive / rays_t (ascending slot order); alive_counter += #survivors"""
    ws = L.workspace("compact", L.lib().ntx_compact_rays_workspace_bytes(int(n_alive)), rays_alive.device)
    L.call("ntx_compact_rays", int(n_alive), L.ptr(rays_alive, _I32), L.ptr(rays_alive_old, _I32), L.ptr(rays_t, _F32), L.ptr(rays_t_old, _F32),
           L.ptr(alive_counter, _I32), L.ptr(ws), L.stream())
    return tuple()
