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
import torch
from torch.autograd import Function

from nerf_texture_b200 import _lib as L

__all__ = ["near_far_from_aabb", "polar_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
           "march_rays_train_differentiable", "composite_rays_train", "march_rays", "composite_rays", "compact_rays"]

_fwd32 = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = torch.amp.custom_bwd(device_type="cuda")


def _f32c(t):
    t = t if t.is_cuda else t.cuda()
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


# ---------------------------------------------------------------------------------------------------- utils
class _near_far_from_aabb(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """rays_o/d [N,3], aabb [6] (xmin,ymin,zmin,xmax,ymax,zmax) -> nears, fars [N]"""
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        aabb = _f32c(aabb)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        L.call("ntx_near_far_from_aabb", L.ptr(rays_o), L.ptr(rays_d), L.ptr(aabb), N, float(min_near), L.ptr(nears), L.ptr(fars), L.stream())
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _polar_from_ray(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, radius):
        """(theta, phi) in [-1,1]^2 of the far intersection with the sphere of `radius` -> coords [N,2]"""
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        L.call("ntx_polar_from_ray", L.ptr(rays_o), L.ptr(rays_d), float(radius), N, L.ptr(coords), L.stream())
        return coords


polar_from_ray = _polar_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """coords [N,3] int32 in [0,1024) -> Morton codes [N] int32"""
        if not coords.is_cuda:
            coords = coords.cuda()
        coords = coords.int().contiguous()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        L.call("ntx_morton3D", L.ptr(coords), N, L.ptr(indices), L.stream())
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        if not indices.is_cuda:
            indices = indices.cuda()
        indices = indices.int().contiguous()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        L.call("ntx_morton3D_invert", L.ptr(indices), N, L.ptr(coords), L.stream())
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, grid, thresh, bitfield=None):
        """grid [C, H^3] float -> bitfield [C*H^3/8] uint8, bit i of byte n = grid[8n+i] > thresh"""
        grid = _f32c(grid)
        N = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        L.call("ntx_packbits", L.ptr(grid), N, float(thresh), L.ptr(bitfield), L.stream())
        _mip_cache.pop(bitfield.data_ptr(), None)   # the bit-field changed behind torch's version counter
        return bitfield


packbits = _packbits.apply


# ---------------------------------------------------------------------------------------------------- training
def _march_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, mean_count, perturb, align, force_all_rays,
                 dt_gamma, max_steps, want_ts):
    rays_o = _f32c(rays_o).view(-1, 3)
    rays_d = _f32c(rays_d).view(-1, 3)
    density_bitfield = (density_bitfield if density_bitfield.is_cuda else density_bitfield.cuda()).contiguous()
    nears, fars = _f32c(nears), _f32c(fars)
    N = rays_o.shape[0]
    M = N * max_steps
    if not force_all_rays and mean_count > 0:
        if align > 0:
            mean_count += align - mean_count % align
        M = mean_count
    dev = rays_o.device
    xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
    rays_ts = torch.zeros(M, 1, dtype=rays_o.dtype, device=dev) if want_ts else None
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
    ws = L.workspace("march_train", L.lib().ntx_march_rays_train_workspace_bytes(N), dev)
    L.call("ntx_march_rays_train", L.ptr(rays_o), L.ptr(rays_d), L.ptr(density_bitfield), float(bound), float(dt_gamma), int(max_steps), N,
           int(C), int(H), M, L.ptr(nears), L.ptr(fars), L.ptr(xyzs), L.ptr(dirs), L.ptr(deltas), L.ptr(rays_ts), L.ptr(rays),
           step_counter.data_ptr(), int(bool(perturb)), L.ptr(ws), L.stream())
    if force_all_rays or mean_count <= 0:
        m = step_counter[0].item()  # D2H copy, only in the first epochs
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    return xyzs, dirs, deltas, rays, rays_ts, N, M


class _march_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1,
                force_all_rays=False, dt_gamma=0, max_steps=1024):
        """-> xyzs [M,3], dirs [M,3], deltas [M,2] (dt, real delta), rays [N,3] int32 (ray id, sample offset, sample count)"""
        xyzs, dirs, deltas, rays, _, _, _ = _march_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, mean_count,
                                                         perturb, align, force_all_rays, dt_gamma, max_steps, False)
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _march_rays_train_differentiable(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1,
                force_all_rays=False, dt_gamma=0, max_steps=1024):
        xyzs, dirs, deltas, rays, rays_ts, N, _ = _march_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter,
                                                               mean_count, perturb, align, force_all_rays, dt_gamma, max_steps, True)
        ctx.N, ctx.max_steps = N, max_steps
        ctx.save_for_backward(rays_ts)
        return xyzs, dirs, deltas, rays

    @staticmethod
    def backward(ctx, grad_xyzs, grad_dirs, grad_deltas, grad_rays):
        # d xyz / d o = I, d xyz / d d = t  (raymarching.py:275-286: assumes max_steps slots per ray)
        rays_ts = ctx.saved_tensors[0]
        total = ctx.N * ctx.max_steps
        g = torch.zeros(total, 3, device=grad_xyzs.device, dtype=grad_xyzs.dtype)
        g[:grad_xyzs.shape[0]] = grad_xyzs
        t = torch.zeros(total, 1, device=grad_xyzs.device, dtype=grad_xyzs.dtype)
        t[:rays_ts.shape[0]] = rays_ts
        g = g.reshape(ctx.N, -1, 3)
        t = t.reshape(ctx.N, -1, 1)
        return (g.sum(dim=1), (g * t).sum(dim=1)) + (None,) * 13


march_rays_train_differentiable = _march_rays_train_differentiable.apply


class _composite_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, rgbs, deltas, rays):
        """sigmas [M], rgbs [M,3], deltas [M,2], rays [N,3] -> weights_sum [N], depth [N], image [N,3]"""
        sigmas, rgbs, deltas = _f32c(sigmas), _f32c(rgbs), _f32c(deltas)
        M, N = sigmas.shape[0], rays.shape[0]
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        L.call("ntx_composite_rays_train_forward", L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas), L.ptr(rays), M, N, L.ptr(weights_sum), L.ptr(depth),
               L.ptr(image), L.stream())
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N]
        return weights_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is ignored, as in the reference (raymarching.py:330)
        grad_weights_sum = grad_weights_sum.contiguous()
        grad_image = grad_image.contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        L.call("ntx_composite_rays_train_backward", L.ptr(grad_weights_sum), L.ptr(grad_image), L.ptr(sigmas), L.ptr(rgbs), L.ptr(deltas),
               L.ptr(rays), L.ptr(weights_sum), L.ptr(image), M, N, L.ptr(grad_sigmas), L.ptr(grad_rgbs), L.stream())
        return grad_sigmas, grad_rgbs, None, None


composite_rays_train = _composite_rays_train.apply


# ---------------------------------------------------------------------------------------------------- inference
_mip_cache = {}


def _occupancy_mip(bitfield, C, H):
    """conservative occupancy mip of `bitfield` (see ntx_build_occupancy_mip), cached per (storage, version): rebuilt whenever torch
    has seen the bit-field change; our own packbits() (which writes through a raw pointer) invalidates it explicitly"""
    if H < 16 or (H & (H - 1)) != 0 or (bitfield.data_ptr() & 15) != 0:
        return None
    key = (bitfield.data_ptr(), bitfield._version, C, H)
    hit = _mip_cache.get(bitfield.data_ptr())
    if hit is None or hit[0] != key:
        mip = torch.empty(L.lib().ntx_occupancy_mip_bytes(C, H), dtype=torch.uint8, device=bitfield.device)
        L.call("ntx_build_occupancy_mip", L.ptr(bitfield), C, H, L.ptr(mip), L.stream())
        _mip_cache[bitfield.data_ptr()] = (key, mip)
        hit = _mip_cache[bitfield.data_ptr()]
    return hit[1].data_ptr()


class _march_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1, perturb=False,
                dt_gamma=0, max_steps=1024):
        """-> xyzs, dirs [n_alive*n_step (+pad to `align`), 3], deltas [.., 2]; unused slots are zero (delta 0 = sentinel)"""
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        dev = rays_o.device
        xyzs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
        dirs = torch.empty(M, 3, dtype=rays_o.dtype, device=dev)
        deltas = torch.empty(M, 2, dtype=rays_o.dtype, device=dev)
        L.call("ntx_march_rays", int(n_alive), int(n_step), L.ptr(rays_alive), L.ptr(rays_t), L.ptr(rays_o), L.ptr(rays_d), float(bound),
               float(dt_gamma), int(max_steps), int(C), int(H), L.ptr(density_bitfield), L.ptr(near), L.ptr(far), L.ptr(xyzs), L.ptr(dirs),
               L.ptr(deltas), int(perturb), 1, M, _occupancy_mip(density_bitfield, int(C), int(H)), L.stream())
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
        """continues the front-to-back accumulation IN PLACE in weights_sum / depth / image; dead rays get rays_t = -1"""
        L.call("ntx_composite_rays", int(n_alive), int(n_step), L.ptr(rays_alive), L.ptr(rays_t), L.ptr(_f32c(sigmas)),
               L.ptr(_f32c(rgbs)), L.ptr(_f32c(deltas)), L.ptr(weights_sum), L.ptr(depth), L.ptr(image), L.stream())
        return tuple()


composite_rays = _composite_rays.apply


class _compact_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter):
        """survivors (rays_t_old >= 0) -> rays_alive / rays_t (ascending slot order); alive_counter += #survivors"""
        ws = L.workspace("compact", L.lib().ntx_compact_rays_workspace_bytes(int(n_alive)), rays_alive.device)
        L.call("ntx_compact_rays", int(n_alive), L.ptr(rays_alive), L.ptr(rays_alive_old), L.ptr(rays_t), L.ptr(rays_t_old),
               L.ptr(alive_counter), L.ptr(ws), L.stream())
        return tuple()


compact_rays = _compact_rays.apply
