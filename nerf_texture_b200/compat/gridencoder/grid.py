"""Drop-in `gridencoder.grid` on libntx (B200).

API mirror of the reference's gridencoder/grid.py (`grid_encode` :90, `GridEncoder` :93): same constructor arguments,
state (`embeddings` [n_entries, level_dim] fp32 Parameter, `offsets` [L+1] int32 buffer), autocast behaviour and
gradients.  What differs is underneath: one kernel launch writes `[B, L*C]` directly (no `[L,B,C]` + permute copy,
grid.py:42-52) and the backward consumes the `[B, L*C]` gradient in place (no permute().contiguous(), grid.py:72).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from nerf_texture_b200 import _lib as L

_gridtype_to_id = {"hash": 0, "tiled": 1}


class _grid_encode(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False):
        # inputs [B, D] float in [0, 1]; embeddings [sO, C]; offsets [L + 1] int32 -> [B, L * C]
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        B, D = inputs.shape
        n_levels = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)

        # autocast: half-precision table only (inputs stay fp32); odd C stays fp32 (reference: grid.py:36-39)
        if torch.is_autocast_enabled() and C % 2 == 0:
            embeddings = embeddings.to(torch.half)
        embeddings = embeddings.contiguous()

        outputs = torch.empty(B, n_levels * C, device=inputs.device, dtype=embeddings.dtype)
        if calc_grad_inputs:
            dy_dx = torch.empty(B, n_levels * D * C, device=inputs.device, dtype=embeddings.dtype)
        else:
            dy_dx = torch.empty(1, device=inputs.device, dtype=embeddings.dtype)

        L.call("ntx_grid_encode_forward", L.ptr(inputs), L.ptr(embeddings), L.ptr(offsets), L.ptr(outputs), B, D, C, n_levels, S, H,
               int(calc_grad_inputs), L.ptr(dy_dx), int(gridtype), int(align_corners), L.dtype_id(embeddings.dtype), L.LAYOUT_BLC,
               L.stream())

        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = [B, D, C, n_levels, S, H, gridtype]
        ctx.calc_grad_inputs = calc_grad_inputs
        ctx.align_corners = align_corners
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, n_levels, S, H, gridtype = ctx.dims
        calc_grad_inputs = ctx.calc_grad_inputs

        grad = grad.contiguous()  # [B, L*C], consumed as is
        if grad.dtype != embeddings.dtype:
            grad = grad.to(embeddings.dtype)
        grad_embeddings = torch.zeros_like(embeddings)
        if calc_grad_inputs:
            grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype)
        else:
            grad_inputs = torch.zeros(1, device=inputs.device, dtype=embeddings.dtype)

        L.call("ntx_grid_encode_backward", L.ptr(grad), L.ptr(inputs), L.ptr(embeddings), L.ptr(offsets), L.ptr(grad_embeddings), B, D, C,
               n_levels, S, H, int(calc_grad_inputs), L.ptr(dy_dx), L.ptr(grad_inputs), int(gridtype), int(ctx.align_corners),
               L.dtype_id(embeddings.dtype), L.LAYOUT_BLC, L.stream())

        if calc_grad_inputs:
            return grad_inputs.to(inputs.dtype), grad_embeddings, None, None, None, None, None, None
        return None, grad_embeddings, None, None, None, None, None, None


grid_encode = _grid_encode.apply


def level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    """Entries per level, as the reference sizes them (grid.py:113-124): min(2^T, res^D or (res+1)^D) rounded up to 8."""
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return offsets


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype="hash", align_corners=False):
        super().__init__()
        # the finest resolution desired at the last level, if provided, overrides per_level_scale
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))

        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.align_corners = align_corners
        self.max_params = 2 ** log2_hashmap_size

        offsets = level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners)
        self.register_buffer("offsets", torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offsets[-1], level_dim))
        self.reset_parameters()

    def reset_parameters(self, std=1e-4):
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners}")

    def forward(self, inputs, bound=1):
        # inputs: [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, inputs.requires_grad,
                              self.gridtype_id, self.align_corners)
        return outputs.view(prefix_shape + [self.output_dim])
