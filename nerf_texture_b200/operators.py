"""Autograd operators and modules of the hot path on libntx: hash-grid encoder, fully-fused MLP, spherical harmonics.

These are the implementations behind the drop-in packages in `nerf_texture_b200/compat/` (gridencoder, ffmlp, shencoder), which
only re-export them under the reference's module paths.  Public names, constructor arguments, state-dict entries, autocast
behaviour and error messages are the reference's (cited below) so that its model code and checkpoints work unchanged; the
bodies are organised around one pattern: an `Op` records the geometry of a call once (`_Geom`), forward/backward hand raw
pointers to the C ABI, and there is no CPU fallback (`_lib.call` raises RuntimeError).

What differs underneath, invisible to callers:
  * grid: one launch writes `[B, L*C]` directly and the backward consumes that layout in place (the reference produces `[L,B,C]`
    and pays a permute copy each way, gridencoder/grid.py:42-52,72);
  * mlp: ragged batches are masked in-kernel (the reference pads the batch to a multiple of 128 with a concatenated zero block,
    ffmlp/ffmlp.py:155-160); no side streams (`allocate_splitk` is a no-op);
  * everything runs on the current torch stream.
"""
import math
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as L

_HALF, _F32 = torch.half, torch.float32


def _dense(t, dtype=None):
    """contiguous tensor of the wanted dtype (no copy when it already is)"""
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _scratch(like, *shape):
    return torch.empty(*shape, device=like.device, dtype=like.dtype)


# ====================================================================================================================== hash grid
_GridGeom = namedtuple("_GridGeom", "B D C L S H gridtype align want_dx")
GRIDTYPE_ID = {"hash": 0, "tiled": 1}          # gridencoder/grid.py:14-17


def _half_table(embeddings):
    """fp16 copy of the table, kept on the tensor object and re-made only when torch has seen the table change (version counter) or
    move (data pointer).  The reference casts all 12.2 M entries on EVERY call under autocast (gridencoder/grid.py:38-39) — 43 times
    per rendered frame; an optimizer step bumps the version, so training still casts once per step, exactly when it has to."""
    if embeddings.dtype == _HALF:
        return embeddings
    try:
        key = (embeddings._version, embeddings.data_ptr(), tuple(embeddings.shape))
    except Exception:                       # inference-mode tensors have no version counter: no caching
        return embeddings.to(_HALF)
    hit = getattr(embeddings, "_ntx_half_table", None)
    if hit is None or hit[0] != key:
        hit = (key, embeddings.detach().to(_HALF))
        try:
            embeddings._ntx_half_table = hit
        except Exception:
            pass
    return hit[1]


class HashGridOp(Function):
    """grid_encode (gridencoder/grid.py:19-87): x in [0,1]^D, table [n_entries, C], offsets [L+1] -> features [B, L*C]"""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0, align_corners=False):
        x = _dense(inputs, _F32)
        # under autocast the table (not the coordinates) goes to half, and only for even C (gridencoder/grid.py:36-39)
        table = _half_table(embeddings) if (torch.is_autocast_enabled() and embeddings.shape[1] % 2 == 0) else embeddings
        table = table.contiguous()
        g = _GridGeom(x.shape[0], x.shape[1], table.shape[1], offsets.shape[0] - 1, float(np.log2(per_level_scale)), int(base_resolution),
                      int(gridtype), int(align_corners), bool(calc_grad_inputs))
        feats = _scratch(table, g.B, g.L * g.C)
        jac = _scratch(table, g.B, g.L * g.D * g.C) if g.want_dx else _scratch(table, 1)
        L.call("ntx_grid_encode_forward", L.ptr(x), L.ptr(table), L.ptr(offsets), L.ptr(feats), g.B, g.D, g.C, g.L, g.S, g.H, int(g.want_dx), L.ptr(jac),
               g.gridtype, g.align, L.dtype_id(table.dtype), L.LAYOUT_BLC, L.stream())
        ctx.geom = g
        ctx.save_for_backward(x, table, offsets, jac)
        return feats

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        x, table, offsets, jac = ctx.saved_tensors
        g = ctx.geom
        dy = _dense(grad, table.dtype)                                  # [B, L*C], used as it is
        d_table = torch.zeros_like(table)
        d_x = torch.zeros_like(x, dtype=table.dtype) if g.want_dx else torch.zeros(1, device=x.device, dtype=table.dtype)
        L.call("ntx_grid_encode_backward", L.ptr(dy), L.ptr(x), L.ptr(table), L.ptr(offsets), L.ptr(d_table), g.B, g.D, g.C, g.L, g.S, g.H, int(g.want_dx),
               L.ptr(jac), L.ptr(d_x), g.gridtype, g.align, L.dtype_id(table.dtype), L.LAYOUT_BLC, L.stream())
        return (d_x.to(x.dtype) if g.want_dx else None), d_table, None, None, None, None, None, None


grid_encode = HashGridOp.apply


def hashgrid_level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    """First table entry of every level (+ the total): a level holds min(2^T, R^D) entries rounded up to a multiple of 8, with
    R = res or res + 1 (align_corners) and res = ceil(base * scale^level)  — the sizing rule of gridencoder/grid.py:113-124."""
    cap = 2 ** log2_hashmap_size
    sizes = []
    for level in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** level))
        side = res if align_corners else res + 1
        sizes.append(8 * int(np.ceil(min(cap, side ** input_dim) / 8)))
    return [int(v) for v in np.concatenate([[0], np.cumsum(sizes)])]


class GridEncoder(nn.Module):
    """gridencoder.GridEncoder (gridencoder/grid.py:93-152): state = `embeddings` [n_entries, level_dim] fp32, `offsets` [L+1] int32."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype="hash", align_corners=False):
        super().__init__()
        if desired_resolution is not None:      # geometric progression from base_resolution to desired_resolution
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution, self.log2_hashmap_size = per_level_scale, base_resolution, log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.gridtype, self.gridtype_id = gridtype, GRIDTYPE_ID[gridtype]
        self.align_corners = align_corners
        self.max_params = 2 ** log2_hashmap_size
        table_offsets = hashgrid_level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners)
        self.register_buffer("offsets", torch.tensor(table_offsets, dtype=torch.int32))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(table_offsets[-1], level_dim))
        self.reset_parameters()

    def reset_parameters(self, std=1e-4):
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        finest = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} resolution={self.base_resolution} -> {finest} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners}")

    def forward(self, inputs, bound=1):
        """[..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]"""
        unit = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
        feats = grid_encode(unit, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, unit.requires_grad, self.gridtype_id,
                            self.align_corners)
        return feats.view(*inputs.shape[:-1], self.output_dim)


# ====================================================================================================================== fused MLP
_MlpGeom = namedtuple("_MlpGeom", "B n_in n_out width depth act out_act want_dx")
ACTIVATION_ID = {"relu": 0, "exponential": 1, "sine": 2, "sigmoid": 3, "squareplus": 4, "softplus": 5}   # ffmlp/ffmlp.py:106-119; anything else = none


def convert_activation(act):
    return ACTIVATION_ID.get(act, 6)


class FusedMLPOp(Function):
    """ffmlp_forward (ffmlp/ffmlp.py:17-86): fp16 bias-free MLP, weights flat [width x n_in | (depth-1) x width x width | 16 x width]"""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=_HALF)
    def forward(ctx, inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, inference=False, calc_grad_inputs=False):
        x, w = _dense(inputs, _HALF), _dense(weights, _HALF)
        g = _MlpGeom(x.shape[0], input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, bool(calc_grad_inputs))
        y = _scratch(x, g.B, g.n_out)
        if inference:
            L.call("ntx_ffmlp_inference", L.ptr(x), L.ptr(w), g.B, g.n_in, g.n_out, g.width, g.depth, g.act, g.out_act, None, L.ptr(y), L.stream())
            return y
        hidden = _scratch(x, g.depth, g.B, g.width)                      # activations of every layer, kept for the backward
        L.call("ntx_ffmlp_forward", L.ptr(x), L.ptr(w), g.B, g.n_in, g.n_out, g.width, g.depth, g.act, g.out_act, L.ptr(hidden), L.ptr(y), L.stream())
        ctx.geom = g
        ctx.save_for_backward(x, w, y, hidden)
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        x, w, _, hidden = ctx.saved_tensors
        g = ctx.geom
        dy = _dense(grad, _HALF)
        d_w = torch.zeros_like(w)
        d_x = torch.zeros_like(x) if g.want_dx else torch.zeros(1, device=dy.device, dtype=dy.dtype)
        d_hidden = _scratch(dy, g.depth, g.B, g.width)
        ws = torch.empty(L.lib().ntx_ffmlp_backward_workspace_bytes(g.n_in, g.n_out, g.width, g.depth), dtype=torch.uint8, device=dy.device)
        L.call("ntx_ffmlp_backward", L.ptr(dy), L.ptr(x), L.ptr(w), L.ptr(hidden), g.B, g.n_in, g.n_out, g.width, g.depth, g.act, g.out_act, int(g.want_dx),
               L.ptr(d_hidden), L.ptr(d_x), L.ptr(d_w), L.ptr(ws), L.stream())
        return (d_x if g.want_dx else None), d_w, None, None, None, None, None, None, None, None


ffmlp_forward = FusedMLPOp.apply


class FFMLP(nn.Module):
    """ffmlp.FFMLP (ffmlp/ffmlp.py:99-170): state = flat fp32 `weights`; output padded to 16 inside, sliced outside."""

    def __init__(self, input_dim, output_dim, hidden_dim, num_layers, activation="relu"):
        super().__init__()
        self.input_dim, self.output_dim, self.hidden_dim, self.num_layers = input_dim, output_dim, hidden_dim, num_layers
        self.activation, self.output_activation = convert_activation(activation), convert_activation("none")
        self.tensorcore_width = 16
        # the reference's argument checks and messages (ffmlp/ffmlp.py:124-127)
        assert hidden_dim in [16, 32, 64, 128, 256], f"FFMLP only support hidden_dim in [16, 32, 64, 128, 256], but got {hidden_dim}"
        assert input_dim > 0 and input_dim % 16 == 0, f"FFMLP input_dim should be 16 * m (m  > 0), but got {input_dim}"
        assert output_dim <= 16, f"FFMLP current only supports output dim <= 16, but got {output_dim}"
        assert num_layers >= 2, f"FFMLP num_layers should be larger than 2 (3 matmuls), but got {num_layers}"
        self.padded_output_dim = 16 * int(math.ceil(output_dim / 16))
        self.num_parameters = hidden_dim * (input_dim + hidden_dim * (num_layers - 1) + self.padded_output_dim)
        self.weights = nn.Parameter(torch.zeros(self.num_parameters))
        self.reset_parameters()
        if torch.cuda.is_available():
            L.call("ntx_allocate_splitk", self.num_layers + 1)          # kept for API parity; libntx uses no side streams

    def cleanup(self):
        L.call("ntx_free_splitk")

    def __repr__(self):
        return f"FFMLP: input_dim={self.input_dim} output_dim={self.output_dim} hidden_dim={self.hidden_dim} num_layers={self.num_layers} activation={self.activation}"

    def reset_parameters(self):
        torch.manual_seed(42)                                            # like the reference (ffmlp/ffmlp.py:142): reseeds the global generator
        bound = math.sqrt(3 / self.hidden_dim)
        self.weights.data.uniform_(-bound, bound)

    def forward(self, inputs, force_grad=False):
        """[B, input_dim] -> [B, output_dim]; inference kernel unless training (or force_grad)"""
        y = ffmlp_forward(inputs, self.weights, self.input_dim, self.padded_output_dim, self.hidden_dim, self.num_layers, self.activation,
                          self.output_activation, not (self.training or force_grad), inputs.requires_grad)
        return y if self.padded_output_dim == self.output_dim else y[:, :self.output_dim]


# ====================================================================================================================== spherical harmonics
class SphericalHarmonicsOp(Function):
    """sh_encode (shencoder/sphere_harmonics.py:14-57): unit directions [B, 3] -> real SH basis [B, degree^2], always fp32"""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=_F32)
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        d = _dense(inputs, _F32)
        n, dim = d.shape
        basis = _scratch(d, n, degree * degree)
        jac = _scratch(d, n, dim * degree * degree) if calc_grad_inputs else _scratch(d, 1)
        L.call("ntx_sh_encode_forward", L.ptr(d), L.ptr(basis), n, dim, degree, int(calc_grad_inputs), L.ptr(jac), L.stream())
        ctx.meta = (n, dim, degree, bool(calc_grad_inputs))
        ctx.save_for_backward(d, jac)
        return basis

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        n, dim, degree, want_dx = ctx.meta
        if not want_dx:
            return None, None, None
        d, jac = ctx.saved_tensors
        d_dir = torch.zeros_like(d)
        L.call("ntx_sh_encode_backward", L.ptr(_dense(grad, _F32)), L.ptr(d), n, dim, degree, L.ptr(jac), L.ptr(d_dir), L.stream())
        return d_dir, None, None


sh_encode = SphericalHarmonicsOp.apply


class SHEncoder(nn.Module):
    """shencoder.SHEncoder (shencoder/sphere_harmonics.py:61-90)"""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert self.degree > 0 and self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        """[..., 3] in [-size, size] -> [..., degree^2]"""
        dirs = (inputs / size).reshape(-1, self.input_dim)
        return sh_encode(dirs, self.degree, dirs.requires_grad).reshape(*inputs.shape[:-1], self.output_dim)
