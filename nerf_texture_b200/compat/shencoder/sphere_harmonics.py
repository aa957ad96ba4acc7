"""Drop-in `shencoder.sphere_harmonics` on libntx (reference: shencoder/sphere_harmonics.py `sh_encode` :57, `SHEncoder` :61)."""
import torch
import torch.nn as nn
from torch.autograd import Function

from nerf_texture_b200 import _lib as L


class _sh_encoder(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)  # fp32 always, like the reference
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        B, input_dim = inputs.shape
        output_dim = degree ** 2
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        dy_dx = torch.empty(B, input_dim * output_dim, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs else \
            torch.empty(1, dtype=inputs.dtype, device=inputs.device)
        L.call("ntx_sh_encode_forward", L.ptr(inputs), L.ptr(outputs), B, input_dim, degree, int(calc_grad_inputs), L.ptr(dy_dx), L.stream())
        ctx.save_for_backward(inputs, dy_dx)
        ctx.dims = [B, input_dim, degree]
        ctx.calc_grad_inputs = calc_grad_inputs
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        if not ctx.calc_grad_inputs:
            return None, None, None
        grad = grad.contiguous().float()
        inputs, dy_dx = ctx.saved_tensors
        B, input_dim, degree = ctx.dims
        grad_inputs = torch.zeros_like(inputs)
        L.call("ntx_sh_encode_backward", L.ptr(grad), L.ptr(inputs), B, input_dim, degree, L.ptr(dy_dx), L.ptr(grad_inputs), L.stream())
        return grad_inputs, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert self.degree > 0 and self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        # inputs [..., 3] in [-size, size] -> [..., degree^2]
        inputs = inputs / size
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(inputs, self.degree, inputs.requires_grad)
        return outputs.reshape(prefix_shape + [self.output_dim])
