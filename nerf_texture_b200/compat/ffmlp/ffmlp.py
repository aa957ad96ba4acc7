"""Drop-in `ffmlp.ffmlp` on libntx (B200 tcgen05 fully-fused MLP).

API mirror of the reference's ffmlp/ffmlp.py (`ffmlp_forward` :86, `FFMLP` :99): bias-free MLP with a flat fp32 `weights`
Parameter laid out [hidden x in | (num_layers-1) x hidden x hidden | 16 x hidden], each matrix row-major [out, in].
Differences that callers cannot observe: no `from turtle import ...` (needs tkinter), no per-call zero-row concatenation
(the kernel masks the ragged last 128-row tile itself), no side streams (allocate_splitk is a no-op).
"""
import math

import torch
import torch.nn as nn
from torch.autograd import Function

from nerf_texture_b200 import _lib as L


class _ffmlp_forward(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.half)
    def forward(ctx, inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, inference=False,
                calc_grad_inputs=False):
        B = inputs.shape[0]
        inputs = inputs.contiguous()
        weights = weights.contiguous()
        if inputs.dtype != torch.half:
            inputs = inputs.half()
        if weights.dtype != torch.half:
            weights = weights.half()
        outputs = torch.empty(B, output_dim, device=inputs.device, dtype=inputs.dtype)
        if not inference:
            forward_buffer = torch.empty(num_layers, B, hidden_dim, device=inputs.device, dtype=inputs.dtype)
            L.call("ntx_ffmlp_forward", L.ptr(inputs), L.ptr(weights), B, input_dim, output_dim, hidden_dim, num_layers, activation,
                   output_activation, L.ptr(forward_buffer), L.ptr(outputs), L.stream())
            ctx.save_for_backward(inputs, weights, outputs, forward_buffer)
            ctx.dims = (input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs)
        else:
            L.call("ntx_ffmlp_inference", L.ptr(inputs), L.ptr(weights), B, input_dim, output_dim, hidden_dim, num_layers, activation,
                   output_activation, None, L.ptr(outputs), L.stream())
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        B = grad.shape[0]
        grad = grad.contiguous()
        if grad.dtype != torch.half:
            grad = grad.half()
        inputs, weights, outputs, forward_buffer = ctx.saved_tensors
        input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs = ctx.dims
        grad_inputs = torch.zeros_like(inputs) if calc_grad_inputs else torch.zeros(1, device=grad.device, dtype=grad.dtype)
        grad_weights = torch.zeros_like(weights)
        backward_buffer = torch.empty(num_layers, B, hidden_dim, device=grad.device, dtype=grad.dtype)
        nbytes = L.lib().ntx_ffmlp_backward_workspace_bytes(input_dim, output_dim, hidden_dim, num_layers)
        workspace = torch.zeros(nbytes, dtype=torch.uint8, device=grad.device)
        L.call("ntx_ffmlp_backward", L.ptr(grad), L.ptr(inputs), L.ptr(weights), L.ptr(forward_buffer), B, input_dim, output_dim, hidden_dim,
               num_layers, activation, output_activation, int(calc_grad_inputs), L.ptr(backward_buffer), L.ptr(grad_inputs),
               L.ptr(grad_weights), L.ptr(workspace), L.stream())
        if calc_grad_inputs:
            return grad_inputs, grad_weights, None, None, None, None, None, None, None, None
        return None, grad_weights, None, None, None, None, None, None, None, None


ffmlp_forward = _ffmlp_forward.apply


def convert_activation(act):
    return {"relu": 0, "exponential": 1, "sine": 2, "sigmoid": 3, "squareplus": 4, "softplus": 5}.get(act, 6)


class FFMLP(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_dim, num_layers, activation="relu"):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.activation = convert_activation(activation)
        self.output_activation = convert_activation("none")
        self.tensorcore_width = 16

        assert hidden_dim in [16, 32, 64, 128, 256], f"FFMLP only support hidden_dim in [16, 32, 64, 128, 256], but got {hidden_dim}"
        assert input_dim > 0 and input_dim % 16 == 0, f"FFMLP input_dim should be 16 * m (m  > 0), but got {input_dim}"
        assert output_dim <= 16, f"FFMLP current only supports output dim <= 16, but got {output_dim}"
        assert num_layers >= 2, f"FFMLP num_layers should be larger than 2 (3 matmuls), but got {num_layers}"

        self.padded_output_dim = int(math.ceil(output_dim / 16)) * 16
        self.num_parameters = hidden_dim * (input_dim + hidden_dim * (num_layers - 1) + self.padded_output_dim)
        self.weights = nn.Parameter(torch.zeros(self.num_parameters))
        self.reset_parameters()
        L.call("ntx_allocate_splitk", self.num_layers + 1) if torch.cuda.is_available() else None

    def cleanup(self):
        L.call("ntx_free_splitk")

    def __repr__(self):
        return (f"FFMLP: input_dim={self.input_dim} output_dim={self.output_dim} hidden_dim={self.hidden_dim} "
                f"num_layers={self.num_layers} activation={self.activation}")

    def reset_parameters(self):
        torch.manual_seed(42)  # the reference reseeds the global generator here too (ffmlp.py:142)
        std = math.sqrt(3 / self.hidden_dim)
        self.weights.data.uniform_(-std, std)

    def forward(self, inputs, force_grad=False):
        # inputs [B, input_dim] -> [B, output_dim]
        B, C = inputs.shape
        outputs = ffmlp_forward(inputs, self.weights, self.input_dim, self.padded_output_dim, self.hidden_dim, self.num_layers,
                                self.activation, self.output_activation, (not self.training) and (not force_grad), inputs.requires_grad)
        if self.padded_output_dim != self.output_dim:
            outputs = outputs[:, :self.output_dim]
        return outputs
