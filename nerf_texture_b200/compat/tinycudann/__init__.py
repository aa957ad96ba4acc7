"""Minimal `tinycudann` stand-in on libntx, so that nerf/network_curvedfield.py, nerf/network_tcnn.py, tools/map.py and the light
models of the reference (which import the un-vendored, un-pinned tiny-cuda-nn bindings: readme.md:44) can run on the
B200 kernels.  Covers what those callers use:

    tcnn.Network(n_input_dims, n_output_dims, network_config)       otype FullyFusedMLP | CutlassMLP, n_neurons in {16,32,64,128},
                                                                     activation ReLU|None|Exponential|Sine|Sigmoid|Squareplus|Softplus
    tcnn.Encoding(n_input_dims, encoding_config)                    otype SphericalHarmonics{degree} | HashGrid{...} | Identity
    tcnn.NetworkWithInputEncoding(n_input_dims, n_output_dims, encoding_config, network_config)

PARITY IS UNPINNED: tiny-cuda-nn is not in /root/reference and no version is recorded, so this shim follows tcnn's documented
conventions (inputs of SH / HashGrid in [0,1]; bias-free fully-fused MLP; n_hidden_layers hidden layers = n_hidden_layers + 1
matmuls; fp16 outputs) on top of the in-tree ops' arithmetic, and is tested only for self-consistency.
"""
import math

import torch
import torch.nn as nn

_ACT = {"relu": 0, "exponential": 1, "sine": 2, "sigmoid": 3, "squareplus": 4, "softplus": 5, "none": 6}


def _act_id(name):
    return _ACT.get(str(name).lower(), 6)


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        from ffmlp.ffmlp import ffmlp_forward
        self._fwd = ffmlp_forward
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.hidden = int(network_config.get("n_neurons", 64))
        self.n_hidden_layers = int(network_config.get("n_hidden_layers", 2))
        assert self.hidden in (16, 32, 64, 128, 256), "n_neurons must be 16, 32, 64, 128 or 256"
        assert self.n_hidden_layers >= 1 and n_output_dims <= 16, "shim supports >= 1 hidden layer and <= 16 outputs"
        self.activation = _act_id(network_config.get("activation", "ReLU"))
        self.output_activation = _act_id(network_config.get("output_activation", "None"))
        self.padded_in = int(math.ceil(n_input_dims / 16)) * 16
        n = self.hidden * (self.padded_in + self.hidden * (self.n_hidden_layers - 1) + 16)
        g = torch.Generator().manual_seed(seed)
        params = torch.empty(n)
        # xavier-uniform per matrix (tcnn's default initialisation)
        o = 0
        for fan_in, fan_out in [(self.padded_in, self.hidden)] + [(self.hidden, self.hidden)] * (self.n_hidden_layers - 1) + [(self.hidden, 16)]:
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            params[o:o + fan_in * fan_out] = (torch.rand(fan_in * fan_out, generator=g) * 2 - 1) * lim
            o += fan_in * fan_out
        self.params = nn.Parameter(params)

    def forward(self, x):
        prefix = x.shape[:-1]
        x = x.reshape(-1, self.n_input_dims)
        if self.padded_in != self.n_input_dims:
            x = torch.nn.functional.pad(x, (0, self.padded_in - self.n_input_dims))
        inference = not (torch.is_grad_enabled() and (self.params.requires_grad or x.requires_grad))
        y = self._fwd(x, self.params, self.padded_in, 16, self.hidden, self.n_hidden_layers, self.activation, self.output_activation, inference, x.requires_grad)
        return y[:, :self.n_output_dims].reshape(*prefix, self.n_output_dims)


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, dtype=torch.half):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.dtype = dtype
        otype = str(encoding_config.get("otype", "Identity")).lower()
        self.otype = otype
        if otype == "sphericalharmonics":
            from shencoder import SHEncoder
            self.enc = SHEncoder(input_dim=n_input_dims, degree=int(encoding_config.get("degree", 4)))
            self.n_output_dims = self.enc.output_dim
        elif otype in ("hashgrid", "grid"):
            from gridencoder import GridEncoder
            self.enc = GridEncoder(input_dim=n_input_dims, num_levels=int(encoding_config.get("n_levels", 16)),
                                   level_dim=int(encoding_config.get("n_features_per_level", 2)),
                                   per_level_scale=float(encoding_config.get("per_level_scale", 2.0)),
                                   base_resolution=int(encoding_config.get("base_resolution", 16)),
                                   log2_hashmap_size=int(encoding_config.get("log2_hashmap_size", 19)), gridtype="hash", align_corners=False)
            self.n_output_dims = self.enc.output_dim
        elif otype == "identity":
            self.enc = None
            self.n_output_dims = n_input_dims
        else:
            raise NotImplementedError("tinycudann shim: encoding otype %r" % encoding_config.get("otype"))

    def forward(self, x):
        if self.otype == "sphericalharmonics":
            return self.enc(x * 2 - 1).to(self.dtype)            # tcnn takes directions in [0,1] (network_curvedfield.py:319)
        if self.otype in ("hashgrid", "grid"):
            return self.enc(x * 2 - 1, bound=1).to(self.dtype)   # GridEncoder maps [-1,1] -> [0,1] itself
        return x.to(self.dtype)


class NetworkWithInputEncoding(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        self.encoding = Encoding(n_input_dims, encoding_config)
        self.network = Network(self.encoding.n_output_dims, n_output_dims, network_config, seed=seed)
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims

    def forward(self, x):
        return self.network(self.encoding(x))
