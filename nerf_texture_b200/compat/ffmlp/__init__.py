"""`ffmlp` drop-in package (libntx): fully-fused fp16 MLP on tcgen05."""
from nerf_texture_b200.operators import FFMLP, convert_activation, ffmlp_forward  # noqa: F401
