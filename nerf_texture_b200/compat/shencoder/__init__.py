"""`shencoder` drop-in package (libntx): real spherical-harmonics direction encoding."""
from nerf_texture_b200.operators import SHEncoder, sh_encode  # noqa: F401
