"""`shencoder.sphere_harmonics` of the drop-in package: the implementation lives in nerf_texture_b200/operators.py."""
from nerf_texture_b200.operators import SHEncoder, sh_encode  # noqa: F401
