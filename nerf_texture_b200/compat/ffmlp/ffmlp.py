"""`ffmlp.ffmlp` of the drop-in package: the implementation lives in nerf_texture_b200/operators.py (FusedMLPOp, FFMLP)."""
from nerf_texture_b200.operators import FFMLP, convert_activation, ffmlp_forward  # noqa: F401
