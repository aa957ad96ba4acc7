"""`gridencoder.grid` of the drop-in package: the implementation lives in nerf_texture_b200/operators.py (HashGridOp, GridEncoder)."""
from nerf_texture_b200.operators import GRIDTYPE_ID, GridEncoder, grid_encode, hashgrid_level_offsets  # noqa: F401

level_offsets = hashgrid_level_offsets
