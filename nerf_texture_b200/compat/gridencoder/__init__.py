"""`gridencoder` drop-in package (libntx): multiresolution hash / tiled grid encoder."""
from nerf_texture_b200.operators import GridEncoder, grid_encode  # noqa: F401
from .grid_clustering import GridEncoder_clustering  # noqa: F401
