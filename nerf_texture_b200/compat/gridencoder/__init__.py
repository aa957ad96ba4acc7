from .grid import GridEncoder, grid_encode
from .grid_clustering import GridEncoder_clustering
