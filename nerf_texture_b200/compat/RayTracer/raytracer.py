"""`RayTracer(vertices, triangles).trace(rays_o, rays_d, inplace=False)` with the reference's signature and return values
(external/RayTracer/RayTracer/raytracer.py:7-68) on libntx's mesh handle (nerf_texture_b200/mesh.py, csrc/mesh.cu).

Differences, on purpose: meshes of <= 8 triangles are not padded with far-away dummy faces (raytracer.py:17-24 works around a limit of
the reference's 4-wide BVH that this tree does not have — the results are the same, the dummies can never be hit); the launch goes to
the current torch stream."""
from nerf_texture_b200.mesh import Mesh


class RayTracer:
    def __init__(self, vertices, triangles):
        # vertices: np.ndarray or tensor [N, 3]; triangles: np.ndarray or tensor [M, 3]
        self.impl = Mesh(vertices, triangles)

    def trace(self, rays_o, rays_d, inplace=False):
        # rays_o, rays_d: float tensors [..., 3] (moved to the mesh's GPU if they are not there)
        # returns positions [..., 3], face_normals [..., 3], depth [...], face_idx [N] (int64, -1 where nothing was hit)
        return self.impl.trace(rays_o, rays_d, inplace=inplace)
