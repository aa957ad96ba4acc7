"""`raymarching` drop-in package (libntx): the eleven functionals of the reference's raymarching extension."""
from .raymarching import (compact_rays, composite_rays, composite_rays_train, march_rays, march_rays_train,  # noqa: F401
                          march_rays_train_differentiable, morton3D, morton3D_invert, near_far_from_aabb, packbits, polar_from_ray)
from .raymarching import __all__  # noqa: F401
