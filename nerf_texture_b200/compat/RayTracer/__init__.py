"""Drop-in for the reference's `RayTracer` package (external/RayTracer/RayTracer/__init__.py): `from RayTracer import RayTracer`."""
from .raytracer import RayTracer

__all__ = ["RayTracer"]
