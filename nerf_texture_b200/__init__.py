"""nerf_texture_b200 — B200 (sm_100a) kernels for NeRF-Texture's per-ray-sample hot path.

    import nerf_texture_b200 as ntx
    ntx.install()            # puts drop-in `gridencoder`, `ffmlp`, `shencoder`, `raymarching`, `tinycudann`
                             # packages (nerf_texture_b200/compat) at the front of sys.path
    from gridencoder import GridEncoder      # now the B200 implementation

The compute lives in lib/libntx.so (C ABI, include/ntx.h), built by `python -m nerf_texture_b200.build`.
"""
import os
import sys

__version__ = "0.1.0"
COMPAT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")


def install():
    """Make the reference's operator package names resolve to the B200 implementations."""
    if COMPAT_DIR not in sys.path:
        sys.path.insert(0, COMPAT_DIR)
    return COMPAT_DIR
