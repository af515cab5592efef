"""slide_b200 -- SLIDE's diffusion-sampling + autoencoder-decode hot path as hand-written sm_100a CUDA.

Layout:
  csrc/, libslide_b200.so   CUDA kernels + the C ABI of include/slide_b200.h (built by slide_b200.build)
  lib.py                    ctypes binding of that ABI
  dropin/pointnet2_ops      drop-in for the reference's `pointnet2_ops` package (same python API)
  dropin/pytorch3d          the few pytorch3d 0.7.0 entry points the path imports
  program.py / engine.py    host side of the fused network programs (denoiser, DDPM loops, decode)
"""
import os
import sys

__version__ = "0.1.0"

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install_dropin():
    """Put the drop-in `pointnet2_ops` and `pytorch3d` packages first on sys.path, so that the reference's
    unmodified model / sampling scripts import them instead of pointnet2_ops_lib and pytorch3d."""
    for name in ("pointnet2_ops", "pytorch3d"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(DROPIN_DIR):
            raise RuntimeError("%s is already imported from %s" % (name, getattr(mod, "__file__", "?")))
    if DROPIN_DIR not in sys.path:
        sys.path.insert(0, DROPIN_DIR)
    return DROPIN_DIR
