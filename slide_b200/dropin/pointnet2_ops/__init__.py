"""Drop-in replacement for the reference's `pointnet2_ops` package (pointnet2_ops_lib/pointnet2_ops):
same sub-modules, class names, constructor keywords, state-dict keys and call signatures; the native
layer underneath is libslide_b200.so (hand-written sm_100a CUDA) instead of the Kepler-era _ext.
Put `slide_b200/dropin` ahead of the reference's package on sys.path (slide_b200.install_dropin())."""
from . import _ext, attention, pointnet2_modules, pointnet2_utils  # noqa: F401
from ._version import __version__  # noqa: F401
