"""Minimal stand-in for pytorch3d==0.7.0 (the reference's pinned dependency, environment.yml:117):
only the entry points SLIDE's sampling/decode path imports, backed by libslide_b200.so."""
__version__ = "0.7.0+slide_b200"
from . import ops, structures  # noqa: F401
