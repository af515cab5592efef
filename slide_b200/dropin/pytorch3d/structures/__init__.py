from .pointclouds import Pointclouds  # noqa: F401
