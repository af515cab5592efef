from . import knn, utils  # noqa: F401
from .knn import knn_gather, knn_points  # noqa: F401
from .sample_farthest_points import sample_farthest_points  # noqa: F401
