"""sample_farthest_points with pytorch3d 0.7.0's signature and RNG behaviour
(call site: models/point_upsample_decoder.py:178-180)."""
import torch

from slide_b200 import lib as _l
from .utils import masked_gather


def sample_farthest_points(points, lengths=None, K=50, random_start_point=False):
    N, P, D = points.shape
    device = points.device
    if not points.is_cuda:
        raise RuntimeError("CPU not supported")
    if lengths is None:
        lengths_host = torch.full((N,), P, dtype=torch.int64)
        lengths_dev = None
    else:
        if lengths.shape != (N,):
            raise ValueError("points and lengths must have same batch dimension.")
        if int(lengths.max()) > P:
            raise ValueError("A value in lengths was too large.")
        lengths_host = lengths.detach().to("cpu", torch.int64)
        lengths_dev = lengths.to(device=device, dtype=torch.int64).contiguous()
    if isinstance(K, int):
        K_dev, maxK = None, K
    else:
        Kt = torch.tensor(K, dtype=torch.int64) if isinstance(K, list) else K.to(torch.int64)
        if Kt.shape != (N,):
            raise ValueError("K and points must have the same batch dimension")
        K_dev, maxK = Kt.to(device).contiguous(), int(Kt.max())
    start_dev = None
    if random_start_point:
        # pytorch3d 0.7.0 draws one CPU randint per cloud, in batch order, from the default generator
        start = torch.tensor([int(torch.randint(high=int(l), size=(1,)).item()) for l in lengths_host],
                             dtype=torch.int64)
        start_dev = start.to(device)
    pts = points.contiguous().float()
    idx = torch.empty(N, maxK, device=device, dtype=torch.int64)
    lib = _l.load()
    tmp = torch.empty(N, P, device=device) if P > lib.slide_fps_resident_max_points() else None  # large clouds only
    with torch.cuda.device(device):
        _l.check(lib.slide_sample_farthest_points_ws(_l.ptr(pts), N, P, D, _l.ptr(lengths_dev), _l.ptr(K_dev),
                                                     _l.ptr(start_dev), int(maxK), _l.ptr(idx), _l.ptr(tmp),
                                                     _l.stream_of(pts)), "sample_farthest_points")
    return masked_gather(points, idx), idx
