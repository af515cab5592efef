import torch


def masked_gather(points, idx):
    """points (N,P,D), idx (N,K) or (N,M,K) with -1 padding -> gathered points, padded entries 0
    (pytorch3d.ops.utils.masked_gather; used at models/point_upsample_decoder.py:180)."""
    if points.shape[0] != idx.shape[0]:
        raise ValueError("points and idx must have the same batch dimension")
    D = points.shape[2]
    mask = idx.eq(-1)
    safe = idx.masked_fill(mask, 0)
    if idx.dim() == 3:
        N, M, K = idx.shape
        out = points[:, :, None].expand(-1, -1, K, -1).gather(1, safe[..., None].expand(-1, -1, -1, D))
    elif idx.dim() == 2:
        out = points.gather(1, safe[..., None].expand(-1, -1, D))
    else:
        raise ValueError("idx format is not supported %s" % repr(idx.shape))
    if mask.any():
        out = out.masked_fill(mask[..., None], 0.0)
    return out
