"""knn_points / knn_gather with pytorch3d 0.7.0's python signatures
(call sites: pointnet2_ops/pointnet2_utils.py:370,506-507)."""
from collections import namedtuple

import torch

from slide_b200 import lib as _l

_KNN = namedtuple("KNN", "dists idx knn")


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False,
               return_sorted=True):
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same point dimension.")
    if norm != 2:
        raise ValueError("slide_b200's knn_points implements norm=2 only")
    if not p1.is_cuda:
        raise RuntimeError("CPU not supported")
    p1c = p1.contiguous().float()
    p2c = p2.contiguous().float()
    B, P1, D = p1c.shape
    P2 = p2c.shape[1]
    l1 = lengths1.to(device=p1.device, dtype=torch.int64).contiguous() if lengths1 is not None else None
    l2 = lengths2.to(device=p1.device, dtype=torch.int64).contiguous() if lengths2 is not None else None
    dists = torch.empty(B, P1, K, device=p1.device, dtype=torch.float32)
    idx = torch.empty(B, P1, K, device=p1.device, dtype=torch.int64)
    with torch.cuda.device(p1.device):
        _l.check(_l.load().slide_knn_points(_l.ptr(p1c), _l.ptr(p2c), B, P1, P2, D, _l.ptr(l1), _l.ptr(l2), int(K),
                                            _l.ptr(dists), _l.ptr(idx), _l.stream_of(p1c)), "knn_points")
    nn = knn_gather(p2, idx, lengths2) if return_nn else None
    if torch.is_grad_enabled() and (p1.requires_grad or p2.requires_grad):
        # pytorch3d's knn_points is differentiable w.r.t. p1 / p2 through dists (group_knn feeds d2 and 1/(d2+1e-8)
        # into the features, pointnet2_utils.py:506-517).  The kernel's distances carry no graph, so the backward is
        # attached here: grad_p1 = 2 (p1 - p2[idx]) g, grad_p2 = scatter of -2 (p1 - p2[idx]) g; forward values stay
        # the kernel's (bit-exact with the no-grad path).
        dists = _KnnDists.apply(p1, p2, idx, dists, l2, l1)
    return _KNN(dists=dists, idx=idx, knn=nn)


class _KnnDists(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p1, p2, idx, dists, lengths2, lengths1=None):
        ctx.save_for_backward(p1, p2, idx)
        ctx.lengths2, ctx.lengths1 = lengths2, lengths1
        return dists.clone()

    @staticmethod
    def backward(ctx, g):
        p1, p2, idx = ctx.saved_tensors
        N, P1, K = idx.shape
        D = p1.shape[2]
        if ctx.lengths2 is not None:  # padded neighbour slots (k >= lengths2) carry no gradient
            g = g * (torch.arange(K, device=g.device)[None, None, :] < ctx.lengths2[:, None, None]).to(g.dtype)
        if ctx.lengths1 is not None:  # query rows beyond lengths1 are padding (dists 0, idx 0): no gradient either
            g = g * (torch.arange(P1, device=g.device)[None, :, None] < ctx.lengths1[:, None, None]).to(g.dtype)
        nb = p2[:, :, None].expand(-1, -1, K, -1).gather(1, idx[:, :, :, None].expand(-1, -1, -1, D))
        diff = 2.0 * (p1[:, :, None, :] - nb) * g[:, :, :, None]            # (N,P1,K,D)
        g1 = diff.sum(2) if ctx.needs_input_grad[0] else None
        g2 = None
        if ctx.needs_input_grad[1]:
            g2 = torch.zeros_like(p2).scatter_add_(1, idx.reshape(N, P1 * K, 1).expand(-1, -1, D),
                                                   (-diff).reshape(N, P1 * K, D))
        return g1, g2, None, None, None, None


def knn_gather(x, idx, lengths=None):
    """x (N,M,U), idx (N,L,K) -> (N,L,K,U)."""
    N, M, U = x.shape
    _N, L, K = idx.shape
    if N != _N:
        raise ValueError("x and idx must have same batch dimension.")
    out = x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[:, :, :, None].expand(-1, -1, -1, U))
    if lengths is not None and bool((lengths < K).any()):
        mask = lengths[:, None] <= torch.arange(K, device=x.device)[None]
        out = out.masked_fill(mask[:, None, :, None].expand(-1, L, -1, U), 0.0)
    return out
