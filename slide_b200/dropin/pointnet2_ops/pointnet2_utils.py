"""`pointnet2_ops.pointnet2_utils` API (reference: pointnet2_ops_lib/pointnet2_ops/pointnet2_utils.py).

Autograd wrappers over the native ops (:62-304 there), QueryAndGroup (:307-448), GroupAll (:451-494),
group_knn (:497-524), count_to_mask / average_feature (:36-60).  Signatures, argument order and returned
shapes are the reference's; the bodies are written against libslide_b200.so through `_ext`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from pytorch3d.ops import knn
from . import _ext


def count_to_mask(count, K):
    """count (B,npoint) -> bool mask (B,npoint,K), True for the first count[b,i] slots."""
    slots = torch.arange(K, device=count.device, dtype=count.dtype)
    return slots.view(1, 1, K) < count.unsqueeze(-1)


def average_feature(feature, count, K):
    """Mean over the neighbour axis of (B,C,npoint,K); `count` is 'all' or (B,npoint) valid-neighbour counts."""
    if isinstance(count, str) and count == 'all':
        return feature.mean(dim=3)
    count = torch.clamp(count, min=1)
    keep = count_to_mask(count, K).unsqueeze(1)
    return (feature * keep).sum(dim=-1) / count.unsqueeze(1)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        idx = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad_out):
        return ()


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n_src = features.size(2)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n_src), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, grad_dist, grad_idx):
        return ()


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.n_src = features.size(2)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        g = _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.n_src)
        return g, torch.zeros_like(idx), torch.zeros_like(weight)


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n_src = features.size(2)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.group_points_grad(grad_out.contiguous(), idx, ctx.n_src), torch.zeros_like(idx)


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        # note the flipped order: the python API takes (xyz, new_xyz), the native op (new_xyz, xyz)
        idx, counts = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(idx)
        return idx, counts

    @staticmethod
    def backward(ctx, grad_out):
        return ()


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Neighbourhood lookup ('radius' ball query or 'nn' k-nearest) + grouping of xyz / features.

    forward(xyz (B,N,3), new_xyz (B,npoint,3), features (B,C,N)) -> (B, C + 3*f, npoint, K) where the
    coordinate block is [relative, (absolute), (centre)] in that order, features first."""

    def __init__(self, radius, nsample, use_xyz=True, include_abs_coordinate=False,
                 include_center_coordinate=False, neighbor_def='radius'):
        super().__init__()
        assert neighbor_def in ('radius', 'nn')
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.include_abs_coordinate = include_abs_coordinate
        self.include_center_coordinate = include_center_coordinate
        self.neighbor_def = neighbor_def
        self.neighbor_stats = None
        self.neighbor_num_quantile = None
        self.quantile = torch.tensor([0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1])

    def _neighbours(self, xyz, new_xyz, length):
        if self.neighbor_def == 'radius':
            if length is not None:
                raise Exception('radius neighbor definition has not supported point clouds with different lengths')
            return ball_query(self.radius, self.nsample, xyz, new_xyz)
        k = min(self.nsample, xyz.shape[1])
        idx = knn.knn_points(new_xyz, xyz, K=k, lengths2=length).idx.int()
        counts = torch.full(idx.shape[:2], float(k), device=new_xyz.device)
        if length is not None:
            counts = torch.minimum(counts, length.unsqueeze(1))
        return idx, counts

    def forward(self, xyz, new_xyz, features=None, subset=True, record_neighbor_stats=False,
                return_counts=False, length=None):
        idx, counts = self._neighbours(xyz, new_xyz, length)
        centre = new_xyz.transpose(1, 2).unsqueeze(-1)                      # (B,3,npoint,1)
        absolute = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,K)
        # a query outside the cloud may have an empty ball: it then stands for itself with zero features
        patch_empty = (not subset) and self.neighbor_def == 'radius'
        if patch_empty:
            found = (counts > 0).float()[:, None, :, None].detach()
            absolute = found * absolute + (1 - found) * centre
        coords = [absolute - centre]
        if self.include_abs_coordinate:
            coords.append(absolute)
        if self.include_center_coordinate:
            coords.append(centre.expand(-1, -1, -1, absolute.shape[3]))
        grouped_xyz = coords[0] if len(coords) == 1 else torch.cat(coords, dim=1)

        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            out = grouped_xyz
        else:
            grouped = grouping_operation(features, idx)
            if patch_empty:
                grouped = found * grouped
            out = torch.cat([grouped, grouped_xyz], dim=1) if self.use_xyz else grouped

        if record_neighbor_stats:
            with torch.no_grad():
                c = counts.float()
                self.neighbor_stats = torch.stack([c.min(), c.mean(), c.max()])
                self.neighbor_num_quantile = torch.quantile(c, self.quantile.to(c.device)).long()
        return (out, counts) if return_counts else out


class GroupAll(nn.Module):
    """Treat the whole cloud as one group: (B,C,N) -> (B,C(+3),1,N)."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        g_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return g_xyz
        g_feat = features.unsqueeze(2)
        return torch.cat([g_feat, g_xyz], dim=1) if self.use_xyz else g_feat


def group_knn(x, y, features_at_y, K, transpose=False):
    """For every point of x (B,N1,3) collect its K nearest points of y (B,N2,3) and build
    [feature_j, d2, w, y_j, y_j - x_i, x_i] (C+11 channels; d2 is the SQUARED distance pytorch3d returns,
    w = 1/(d2+1e-8) normalised over K).  transpose=True: features_at_y is (B,C,N2) and the result is
    (B,C+11,N1,K); otherwise (B,N2,C) -> (B,N1,K,C+11)."""
    feats = features_at_y.transpose(1, 2).contiguous() if transpose else features_at_y
    d2, idx, y_nn = knn.knn_points(x, y, K=K, return_nn=True)
    f_nn = knn.knn_gather(feats, idx)
    centre = x.unsqueeze(2).expand(-1, -1, K, -1)
    d2 = d2.unsqueeze(3)
    inv = 1.0 / (d2 + 1e-8)
    w = inv / inv.sum(dim=2, keepdim=True)
    out = torch.cat([f_nn, d2, w, y_nn, y_nn - centre, centre], dim=3)
    return out.permute(0, 3, 1, 2) if transpose else out
