"""`pointnet2_ops.pointnet2_modules` API (reference: pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py).

Set-abstraction (PointnetSAModule[MSG], :212-454), three-NN and kNN feature propagation
(PointnetFPModule :457-588, PointnetKnnFPModule :666-873), cross-set FeatureMapModule (:591-663) and the
shared per-point MLP with timestep / condition injection (Mlp_plus_t_emb :72-176).  Constructor keywords,
attribute names (hence state-dict keys) and forward signatures follow the reference so that its model
files and checkpoints load unchanged; neighbour search, sampling and grouping run in libslide_b200.so.
"""
import copy
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from pointnet2_ops import pointnet2_utils
from pointnet2_ops.attention import AttentionModule, GlobalAttentionModule


def swish(x):
    return x * torch.sigmoid(x)


class Swish(nn.Module):
    def forward(self, x):
        return swish(x)


class MyGroupNorm(nn.Module):
    """GroupNorm on the leading floor(C/G)*G channels, identity on the rest (appended coordinates)."""

    def __init__(self, num_groups, num_channels):
        super().__init__()
        assert num_channels >= num_groups
        self.num_groups = num_groups
        self.num_channels = num_channels - num_channels % num_groups
        self.group_norm = nn.GroupNorm(self.num_groups, self.num_channels)

    def forward(self, x):
        n = self.num_channels
        if x.shape[1] == n:
            return self.group_norm(x)
        return torch.cat([self.group_norm(x[:, :n]), x[:, n:]], dim=1)


def _act(name):
    return nn.ReLU(True) if name == 'relu' else Swish()


def build_shared_mlp(mlp_spec: List[int], bn: bool = True, bn_first: bool = False, bias: bool = False,
                     activation: str = 'relu'):
    """1x1-conv stack; each stage is conv->norm->act, or norm->act->conv when bn_first."""
    assert activation in ['relu', 'swish']
    layers = []
    for c_in, c_out in zip(mlp_spec[:-1], mlp_spec[1:]):
        conv = nn.Conv2d(c_in, c_out, kernel_size=1, bias=bias)
        c_norm = c_in if bn_first else c_out
        tail = ([MyGroupNorm(min(32, c_norm), c_norm)] if bn else []) + [_act(activation)]
        layers += (tail + [conv]) if bn_first else ([conv] + tail)
    return nn.Sequential(*layers)


class Mlp_plus_t_emb(nn.Module):
    """Shared MLP over (B,C,npoint,K) with additive injections: fc(t_emb) after stage 1, fc_condition after
    stage 2, fc_second_condition after the last stage, optional input conv and residual branch."""

    def __init__(self, mlp_spec, bn, t_dim=128, include_t=True, bn_first=False, bias=False, first_conv=False,
                 first_conv_in_channel=0, res_connect=False, include_condition=False, condition_dim=128,
                 include_second_condition=False, second_condition_dim=128, activation='relu'):
        super().__init__()
        assert len(mlp_spec) >= 3
        if include_second_condition:
            assert len(mlp_spec) >= 4
        self.include_t = include_t
        self.include_condition = include_condition
        self.include_second_condition = include_second_condition
        self.first_conv_bool = first_conv
        self.res_connect_bool = res_connect
        if include_t:
            self.fc = nn.Linear(t_dim, mlp_spec[1])
        if include_condition:
            self.fc_condition = nn.Linear(condition_dim, mlp_spec[2])
        if include_second_condition:
            self.fc_second_condition = nn.Linear(second_condition_dim, mlp_spec[-1])
        if first_conv:
            self.first_conv = nn.Conv2d(first_conv_in_channel, mlp_spec[0], kernel_size=1, bias=bias)
        if res_connect:
            same = mlp_spec[0] == mlp_spec[-1]
            self.res_connect = None if same else nn.Conv2d(mlp_spec[0], mlp_spec[-1], kernel_size=1, bias=bias)
        kw = dict(bn_first=bn_first, bias=bias, activation=activation)
        self.first_mlp = build_shared_mlp(mlp_spec[0:2], bn, **kw)
        self.second_mlp = build_shared_mlp(mlp_spec[1:3], bn, **kw)
        self.rest_mlp = build_shared_mlp(mlp_spec[2:], bn, **kw) if len(mlp_spec) > 3 else None

    @staticmethod
    def _inject(h, fc, emb, want, what):
        if want:
            if emb is None:
                raise Exception('Should pass %s to the forward function' % what)
            return h + fc(emb)[:, :, None, None]
        if emb is not None:
            raise Exception('This module does not include %s but it is given' % what)
        return h

    def forward(self, feature, t_emb=None, condition_emb=None, second_condition_emb=None):
        if self.first_conv_bool:
            feature = self.first_conv(feature)
        h = self.first_mlp(feature)
        h = self._inject(h, getattr(self, 'fc', None), t_emb, self.include_t, 't_emb')
        h = self.second_mlp(h)
        h = self._inject(h, getattr(self, 'fc_condition', None), condition_emb, self.include_condition,
                         'condition_emb')
        if self.rest_mlp is not None:
            h = self.rest_mlp(h)
        h = self._inject(h, getattr(self, 'fc_second_condition', None), second_condition_emb,
                         self.include_second_condition, 'second_condition_emb')
        if self.res_connect_bool:
            h = h + (feature if self.res_connect is None else self.res_connect(feature))
        return h


def pooling_features(feature, count=None, pooling='max'):
    """(B,C,npoint,K) -> (B,C,npoint) by max / masked mean / half-max-half-mean over K."""
    assert pooling in ['max', 'avg', 'avg_max', 'max_avg']
    K = feature.size(3)
    if pooling == 'max':
        return feature.max(dim=3)[0]
    if pooling == 'avg':
        return pointnet2_utils.average_feature(feature, count, K)
    half = int(feature.shape[1] / 2)
    return torch.cat([feature[:, :half].max(dim=3)[0],
                      pointnet2_utils.average_feature(feature[:, half:], count, K)], dim=1)


def _xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate):
    return (3 + (3 if include_abs_coordinate else 0) + (3 if include_center_coordinate else 0)) if use_xyz else 0


def _attention_from(setting, c_query, c_key, c_out):
    return AttentionModule(c_query, c_key, c_query, c_key, c_out, attention_bn=setting['attention_bn'],
                           transform_grouped_feat_out=setting['transform_grouped_feat_out'],
                           last_activation=setting['last_activation'])


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward(self, xyz: torch.Tensor, features: Optional[torch.Tensor], t_emb: torch.Tensor = None,
                condition_emb: torch.Tensor = None, second_condition_emb: torch.Tensor = None,
                subset: bool = True, record_neighbor_stats: bool = False, pooling: str = 'max',
                length: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """xyz (B,N,3), features (B,C,N) -> new_xyz (B,npoint,3), new_features (B,sum_k mlps[k][-1],npoint)."""
        assert self.npoint is not None
        centre_feat = None
        if xyz.shape[1] <= self.npoint:
            new_xyz = xyz  # nothing to drop: keep every point, in order
            centre_feat = features
        else:
            pick = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
            new_xyz = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), pick)
            new_xyz = new_xyz.transpose(1, 2).contiguous()
            if self.use_attention_module:
                centre_feat = pointnet2_utils.gather_operation(features, pick)

        t_emb = t_emb if self.include_t else None
        condition_emb = condition_emb if self.include_condition else None
        second_condition_emb = second_condition_emb if self.include_second_condition else None
        outs = []
        for i, grouper in enumerate(self.groupers):
            grouped, count = grouper(xyz, new_xyz, features, subset=subset,
                                     record_neighbor_stats=record_neighbor_stats, return_counts=True, length=length)
            h = self.mlps[i](grouped, t_emb=t_emb, condition_emb=condition_emb,
                             second_condition_emb=second_condition_emb)
            if self.use_attention_module:
                h = self.attention_modules[i](centre_feat, grouped, h, count)
            else:
                h = pooling_features(h, count=count, pooling=pooling)
            if self.use_global_attention_module:
                h = self.global_attention_modules[i](torch.cat([h, new_xyz.transpose(1, 2)], dim=1))
            outs.append(h)
        return new_xyz, torch.cat(outs, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set abstraction with one grouper + MLP (+ attention) per scale."""

    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True, t_dim=128, include_t=False,
                 include_abs_coordinate=False, include_center_coordinate=False, bn_first=False, bias=False,
                 first_conv=False, first_conv_in_channel=0, res_connect=False, include_condition=False,
                 condition_dim=128, include_second_condition=False, second_condition_dim=128,
                 neighbor_def='radius', activation='relu', attention_setting=None, global_attention_setting=None):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.include_t, self.t_dim = include_t, t_dim
        self.include_condition, self.condition_dim = include_condition, condition_dim
        self.include_second_condition, self.second_condition_dim = include_second_condition, second_condition_dim
        self.use_attention_module = bool(attention_setting and attention_setting['use_attention_module'])
        self.use_global_attention_module = bool(
            global_attention_setting and global_attention_setting['use_global_attention_module'])
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        self.attention_modules = nn.ModuleList() if self.use_attention_module else None
        self.global_attention_modules = nn.ModuleList() if self.use_global_attention_module else None
        extra = _xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate)
        for radius, nsample, spec in zip(radii, nsamples, mlps):
            self.groupers.append(
                pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz,
                                              include_abs_coordinate=include_abs_coordinate,
                                              include_center_coordinate=include_center_coordinate,
                                              neighbor_def=neighbor_def)
                if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            c_query = first_conv_in_channel if first_conv else spec[0]
            # like the reference, the coordinate channels accumulate into the caller's spec / running width
            if first_conv:
                first_conv_in_channel += extra
            else:
                spec[0] += extra
            c_key = first_conv_in_channel if first_conv else spec[0]
            self.mlps.append(Mlp_plus_t_emb(
                spec, bn, t_dim=t_dim, include_t=include_t, bn_first=bn_first, bias=bias, first_conv=first_conv,
                first_conv_in_channel=first_conv_in_channel, res_connect=res_connect,
                include_condition=include_condition, condition_dim=condition_dim,
                include_second_condition=include_second_condition, second_condition_dim=second_condition_dim,
                activation=activation))
            if self.use_attention_module:
                self.attention_modules.append(_attention_from(attention_setting, c_query, c_key, spec[-1]))
            if self.use_global_attention_module:
                self.global_attention_modules.append(GlobalAttentionModule(
                    spec[-1], additional_dim=3, attention_bn=global_attention_setting['attention_bn'],
                    last_activation=global_attention_setting['last_activation']))


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction."""

    def __init__(self, mlp, npoint=None, radius=None, nsample=None, bn=True, use_xyz=True, t_dim=128,
                 include_t=False, include_abs_coordinate=False, include_center_coordinate=False, bn_first=False,
                 bias=False, first_conv=False, first_conv_in_channel=0, res_connect=False,
                 include_condition=False, condition_dim=128, include_second_condition=False,
                 second_condition_dim=128, neighbor_def='radius', activation='relu', attention_setting=None,
                 global_attention_setting=None):
        super().__init__(
            npoint=npoint, radii=[radius], nsamples=[nsample], mlps=[mlp], bn=bn, use_xyz=use_xyz, t_dim=t_dim,
            include_t=include_t, include_abs_coordinate=include_abs_coordinate,
            include_center_coordinate=include_center_coordinate, bn_first=bn_first, bias=bias,
            first_conv=first_conv, first_conv_in_channel=first_conv_in_channel, res_connect=res_connect,
            include_condition=include_condition, condition_dim=condition_dim,
            include_second_condition=include_second_condition, second_condition_dim=second_condition_dim,
            neighbor_def=neighbor_def, activation=activation, attention_setting=attention_setting,
            global_attention_setting=global_attention_setting)


class PointnetFPModule(nn.Module):
    """Feature propagation by inverse-distance interpolation over the three nearest known points."""

    def __init__(self, mlp, bn=True, t_dim=128, include_t=False, bn_first=False, bias=False, first_conv=False,
                 first_conv_in_channel=0, res_connect=False, include_condition=False, condition_dim=128,
                 include_second_condition=False, second_condition_dim=128, include_grouper=False, radius=0,
                 nsample=32, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=False,
                 neighbor_def='radius', activation='relu'):
        super().__init__()
        self.include_t, self.t_dim = include_t, t_dim
        self.include_condition, self.condition_dim = include_condition, condition_dim
        self.include_second_condition, self.second_condition_dim = include_second_condition, second_condition_dim
        self.include_grouper = include_grouper
        if include_grouper:
            extra = _xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate)
            if first_conv:
                first_conv_in_channel += extra
            else:
                mlp[0] += extra
            self.grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz,
                                                         include_abs_coordinate=include_abs_coordinate,
                                                         include_center_coordinate=include_center_coordinate,
                                                         neighbor_def=neighbor_def)
        self.mlp = Mlp_plus_t_emb(
            mlp, bn, t_dim=t_dim, include_t=include_t, bn_first=bn_first, bias=bias, first_conv=first_conv,
            first_conv_in_channel=first_conv_in_channel, res_connect=res_connect,
            include_condition=include_condition, condition_dim=condition_dim,
            include_second_condition=include_second_condition, second_condition_dim=second_condition_dim,
            activation=activation)

    def forward(self, unknown, known, unknow_feats, known_feats, t_emb=None, condition_emb=None,
                second_condition_emb=None, record_neighbor_stats=False, pooling='max'):
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m) -> (B,mlp[-1],n)."""
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            inv = 1.0 / (dist + 1e-8)
            weight = inv / inv.sum(dim=2, keepdim=True)
            spread = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            spread = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        h = torch.cat([spread, unknow_feats], dim=1) if unknow_feats is not None else spread
        count = None
        if self.include_grouper:
            h, count = self.grouper(unknown, unknown, h, subset=True, record_neighbor_stats=record_neighbor_stats,
                                    return_counts=True)
        else:
            h = h.unsqueeze(-1)
        h = self.mlp(h, t_emb=t_emb if self.include_t else None,
                     condition_emb=condition_emb if self.include_condition else None,
                     second_condition_emb=second_condition_emb if self.include_second_condition else None)
        if self.include_grouper:
            return pooling_features(h, count=count, pooling=pooling)
        return h.squeeze(-1)


class FeatureMapModule(nn.Module):
    """Carry features living on `xyz` over to another point set `new_xyz` (neighbour MLP + attention/pool)."""

    def __init__(self, mlp, radius, K, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=False,
                 bn=True, bn_first=True, bias=True, res_connect=True, first_conv=False, first_conv_in_channel=0,
                 neighbor_def='radius', activation='relu', attention_setting=None, query_feature_dim=None):
        super().__init__()
        self.use_attention_module = bool(attention_setting and attention_setting['use_attention_module'])
        extra = _xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate)
        if first_conv:
            first_conv_in_channel += extra
        else:
            mlp[0] += extra
        self.mlp = Mlp_plus_t_emb(mlp, bn, include_t=False, bn_first=bn_first, bias=bias, first_conv=first_conv,
                                  first_conv_in_channel=first_conv_in_channel, res_connect=res_connect,
                                  include_condition=False, activation=activation)
        self.mapper = pointnet2_utils.QueryAndGroup(radius, K, use_xyz=use_xyz,
                                                    include_abs_coordinate=include_abs_coordinate,
                                                    include_center_coordinate=include_center_coordinate,
                                                    neighbor_def=neighbor_def)
        if self.use_attention_module:
            c_key = first_conv_in_channel if first_conv else mlp[0]
            self.attention_module = _attention_from(attention_setting, query_feature_dim, c_key, mlp[-1])

    def forward(self, xyz, features, new_xyz, subset=False, record_neighbor_stats=True, pooling='max',
                features_at_new_xyz=None):
        """xyz (B,N,3), features (B,C,N), new_xyz (B,npoint,3) [, features_at_new_xyz (B,C',npoint) = query]
        -> (B,mlp[-1],npoint)."""
        grouped, count = self.mapper(xyz, new_xyz, features, subset=subset,
                                     record_neighbor_stats=record_neighbor_stats, return_counts=True)
        h = self.mlp(grouped)
        if self.use_attention_module:
            return self.attention_module(features_at_new_xyz, grouped, h, count)
        return pooling_features(h, count=count, pooling=pooling)


class PointnetKnnFPModule(nn.Module):
    """Feature propagation through the K nearest known points: mlp1 on the augmented neighbour features,
    attention (query = skip features) or pooling over K, then mlp2 on [propagated, skip, xyz]."""

    def __init__(self, mlp1, mlp2, K, bn=True, t_dim=128, include_t=False, bn_first=False, bias=False,
                 first_conv=False, first_conv_in_channel1=0, first_conv_in_channel2=0, res_connect=False,
                 include_condition=False, condition_dim=128, include_second_condition=False,
                 second_condition_dim=128, include_grouper=False, radius=0, nsample=32, use_xyz=True,
                 include_abs_coordinate=True, include_center_coordinate=False, neighbor_def='radius',
                 activation='relu', attention_setting=None, global_attention_setting=None):
        super().__init__()
        self.include_t, self.t_dim = include_t, t_dim
        self.include_condition, self.condition_dim = include_condition, condition_dim
        self.include_second_condition, self.second_condition_dim = include_second_condition, second_condition_dim
        self.K = K
        if first_conv:
            first_conv_in_channel1 += 11
        else:
            mlp1[0] += 11
        # mlp1 never sees t; its single optional condition slot carries the SECOND condition
        self.mlp1 = Mlp_plus_t_emb(mlp1, bn, t_dim=t_dim, include_t=False, bn_first=bn_first, bias=bias,
                                   first_conv=first_conv, first_conv_in_channel=first_conv_in_channel1,
                                   res_connect=res_connect, include_condition=include_second_condition,
                                   condition_dim=second_condition_dim, activation=activation)
        self.use_attention_module = bool(attention_setting and attention_setting['use_attention_module'])
        if self.use_attention_module:
            c_skip = (first_conv_in_channel2 if first_conv else mlp2[0]) - mlp1[-1]
            c_key = first_conv_in_channel1 if first_conv else mlp1[0]
            self.attention_module = _attention_from(attention_setting, c_skip, c_key, mlp1[-1])
        self.include_grouper = include_grouper
        extra = _xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate) if include_grouper else 3
        if first_conv:
            first_conv_in_channel2 += extra
        else:
            mlp2[0] += extra
        if include_grouper:
            self.grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz,
                                                         include_abs_coordinate=include_abs_coordinate,
                                                         include_center_coordinate=include_center_coordinate,
                                                         neighbor_def=neighbor_def)
        self.mlp2 = Mlp_plus_t_emb(mlp2, bn, t_dim=t_dim, include_t=include_t, bn_first=bn_first, bias=bias,
                                   first_conv=first_conv, first_conv_in_channel=first_conv_in_channel2,
                                   res_connect=res_connect, include_condition=include_condition,
                                   condition_dim=condition_dim, activation=activation)
        self.use_global_attention_module = bool(
            global_attention_setting and global_attention_setting['use_global_attention_module'])
        if self.use_global_attention_module:
            self.global_attention_module = GlobalAttentionModule(
                mlp2[-1], additional_dim=3, attention_bn=global_attention_setting['attention_bn'],
                last_activation=global_attention_setting['last_activation'])

    def forward(self, unknown, known, unknow_feats, known_feats, t_emb=None, condition_emb=None,
                second_condition_emb=None, record_neighbor_stats=False, pooling='max'):
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m) -> (B,mlp2[-1],n)."""
        if self.use_attention_module or self.use_global_attention_module:
            assert known is not None and unknown is not None
            if self.use_global_attention_module:
                assert not self.include_grouper
        if known is not None:
            grouped = pointnet2_utils.group_knn(unknown, known, known_feats, self.K, transpose=True)
            h = self.mlp1(grouped, t_emb=None,
                          condition_emb=second_condition_emb if self.include_second_condition else None)
            if self.use_attention_module:
                spread = self.attention_module(unknow_feats, grouped, h, count='all')
            else:
                spread = pooling_features(h, count='all', pooling=pooling)
        else:
            spread = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        h = torch.cat([spread, unknow_feats], dim=1) if unknow_feats is not None else spread
        count = None
        if self.include_grouper:
            h, count = self.grouper(unknown, unknown, h, subset=True, record_neighbor_stats=record_neighbor_stats,
                                    return_counts=True)
        else:
            h = torch.cat([h, unknown.transpose(1, 2)], dim=1).unsqueeze(-1)
        h = self.mlp2(h, t_emb=t_emb if self.include_t else None,
                      condition_emb=condition_emb if self.include_condition else None)
        if self.include_grouper:
            return pooling_features(h, count=count, pooling=pooling)
        h = h.squeeze(-1)
        if self.use_global_attention_module:
            h = self.global_attention_module(torch.cat([h, unknown.transpose(1, 2)], dim=1))
        return h
