"""`pointnet2_ops.attention` API (reference: pointnet2_ops_lib/pointnet2_ops/attention.py:6-155)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MyGroupNorm(nn.Module):
    """GroupNorm over the leading floor(C/G)*G channels; trailing (position) channels pass through."""

    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups = num_groups
        self.num_channels = num_channels - num_channels % num_groups
        self.group_norm = nn.GroupNorm(self.num_groups, self.num_channels)

    def forward(self, x):
        n = self.num_channels
        if x.shape[1] == n:
            return self.group_norm(x)
        return torch.cat([self.group_norm(x[:, :n]), x[:, n:]], dim=1)


def count_to_mask(count, K):
    slots = torch.arange(K, device=count.device, dtype=count.dtype)
    return slots.view(1, 1, K) < count.unsqueeze(-1)


def _score_net(c_in, c_mid, c_out, norm):
    layers = [nn.ReLU(inplace=True)]
    if norm:
        layers.append(MyGroupNorm(min(32, c_in), c_in))
    layers += [nn.Conv2d(c_in, c_mid, kernel_size=1), nn.ReLU(inplace=True)]
    if norm:
        layers.append(MyGroupNorm(min(32, c_mid), c_mid))
    layers.append(nn.Conv2d(c_mid, c_out, kernel_size=1))
    return nn.Sequential(*layers)


def _value_net(c_in, c_out, norm, act):
    layers = [nn.Conv2d(c_in, c_out, kernel_size=1)]
    if act:
        if norm:
            layers.append(MyGroupNorm(min(32, c_out), c_out))
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class AttentionModule(nn.Module):
    """Per-channel softmax attention over the K neighbours of every point.

    feat (B,C_in1,N) is the query, grouped_feat (B,C_in2,N,K) the keys, grouped_feat_out (B,C_out,N,K) the
    values; count is 'all' or (B,N) numbers of valid neighbours.  Returns (B,C_out,N)."""

    def __init__(self, C_in1, C_in2, C1, C2, C_out, attention_bn=True, transform_grouped_feat_out=True,
                 last_activation=True):
        super().__init__()
        C1, C2 = max(C1, 32), max(C2, 32)
        self.feat_conv = nn.Conv2d(C_in1, C1, kernel_size=1)
        self.grouped_feat_conv = nn.Conv2d(C_in2, C2, kernel_size=1)
        self.weight_conv = _score_net(C1 + C2, min(C1 + C2, C_out), C_out, attention_bn)
        self.transform_grouped_feat_out = transform_grouped_feat_out
        if transform_grouped_feat_out:
            self.feat_out_conv = _value_net(C_out, C_out, attention_bn, last_activation)

    def forward(self, feat, grouped_feat, grouped_feat_out, count):
        K = grouped_feat.shape[-1]
        q = self.feat_conv(feat.unsqueeze(-1)).expand(-1, -1, -1, K)
        k = self.grouped_feat_conv(grouped_feat)
        scores = self.weight_conv(torch.cat([q, k], dim=1))
        if not (isinstance(count, str) and count == 'all'):
            keep = count_to_mask(torch.clamp(count, min=1), K).unsqueeze(1).float()
            scores = scores * keep + (-1e9) * (1 - keep)
        w = F.softmax(scores, dim=-1)
        v = self.feat_out_conv(grouped_feat_out) if self.transform_grouped_feat_out else grouped_feat_out
        return (v * w).sum(dim=-1)


class GlobalAttentionModule(nn.Module):
    """All-pairs variant: every point attends to every point of the same cloud. feat (B,C+add,N) -> (B,C,N)."""

    def __init__(self, C, additional_dim=0, attention_bn=True, last_activation=True):
        super().__init__()
        self.key_conv = nn.Conv2d(C + additional_dim, C, kernel_size=1)
        self.query_conv = nn.Conv2d(C + additional_dim, C, kernel_size=1)
        self.value_conv = _value_net(C + additional_dim, C, attention_bn, last_activation)
        self.weight_conv = _score_net(2 * C, C, C, attention_bn)

    def forward(self, feat):
        N = feat.shape[2]
        x = feat.unsqueeze(-1)
        key = self.key_conv(x).squeeze(-1)
        query = self.query_conv(x).squeeze(-1)
        value = self.value_conv(x).squeeze(-1)
        pair = torch.cat([query.unsqueeze(-1).expand(-1, -1, -1, N), key.unsqueeze(-2).expand(-1, -1, N, -1)], dim=1)
        w = F.softmax(self.weight_conv(pair), dim=-1)
        # value is indexed by the QUERY point (reference attention.py:153), so this reduces to `value`
        return (value.unsqueeze(-1) * w).sum(dim=-1)
