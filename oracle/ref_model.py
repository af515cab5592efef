"""CPU restatement (plain torch ops on CPU tensors, functional, state-dict driven) of the reference's
network forwards and samplers on the hot path.

TEST INFRASTRUCTURE ONLY -- never imported by slide_b200/.  It exists because the reference's python tree
does not travel to the GPU box; tests/golden/make_golden.py checks every function here against the real
reference modules (imported from /root/reference in the build container) on seeded inputs, and the GPU
parity tests then compare the CUDA path with this file on the same inputs.

Citations (relative to /root/reference; OPS = pointnet2_ops_lib/pointnet2_ops):
  my_group_norm        OPS/pointnet2_modules.py:24-42, OPS/attention.py:6-23
  shared_stage / mlp   OPS/pointnet2_modules.py:44-176   (bn_first=False configurations only)
  attention            OPS/attention.py:35-96
  query_and_group_nn   OPS/pointnet2_utils.py:307-448 ('nn' neighbour definition)
  group_knn            OPS/pointnet2_utils.py:497-524
  sa_module            OPS/pointnet2_modules.py:212-292
  knn_fp_module        OPS/pointnet2_modules.py:771-873
  feature_map_module   OPS/pointnet2_modules.py:640-663
  calc_t_emb           pointnet2/models/pointnet2_ssg_sem.py:14-31
  cloud_condition_net  pointnet2/models/pointnet2_with_pcld_condition.py:286-489 (no condition cloud)
  point_upsample       pointnet2/models/point_upsample_module.py:4-46
  upsample_points / propagate_feature / decode
                       pointnet2/models/point_upsample_decoder.py:106-190, keypoint_decoder.py:25-36,
                       autoencoder.py:42-45
  pnet2stage           pointnet2/models/pnet.py:7-40
  encoder_net          pointnet2/models/pointnet2_feature_extractor.py:143-218 (PointNet2Encoder.forward)
  kl_latent            pointnet2/data_utils/distributions.py:4-43 + point_upsample_decoder.py:95-104
  encode               pointnet2/models/autoencoder.py:37-40, point_upsample_decoder.py:106-147 (KL branch)
  position_schedule / position_sampling   pointnet2/util.py:167-259
  latent_schedule / denoising_step / denoise_and_reconstruct
                       pointnet2/diffusion_utils/diffusion.py:12-39,58-95,158-208,346-404
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops


class Params(object):
    """Prefix view over a state dict."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, key):
        return self.sd[self.prefix + key]

    def has(self, key):
        return (self.prefix + key) in self.sd

    def sub(self, name):
        return Params(self.sd, self.prefix + name + ".")


def _act(x, name):
    return F.relu(x) if name == "relu" else x * torch.sigmoid(x)


def my_group_norm(x, P):
    """x (B,C,H,W); P holds group_norm.{weight,bias}; groups = min(32, C), tail channels untouched."""
    w, b = P["group_norm.weight"], P["group_norm.bias"]
    C = x.shape[1]
    G = min(32, C)
    n = w.numel()
    assert n == C - C % G
    if n == C:
        return F.group_norm(x, G, w, b, 1e-5)
    return torch.cat([F.group_norm(x[:, :n], G, w, b, 1e-5), x[:, n:]], dim=1)


def conv1x1(x, P):
    return F.conv2d(x, P["weight"], P["bias"] if P.has("bias") else None)


def shared_stack(x, P, act):
    """build_shared_mlp with bn_first=False: repeated (conv, [MyGroupNorm], act)."""
    i = 0
    while P.has("%d.weight" % i):
        x = conv1x1(x, P.sub(str(i)))
        i += 1
        if P.has("%d.group_norm.weight" % i):
            x = my_group_norm(x, P.sub(str(i)))
            i += 1
        x = _act(x, act)
        i += 1
    return x


def mlp_plus_t_emb(x, P, res_connect, act, t_emb=None, cond=None, cond2=None):
    if P.has("first_conv.weight"):
        x = conv1x1(x, P.sub("first_conv"))
    h = shared_stack(x, P.sub("first_mlp"), act)
    if P.has("fc.weight"):
        h = h + F.linear(t_emb, P["fc.weight"], P["fc.bias"])[:, :, None, None]
    h = shared_stack(h, P.sub("second_mlp"), act)
    if P.has("fc_condition.weight"):
        h = h + F.linear(cond, P["fc_condition.weight"], P["fc_condition.bias"])[:, :, None, None]
    if P.has("rest_mlp.0.weight"):
        h = shared_stack(h, P.sub("rest_mlp"), act)
    if P.has("fc_second_condition.weight"):
        h = h + F.linear(cond2, P["fc_second_condition.weight"], P["fc_second_condition.bias"])[:, :, None, None]
    if res_connect:
        h = h + (conv1x1(x, P.sub("res_connect")) if P.has("res_connect.weight") else x)
    return h


def attention(feat, grouped, value_in, P, last_activation):
    """feat (B,C1,N), grouped (B,C2,N,K), value_in (B,Co,N,K) -> (B,Co,N); attention_bn=True,
    transform_grouped_feat_out=True, count == K everywhere ('nn' neighbours), so the mask is a no-op."""
    K = grouped.shape[-1]
    q = conv1x1(feat.unsqueeze(-1), P.sub("feat_conv")).expand(-1, -1, -1, K)
    k = conv1x1(grouped, P.sub("grouped_feat_conv"))
    s = F.relu(torch.cat([q, k], dim=1))
    s = my_group_norm(s, P.sub("weight_conv.1"))
    s = conv1x1(s, P.sub("weight_conv.2"))
    s = my_group_norm(F.relu(s), P.sub("weight_conv.4"))
    s = conv1x1(s, P.sub("weight_conv.5"))
    w = F.softmax(s, dim=-1)
    v = conv1x1(value_in, P.sub("feat_out_conv.0"))
    if last_activation:
        v = F.relu(my_group_norm(v, P.sub("feat_out_conv.1")))
    return (v * w).sum(dim=-1)


def _gather_cols(feat, idx):
    """feat (B,C,N), idx (B,np,K) int -> (B,C,np,K)."""
    B, C, N = feat.shape
    _, npnt, K = idx.shape
    flat = idx.reshape(B, 1, npnt * K).expand(-1, C, -1).long()
    return feat.gather(2, flat).reshape(B, C, npnt, K)


def query_and_group_nn(xyz, new_xyz, features, nsample, include_abs=True, include_center=True):
    """-> (B, C+3(+3)(+3), npoint, K): [features_j, xyz_j - c_i, xyz_j, c_i]."""
    K = min(nsample, xyz.shape[1])
    idx = ops.knn_points(new_xyz, xyz, K=K).idx.int()
    absolute = _gather_cols(xyz.transpose(1, 2).contiguous(), idx)
    centre = new_xyz.transpose(1, 2).unsqueeze(-1)
    parts = [absolute - centre]
    if include_abs:
        parts.append(absolute)
    if include_center:
        parts.append(centre.expand(-1, -1, -1, K))
    g_xyz = torch.cat(parts, dim=1)
    if features is None:
        return g_xyz
    return torch.cat([_gather_cols(features, idx), g_xyz], dim=1)


def group_knn(x, y, feats_at_y, K):
    """x (B,N1,3), y (B,N2,3), feats_at_y (B,C,N2) -> (B,C+11,N1,K)."""
    d2, idx, _ = ops.knn_points(x, y, K=K)
    y_nn = ops.knn_gather(y, idx)                                        # (B,N1,K,3)
    f_nn = ops.knn_gather(feats_at_y.transpose(1, 2).contiguous(), idx)  # (B,N1,K,C)
    centre = x.unsqueeze(2).repeat(1, 1, K, 1)
    d2 = d2.unsqueeze(3)
    inv = 1.0 / (d2 + 1e-8)
    w = inv / torch.sum(inv, dim=2, keepdim=True)
    out = torch.cat([f_nn, d2, w, y_nn, y_nn - centre, centre], dim=3)
    return out.transpose(2, 3).transpose(1, 2)


def sa_module(xyz, features, P, npoint, nsample, cfg, t_emb, cond, cond2=None):
    act = cfg.get("activation", "relu")
    if xyz.shape[1] <= npoint:
        new_xyz, q_feat = xyz, features
    else:
        pick = ops.furthest_point_sampling(xyz.contiguous(), npoint).long()
        new_xyz = xyz.gather(1, pick[:, :, None].expand(-1, -1, 3)).contiguous()
        q_feat = features.gather(2, pick[:, None, :].expand(-1, features.shape[1], -1))
    grouped = query_and_group_nn(xyz, new_xyz, features, nsample, cfg["include_abs_coordinate"],
                                 cfg.get("include_center_coordinate", False))
    h = mlp_plus_t_emb(grouped, P.sub("mlps.0"), cfg["res_connect"], act, t_emb=t_emb, cond=cond, cond2=cond2)
    out = attention(q_feat, grouped, h, P.sub("attention_modules.0"), cfg["attention_setting"]["last_activation"])
    return new_xyz, out


def knn_fp_module(unknown, known, unknow_feats, known_feats, P, K, cfg, t_emb, cond):
    act = cfg.get("activation", "relu")
    grouped = group_knn(unknown, known, known_feats, K)
    h = mlp_plus_t_emb(grouped, P.sub("mlp1"), cfg["res_connect"], act)
    spread = attention(unknow_feats, grouped, h, P.sub("attention_module"),
                       cfg["attention_setting"]["last_activation"])
    h = torch.cat([spread, unknow_feats, unknown.transpose(1, 2)], dim=1).unsqueeze(-1)
    h = mlp_plus_t_emb(h, P.sub("mlp2"), cfg["res_connect"], act, t_emb=t_emb, cond=cond)
    return h.squeeze(-1)


def feature_map_module(xyz, features, new_xyz, q_feat, P, nsample, cfg):
    act = cfg.get("activation", "relu")
    grouped = query_and_group_nn(xyz, new_xyz, features, nsample, cfg["include_abs_coordinate"],
                                 cfg.get("include_center_coordinate", False))
    h = mlp_plus_t_emb(grouped, P.sub("mlp"), cfg["res_connect"], act)
    return attention(q_feat, grouped, h, P.sub("attention_module"), cfg["attention_setting"]["last_activation"])


def calc_t_emb(ts, t_emb_dim):
    half = t_emb_dim // 2
    freq = torch.exp(torch.arange(half) * -(np.log(10000) / (half - 1)))
    arg = ts.unsqueeze(1) * freq
    return torch.cat((torch.sin(arg), torch.cos(arg)), 1)


def _swish(x):
    return x * torch.sigmoid(x)


def cloud_condition_net(pointcloud, P, cfg, ts=None, label=None):
    """PointNet2CloudCondition.forward without a condition cloud (include_local_feature=False,
    include_global_feature=False): the DDPM denoisers and the decoder-level feature extractors."""
    assert not cfg.get("include_local_feature", True) and not cfg.get("include_global_feature", False)
    assert not cfg["bn_first"] and cfg.get("bn", True)
    arch = cfg["architecture"]
    assert arch["neighbor_definition"] == "nn" and arch.get("use_knn_FP", False)
    if cfg["attach_position_to_input_feature"]:
        pointcloud = torch.cat([pointcloud, pointcloud[:, :, 0:3]], dim=2)
    xyz = pointcloud[..., 0:3].contiguous()
    features = pointcloud[..., 3:].transpose(1, 2).contiguous() if pointcloud.size(-1) > 3 else None
    t_emb = None
    if ts is not None and cfg["include_t"]:
        t_emb = calc_t_emb(ts, cfg["t_dim"])
        t_emb = _swish(F.linear(t_emb, P["fc_t1.weight"], P["fc_t1.bias"]))
        t_emb = _swish(F.linear(t_emb, P["fc_t2.weight"], P["fc_t2.bias"]))
    cond = None
    if label is not None and cfg["include_class_condition"]:
        cond = F.embedding(label, P["class_emb.weight"])
    l_xyz, l_feat = [xyz], [features]
    for i, (npoint, nsample) in enumerate(zip(arch["npoint"], arch["nsample"])):
        nx, nf = sa_module(l_xyz[i], l_feat[i], P.sub("SA_modules.%d" % i), npoint, nsample, cfg, t_emb, cond)
        l_xyz.append(nx)
        l_feat.append(nf)
    n_fp = len(arch["decoder_feature_dim"]) - 1
    for i in range(-1, -(n_fp + 1), -1):
        l_feat[i - 1] = knn_fp_module(l_xyz[i - 1], l_xyz[i], l_feat[i - 1], l_feat[i],
                                      P.sub("FP_modules.%d" % (n_fp + i)), arch.get("K", 3), cfg, t_emb, cond)
    out = l_feat[0]
    if cfg.get("transform_output", True):
        out = torch.cat([out, xyz.transpose(1, 2)], dim=1)
        out = F.conv1d(out, P["fc_lyaer.0.weight"], P["fc_lyaer.0.bias"] if P.has("fc_lyaer.0.bias") else None)
        out = F.group_norm(out, 32, P["fc_lyaer.1.weight"], P["fc_lyaer.1.bias"], 1e-5)
        out = _act(out, cfg.get("activation", "relu"))
        out = F.conv1d(out, P["fc_lyaer.3.weight"], P["fc_lyaer.3.bias"])
    return out.transpose(1, 2)


# ------------------------------------------------------------------------------------ autoencoder decode
def point_upsample(coarse, displacement, factor, scale):
    """first_refine_coarse_points=False branch: every coarse point spawns `factor` children."""
    B, N, Fdim = coarse.shape
    grid = (displacement * (1 / np.sqrt(factor))).view(B, N, factor, Fdim)
    return (coarse.unsqueeze(2) + grid * scale).reshape(B, -1, Fdim).contiguous()


def upsample_points(final_feature, new_xyz, P, cfg, start_idx=None):
    """PointUpsampleDecoder.upsample_points; start_idx (B,) int64 pins pytorch3d's random FPS start."""
    up = cfg["upsampling_setting"]
    assert not up["first_refine_coarse_points"]
    x = torch.cat([final_feature, new_xyz], dim=2).transpose(1, 2)
    split = F.conv1d(x, P["fc_layer.weight"], P["fc_layer.bias"]).transpose(1, 2)
    in_dim = cfg.get("in_position_and_normal_dim", cfg["out_dim"])
    coarse = new_xyz[:, :, 0:in_dim]
    if in_dim < cfg["out_dim"]:
        pad = torch.zeros(coarse.shape[0], coarse.shape[1], cfg["out_dim"] - in_dim)
        coarse = torch.cat([coarse, pad], dim=2)
    pts = point_upsample(coarse, split, up["point_upsample_factor"], up["output_scale_factor"])
    n_out = up["num_output_points"]
    if pts.shape[1] > n_out:
        _, sel = ops.sample_farthest_points(pts[:, :, 0:3].contiguous(), K=n_out, random_start_point=True,
                                            start_idx=start_idx)
        pts = ops.masked_gather(pts, sel)
    return pts


def propagate_feature(xyz, features, new_xyz, P, cfg, label):
    """PointUpsampleDecoder.propagate_feature for decoder levels (PointNet2CloudCondition extractor, no KL)."""
    out = cloud_condition_net(new_xyz, P.sub("feature_extractor"), cfg, ts=None, label=label)
    mapped = feature_map_module(xyz, features.transpose(1, 2).contiguous(), new_xyz[:, :, 0:3].contiguous(),
                                out.transpose(1, 2), P.sub("feature_mapper"),
                                cfg["feature_mapper_setting"]["nsample"], cfg)
    return torch.cat([out, mapped.transpose(1, 2)], dim=2)


def decode(keypoint, feature, P, decoder_cfgs, label, start_idx_list=None):
    """PointAutoencoder.decode: keypoints (B,16,3) + latent features (B,16,48) -> (B,2048,6).
    decoder_cfgs = [level1, level2, level3] pointnet_config dicts; start_idx_list = one (B,) int64 tensor per
    level (None = draw like pytorch3d does)."""
    sl = start_idx_list or [None] * len(decoder_cfgs)
    new_xyz = upsample_points(feature, keypoint, P.sub("keypoint_encoder"), decoder_cfgs[0], sl[0])
    xyzs, feats = [keypoint[:, :, 0:3], new_xyz], [feature]
    for i, cfg in enumerate(decoder_cfgs[1:]):
        Pd = P.sub("decoder.decoders.%d" % i)
        f = propagate_feature(xyzs[i][:, :, 0:3], feats[i], xyzs[i + 1], Pd, cfg, label)
        xyzs.append(upsample_points(f, xyzs[i + 1], Pd, cfg, sl[i + 1]))
        feats.append(f)
    return xyzs[-1], xyzs


# ------------------------------------------------------------------------------------ autoencoder encode
def pnet2stage(x, P, act="relu"):
    """Pnet2Stage.forward with remove_last_activation=False: x (B,C,N) -> global feature (B, mlp2[-1])."""
    f = mlp_plus_t_emb(x.unsqueeze(-1), P.sub("mlp1"), False, act)
    g = F.max_pool2d(f, kernel_size=[f.size(2), 1]).expand(-1, -1, f.size(2), -1)
    f = mlp_plus_t_emb(torch.cat([f, g], dim=1), P.sub("mlp2"), False, act)
    return F.max_pool2d(f, kernel_size=[f.size(2), 1]).squeeze(-1).squeeze(-1)


def encoder_net(pointcloud, P, cfg, label=None):
    """PointNet2Encoder.forward (no timestep): -> (out (B, np_last, C_last), l_xyz, l_features)."""
    assert not cfg["bn_first"] and cfg.get("bn", True) and not cfg["include_t"]
    arch = cfg["architecture"]
    assert arch["neighbor_definition"] == "nn"
    in_fea = cfg["in_fea_dim"]
    if cfg["attach_position_to_input_feature"]:
        pointcloud = torch.cat([pointcloud, pointcloud[:, :, 0:3]], dim=2)
    xyz = pointcloud[..., 0:3].contiguous()
    features = pointcloud[..., 3:].transpose(1, 2).contiguous() if pointcloud.size(-1) > 3 else None
    class_emb = F.embedding(label, P["class_emb.weight"]) if (label is not None and cfg["include_class_condition"]) else None
    if cfg.get("include_global_feature", False):
        assert not cfg.get("global_feature_remove_last_activation", True)
        g_in = torch.cat([xyz, pointcloud[:, :, 3:3 + in_fea]], dim=2) if in_fea > 0 else xyz
        cond, cond2 = pnet2stage(g_in.transpose(1, 2), P.sub("global_pnet")), class_emb
    else:
        cond, cond2 = class_emb, None
    l_xyz, l_feat = [xyz], [features]
    for i, (npoint, nsample) in enumerate(zip(arch["npoint"], arch["nsample"])):
        nx, nf = sa_module(l_xyz[i], l_feat[i], P.sub("SA_modules.%d" % i), npoint, nsample, cfg, None, cond, cond2)
        l_xyz.append(nx)
        l_feat.append(nf)
    return l_feat[-1].transpose(1, 2), l_xyz, l_feat


def kl_latent(params, noise=None):
    """DiagonalGaussianDistribution over the channel dim of (B,N,2C): mode (noise None) or mean + std * noise."""
    mean, logvar = torch.chunk(params, 2, dim=2)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def encode(pointcloud, keypoint, P, enc_cfg, kp_cfg, label, noises=None):
    """PointAutoencoder.encode with apply_kl_regularization=True.  noises = (n1 (B,16,C1), n2 (B,16,C2)) replaces the
    two CPU torch.randn draws of DiagonalGaussianDistribution.sample; None = posterior mode (sample_posterior=False)."""
    out, l_xyz, _ = encoder_net(pointcloud, P.sub("encoder"), enc_cfg, label)
    Pk = P.sub("keypoint_encoder")
    cfgx = dict(kp_cfg)
    f1, _, _ = encoder_net(keypoint, Pk.sub("feature_extractor"), cfgx, label)
    f1 = kl_latent(f1, None if noises is None else noises[0])
    mapped = feature_map_module(l_xyz[-1], out.transpose(1, 2).contiguous(), keypoint[:, :, 0:3].contiguous(),
                                f1.transpose(1, 2), Pk.sub("feature_mapper"), kp_cfg["feature_mapper_setting"]["nsample"],
                                kp_cfg).transpose(1, 2)
    mapped = kl_latent(mapped, None if noises is None else noises[1])
    return torch.cat([f1, mapped], dim=2)


# ------------------------------------------------------------------------------------ samplers
def position_schedule(T, beta_0, beta_T):
    """calc_diffusion_hyperparams: fp32 torch, sequential in-place products."""
    Beta = torch.linspace(beta_0, beta_T, T)
    Alpha = 1 - Beta
    Alpha_bar = Alpha + 0
    Beta_tilde = Beta + 0
    for t in range(1, T):
        Alpha_bar[t] *= Alpha_bar[t - 1]
        Beta_tilde[t] *= (1 - Alpha_bar[t - 1]) / (1 - Alpha_bar[t])
    return {"T": T, "Beta": Beta, "Alpha": Alpha, "Alpha_bar": Alpha_bar, "Sigma": torch.sqrt(Beta_tilde)}


def position_sampling(net_fn, x_T, noises, dh, t_start=None, n_steps=None):
    """util.sampling: x_T (B,N,3); noises[t] is the std_normal added after step t (t > 0)."""
    Alpha, Alpha_bar, Sigma = dh["Alpha"], dh["Alpha_bar"], dh["Sigma"]
    x = x_T
    t_start = dh["T"] - 1 if t_start is None else t_start
    stop = -1 if n_steps is None else t_start - n_steps
    for t in range(t_start, stop, -1):
        ts = t * torch.ones((x.shape[0],))
        eps = net_fn(x, ts)
        x = (x - (1 - Alpha[t]) / torch.sqrt(1 - Alpha_bar[t]) * eps) / torch.sqrt(Alpha[t])
        if t > 0:
            x = x + Sigma[t] * noises[t]
    return x


# ---- FastDPM STEP sampler of the position DDPM (pointnet2/util_fastdpmv2.py) ---------------------------------
def fast_step_steps(S, dcfg, schedule="linear"):
    """get_STEP_step, util_fastdpmv2.py:239-258."""
    if schedule == "linear":
        c = (dcfg["T"] - 1.0) / (S - 1.0)
        list_tau = [np.floor(i * c) for i in range(S)]
    else:
        assert schedule == "quadratic"
        list_tau = np.linspace(0, np.sqrt(dcfg["T"] * 0.8), S) ** 2
    return [int(s) for s in list_tau]


def fast_sampling(net_fn, x_T, noises, dcfg, method, length, schedule, kappa):
    """fast_sampling_function_v2 -> STEP_sampling (util_fastdpmv2.py:384-452, 455-478).  noises[i] is the std_normal
    drawn in iteration i (drawn in every iteration, also the last where sigma = 0).  (method 'var': the reference's
    VAR_sampling trips its own `assert abs(tau) < 0.1` with the shipped schedule, so there is nothing to restate.)"""
    assert method == "step"
    dh = position_schedule(dcfg["T"], dcfg["beta_0"], dcfg["beta_T"])
    Alpha_bar = dh["Alpha_bar"]
    taus = sorted(fast_step_steps(length, dcfg, schedule), reverse=True)
    x = x_T.clone()
    for i, tau in enumerate(taus):
        eps = net_fn(x, tau * torch.ones((x.shape[0],)))
        if i == length - 1:
            alpha_next, sigma = torch.tensor(1.0), torch.tensor(0.0)
        else:
            alpha_next = Alpha_bar[taus[i + 1]]
            sigma = kappa * torch.sqrt((1 - alpha_next) / (1 - Alpha_bar[tau]) * (1 - Alpha_bar[tau] / alpha_next))
        x *= torch.sqrt(alpha_next / Alpha_bar[tau])
        c = torch.sqrt(1 - alpha_next - sigma ** 2) - torch.sqrt(1 - Alpha_bar[tau]) * torch.sqrt(alpha_next / Alpha_bar[tau])
        x += c * eps + sigma * noises[i]
    return x, taus


def latent_schedule(cfg):
    """Diffusion.init_diffusion_parameters: float64 numpy, linear betas, fixedsmall log-variance."""
    assert cfg["beta_schedule"] == "linear" and cfg.get("model_var_type", "fixedsmall") == "fixedsmall"
    betas = np.linspace(cfg["beta_start"], cfg["beta_end"], cfg["num_diffusion_timesteps"], dtype=np.float64)
    alphas = 1.0 - betas
    acp = np.cumprod(alphas, axis=0)
    acp_prev = np.append(1.0, acp[:-1])
    post_var = betas * (1.0 - acp_prev) / (1.0 - acp)
    return {
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / acp),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / acp - 1),
        "posterior_mean_coef1": betas * np.sqrt(acp_prev) / (1.0 - acp),
        "posterior_mean_coef2": (1.0 - acp_prev) * np.sqrt(alphas) / (1.0 - acp),
        "logvar": np.log(np.maximum(post_var, 1e-20)),
        "T": betas.shape[0],
        "data_clamp_range": cfg["data_clamp_range"],
    }


def _extract(a, t, ndim):
    out = torch.gather(torch.tensor(a).float(), 0, t.long())
    return out.reshape((t.shape[0],) + (1,) * (ndim - 1))


def denoising_step(x, t, net_fn, sch, noise, complete_x0=None, keypoint_mask=None):
    eps = net_fn(x, t)
    x0 = _extract(sch["sqrt_recip_alphas_cumprod"], t, x.dim()) * x - \
        _extract(sch["sqrt_recipm1_alphas_cumprod"], t, x.dim()) * eps
    if sch["data_clamp_range"] > 0:
        x0 = torch.clamp(x0, -sch["data_clamp_range"], sch["data_clamp_range"])
    if complete_x0 is not None:
        m = keypoint_mask
        while m.dim() < complete_x0.dim():
            m = m.unsqueeze(2)
        x0 = x0 * m + complete_x0 * (1 - m)
    mean = _extract(sch["posterior_mean_coef1"], t, x.dim()) * x0 + \
        _extract(sch["posterior_mean_coef2"], t, x.dim()) * x
    logvar = _extract(sch["logvar"], t, x.dim())
    mask = (1 - (t == 0).float()).reshape((x.shape[0],) + (1,) * (x.dim() - 1))
    return (mean + mask * torch.exp(0.5 * logvar) * noise).float(), x0


def latent_denoise(net_fn, x_T, keypoint, noises, sch, t_start=None, n_steps=None, complete_x0=None,
                   keypoint_mask=None):
    """The loop of LatentDiffusion.denoise_and_reconstruct (keypoint_conditional=True); noises[t] replaces the
    randn_like of step t.  Returns x (B,N,3+F) with the keypoints written back."""
    kd = keypoint.shape[2]
    x = x_T
    t_start = sch["T"] - 1 if t_start is None else t_start
    stop = -1 if n_steps is None else t_start - n_steps
    for i in range(t_start, stop, -1):
        t = torch.ones(x.shape[0]) * i
        x = torch.cat([keypoint, x[:, :, kd:]], dim=2)
        x, _ = denoising_step(x, t, net_fn, sch, noises[i], complete_x0, keypoint_mask)
    return torch.cat([keypoint, x[:, :, kd:]], dim=2)
