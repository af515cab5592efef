"""Golden vectors for the parts of the reference's `pointnet2_ops` python package that the shipped sampling configs do
not exercise and for the clamp branch of the feature-DDPM update (a5), produced by the
REAL reference classes (build container only: needs /root/reference; the C oracle stands in for `_ext`):

    python tests/golden/make_golden_dropin.py      ->  tests/golden/golden_dropin.npz

  qg_*     QueryAndGroup(neighbor_def='radius'), subset=False with queries whose ball is empty, and subset=True
           (pointnet2_utils.py:320-400)
  ga_*     GroupAll with and without features (pointnet2_utils.py:403-440)
  pool_*   pooling_features 'max' / 'avg' / 'avg_max' + average_feature (pointnet2_modules.py:179-208)
  fp_*     PointnetFPModule forward, without and with its ball-query grouper (pointnet2_modules.py:457-588)
  grad_*   autograd through GroupingOperation and ThreeInterpolate (group_points_grad, three_interpolate_grad)
  clamp_*  denoising_step with data_clamp_range > 0 (diffusion_utils/diffusion.py:74-75)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops, ref_model  # noqa: E402

ops.install_reference_stubs()
from pointnet2_ops import pointnet2_utils as ref_utils  # noqa: E402
from pointnet2_ops import pointnet2_modules as ref_modules  # noqa: E402
from diffusion_utils import diffusion as ref_diffusion  # noqa: E402
from slide_b200 import weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

FP_KW = dict(bn=True, t_dim=32, include_t=True, bn_first=False, bias=True, first_conv=False, res_connect=True,
             include_condition=True, condition_dim=24, radius=0.3, nsample=8, use_xyz=True,
             include_abs_coordinate=True, include_center_coordinate=True, neighbor_def="radius", activation="relu")


def seeded_state(module, g):
    sd = {}
    for k, v in module.state_dict().items():
        sd[k] = (torch.randn(v.shape, generator=g) * (0.3 if v.dim() > 1 else 0.1) + (1.0 if k.endswith("norm.weight") or ".weight" in k and v.dim() == 1 else 0.0)).float()
    module.load_state_dict(sd, strict=True)
    return sd


def main():
    g = torch.Generator().manual_seed(77)
    gold = {}
    B, N, M, C = 2, 64, 16, 5
    xyz = torch.rand(B, N, 3, generator=g) - 0.5
    new_xyz = torch.rand(B, M, 3, generator=g) - 0.5
    new_xyz[:, -3:] += 5.0  # three queries far outside the cloud: their ball is empty
    feats = torch.randn(B, C, N, generator=g)
    gold.update(xyz=xyz.numpy(), new_xyz=new_xyz.numpy(), feats=feats.numpy())

    # ---- QueryAndGroup, radius
    qg = ref_utils.QueryAndGroup(0.25, 8, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=True,
                                 neighbor_def="radius")
    out, cnt = qg(xyz, new_xyz, feats, subset=False, return_counts=True)
    gold["qg_open_out"], gold["qg_open_count"] = out.numpy(), cnt.numpy()
    assert (cnt[:, -3:] == 0).all() and (cnt[:, :-3] > 0).any()
    sub_xyz = xyz[:, :M].contiguous()
    out, cnt = qg(xyz, sub_xyz, feats, subset=True, return_counts=True)
    gold["qg_subset_out"], gold["qg_subset_count"] = out.numpy(), cnt.numpy()
    out = ref_utils.QueryAndGroup(0.25, 8, use_xyz=True, neighbor_def="radius")(xyz, sub_xyz, None)
    gold["qg_xyz_only_out"] = out.numpy()

    # ---- GroupAll
    gold["ga_feat_out"] = ref_utils.GroupAll(use_xyz=True)(xyz, None, feats).numpy()
    gold["ga_nofeat_out"] = ref_utils.GroupAll(use_xyz=True)(xyz, None, None).numpy()
    gold["ga_noxyz_out"] = ref_utils.GroupAll(use_xyz=False)(xyz, None, feats).numpy()

    # ---- pooling_features / average_feature
    pf = torch.randn(B, 6, M, 8, generator=g)
    pcount = gold["qg_open_count"]
    pc = torch.from_numpy(pcount)
    gold["pool_in"] = pf.numpy()
    for mode in ("max", "avg", "avg_max"):
        gold["pool_" + mode] = ref_modules.pooling_features(pf, count=pc, pooling=mode).numpy()
    gold["pool_avg_all"] = ref_modules.pooling_features(pf, count="all", pooling="avg").numpy()

    # ---- PointnetFPModule
    C1, C2, n_unknown, n_known = 6, 7, 32, 12
    unknown = torch.rand(B, n_unknown, 3, generator=g) - 0.5
    known = torch.rand(B, n_known, 3, generator=g) - 0.5
    unknow_feats = torch.randn(B, C1, n_unknown, generator=g)
    known_feats = torch.randn(B, C2, n_known, generator=g)
    t_emb = torch.randn(B, 32, generator=g)
    cond = torch.randn(B, 24, generator=g)
    gold.update(fp_unknown=unknown.numpy(), fp_known=known.numpy(), fp_unknow_feats=unknow_feats.numpy(),
                fp_known_feats=known_feats.numpy(), fp_t_emb=t_emb.numpy(), fp_cond=cond.numpy())
    for tag, grouper in (("plain", False), ("grouper", True)):
        fp = ref_modules.PointnetFPModule(mlp=[C1 + C2, 16, 16, 20], include_grouper=grouper, **FP_KW).eval()
        sd = seeded_state(fp, g)
        for k, v in sd.items():
            gold["fp_%s_sd.%s" % (tag, k)] = v.numpy()
        with torch.no_grad():
            for pooling in (("max", "avg") if grouper else ("max",)):
                y = fp(unknown, known, unknow_feats, known_feats, t_emb=t_emb, condition_emb=cond, pooling=pooling)
                gold["fp_%s_out_%s" % (tag, pooling)] = y.numpy()

    # ---- autograd wrappers
    idx = torch.randint(0, N, (B, M, 8), generator=g).int()
    f = feats.clone().requires_grad_(True)
    w = torch.randn(B, C, M, 8, generator=g)
    (ref_utils.grouping_operation(f, idx) * w).sum().backward()
    gold.update(grad_group_idx=idx.numpy(), grad_group_w=w.numpy(), grad_group_features=f.grad.numpy())
    idx3 = torch.randint(0, n_known, (B, n_unknown, 3), generator=g).int()
    w3 = torch.rand(B, n_unknown, 3, generator=g)
    w3 = w3 / w3.sum(dim=2, keepdim=True)
    kf = known_feats.clone().requires_grad_(True)
    wo = torch.randn(B, C2, n_unknown, generator=g)
    (ref_utils.three_interpolate(kf, idx3, w3) * wo).sum().backward()
    gold.update(grad_interp_idx=idx3.numpy(), grad_interp_weight=w3.numpy(), grad_interp_w=wo.numpy(),
                grad_interp_features=kf.grad.numpy())
    f = feats.clone().requires_grad_(True)
    gidx = torch.randint(0, N, (B, M), generator=g).int()
    wg = torch.randn(B, C, M, generator=g)
    (ref_utils.gather_operation(f, gidx) * wg).sum().backward()
    gold.update(grad_gather_idx=gidx.numpy(), grad_gather_w=wg.numpy(), grad_gather_features=f.grad.numpy())

    # ---- feature-DDPM update with the clamp active
    cfg = dict(weights.load_json("pipeline_airplane.json")["latent_ddpm"]["standard_diffusion_config"])
    cfg["data_clamp_range"] = 0.6
    D = ref_diffusion.Diffusion(cfg, device=torch.device("cpu"))
    T = D.num_timesteps
    sch = ref_model.latent_schedule(cfg)
    Bc, Cc = 3, 51
    x = torch.randn(Bc, 16, Cc, generator=g)
    kp = torch.rand(Bc, 16, 3, generator=g) - 0.5
    x = torch.cat([kp, x[:, :, 3:]], dim=2)
    gold["clamp_x"] = x.numpy()
    gold["clamp_range"] = np.float32(0.6)
    model = lambda xx, ts=None, label=None: 0.5 * torch.tanh(xx) + 0.01 * (ts / T).reshape(-1, 1, 1)  # noqa: E731
    orig = torch.randn_like
    for t in (999, 500, 0):
        nz = torch.randn(Bc, 16, Cc, generator=g)
        torch.randn_like = lambda like, _n=nz: _n
        y, x0 = ref_diffusion.denoising_step(
            x, t=torch.ones(Bc) * t, model=model, logvar=D.logvar, sqrt_recip_alphas_cumprod=D.sqrt_recip_alphas_cumprod,
            sqrt_recipm1_alphas_cumprod=D.sqrt_recipm1_alphas_cumprod, posterior_mean_coef1=D.posterior_mean_coef1,
            posterior_mean_coef2=D.posterior_mean_coef2, return_pred_xstart=True, label=None,
            data_clamp_range=D.data_clamp_range)
        torch.randn_like = orig
        assert float((x0.abs() >= 0.6 - 1e-7).float().mean()) > 0.05, "the clamp must be active in this fixture"
        mine, _ = ref_model.denoising_step(x, torch.ones(Bc) * t, lambda xx, tt: model(xx, ts=tt), sch, nz)
        assert torch.equal(y, mine), "oracle/ref_model.denoising_step (clamp) deviates from the reference at t=%d" % t
        gold["clamp_noise_t%d" % t] = nz.numpy()
        gold["clamp_out_t%d" % t] = y.numpy()
    np.savez_compressed(os.path.join(OUT, "golden_dropin.npz"), **gold)
    print("wrote golden_dropin.npz: %d arrays" % len(gold))


if __name__ == "__main__":
    main()
