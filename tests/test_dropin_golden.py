"""The drop-in `pointnet2_ops` python package against golden vectors produced by the REAL reference classes
(tests/golden/make_golden_dropin.py -> golden_dropin.npz): the branches the shipped sampling configs never take
(SURVEY 8 rows a12 'radius' grouping + empty-ball patch, GroupAll; a14 pooling; a15 PointnetFPModule; a16 the three
backward ops) and the clamp branch of the feature-DDPM update (a5).  GPU tests go through the C ABI; the pure-torch parts
(GroupAll, pooling_features) are also checked on the CPU."""
import os

import numpy as np
import pytest
import torch

import slide_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gd():
    return {k: v for k, v in np.load(os.path.join(ROOT, "tests", "golden", "golden_dropin.npz")).items()}


def T(a, dev="cpu"):
    return torch.from_numpy(np.asarray(a)).to(dev)


def close(got, want, tol=1e-5):
    want = T(want)
    return float((got.detach().cpu().float() - want).abs().max()) <= tol * max(1.0, float(want.abs().max()))


# ---------------------------------------------------------------------------------------------- CPU (pure torch parts)
def test_group_all_and_pooling_cpu(gd):
    slide_b200.install_dropin()
    from pointnet2_ops import pointnet2_utils as U
    from pointnet2_ops import pointnet2_modules as M
    xyz, feats = T(gd["xyz"]), T(gd["feats"])
    assert torch.equal(U.GroupAll(use_xyz=True)(xyz, None, feats), T(gd["ga_feat_out"]))
    assert torch.equal(U.GroupAll(use_xyz=True)(xyz, None, None), T(gd["ga_nofeat_out"]))
    assert torch.equal(U.GroupAll(use_xyz=False)(xyz, None, feats), T(gd["ga_noxyz_out"]))
    pf, cnt = T(gd["pool_in"]), T(gd["qg_open_count"])
    assert torch.equal(M.pooling_features(pf, count=cnt, pooling="max"), T(gd["pool_max"]))
    assert close(M.pooling_features(pf, count=cnt, pooling="avg"), gd["pool_avg"], 1e-6)
    assert close(M.pooling_features(pf, count=cnt, pooling="avg_max"), gd["pool_avg_max"], 1e-6)
    assert close(M.pooling_features(pf, count="all", pooling="avg"), gd["pool_avg_all"], 1e-6)
    with pytest.raises(AssertionError):
        M.pooling_features(pf, count=cnt, pooling="median")


# ---------------------------------------------------------------------------------------------- GPU (through the C ABI)
@pytest.mark.gpu
def test_query_and_group_radius_matches_reference(gd):
    """neighbor_def='radius': ball query + grouping; subset=False patches queries with an empty ball (they stand for
    themselves with zero features, pointnet2_utils.py:355-358,385-395)."""
    slide_b200.install_dropin()
    from pointnet2_ops import pointnet2_utils as U
    xyz, new_xyz, feats = T(gd["xyz"], "cuda"), T(gd["new_xyz"], "cuda"), T(gd["feats"], "cuda")
    qg = U.QueryAndGroup(0.25, 8, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=True,
                         neighbor_def="radius")
    out, cnt = qg(xyz, new_xyz, feats, subset=False, return_counts=True)
    assert torch.equal(cnt.cpu().float(), T(gd["qg_open_count"]).float())
    assert torch.equal(out.cpu(), T(gd["qg_open_out"]))
    sub = xyz[:, :16].contiguous()
    out, cnt = qg(xyz, sub, feats, subset=True, return_counts=True)
    assert torch.equal(cnt.cpu().float(), T(gd["qg_subset_count"]).float())
    assert torch.equal(out.cpu(), T(gd["qg_subset_out"]))
    out = U.QueryAndGroup(0.25, 8, use_xyz=True, neighbor_def="radius")(xyz, sub, None)
    assert torch.equal(out.cpu(), T(gd["qg_xyz_only_out"]))
    ga = U.GroupAll(use_xyz=True)(xyz, None, feats)
    assert torch.equal(ga.cpu(), T(gd["ga_feat_out"]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["plain", "grouper"])
def test_pointnet_fp_module_matches_reference(gd, tag):
    """PointnetFPModule.forward: three_nn + three_interpolate + Mlp_plus_t_emb (+ ball-query grouper and pooling)."""
    slide_b200.install_dropin()
    from pointnet2_ops.pointnet2_modules import PointnetFPModule
    kw = dict(bn=True, t_dim=32, include_t=True, bn_first=False, bias=True, first_conv=False, res_connect=True,
              include_condition=True, condition_dim=24, radius=0.3, nsample=8, use_xyz=True,
              include_abs_coordinate=True, include_center_coordinate=True, neighbor_def="radius", activation="relu")
    fp = PointnetFPModule(mlp=[13, 16, 16, 20], include_grouper=(tag == "grouper"), **kw)
    pre = "fp_%s_sd." % tag
    fp.load_state_dict({k[len(pre):]: T(v) for k, v in gd.items() if k.startswith(pre)}, strict=True)
    fp = fp.cuda().eval()
    args = [T(gd[k], "cuda") for k in ("fp_unknown", "fp_known", "fp_unknow_feats", "fp_known_feats")]
    torch.backends.cudnn.allow_tf32 = False  # the golden is fp32 (CPU); compare like with like
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        for pooling in (("max", "avg") if tag == "grouper" else ("max",)):
            y = fp(*args, t_emb=T(gd["fp_t_emb"], "cuda"), condition_emb=T(gd["fp_cond"], "cuda"), pooling=pooling)
            assert close(y, gd["fp_%s_out_%s" % (tag, pooling)], 2e-5), (tag, pooling)


@pytest.mark.gpu
def test_backward_ops_match_reference(gd):
    """gather_points_grad, group_points_grad, three_interpolate_grad through the autograd wrappers."""
    slide_b200.install_dropin()
    from pointnet2_ops import pointnet2_utils as U
    f = T(gd["feats"], "cuda").clone().requires_grad_(True)
    (U.grouping_operation(f, T(gd["grad_group_idx"], "cuda")) * T(gd["grad_group_w"], "cuda")).sum().backward()
    assert close(f.grad, gd["grad_group_features"], 1e-5)
    kf = T(gd["fp_known_feats"], "cuda").clone().requires_grad_(True)
    (U.three_interpolate(kf, T(gd["grad_interp_idx"], "cuda"), T(gd["grad_interp_weight"], "cuda")) *
     T(gd["grad_interp_w"], "cuda")).sum().backward()
    assert close(kf.grad, gd["grad_interp_features"], 1e-5)
    f = T(gd["feats"], "cuda").clone().requires_grad_(True)
    (U.gather_operation(f, T(gd["grad_gather_idx"], "cuda")) * T(gd["grad_gather_w"], "cuda")).sum().backward()
    assert close(f.grad, gd["grad_gather_features"], 1e-5)


@pytest.mark.gpu
def test_update_kernel_clamp_branch_matches_reference(gd, pipeline_cfg):
    """DDPM_UPDATE mode 1 with data_clamp_range > 0 (diffusion.py:74-75): bit-exact against the real denoising_step."""
    from slide_b200 import engine
    from slide_b200.program import KIND, Program
    from tests import common
    B, Tn = 3, 1000
    lat = pipeline_cfg["latent_ddpm"]
    dcfg = dict(lat["standard_diffusion_config"])
    dcfg["data_clamp_range"] = float(gd["clamp_range"])
    b, h = engine.build_ddpm(lat["pointnet_config"], common.state_dict("lat"), B, Tn, engine.latent_table(dcfg), 1,
                             keep_cols=0, clamp=float(gd["clamp_range"]))
    upd = [i for i, op in enumerate(b.ops) if op[0] == KIND["SLIDE_OP_DDPM_UPDATE"]][0]
    prog = Program(b)
    x = T(gd["clamp_x"])
    nz = prog.view(h["noise"]).view(Tn, B * 16, h["C"])
    for t in (999, 500, 0):
        prog.upload(h["x"], x.reshape(B * 16, -1))
        nz[t].copy_(T(gd["clamp_noise_t%d" % t]).reshape(B * 16, -1))
        eps = 0.5 * torch.tanh(x) + 0.01 * (torch.ones(B) * t / Tn).reshape(-1, 1, 1)
        prog.upload(h["eps"], eps.reshape(B * 16, -1))
        prog.set_step(t)
        prog.run(upd, 1)
        got = prog.download(h["x"]).cpu().reshape(B, 16, -1).numpy()
        assert np.array_equal(got, gd["clamp_out_t%d" % t]), t
