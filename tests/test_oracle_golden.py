"""The CPU oracle against the golden vectors produced by the REAL reference (tests/golden/make_golden.py)."""
import numpy as np
import torch

from oracle import ops, ref_model
from tests import common


def test_c_oracle_config1(golden):
    xyz = torch.from_numpy(golden["c1_xyz"])
    fps = ops.furthest_point_sampling(xyz, 1024)
    assert np.array_equal(fps.numpy(), golden["c1_fps"])
    new_xyz = xyz[0][fps[0].long()][None].contiguous()
    idx, cnt = ops.ball_query(new_xyz, xyz, 0.2, 32)
    assert np.array_equal(idx.numpy(), golden["c1_bq_idx"]) and np.array_equal(cnt.numpy(), golden["c1_bq_cnt"])
    # independent numpy check of ball query semantics (first nsample hits in index order, padded with the first)
    d2 = ((new_xyz[0][:, None, :] - xyz[0][None, :, :]) ** 2).sum(-1).numpy()
    for j in (0, 17, 1023):
        hits = np.nonzero(d2[j] < np.float32(0.2) * np.float32(0.2))[0][:32]
        assert cnt[0, j] == len(hits)
        assert np.array_equal(idx[0, j, :len(hits)].numpy(), hits)
        assert (idx[0, j, len(hits):].numpy() == hits[0]).all()


def test_fps_matches_numpy_argmax():
    """Without exact ties or near-origin points, FPS is the textbook arg-max chain."""
    g = torch.Generator().manual_seed(3)
    xyz = torch.rand(2, 300, 3, generator=g) + 1.0
    got = ops.furthest_point_sampling(xyz, 40).numpy()
    for b in range(2):
        p = xyz[b].numpy()
        dist = np.full(300, 1e10, dtype=np.float32)
        cur, want = 0, [0]
        for _ in range(39):
            d = ((p - p[cur]) ** 2).sum(1).astype(np.float32)
            dist = np.minimum(dist, d)
            cur = int(dist.argmax())
            want.append(cur)
        assert got[b].tolist() == want


def test_knn_sorted_and_tie_order():
    g = torch.Generator().manual_seed(4)
    p2 = torch.rand(1, 50, 3, generator=g)
    p2[0, 10] = p2[0, 3]  # exact duplicate -> equal distances, ascending index expected
    p1 = torch.rand(1, 7, 3, generator=g)
    res = ops.knn_points(p1, p2, K=50)
    d = res.dists[0].numpy()
    assert (np.diff(d, axis=1) >= 0).all()
    for i in range(7):
        row = res.idx[0, i].tolist()
        assert row.index(3) + 1 == row.index(10)
    ref = ((p1[0][:, None] - p2[0][None]) ** 2).sum(-1)
    assert torch.allclose(res.dists[0], ref.sort(dim=1)[0], atol=1e-6)


def test_ref_model_denoisers(golden, pipeline_cfg):
    label = torch.from_numpy(golden["label"]).long()
    for which, key in (("pos", "position_ddpm"), ("lat", "latent_ddpm")):
        pc = pipeline_cfg[key]["pointnet_config"]
        sd = common.state_dict(which)
        x = torch.from_numpy(golden[which + "_x"])
        for t in (999, 500, 0):
            with torch.no_grad():
                y = ref_model.cloud_condition_net(x, ref_model.Params(sd), pc, ts=torch.ones(2) * t, label=label)
            assert np.array_equal(y.numpy(), golden["%s_eps_t%d" % (which, t)])


def test_ref_model_decode(golden, pipeline_cfg):
    sd = common.state_dict("ae")
    starts = [torch.from_numpy(s).long() for s in golden["dec_starts"]]
    with torch.no_grad():
        out, levels = ref_model.decode(torch.from_numpy(golden["dec_kp"]), torch.from_numpy(golden["dec_feat"]),
                                       ref_model.Params(sd), pipeline_cfg["autoencoder"]["decoders"],
                                       torch.from_numpy(golden["label"]).long(), start_idx_list=starts)
    assert np.array_equal(out.numpy(), golden["dec_out"])


def test_ref_model_encode(golden, pipeline_cfg):
    sd = common.state_dict("ae")
    aec = pipeline_cfg["autoencoder"]
    args = (torch.from_numpy(golden["enc_cloud"]), torch.from_numpy(golden["enc_kp"]), ref_model.Params(sd),
            aec["encoder"], aec["decoders"][0], torch.from_numpy(golden["label"]).long())
    with torch.no_grad():
        assert np.array_equal(ref_model.encode(*args).numpy(), golden["enc_mode"])
        noises = (torch.from_numpy(golden["enc_n1"]), torch.from_numpy(golden["enc_n2"]))
        assert np.array_equal(ref_model.encode(*args, noises=noises).numpy(), golden["enc_sample"])


def test_schedules_match_reference_formulas(pipeline_cfg):
    from slide_b200 import engine
    d = pipeline_cfg["position_ddpm"]["diffusion_config"]
    tab = engine.position_table(d["T"], d["beta_0"], d["beta_T"])
    dh = ref_model.position_schedule(d["T"], d["beta_0"], d["beta_T"])
    t = 417
    assert tab[t, 0] == float((1 - dh["Alpha"][t]) / torch.sqrt(1 - dh["Alpha_bar"][t]))
    assert tab[t, 1] == float(torch.sqrt(dh["Alpha"][t])) and tab[t, 2] == float(dh["Sigma"][t])
    sch = ref_model.latent_schedule(pipeline_cfg["latent_ddpm"]["standard_diffusion_config"])
    tab = engine.latent_table(pipeline_cfg["latent_ddpm"]["standard_diffusion_config"])
    tt = torch.tensor([t])
    assert tab[t, 0] == float(ref_model._extract(sch["sqrt_recip_alphas_cumprod"], tt, 1))
    assert tab[t, 3] == float(ref_model._extract(sch["posterior_mean_coef2"], tt, 1))
    assert tab[t, 4] == float(torch.exp(0.5 * ref_model._extract(sch["logvar"], tt, 1)))


FPS_CASES = ["full", "ragged", "dup", "short", "decode"]


def _fps_golden():
    import os
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_fps.npz")))


def test_p3d_fps_oracle_matches_vendored_pytorch3d_reference():
    """golden_fps.npz = pointnet2/data_utils/points_sampling.py::sample_farthest_points_naive (the copy of pytorch3d's
    reference implementation inside the SLIDE tree) on full / ragged / duplicated / short clouds."""
    gold = _fps_golden()
    for name in FPS_CASES:
        p = torch.from_numpy(gold[name + "_points"])
        _, idx = ops.sample_farthest_points(p, torch.from_numpy(gold[name + "_lengths"]), list(gold[name + "_K"]))
        assert np.array_equal(idx.numpy(), gold[name + "_idx"]), name
