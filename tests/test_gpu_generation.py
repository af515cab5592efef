"""SURVEY.md 8(f) rows f1 / f2 on the GPU, through the C ABI: the local-resampling update kernel bit-exact against the
reference's golden vectors, and the generation driver (external keypoints, local resampling, result files)."""
import copy
import os

import numpy as np
import pytest
import torch

from slide_b200 import engine, generation, lib, pipeline
from slide_b200.program import KIND, Program
from tests import common

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tag,t_list", [("top", [999, 998, 997]), ("bottom", [2, 1, 0])])
@pytest.mark.parametrize("local", [False, True])
def test_update_kernel_matches_reference_golden(tag, t_list, local, pipeline_cfg):
    gs = dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_sampler.npz")))
    B, T = 3, 1000
    lat = pipeline_cfg["latent_ddpm"]
    b, h = engine.build_ddpm(lat["pointnet_config"], common.state_dict("lat"), B, T,
                             engine.latent_table(lat["standard_diffusion_config"]), 1, keep_cols=3,
                             local_resampling=local)
    upd = [i for i, op in enumerate(b.ops) if op[0] == KIND["SLIDE_OP_DDPM_UPDATE"]][0]
    prog = Program(b)
    kp = torch.from_numpy(gs["keypoint"])
    prog.upload(h["x"], torch.cat([kp, torch.from_numpy(gs["x_T"])[:, :, 3:]], dim=2).reshape(B * 16, -1))
    if local:
        prog.upload(h["x0c"], torch.from_numpy(gs["complete_x0"]).reshape(B * 16, -1))
        prog.upload(h["mask"], torch.from_numpy(gs["mask"]).reshape(-1, 1))
    nz = prog.view(h["noise"]).view(T, B * 16, h["C"])
    for i, t in enumerate(t_list):
        nz[t].copy_(torch.from_numpy(gs["noise_" + tag][i]).reshape(B * 16, -1))
        xin = prog.download(h["x"]).cpu().reshape(B, 16, -1)
        eps = 0.5 * torch.tanh(xin) + 0.01 * (torch.ones(B) * t / T).reshape(-1, 1, 1)  # the golden's stand-in denoiser
        prog.upload(h["eps"], eps.reshape(B * 16, -1))
        prog.set_step(t)
        prog.run(upd, 1)
    got = prog.download(h["x"]).cpu().reshape(B, 16, -1).numpy()
    assert np.array_equal(got, gs["out_%s_%s" % (tag, "local" if local else "plain")])


def _tiny_cfg(cfg):
    cfg = copy.deepcopy(cfg)
    cfg["position_ddpm"]["diffusion_config"]["T"] = 4
    cfg["latent_ddpm"]["standard_diffusion_config"]["num_diffusion_timesteps"] = 4
    return cfg


def test_generation_driver_external_keypoints_and_local_resampling(pipeline_cfg, tmp_path):
    cfg = _tiny_cfg(pipeline_cfg)
    Bl, n = 4, 6  # 6 keypoint sets through a batch-4 pipeline: one full batch + a padded tail
    g = torch.Generator().manual_seed(3)
    kp = torch.rand(n, 16, 3, generator=g) - 0.5
    feat = torch.randn(n, 16, 48, generator=g)
    mask = (torch.rand(n, 16, generator=g) < 0.5).float()
    label = torch.full((n,), cfg["label"], dtype=torch.long)
    pipe = pipeline.SlidePipeline(cfg, Bl, decode_chunk=4, local_resampling=True)
    torch.manual_seed(5)
    torch.cuda.manual_seed(6)
    res = generation.generate_per_rank(pipe, kp, label, category=["02691156"] * n, category_name=["airplane"] * n,
                                       save_dir=str(tmp_path), save_keypoint_feature=True,
                                       complete_x0=torch.cat([kp, feat], dim=2), keypoint_mask=mask, rank=1, world_size=2)
    assert lib.load().slide_tc_error() == 0
    assert res["points"].shape == (n, 2048, 3) and res["normals"].shape == (n, 2048, 3)
    assert np.isfinite(res["points"]).all() and np.isfinite(res["normals"]).all()
    assert np.array_equal(res["keypoint"], kp.numpy()) and len(res["timing"]) == n
    # the loop ran down to t = 0, where the posterior mean is exactly the blended x0: features of the points that
    # were NOT re-sampled are bit-identical to the given ones, the others are not
    keep = mask.numpy() == 0
    assert np.array_equal(res["keypoint_feature"][keep], feat.numpy()[keep])
    assert not np.allclose(res["keypoint_feature"][~keep], feat.numpy()[~keep])
    f = os.path.join(str(tmp_path), "shapenet_psr_generated_data_2048_pts_rank_1.npz")
    data = np.load(f)
    assert np.array_equal(data["points"], res["points"]) and list(data["category_name"]) == ["airplane"] * n
    # the same pipeline without pinned points = plain keypoint-conditional generation (all-ones mask)
    torch.manual_seed(5)
    torch.cuda.manual_seed(6)
    res2 = generation.generate_per_rank(pipe, kp, label)
    assert res2["points"].shape == (n, 2048, 3) and np.isfinite(res2["points"]).all()


@pytest.mark.parametrize("schedule,kappa", [("linear", 0.0), ("quadratic", 0.5), ("quadratic", 1.0)])
def test_fast_sampler_update_kernel_matches_reference_golden(schedule, kappa, pipeline_cfg):
    """the mode-2 update kernel + engine.fast_position_schedule, bit-exact against the reference's
    STEP_sampling (golden_fast.npz, stand-in denoiser)."""
    gf = dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_fast.npz")))
    dcfg = {"T": 1000, "beta_0": 0.0001, "beta_T": 0.02}
    draws = torch.from_numpy(gf["draws"])
    L, B = int(gf["length"]), draws.shape[1]
    ts, table = engine.fast_position_schedule("step", L, schedule, kappa, dcfg)
    pos = pipeline_cfg["position_ddpm"]
    b, h = engine.build_ddpm(pos["pointnet_config"], common.state_dict("pos"), B, L, table, 2, ts_values=ts)
    upd = [i for i, op in enumerate(b.ops) if op[0] == KIND["SLIDE_OP_DDPM_UPDATE"]][0]
    prog = Program(b)
    prog.upload(h["x"], draws[0].reshape(B * 16, 3))
    nz = prog.view(h["noise"]).view(L, B * 16, 3)
    for i in range(L):
        s = L - 1 - i
        nz[s].copy_(draws[1 + i].reshape(B * 16, 3))
        x = prog.download(h["x"]).cpu().reshape(B, 16, 3)
        eps = 0.5 * torch.tanh(x) + 0.01 * (torch.ones(B) * float(ts[s]) / 1000).reshape(-1, 1, 1)
        prog.upload(h["eps"], eps.reshape(B * 16, 3))
        prog.set_step(s)
        prog.run(upd, 1)
    got = prog.download(h["x"]).cpu().reshape(B, 16, 3).numpy()
    assert np.array_equal(got, gf["out_step_%s_%g" % (schedule, kappa)])


def test_pipeline_with_fast_position_sampler(pipeline_cfg):
    cfg = _tiny_cfg(pipeline_cfg)
    cfg["position_ddpm"]["diffusion_config"]["T"] = 1000  # the STEP sampler picks its steps from the full schedule
    B = 4
    pipe = pipeline.SlidePipeline(cfg, B, decode_chunk=4,
                                  position_sampler=dict(method="step", length=6, schedule="quadratic", kappa=0.5))
    assert pipe.T_pos == 6 and pipe.pos.mode == 2
    torch.manual_seed(1)
    pipe.draw_host_inputs(torch.full((B,), cfg["label"], dtype=torch.long))
    torch.cuda.manual_seed(2)
    out = pipe.sample_to_host()
    assert out.shape == (B, 2048, 6) and torch.isfinite(out).all()
    assert lib.load().slide_tc_error() == 0
