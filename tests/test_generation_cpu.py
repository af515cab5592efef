"""SURVEY.md 8(f) rows f1 / f2 on CPU: the local-resampling update (oracle and program record) against golden vectors
made by the REAL reference (tests/golden/make_golden_sampler.py), and the npz formats of the generation driver."""
import os

import numpy as np
import pytest
import torch

from oracle import ir_exec, ref_model
from slide_b200 import engine, generation
from slide_b200.program import KIND
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gs():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_sampler.npz")))


def _stand_in(T):
    return lambda x, ts: 0.5 * torch.tanh(x) + 0.01 * (ts / T).reshape(-1, 1, 1)


CASES = [("top", [999, 998, 997]), ("bottom", [2, 1, 0])]


@pytest.mark.parametrize("tag,t_list", CASES)
@pytest.mark.parametrize("local", [False, True])
def test_oracle_update_matches_reference_golden(tag, t_list, local, gs, pipeline_cfg):
    sch = ref_model.latent_schedule(pipeline_cfg["latent_ddpm"]["standard_diffusion_config"])
    noises = {t: torch.from_numpy(gs["noise_" + tag][i]) for i, t in enumerate(t_list)}
    got = ref_model.latent_denoise(_stand_in(sch["T"]), torch.from_numpy(gs["x_T"]), torch.from_numpy(gs["keypoint"]),
                                   noises, sch, t_start=t_list[0], n_steps=len(t_list),
                                   complete_x0=torch.from_numpy(gs["complete_x0"]) if local else None,
                                   keypoint_mask=torch.from_numpy(gs["mask"]) if local else None)
    assert np.array_equal(got.numpy(), gs["out_%s_%s" % (tag, "local" if local else "plain")])


@pytest.mark.parametrize("tag,t_list", CASES)
@pytest.mark.parametrize("local", [False, True])
def test_update_record_matches_reference_golden(tag, t_list, local, gs, pipeline_cfg):
    """The SLIDE_OP_DDPM_UPDATE record (interpreter of the program the GPU executes) with the stand-in eps uploaded:
    bit-exact against the reference's loop, including the keypoint columns the update must not touch."""
    B, T = 3, 1000
    lat = pipeline_cfg["latent_ddpm"]
    b, h = engine.build_ddpm(lat["pointnet_config"], common.state_dict("lat"), B, T,
                             engine.latent_table(lat["standard_diffusion_config"]), 1, keep_cols=3,
                             local_resampling=local)
    upd = [i for i, op in enumerate(b.ops) if op[0] == KIND["SLIDE_OP_DDPM_UPDATE"]]
    assert len(upd) == 1
    m = ir_exec.Machine(b)
    kp = torch.from_numpy(gs["keypoint"])
    x = torch.cat([kp, torch.from_numpy(gs["x_T"])[:, :, 3:]], dim=2)
    m.upload(h["x"], x)
    if local:
        m.upload(h["x0c"], torch.from_numpy(gs["complete_x0"]))
        m.upload(h["mask"], torch.from_numpy(gs["mask"]).reshape(-1, 1))
    nz = m.view(h["noise"]).reshape(T, B * 16, h["C"])
    model = _stand_in(T)
    for i, t in enumerate(t_list):
        nz[t] = gs["noise_" + tag][i].reshape(B * 16, -1)
        xin = m.download(h["x"]).reshape(B, 16, -1)
        m.upload(h["eps"], model(xin, torch.ones(B) * t))
        m.set_step(t)
        m.run(upd[0], 1)
    got = m.download(h["x"]).reshape(B, 16, -1).numpy()
    assert np.array_equal(got, gs["out_%s_%s" % (tag, "local" if local else "plain")])


def test_all_ones_mask_is_plain_sampling(gs, pipeline_cfg):
    sch = ref_model.latent_schedule(pipeline_cfg["latent_ddpm"]["standard_diffusion_config"])
    noises = {t: torch.from_numpy(gs["noise_top"][i]) for i, t in enumerate([999, 998, 997])}
    args = (_stand_in(sch["T"]), torch.from_numpy(gs["x_T"]), torch.from_numpy(gs["keypoint"]), noises, sch)
    plain = ref_model.latent_denoise(*args, n_steps=3)
    ones = ref_model.latent_denoise(*args, n_steps=3, complete_x0=torch.from_numpy(gs["complete_x0"]),
                                    keypoint_mask=torch.ones(3, 16))
    assert torch.equal(plain, ones)


def test_keypoint_file_and_result_formats(tmp_path):
    B = 5
    g = np.random.RandomState(0)
    kp_file = str(tmp_path / "keypoints.npz")
    np.savez(kp_file, points=g.rand(B, 16, 3).astype(np.float32), label=np.arange(B) % 13,
             category=np.array(["02691156"] * B), category_name=np.array(["airplane"] * B),
             keypoint_feature=g.randn(B, 16, 48).astype(np.float32), keypoint_mask=(g.rand(B, 16) < 0.5).astype(np.float32))
    whole = generation.load_keypoint_file(kp_file, local_resampling=True)
    assert whole["points"].shape == (B, 16, 3) and whole["complete_x0"].shape == (B, 16, 51)
    assert torch.equal(whole["complete_x0"][:, :, :3], whole["points"])
    # GeneralNpzDataset slicing: ceil(B / W) per rank, the last rank takes the remainder
    parts = [generation.load_keypoint_file(kp_file, rank=r, world_size=2, local_resampling=True) for r in range(2)]
    assert [p["points"].shape[0] for p in parts] == [3, 2]
    assert torch.equal(torch.cat([p["points"] for p in parts]), whole["points"])
    assert parts[0]["category"] + parts[1]["category"] == whole["category"]
    # result files: per-rank names, keys and the points / normals split of evaluate_per_rank, then the rank-0 gather
    save_dir = str(tmp_path / "out")
    os.makedirs(save_dir)
    clouds = g.randn(B, 2048, 6).astype(np.float32)
    lo = 0
    for r, p in enumerate(parts):
        n = p["points"].shape[0]
        res = generation.pack_results(clouds[lo:lo + n], p["label"].numpy(), p["category"], p["category_name"],
                                      [0.1] * n, keypoint=p["points"].numpy(),
                                      keypoint_feature=np.zeros((n, 16, 48), np.float32))
        assert res["points"].shape == (n, 2048, 3) and res["normals"].shape == (n, 2048, 3)
        f = generation.result_file(save_dir, 2048, r, 2)
        assert os.path.basename(f) == "shapenet_psr_generated_data_2048_pts_rank_%d.npz" % r
        np.savez(f, **res)
        lo += n
    out = generation.gather_generated_results(save_dir, 2, num_points=2048)
    assert os.path.basename(out) == "shapenet_psr_generated_data_2048_pts.npz"
    assert sorted(os.listdir(save_dir)) == ["shapenet_psr_generated_data_2048_pts.npz"]
    data = np.load(out)
    assert sorted(data.files) == sorted(["points", "normals", "label", "category", "category_name", "timing", "keypoint",
                                         "keypoint_feature"])
    assert np.array_equal(data["points"], clouds[:, :, :3]) and np.array_equal(data["normals"], clouds[:, :, 3:])
    assert np.array_equal(data["keypoint"], whole["points"].numpy())
    assert list(data["category_name"]) == ["airplane"] * B


class _FakePipe(object):
    """Stands in for SlidePipeline (no GPU): clouds = keypoints tiled, features = batch index, so that the driver's
    batching / tail padding / trimming can be checked on CPU."""
    world, rank = 1, 0

    class _Dec(object):
        out_points = 2048

    def __init__(self, Bl):
        self.Bl, self.dec, self.device, self.calls = Bl, self._Dec(), torch.device("cpu"), []

    def draw_host_inputs(self, labels, skip_position=False):
        assert labels.shape == (self.Bl,)
        assert skip_position, "external keypoints: the reference draws no position noise on this path"
        self._labels = labels

    def check_device_errors(self):
        pass

    def stage_inputs(self):
        pass

    def sample_resident(self, keypoints=None, complete_x0=None, keypoint_mask=None):
        assert keypoints.shape == (self.Bl, 16, 3)
        self.calls.append((complete_x0 is not None, keypoint_mask is not None))
        self.keypoint_feature = keypoints.sum(dim=2, keepdim=True).expand(self.Bl, 16, 48).contiguous()
        return torch.cat([keypoints.repeat(1, 128, 1), torch.ones(self.Bl, 2048, 3)], dim=2)


def test_generation_driver_batches_pads_and_trims(tmp_path):
    n, Bl = 7, 3
    kp = torch.arange(n * 16 * 3, dtype=torch.float32).reshape(n, 16, 3)
    label = torch.arange(n) % 13
    pipe = _FakePipe(Bl)
    res = generation.generate_per_rank(pipe, kp, label, category=["c"] * n, category_name=["n"] * n,
                                       save_dir=str(tmp_path), save_keypoint_feature=True,
                                       complete_x0=torch.zeros(n, 16, 51), keypoint_mask=torch.ones(n, 16))
    assert len(pipe.calls) == 3 and all(c == (True, True) for c in pipe.calls)  # 3 + 3 + (1 padded to 3)
    assert res["points"].shape == (n, 2048, 3) and res["normals"].shape == (n, 2048, 3)
    assert np.array_equal(res["points"][:, :16], kp.numpy())  # sample i of the output belongs to keypoint set i
    assert np.array_equal(res["keypoint_feature"][:, :, 0], kp.sum(dim=2).numpy())
    assert np.array_equal(res["label"], label.numpy()) and len(res["timing"]) == n
    data = np.load(os.path.join(str(tmp_path), "shapenet_psr_generated_data_2048_pts.npz"))
    assert np.array_equal(data["keypoint"], kp.numpy())
    assert "gt_points" not in data.files


def test_generation_driver_known_shapes_and_keypoint_noise(tmp_path):
    """evaluate_per_rank's `test_external_keypoint=False` branch (mesh_evaluation.py:86-98,140-142): the shapes the keypoints
    came from travel into the result file as `gt_points`, and keypoint noise is `magnitude * randn_like` per batch on the
    batch's real rows -- the noised keypoints are what the sampler sees and what is saved."""
    n, Bl = 5, 2
    kp = torch.arange(n * 16 * 3, dtype=torch.float32).reshape(n, 16, 3)
    gt = torch.randn(n, 64, 6)
    pipe = _FakePipe(Bl)
    torch.manual_seed(7)
    res = generation.generate_per_rank(pipe, kp, torch.zeros(n, dtype=torch.int64), save_dir=str(tmp_path), gt_points=gt,
                                       keypoint_noise_magnitude=0.04)
    torch.manual_seed(7)
    want = torch.cat([kp[b0:b0 + Bl] + 0.04 * torch.randn_like(kp[b0:b0 + Bl]) for b0 in range(0, n, Bl)])
    assert np.array_equal(res["keypoint"], want.numpy())                  # the reference's draws: (2,16,3), (2,16,3), (1,16,3)
    assert np.array_equal(res["points"][:, :16], want.numpy())            # ... and they are what was conditioned on
    assert np.array_equal(res["gt_points"], gt.numpy())
    data = np.load(os.path.join(str(tmp_path), "shapenet_psr_generated_data_2048_pts.npz"))
    assert sorted(data.files) == sorted(["points", "normals", "label", "category", "category_name", "timing", "keypoint",
                                         "gt_points"])


def test_position_sampler_matches_reference_golden(gs, pipeline_cfg):
    """The position DDPM's full 1000-step chain (pointnet2/util.py::sampling run by make_golden_sampler.py with a
    stand-in denoiser): the oracle's loop and the mode-0 SLIDE_OP_DDPM_UPDATE record with engine.position_table are
    both bit-exact, step after step."""
    d = pipeline_cfg["position_ddpm"]["diffusion_config"]
    T = d["T"]
    draws = torch.from_numpy(gs["pos_draws"])
    B = draws.shape[1]
    model = _stand_in(T)
    noises = {t: draws[T - t] for t in range(T - 1, 0, -1)}
    got = ref_model.position_sampling(model, draws[0], noises, ref_model.position_schedule(T, d["beta_0"], d["beta_T"]))
    assert np.array_equal(got.numpy(), gs["pos_out"])
    b, h = engine.build_ddpm(pipeline_cfg["position_ddpm"]["pointnet_config"], common.state_dict("pos"), B, T,
                             engine.position_table(T, d["beta_0"], d["beta_T"]), 0)
    upd = [i for i, op in enumerate(b.ops) if op[0] == KIND["SLIDE_OP_DDPM_UPDATE"]][0]
    m = ir_exec.Machine(b)
    m.upload(h["x"], draws[0])
    nz = m.view(h["noise"]).reshape(T, B * 16, 3)
    for t in range(T - 1, -1, -1):
        if t > 0:
            nz[t] = noises[t].reshape(B * 16, 3).numpy()
        x = m.download(h["x"]).reshape(B, 16, 3)
        m.upload(h["eps"], model(x, torch.ones(B) * t))
        m.set_step(t)
        m.run(upd, 1)
    assert np.array_equal(m.download(h["x"]).reshape(B, 16, 3).numpy(), gs["pos_out"])


@pytest.mark.parametrize("ci", range(6))
def test_update_record_under_the_reference_s_other_schedules(ci, pipeline_cfg):
    """engine.latent_table for get_beta_schedule 'quad' / 'const' / 'jsd', model_var_type 'fixedlarge' and other step counts /
    beta ranges (no shipped config selects them) + the mode-1 update record: bit-exact against the REAL Diffusion class and
    denoising_step (tests/golden/make_golden_schedules.py), including the inf / nan the reference itself produces at the
    last 'jsd' step."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gsch = np.load(os.path.join(root, "tests", "golden", "golden_schedules.npz"))
    case = json.loads(str(gsch["cases_json"]))[ci]
    T, B = case["num_diffusion_timesteps"], 2
    dcfg = dict(case, data_clamp_range=-1)
    lat = pipeline_cfg["latent_ddpm"]
    b, h = engine.build_ddpm(lat["pointnet_config"], common.state_dict("lat"), B, T, engine.latent_table(dcfg), 1, keep_cols=3)
    upd = [i for i, op in enumerate(b.ops) if op[0] == KIND["SLIDE_OP_DDPM_UPDATE"]][0]
    m = ir_exec.Machine(b)
    x = torch.cat([torch.from_numpy(gsch["keypoint"]), torch.from_numpy(gsch["x"])[:, :, 3:]], dim=2)
    nz = m.view(h["noise"]).reshape(T, B * 16, h["C"])
    model = _stand_in(T)
    for j, t in enumerate([T - 1, T // 2, 1, 0]):
        m.upload(h["x"], x)
        nz[t] = gsch["c%d_noise" % ci][j].reshape(B * 16, -1)
        m.upload(h["eps"], model(x, torch.ones(B) * t))
        m.set_step(t)
        m.run(upd, 1)
        got = m.download(h["x"]).reshape(B, 16, -1).numpy()
        assert np.array_equal(got, gsch["c%d_out_t%d" % (ci, t)], equal_nan=True), (case, t)


def test_unknown_schedules_are_refused():
    with pytest.raises(NotImplementedError):
        engine.beta_schedule("warmup10", 1e-4, 0.02, 10)     # NameError in the reference itself (undefined _warmup_beta)
    with pytest.raises(NotImplementedError):
        engine.latent_table(dict(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=10,
                                 model_var_type="learned"))
