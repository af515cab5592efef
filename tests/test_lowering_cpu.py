"""The lowering (slide_b200/nets.py -> records) interpreted on CPU (oracle/ir_exec.py) against the golden vectors
of the reference.  No GPU: this pins offsets, weight packing, GroupNorm bookkeeping and the q/k split."""
import numpy as np
import pytest
import torch

from oracle import ir_exec, ref_model
from slide_b200 import engine
from tests import common


@pytest.mark.parametrize("which", ["pos", "lat"])
def test_denoiser_records_match_reference(which, golden, pipeline_cfg):
    b, h, pc, sd = common.ddpm_program(pipeline_cfg, which, 2)
    m = ir_exec.Machine(b)
    common.init_machine(m, h, golden["label"])
    m.run_segment("setup")
    for t in (999, 500, 0):
        m.upload(h["x"], golden[which + "_x"])
        m.set_step(t + 1)
        m.run_segment("forward")
        eps = m.download(h["eps"]).numpy().reshape(2, 16, -1)
        want = golden["%s_eps_t%d" % (which, t)]
        assert np.abs(eps - want).max() < 2e-5 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("which,steps", [("pos", 3), ("lat", 3)])
def test_sampler_records_match_reference_loop(which, steps, golden, pipeline_cfg):
    """step segment (net + update + step counter) against oracle/ref_model's restatement of the samplers."""
    B, T = 2, 1000
    b, h, pc, sd = common.ddpm_program(pipeline_cfg, which, B, with_noise=True)
    m = ir_exec.Machine(b)
    label = torch.from_numpy(golden["label"]).long()
    common.init_machine(m, h, golden["label"])
    m.run_segment("setup")
    g = torch.Generator().manual_seed(5)
    C = h["C"]
    x_T = torch.randn(B, 16, C, generator=g)
    noises = {t: torch.randn(B, 16, C, generator=g) for t in range(T - 1, T - 1 - steps, -1)}
    nz = m.view(h["noise"]).reshape(T, B * 16, C)
    for t, v in noises.items():
        nz[t] = v.reshape(B * 16, C).numpy()
    net = lambda x, ts: ref_model.cloud_condition_net(x, ref_model.Params(sd), pc, ts=ts, label=label)
    with torch.no_grad():
        if which == "pos":
            d = pipeline_cfg["position_ddpm"]["diffusion_config"]
            dh = ref_model.position_schedule(d["T"], d["beta_0"], d["beta_T"])
            want = ref_model.position_sampling(net, x_T, noises, dh, n_steps=steps)
            m.upload(h["x"], x_T)
        else:
            sch = ref_model.latent_schedule(pipeline_cfg["latent_ddpm"]["standard_diffusion_config"])
            kp = torch.rand(B, 16, 3, generator=g) - 0.5
            want = ref_model.latent_denoise(net, x_T, kp, noises, sch, n_steps=steps)
            m.upload(h["x"], torch.cat([kp, x_T[:, :, 3:]], dim=2))
    m.set_step(T)
    for _ in range(steps):
        m.run_segment("step")
    assert m.step() == T - steps
    got = m.download(h["x"]).reshape(B, 16, C)
    assert (got - want).abs().max() < 5e-5 * max(1.0, want.abs().max())


def _chamfer(a, b):
    d = torch.cdist(a, b)
    return max(d.min(1)[0].max().item(), d.min(0)[0].max().item())


def test_decode_records_match_reference(golden, pipeline_cfg):
    B = 2
    sd = common.state_dict("ae")
    decs = pipeline_cfg["autoencoder"]["decoders"]
    b, h = engine.build_decode(decs, sd, B)
    m = ir_exec.Machine(b)
    engine.init_constants(m, h)
    m.upload(h["labels"], golden["label"].astype(np.int32))
    m.upload(h["keypoint"], golden["dec_kp"])
    m.upload(h["feature"], golden["dec_feat"])
    for t, s in zip(h["starts"], golden["dec_starts"]):
        m.upload(t, s.astype(np.int32))
    m.run_segment("setup")
    m.run_segment("decode")
    l1 = m.download(h["levels"][1]).reshape(B, 256, 6)
    assert np.abs(l1.numpy() - golden["dec_l1"]).max() < 1e-6
    # FPS over near-coincident children is chaotic under fp32 re-association: later levels agree as point SETS
    out = m.download(h["out"]).reshape(B, 2048, 6)
    want = torch.from_numpy(golden["dec_out"])
    for i in range(B):
        assert _chamfer(out[i, :, :3], want[i, :, :3]) < 2e-3


@pytest.mark.parametrize("sample", [False, True])
def test_encode_records_match_reference(sample, golden, pipeline_cfg):
    B = 2
    sd = common.state_dict("ae")
    aec = pipeline_cfg["autoencoder"]
    b, h = engine.build_encode(aec["encoder"], aec["decoders"][0], sd, B, 2048, sample_posterior=sample)
    m = ir_exec.Machine(b)
    common.load_encode_inputs(m, h, golden, sample)
    m.run_segment("encode")
    got = m.download(h["out"]).numpy().reshape(B, 16, 48)
    want = golden["enc_sample" if sample else "enc_mode"]
    assert np.abs(got - want).max() < 2e-5 * max(1.0, np.abs(want).max())


def test_frozen_coordinates_hoist_the_neighbour_search(golden, pipeline_cfg):
    """engine.build_ddpm(frozen_xyz=True): the kNN records leave the step for a per-chain "geometry" segment; the chain
    is unchanged bit for bit as long as the coordinates really do not move (keep_cols = 3)."""
    B, T, steps = 2, 1000, 3
    pc = pipeline_cfg["latent_ddpm"]["pointnet_config"]
    table = engine.latent_table(pipeline_cfg["latent_ddpm"]["standard_diffusion_config"])
    sd = common.state_dict("lat")
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(B * 16, 3 + pc["in_fea_dim"], generator=g)
    out = []
    for frozen in (False, True):
        b, h = engine.build_ddpm(pc, sd, B, T, table, 1, keep_cols=3, with_noise=True, frozen_xyz=frozen)
        kinds = [ir_exec.KIND_NAME[b.ops[i][0]] for i in range(*[(f, f + c) for f, c in [b.segments["step"]]][0])]
        assert ("SLIDE_OP_KNN" in kinds) == (not frozen)
        m = ir_exec.Machine(b)
        common.init_machine(m, h, golden["label"])
        m.run_segment("setup")
        m.view(h["noise"])[...] = torch.randn(h["noise"].rows, h["C"], generator=torch.Generator().manual_seed(4)).numpy()
        m.upload(h["x"], x0)
        if frozen:
            assert b.segments["geometry"][1] == 5  # four neighbour searches + the attached coordinate columns
            m.run_segment("geometry")
        m.set_step(T)
        for _ in range(steps):
            m.run_segment("step")
        out.append(np.array(m.download(h["x"])))
    assert np.array_equal(out[0], out[1])


def test_automatic_geometry_hoist_only_without_down_sampling(golden, pipeline_cfg):
    """DDPMSampler(frozen_xyz=None) hoists the neighbour searches only when engine.can_freeze_geometry says so: every
    shipped denoiser qualifies; a custom feature denoiser that down-samples (FPS inside the step) keeps them in the step
    instead of failing, and still reproduces the reference network."""
    import copy
    lat = pipeline_cfg["latent_ddpm"]
    assert engine.can_freeze_geometry(lat["pointnet_config"]) and engine.can_freeze_geometry(pipeline_cfg["position_ddpm"]["pointnet_config"])
    pc = copy.deepcopy(lat["pointnet_config"])
    pc["architecture"]["npoint"] = [16, 8]      # the point counts change no parameter shape: same state dict
    assert not engine.can_freeze_geometry(pc)
    sd = common.state_dict("lat")
    table = engine.latent_table(lat["standard_diffusion_config"])
    with pytest.raises(NotImplementedError):
        engine.build_ddpm(pc, sd, 2, 1000, table, 1, keep_cols=3, with_noise=False, frozen_xyz=True)
    b, h = engine.build_ddpm(pc, sd, 2, 1000, table, 1, keep_cols=3, with_noise=False, frozen_xyz=False)
    assert "geometry" not in b.segments
    m = ir_exec.Machine(b)
    common.init_machine(m, h, golden["label"])
    m.run_segment("setup")
    x = torch.from_numpy(golden["lat_x"])
    m.upload(h["x"], x)
    m.set_step(501)
    m.run_segment("forward")
    eps = m.download(h["eps"]).numpy().reshape(2, 16, -1)
    with torch.no_grad():
        want = ref_model.cloud_condition_net(x, ref_model.Params(sd), pc, ts=torch.ones(2) * 500,
                                             label=torch.from_numpy(golden["label"]).long()).numpy()
    assert np.abs(eps - want).max() < 2e-5 * max(1.0, np.abs(want).max())
