"""SURVEY.md 8(f) row f4 on CPU: the FastDPM STEP sampler of the position DDPM (pointnet2/util_fastdpmv2.py) -- oracle,
schedule tables and the mode-2 DDPM_UPDATE record against golden vectors made by the REAL reference
(tests/golden/make_golden_fast.py), and the whole lowered step (denoiser evaluated at the sampler's timesteps)
against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import ir_exec, ref_model
from slide_b200 import engine
from slide_b200.program import KIND
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DCFG = {"T": 1000, "beta_0": 0.0001, "beta_T": 0.02}
CASES = [(s, k) for s in ("linear", "quadratic") for k in (0.0, 0.5, 1.0)]


@pytest.fixture(scope="module")
def gf():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_fast.npz")))


def _stand_in(x, ts):
    return 0.5 * torch.tanh(x) + 0.01 * (ts / DCFG["T"]).reshape(-1, 1, 1)


@pytest.mark.parametrize("schedule,kappa", CASES)
def test_oracle_matches_reference_golden(schedule, kappa, gf):
    draws = torch.from_numpy(gf["draws"])
    L = int(gf["length"])
    got, taus = ref_model.fast_sampling(_stand_in, draws[0], draws[1:], DCFG, "step", L, schedule, kappa)
    key = "step_%s_%g" % (schedule, kappa)
    assert np.array_equal(got.numpy(), gf["out_" + key]) and np.array_equal(np.asarray(taus, float), gf["taus_" + key])


@pytest.mark.parametrize("schedule,kappa", CASES)
def test_update_record_matches_reference_golden(schedule, kappa, gf, pipeline_cfg):
    """engine.fast_position_schedule + the mode-2 SLIDE_OP_DDPM_UPDATE record (program interpreter), stand-in eps
    uploaded per step: bit-exact against the reference's STEP_sampling."""
    draws = torch.from_numpy(gf["draws"])
    L, B = int(gf["length"]), draws.shape[1]
    ts, table = engine.fast_position_schedule("step", L, schedule, kappa, DCFG)
    key = "step_%s_%g" % (schedule, kappa)
    assert np.array_equal(ts[::-1].astype(np.float64), gf["taus_" + key])  # row s <-> iteration L-1-s
    pos = pipeline_cfg["position_ddpm"]
    b, h = engine.build_ddpm(pos["pointnet_config"], common.state_dict("pos"), B, L, table, 2, ts_values=ts)
    upd = [i for i, op in enumerate(b.ops) if op[0] == KIND["SLIDE_OP_DDPM_UPDATE"]][0]
    m = ir_exec.Machine(b)
    m.upload(h["x"], draws[0])
    nz = m.view(h["noise"]).reshape(L, B * 16, 3)
    for i in range(L):
        s = L - 1 - i
        nz[s] = draws[1 + i].reshape(B * 16, 3).numpy()
        x = m.download(h["x"]).reshape(B, 16, 3)
        m.upload(h["eps"], _stand_in(x, torch.ones(B) * float(ts[s])))
        m.set_step(s)
        m.run(upd, 1)
    assert np.array_equal(m.download(h["x"]).reshape(B, 16, 3).numpy(), gf["out_" + key])


def test_var_sampler_is_refused():
    with pytest.raises(NotImplementedError):
        engine.fast_position_schedule("var", 10, "linear", 0.5, DCFG)


def test_lowered_fast_sampler_matches_oracle_loop(golden, pipeline_cfg):
    """The whole step segment of a mode-2 program (timestep embeddings taken at the sampler's steps, network, update,
    step counter) against ref_model.fast_sampling with the oracle's network."""
    B, L, schedule, kappa = 2, 4, "quadratic", 0.5
    pos = pipeline_cfg["position_ddpm"]
    pc, sd = pos["pointnet_config"], common.state_dict("pos")
    ts, table = engine.fast_position_schedule("step", L, schedule, kappa, DCFG)
    b, h = engine.build_ddpm(pc, sd, B, L, table, 2, ts_values=ts)
    m = ir_exec.Machine(b)
    label = torch.from_numpy(golden["label"]).long()
    common.init_machine(m, h, golden["label"])
    m.run_segment("setup")
    g = torch.Generator().manual_seed(9)
    x_T = torch.randn(B, 16, 3, generator=g)
    noises = [torch.randn(B, 16, 3, generator=g) for _ in range(L)]
    nz = m.view(h["noise"]).reshape(L, B * 16, 3)
    for i in range(L):
        nz[L - 1 - i] = noises[i].reshape(B * 16, 3).numpy()
    net = lambda x, t: ref_model.cloud_condition_net(x, ref_model.Params(sd), pc, ts=t, label=label)
    with torch.no_grad():
        want, _ = ref_model.fast_sampling(net, x_T, noises, DCFG, "step", L, schedule, kappa)
    m.upload(h["x"], x_T)
    m.set_step(L)
    for _ in range(L):
        m.run_segment("step")
    got = m.download(h["x"]).reshape(B, 16, 3)
    assert (got - want).abs().max() < 5e-5 * max(1.0, want.abs().max())


def test_host_draw_order_of_the_fast_sampler(pipeline_cfg):
    """x_T first, then one std_normal per iteration (also the last); row s of pos_noise = draw of iteration L-1-s;
    slices do not depend on the world size."""
    from slide_b200 import pipeline
    B, L = 4, 5
    labels = torch.zeros(B, dtype=torch.long)
    torch.manual_seed(3)
    full = pipeline.draw_host_inputs(pipeline_cfg, B, 0, 1, labels, fast_steps=L)
    torch.manual_seed(3)
    x_T = torch.normal(0, 1, size=(B, 16, 3))
    z = [torch.normal(0, 1, size=(B, 16, 3)) for _ in range(L)]
    assert torch.equal(full["pos_xT"], x_T) and full["pos_noise"].shape == (L, B, 16, 3)
    for i in range(L):
        assert torch.equal(full["pos_noise"][L - 1 - i], z[i])
    parts = []
    for r in range(2):
        torch.manual_seed(3)
        parts.append(pipeline.draw_host_inputs(pipeline_cfg, B, r, 2, labels, fast_steps=L))
    assert torch.equal(torch.cat([p["pos_noise"] for p in parts], dim=1), full["pos_noise"])
