"""The sample-resident compiler (slide_b200/resident.py) on CPU: its packed plans, interpreted by
oracle/resident_sim.py with one NaN-filled shared-memory array per CTA, reproduce the record interpreter
(oracle/ir_exec.py) on the same arena -- operand binding, shared-memory liveness / spills, transform placement,
statistics ownership across the cluster and weight packing are all checked without a GPU."""
import numpy as np
import pytest
import torch

from oracle import ir_exec, resident_sim
from slide_b200 import engine, resident
from tests import common


def _program(cfg, cluster, precise, B=2):
    pos = cfg["position_ddpm"]
    d = pos["diffusion_config"]
    sd = common.state_dict("pos")
    b, h = engine.build_ddpm(pos["pointnet_config"], sd, B, 4, engine.position_table(d["T"], d["beta_0"], d["beta_T"]), 0,
                             resident=dict(cluster=cluster, precise=precise))
    assert len(h["resident_plans"]) == 2, h.get("resident_unsupported")
    return b, h


@pytest.mark.parametrize("cluster", [2, 4])
@pytest.mark.parametrize("precise", [True, False])
def test_plan_matches_record_interpreter(cluster, precise, pipeline_cfg):
    B = 2
    b, h = _program(pipeline_cfg, cluster, precise, B)
    m = ir_exec.Machine(b)
    common.init_machine(m, h, np.arange(B) % 13)
    g = torch.Generator().manual_seed(3)
    m.upload(h["x"], torch.randn(B * 16, 3, generator=g))
    m.upload(h["noise"], torch.randn(h["noise"].rows, 3, generator=g))
    m.run_segment("setup")
    m.set_step(3)
    start = m.arena.copy()
    for plan, seg in zip(h["resident_plans"], ("step", "forward")):
        m.arena[:] = start
        m.run_segment(seg)
        want_x, want_eps, want_step = m.download(h["x"]).numpy().copy(), m.download(h["eps"]).numpy().copy(), m.step()
        hdr, rec = plan.pack()
        assert int(hdr[0]["smem_floats"]) <= resident.SMEM_LIMIT_FLOATS
        arena = start.copy()
        resident_sim.ResidentSim(b, hdr, rec).run(arena)
        m.arena[:] = arena
        assert m.step() == want_step == 2
        tol = 1e-5 if precise else 3e-3
        assert np.abs(m.download(h["x"]).numpy() - want_x).max() <= tol * np.abs(want_x).max()
        if seg == "forward":  # eps leaves the range only in the forward-only plan
            assert np.abs(m.download(h["eps"]).numpy() - want_eps).max() <= tol * np.abs(want_eps).max()


def test_spills_are_exercised(pipeline_cfg):
    """cluster 2 does not fit the position denoiser's widest module without spilling; cluster 4 does."""
    _, h2 = _program(pipeline_cfg, 2, False)
    _, h4 = _program(pipeline_cfg, 4, False)
    assert h2["resident_plans"][0].summary()["spills"] > 0
    assert h4["resident_plans"][0].summary()["spills"] == 0


def test_feature_denoiser_is_refused(pipeline_cfg):
    """512-channel pair tensors do not fit a cluster's shared memory: the per-record executor stays in charge."""
    lat = pipeline_cfg["latent_ddpm"]
    b, h = engine.build_ddpm(lat["pointnet_config"], common.state_dict("lat"), 2, 4,
                             engine.latent_table(lat["standard_diffusion_config"]), 1, keep_cols=3,
                             resident=dict(cluster=4))
    assert h["resident_plans"] == [] and "shared memory" in h["resident_unsupported"]
