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


# ------------------------------------------------------------------------------------------ random 16-point architectures
def _random16():
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gr = np.load(os.path.join(root, "tests", "golden", "golden_random_archs.npz"))
    return gr, json.loads(str(gr["meta16_json"]))


@pytest.mark.parametrize("i", range(6))
@pytest.mark.parametrize("cluster", [2, 4])
def test_random_architectures_compile_or_refuse(i, cluster, pipeline_cfg):
    """Six randomly drawn 16-point denoisers (1-3 levels, widths 16-128 incl. non-multiples of the GroupNorm group count,
    K 8 / 16, depths 2-3; tests/golden/make_golden_random_archs.py): the record program reproduces the REAL
    PointNet2CloudCondition, and the sample-resident compiler either refuses the architecture loudly (the pipeline then keeps
    the record executor) or emits plans whose simulation equals the record interpreter -- never a silently different result."""
    from slide_b200 import weights
    gr, metas = _random16()
    meta = metas[i]
    pc, n0 = meta["pointnet_config"], meta["n_points"]
    sd = weights.random_state_dict(meta["schema"], meta["seed"])
    d = pipeline_cfg["position_ddpm"]["diffusion_config"]
    B = 2
    b, h = engine.build_ddpm(pc, sd, B, 4, engine.position_table(d["T"], d["beta_0"], d["beta_T"]), 0, n_points=n0,
                             resident=dict(cluster=cluster, precise=True))
    m = ir_exec.Machine(b)
    engine.init_constants(m, h)
    m.upload(h["labels"], gr["label"].astype(np.int32))
    m.run_segment("setup")
    # the record program against the real module (timestep table of this 4-step program: t = 0 is its last row too)
    m.upload(h["x"], gr["r%d_x" % i])
    m.set_step(1)
    m.run_segment("forward")
    want = gr["r%d_eps_t0" % i]
    eps = m.download(h["eps"]).numpy().reshape(want.shape)
    assert np.abs(eps - want).max() < 2e-5 * max(1.0, np.abs(want).max())
    if not h["resident_plans"]:
        assert h["resident_unsupported"]          # refused with a reason
        return
    g = torch.Generator().manual_seed(3)
    m.upload(h["x"], torch.randn(B * n0, 3, generator=g))
    m.upload(h["noise"], torch.randn(h["noise"].rows, 3, generator=g))
    m.set_step(3)
    start = m.arena.copy()
    for plan, seg in zip(h["resident_plans"], ("step", "forward")):
        m.arena[:] = start
        m.run_segment(seg)
        want_x, want_eps = m.download(h["x"]).numpy().copy(), m.download(h["eps"]).numpy().copy()
        hdr, rec = plan.pack()
        assert int(hdr[0]["smem_floats"]) <= resident.SMEM_LIMIT_FLOATS
        arena = start.copy()
        resident_sim.ResidentSim(b, hdr, rec).run(arena)
        m.arena[:] = arena
        assert np.abs(m.download(h["x"]).numpy() - want_x).max() <= 1e-5 * np.abs(want_x).max()
        if seg == "forward":
            assert np.abs(m.download(h["eps"]).numpy() - want_eps).max() <= 1e-5 * np.abs(want_eps).max()
