"""The sample-resident kernel (slide_b200/csrc/resident.cu) on the GPU against the CPU record interpreter:
one launch per position-DDPM step, clusters of 2 (with spills) and 4 CTAs per sample, TF32 and 3xTF32 (PRECISE)
products, eager and CUDA-graph replay.  All calls go through the C ABI (slide_program_set_resident / _run)."""
import numpy as np
import pytest
import torch

from oracle import ir_exec, ref_model
from slide_b200 import engine, lib, pipeline
from slide_b200.program import Program
from tests import common

pytestmark = pytest.mark.gpu


def _build(cfg, B, cluster, precise, T=4):
    pos = cfg["position_ddpm"]
    d = pos["diffusion_config"]
    sd = common.state_dict("pos")
    b, h = engine.build_ddpm(pos["pointnet_config"], sd, B, T, engine.position_table(d["T"], d["beta_0"], d["beta_T"]), 0,
                             resident=dict(cluster=cluster, precise=precise))
    assert len(h["resident_plans"]) == 2, h.get("resident_unsupported")
    return b, h, sd


@pytest.mark.parametrize("cluster", [4, 2])
@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("B", [5, 32])
def test_resident_step_matches_interpreter(cluster, precise, B, pipeline_cfg):
    b, h, sd = _build(pipeline_cfg, B, cluster, precise)
    m = ir_exec.Machine(b)
    common.init_machine(m, h, np.arange(B) % 13)
    g = torch.Generator().manual_seed(40 + B)
    m.upload(h["x"], torch.randn(B * 16, 3, generator=g))
    m.upload(h["noise"], torch.randn(h["noise"].rows, 3, generator=g))
    m.run_segment("setup")
    m.set_step(3)
    prog = Program(b)
    for plan in h["resident_plans"]:
        prog.set_resident(plan)
    prog.raw_arena().copy_(torch.from_numpy(m.arena))
    lib.reset_launch_count()
    # forward-only range: eps
    prog.run_segment("forward")
    torch.cuda.synchronize()
    assert lib.launch_count() == 1, "the forward range must be ONE kernel launch"
    m.run_segment("forward")
    eps, want = prog.download(h["eps"]).cpu().numpy(), m.download(h["eps"]).numpy()
    assert int(prog.view(b.step).item()) == m.step() == 2
    tol = 2e-4 if precise else 5e-3
    assert np.isfinite(eps).all()
    assert np.abs(eps - want).max() <= tol * max(1.0, np.abs(want).max()), np.abs(eps - want).max()
    # full step (forward + update of x)
    m.set_step(3)
    prog.set_step(3)
    prog.run_segment("step")
    torch.cuda.synchronize()
    m.run_segment("step")
    x, wx = prog.download(h["x"]).cpu().numpy(), m.download(h["x"]).numpy()
    assert int(prog.view(b.step).item()) == 2
    assert np.abs(x - wx).max() <= tol * max(1.0, np.abs(wx).max())


def test_resident_equals_per_record_path_over_steps(pipeline_cfg):
    """Four steps: resident kernel (TF32) vs the per-record executor (fp32 FFMA), same program, same noise."""
    B = 8
    b, h, sd = _build(pipeline_cfg, B, 4, True)
    prog = Program(b)
    for plan in h["resident_plans"]:
        prog.set_resident(plan)
    engine.init_constants(prog, h)
    prog.upload(h["labels"], (torch.arange(B) % 13).int())
    prog.run_segment("setup")
    g = torch.Generator().manual_seed(2)
    x0 = torch.randn(B * 16, 3, generator=g)
    prog.upload(h["noise"], torch.randn(h["noise"].rows, 3, generator=g))
    out = []
    for resident in (True, False):
        prog.use_resident(resident)
        prog.set_gemm_backend("auto" if resident else "simt")
        prog.upload(h["x"], x0)
        prog.set_step(4)
        for _ in range(4):
            prog.run_segment("step")
        torch.cuda.synchronize()
        assert int(prog.view(b.step).item()) == 0
        out.append(prog.download(h["x"]).cpu())
    assert (out[0] - out[1]).abs().max().item() < 2e-4 * max(1.0, out[1].abs().max().item())


def test_sampler_with_resident_plan_graph_replay(pipeline_cfg):
    """DDPMSampler(resident=...) = what the pipeline runs: 40 steps through CUDA-graph replay (2 graphs of 20 launches)
    against the oracle's loop on the same noise."""
    B, T, steps = 6, 1000, 40
    pos = pipeline_cfg["position_ddpm"]
    d = pos["diffusion_config"]
    sd = common.state_dict("pos")
    label = torch.arange(B) % 13
    smp = pipeline.DDPMSampler(pos["pointnet_config"], sd, B, engine.position_table(T, d["beta_0"], d["beta_T"]), 0, 0, T,
                               torch.device("cuda"), resident=dict(cluster=4, precise=True))
    assert smp.resident and smp.launches_per_step() == 1
    smp.set_labels(label.cuda())
    g = torch.Generator().manual_seed(12)
    x_T = torch.randn(B, 16, 3, generator=g)
    noises = {t: torch.randn(B, 16, 3, generator=g) for t in range(T - 1, T - 1 - steps, -1)}
    nz = smp.noise_view()
    for t, v in noises.items():
        nz[t].copy_(v.reshape(B * 16, 3))
    smp.x_view().copy_(x_T.reshape(-1, 3))
    smp.run(steps)
    torch.cuda.synchronize()
    got = smp.x_view().cpu().reshape(B, 16, 3)
    net = lambda x, ts: ref_model.cloud_condition_net(x, ref_model.Params(sd), pos["pointnet_config"], ts=ts, label=label)
    with torch.no_grad():
        want = ref_model.position_sampling(net, x_T, noises, ref_model.position_schedule(T, d["beta_0"], d["beta_T"]),
                                           n_steps=steps)
    assert (got - want).abs().max().item() < 5e-4 * max(1.0, want.abs().max().item())
