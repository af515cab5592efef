"""Whole-chain parity on the GPU:
  * the position DDPM's full 1000-step ancestral chain at batch 32 (BASELINE config 2) through DDPMSampler (CUDA-graph
    replay, default dispatch): per-step eps against the oracle's network evaluated on the GPU's own x_t (tight), final
    x_0 against the oracle's own 1000-step loop (loose: 1000 TF32 denoiser evaluations feed back into x);
  * latent loop -> decode end to end through SlidePipeline.sample_resident against ref_model.latent_denoise + decode;
  * a fresh pipeline's first sample() equals its second (the capture warm-up step must not touch x).
Reference: pointnet2/util.py:197-259, diffusion_utils/diffusion.py:346-404, models/autoencoder.py:42-45.
"""
import numpy as np
import pytest
import torch

from oracle import ref_model
from slide_b200 import engine, lib, pipeline
from tests import common

pytestmark = pytest.mark.gpu

EPS_TOL = {"simt": 5e-5, "auto": 5e-3}   # x max|eps|, one denoiser forward
X0_TOL = {"simt": 2e-3, "auto": 5e-2}    # x max|x_0|, after 1000 steps of feedback


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_position_chain_1000_steps_b32(backend, pipeline_cfg):
    B, T = 32, 1000
    pos = pipeline_cfg["position_ddpm"]
    d = pos["diffusion_config"]
    sd = common.state_dict("pos")
    label = torch.arange(B) % 13
    g = torch.Generator().manual_seed(2024)
    x_T = torch.randn(B, 16, 3, generator=g)
    noises = {t: torch.randn(B, 16, 3, generator=g) for t in range(T - 1, 0, -1)}
    smp = pipeline.DDPMSampler(pos["pointnet_config"], sd, B, engine.position_table(T, d["beta_0"], d["beta_T"]), 0, 0, T,
                               torch.device("cuda"), backend=backend)
    smp.set_labels(label.cuda())
    nz = smp.noise_view()
    nz.zero_()
    for t, v in noises.items():
        nz[t].copy_(v.reshape(B * 16, 3))
    net = lambda x, ts: ref_model.cloud_condition_net(x, ref_model.Params(sd), pos["pointnet_config"], ts=ts, label=label)
    # the chain in segments of graph replays; at each checkpoint compare one forward on the GPU's own state
    smp.x_view().copy_(x_T.reshape(-1, 3))
    first, count = smp.builder.segments["step"]
    ffirst, fcount = smp.builder.segments["forward"]
    smp.run(0)  # warm-up + capture only (0 steps): must leave x untouched
    assert torch.equal(smp.x_view().cpu(), x_T.reshape(-1, 3)), "graph-capture warm-up modified the live state"
    done = 0
    for t_check in (999, 799, 499, 199, 19):
        n = (T - 1 - t_check) - done
        assert n % smp.graph_steps == 0
        smp.prog.replay(0, n // smp.graph_steps)
        done += n
        torch.cuda.synchronize()
        assert int(smp.prog.view(smp.builder.step).item()) == t_check + 1
        x_t = smp.x_view().cpu().reshape(B, 16, 3).clone()
        smp.prog.run(ffirst, fcount)          # forward only: decrements the counter, leaves x alone
        torch.cuda.synchronize()
        eps = smp.prog.download(smp.h["eps"]).cpu().reshape(B, 16, 3)
        smp.prog.set_step(t_check + 1)
        with torch.no_grad():
            want = net(x_t, torch.ones(B) * t_check)
        err = (eps - want).abs().max().item()
        assert err < EPS_TOL[backend] * max(1.0, want.abs().max().item()), (t_check, err)
    assert smp.graph_steps == 20 and done == T - 20
    smp.prog.replay(0, 1)  # the last 20 steps (t = 19 .. 0)
    torch.cuda.synchronize()
    assert int(smp.prog.view(smp.builder.step).item()) == 0
    got = smp.x_view().cpu().reshape(B, 16, 3)
    with torch.no_grad():
        want = ref_model.position_sampling(net, x_T, noises, ref_model.position_schedule(T, d["beta_0"], d["beta_T"]))
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    assert err < X0_TOL[backend] * max(1.0, want.abs().max().item()), err
    assert lib.load().slide_tc_error() == 0


def _hausdorff(a, b):
    d = torch.cdist(a, b)
    return max(d.min(1)[0].max().item(), d.min(0)[0].max().item())


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_latent_loop_then_decode_end_to_end(backend, pipeline_cfg):
    """feature DDPM (external keypoints, first `steps` steps of the 1000-step schedule) -> decode through the public
    pipeline, against the oracle's latent_denoise + decode on the same x_T / noise / start indices."""
    B, steps = 4, 6
    cfg = pipeline_cfg
    sds = {"position": common.state_dict("pos"), "latent": common.state_dict("lat"), "autoencoder": common.state_dict("ae")}
    pipe = pipeline.SlidePipeline(cfg, B, ddpm_steps=steps, backend=backend, state_dicts=sds, decode_chunk=B)
    labels = torch.full((B,), cfg["label"], dtype=torch.long)
    torch.manual_seed(3)
    pipe.draw_host_inputs(labels, skip_position=True)
    pipe.stage_inputs()
    g = torch.Generator().manual_seed(9)
    kp = torch.rand(B, 16, 3, generator=g) - 0.5
    torch.cuda.manual_seed(21)
    out = pipe.sample_resident(keypoints=kp.cuda()).cpu()
    pipe.check_device_errors()
    # the oracle on the same draws: x_T / starts from the pinned host buffers, per-step noise from the device buffer
    T = pipe.T_lat
    C = pipe.lat.C
    nz = pipe.lat.noise_view().cpu().reshape(T, B, 16, C)
    noises = {t: nz[t] for t in range(T - 1, T - 1 - steps, -1)}
    lat = cfg["latent_ddpm"]
    net = lambda x, ts: ref_model.cloud_condition_net(x, ref_model.Params(sds["latent"]), lat["pointnet_config"], ts=ts,
                                                      label=labels)
    with torch.no_grad():
        x = ref_model.latent_denoise(net, pipe._lat_xT_host.clone(), kp, noises,
                                     ref_model.latent_schedule(lat["standard_diffusion_config"]), n_steps=steps)
        feat = x[:, :, 3:]
        got_feat = pipe.keypoint_feature.cpu()
        tol = 2e-4 if backend == "simt" else 2e-2
        assert (got_feat - feat).abs().max().item() < tol * max(1.0, feat.abs().max().item())
        # decode the GPU's own features with the oracle (isolates the decoder; FPS picks are discrete)
        want, _ = ref_model.decode(kp, got_feat, ref_model.Params(sds["autoencoder"]), cfg["autoencoder"]["decoders"], labels,
                                start_idx_list=[pipe._starts_host[l] for l in range(pipe._starts_host.shape[0])])
    assert torch.isfinite(out).all() and out.shape == (B, 2048, 6)
    for i in range(B):
        assert _hausdorff(out[i, :, :3], want[i, :, :3]) < (2e-3 if backend == "simt" else 1e-2)


def test_fresh_pipeline_first_call_equals_second(pipeline_cfg):
    """Same seeds -> same result whether or not the CUDA graphs have been captured yet."""
    B, steps = 4, 20
    pipe = pipeline.SlidePipeline(pipeline_cfg, B, ddpm_steps=steps, decode_chunk=B, backend="simt")
    labels = torch.full((B,), pipeline_cfg["label"], dtype=torch.long)
    outs = []
    for _ in range(2):
        torch.manual_seed(5)
        pipe.draw_host_inputs(labels)
        torch.cuda.manual_seed(6)
        outs.append(pipe.sample_to_host().clone())
        kps = pipe.keypoint.cpu().clone()
        outs.append(kps)
    # keypoints (position DDPM) and clouds of the first call == second call (fp64 statistics atomics: tiny jitter)
    assert (outs[1] - outs[3]).abs().max().item() < 1e-4
    for i in range(B):
        assert _hausdorff(outs[0][i, :, :3], outs[2][i, :, :3]) < 2e-3
