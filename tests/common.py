"""Shared builders for the tests (CPU and GPU)."""
import numpy as np
import torch

from slide_b200 import engine, weights

SEEDS = {"pos": 11, "lat": 12, "ae": 13}


def state_dict(which):
    name = {"pos": "schema_position_ddpm.json", "lat": "schema_latent_ddpm.json", "ae": "schema_autoencoder.json"}[which]
    return weights.random_state_dict(weights.load_json(name), SEEDS[which])


def ddpm_program(cfg, which, B, with_noise=False, T=1000):
    """-> (builder, handles, pointnet_config, state_dict)"""
    if which == "pos":
        pc = cfg["position_ddpm"]["pointnet_config"]
        d = cfg["position_ddpm"]["diffusion_config"]
        table = engine.position_table(d["T"], d["beta_0"], d["beta_T"])
        mode, keep = 0, 0
    else:
        pc = cfg["latent_ddpm"]["pointnet_config"]
        table = engine.latent_table(cfg["latent_ddpm"]["standard_diffusion_config"])
        mode, keep = 1, 3
    sd = state_dict(which)
    b, h = engine.build_ddpm(pc, sd, B, T, table, mode, keep_cols=keep, with_noise=with_noise)
    return b, h, pc, sd


def init_machine(m, h, labels):
    engine.init_constants(m, h)
    m.upload(h["labels"], np.asarray(labels, dtype=np.int32))


def compare_tensors(builder, cpu, gpu_arena, rtol, note="", only=None):
    """Compare every arena tensor of the CPU interpreter with a downloaded GPU arena (uint8 numpy).
    Returns a list of (name, max_abs_err, scale) for tensors that deviate."""
    bad = []
    for t in (builder.tensors if only is None else only):
        dt = {"f32": np.float32, "i32": np.int32, "f64": np.float64}[t.dtype]
        n = t.rows * t.ld
        a = np.frombuffer(cpu.arena, dtype=dt, count=n, offset=t.off).reshape(t.rows, t.ld)[:, :t.C]
        g = np.frombuffer(gpu_arena, dtype=dt, count=n, offset=t.off).reshape(t.rows, t.ld)[:, :t.C]
        if t.dtype == "i32":
            if not np.array_equal(a, g):
                bad.append((t.name, float((a != g).mean()), 1.0))
            continue
        if not np.isfinite(g).all():
            bad.append((t.name, float("nan"), 0.0))
            continue
        scale = float(np.abs(a).max()) if a.size else 0.0
        err = float(np.abs(a.astype(np.float64) - g.astype(np.float64)).max()) if a.size else 0.0
        if err > rtol * max(scale, 1e-6) + 1e-7:
            bad.append((t.name, err, scale))
    return bad


def load_encode_inputs(m, h, golden, sample):
    engine.init_constants(m, h)
    m.upload(h["labels"], np.asarray(golden["label"], dtype=np.int32))
    m.upload(h["cloud"], golden["enc_cloud"])
    m.upload(h["keypoint"], golden["enc_kp"])
    if sample:
        m.upload(h["noises"][0], golden["enc_n1"])
        m.upload(h["noises"][1], golden["enc_n2"])
