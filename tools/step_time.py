"""Device microseconds per sampling step of both DDPMs under CUDA-graph replay (what bench.py's chains run).
usage: python tools/step_time.py [batch] [replays]      (A/B: SLIDE_PDL=0|1, SLIDE_TC_* knobs, SLIDE_B200_LIB)"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from slide_b200 import engine, pipeline, weights  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    cfg = weights.load_json("pipeline_airplane.json")
    sds = pipeline.default_state_dicts()
    dev = torch.device("cuda", 0)
    d = cfg["position_ddpm"]["diffusion_config"]
    lat = cfg["latent_ddpm"]
    samplers = {
        "pos": pipeline.DDPMSampler(cfg["position_ddpm"]["pointnet_config"], sds["position"], B,
                                    engine.position_table(d["T"], d["beta_0"], d["beta_T"]), 0, 0, d["T"], dev),
        "lat": pipeline.DDPMSampler(lat["pointnet_config"], sds["latent"], B, engine.latent_table(lat["standard_diffusion_config"]),
                                    1, 3, lat["standard_diffusion_config"]["num_diffusion_timesteps"], dev),
    }
    out = []
    for name, s in samplers.items():
        s.set_labels(torch.full((B,), cfg["label"], dtype=torch.long, device=dev))
        s.noise_view().normal_()
        s.x_view().normal_()
        s.run(s.graph_steps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s.run(s.graph_steps * reps)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (s.graph_steps * reps)
        out.append("%s %.1f us/step (%d launches)" % (name, us, s.launches_per_step()))
        ok = bool(torch.isfinite(s.x_view()).all().item())
        assert ok, name
    print("B=%d SLIDE_PDL=%s: %s" % (B, os.environ.get("SLIDE_PDL", "default"), "; ".join(out)))


if __name__ == "__main__":
    main()
