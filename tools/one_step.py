"""One diffusion step of each DDPM (eager records, batch 256) + one decode chunk, bracketed by cudaProfilerStart/Stop:
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ... python tools/one_step.py
gives the launch list of exactly the kernels that make up the benchmark's step (each DDPM step repeats 1000x)."""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from slide_b200 import pipeline, weights  # noqa: E402


def main():
    cfg = weights.load_json("pipeline_airplane.json")
    pipe = pipeline.SlidePipeline(cfg, 256, ddpm_steps=1)
    labels = torch.full((256,), cfg["label"], dtype=torch.long)
    torch.manual_seed(0)
    pipe.draw_host_inputs(labels)
    pipe.stage_inputs()
    for s in (pipe.pos, pipe.lat):  # warm-up: one eager step each (also configures kernel attributes)
        s.prog.set_step(s.T)
        s.prog.run(*s.builder.segments["step"])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for s in (pipe.pos, pipe.lat):
        s.prog.set_step(s.T)
        s.prog.run(*s.builder.segments["step"])
    kp = pipe.pos.x_view().view(256, 16, 3)[:pipe.dec.chunk].contiguous()
    feat = pipe.lat.x_view().view(256, 16, pipe.lat.C)[:pipe.dec.chunk, :, 3:].contiguous()
    pipe.dec.run(kp, feat, pipe._in[pipe._slot]['labels'][:pipe.dec.chunk], pipe._in[pipe._slot]['starts'][:, :pipe.dec.chunk], pipe.out[:pipe.dec.chunk])
    if "--config1" in sys.argv:  # BASELINE config 1 through the drop-in _ext: FPS 2048 -> 1024 + ball query r=0.2 ns=32
        from slide_b200 import install_dropin
        install_dropin()
        from pointnet2_ops import _ext
        torch.manual_seed(7)
        xyz = torch.rand(1, 2048, 3, device="cuda")
        idx = _ext.furthest_point_sampling(xyz, 1024)
        new_xyz = _ext.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
        _ext.ball_query(new_xyz, xyz, 0.2, 32)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
