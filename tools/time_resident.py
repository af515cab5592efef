"""Time one position-DDPM step as a sample-resident kernel (clusters of 2 / 4, TF32 / 3xTF32) next to the per-record
executor.  usage: python tools/time_resident.py [B] [cluster:precise,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from slide_b200 import engine, weights
from slide_b200.program import Program


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfg = weights.load_json("pipeline_airplane.json")
    pos = cfg["position_ddpm"]
    d = pos["diffusion_config"]
    sd = weights.random_state_dict(weights.load_json("schema_position_ddpm.json"), 1)
    table = engine.position_table(d["T"], d["beta_0"], d["beta_T"])
    configs = ((4, False), (2, False), (4, True), (0, False))
    if len(sys.argv) > 2:  # e.g. "4:0,2:0"
        configs = tuple((int(c.split(":")[0]), bool(int(c.split(":")[1]))) for c in sys.argv[2].split(","))
    for cluster, precise in configs:
        b, h = engine.build_ddpm(pos["pointnet_config"], sd, B, 1000, table, 0,
                                 resident=dict(cluster=cluster, precise=precise) if cluster else None)
        prog = Program(b)
        for plan in h.get("resident_plans", []):
            prog.set_resident(plan)
        engine.init_constants(prog, h)
        prog.upload(h["labels"], torch.zeros(B, dtype=torch.int32))
        prog.run_segment("setup")
        prog.upload(h["x"], torch.randn(B * 16, 3))
        prog.view(h["noise"]).normal_()
        first, count = b.segments["step"]
        prog.set_step(1000)
        for _ in range(5):
            prog.run(first, count)
        prog.capture(0, first, count, repeat=20)
        prog.set_step(1000)
        prog.replay(0, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prog.set_step(1000)
        e0.record()
        prog.replay(0, 20)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 400
        info = h["resident_plans"][0].summary() if h.get("resident_plans") else {"per_record": True}
        finite = bool(torch.isfinite(prog.download(h["x"])).all())
        print("B=%d cluster=%d precise=%d: %.1f us/step (graph replay, 400 steps) launches/step=%d finite=%s %s"
              % (B, cluster, precise, us, prog.launches(first, count), finite, info), flush=True)
        prog.close()


if __name__ == "__main__":
    main()
