"""Probe: does the position DDPM chain of the NEXT batch overlap with the feature DDPM chain of the current one when the
two run on separate streams?  (B200, batch 256; prints one JSON line: each chain alone, both concurrently.)"""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    import torch
    from slide_b200 import pipeline, weights
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfg = weights.load_json("pipeline_airplane.json")
    p = pipeline.SlidePipeline(cfg, B)
    labels = torch.full((B,), cfg["label"], dtype=torch.long)
    torch.manual_seed(0)
    p.draw_host_inputs(labels)
    p.sample()  # warm-up: captures the graphs
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def pos_alone():
        p.pos.x_view().normal_()
        p.pos.run()

    def lat_alone():
        p.lat.x_view().normal_()
        p.lat.run()

    def both():
        nonlocal s1, s2
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            p.pos.x_view().normal_()
            p.pos.run()
        with torch.cuda.stream(s2):
            p.lat.x_view().normal_()
            p.lat.run()
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    out = {"batch": B, "pos_ms": timed(pos_alone), "lat_ms": timed(lat_alone), "both_ms": timed(both)}
    lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
    out["priority_range"] = [lo, hi]
    for name, (p1, p2) in {"lat_high": (0, -1), "pos_high": (-1, 0), "lat_highest": (0, hi)}.items():
        s1, s2 = torch.cuda.Stream(priority=p1), torch.cuda.Stream(priority=p2)
        out["both_ms_" + name] = timed(both)
    out["sum_ms"] = out["pos_ms"] + out["lat_ms"]
    out["overlap_gain"] = out["sum_ms"] / out["both_ms"]
    out["finite"] = bool(torch.isfinite(p.pos.x_view()).all().item() and torch.isfinite(p.lat.x_view()).all().item())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
