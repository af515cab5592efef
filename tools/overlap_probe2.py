"""Probe: two independent half-batch feature-DDPM chains (2 x 128 shapes) on two streams against one 256-shape chain."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    import torch
    from slide_b200 import pipeline, weights
    cfg = weights.load_json("pipeline_airplane.json")
    pipes = [pipeline.SlidePipeline(cfg, 128) for _ in range(2)]
    big = pipeline.SlidePipeline(cfg, 256)
    for p in pipes + [big]:
        labels = torch.full((p.B,), cfg["label"], dtype=torch.long)
        torch.manual_seed(0)
        p.draw_host_inputs(labels)
        p.sample()
    torch.cuda.synchronize()
    s = [torch.cuda.Stream(), torch.cuda.Stream()]

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def one(p):
        p.lat.x_view().normal_()
        p.lat.run()

    def two():
        cur = torch.cuda.current_stream()
        for st, p in zip(s, pipes):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                one(p)
        for st in s:
            cur.wait_stream(st)

    out = {"lat256_ms": timed(lambda: one(big)), "lat128_ms": timed(lambda: one(pipes[0])), "two_lat128_ms": timed(two)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
