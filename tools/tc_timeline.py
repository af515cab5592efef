"""Phase timeline of the tcgen05 GEMM's CTAs for chosen records of one DDPM step (tuning only).
Builds libslide_b200_tl.so with -DTC_TIMELINE when missing (do that HERE, before gpurun) and must be started with
SLIDE_B200_LIB pointing at it:   SLIDE_B200_LIB=slide_b200/libslide_b200_tl.so python tools/tc_timeline.py pos 256 rec1,rec2
Prints, per record, the median / p90 over CTAs of the cycles between consecutive stamps (slot layout: gemm_tc.cu)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

NAMES = ["start", "prologue", "A-issued", "stage0-pub", "last-pub", "mma-saw0", "last-commit", "accum-ready",
         "resid-tab", "drained", "closing-bar", "end"]


def main():
    if "--build" in sys.argv:
        from slide_b200 import build
        print(build.build_variant("tl", ["-DTC_TIMELINE"]))
        return
    import torch
    from slide_b200 import engine, weights, lib
    from slide_b200.program import Program
    which, B, recs = sys.argv[1], int(sys.argv[2]), sys.argv[3].split(",")
    cfg = weights.load_json("pipeline_airplane.json")
    if which == "pos":
        pc = cfg["position_ddpm"]["pointnet_config"]
        d = cfg["position_ddpm"]["diffusion_config"]
        table, mode, keep = engine.position_table(d["T"], d["beta_0"], d["beta_T"]), 0, 0
        sd = weights.random_state_dict(weights.load_json("schema_position_ddpm.json"), 1)
    else:
        pc = cfg["latent_ddpm"]["pointnet_config"]
        table, mode, keep = engine.latent_table(cfg["latent_ddpm"]["standard_diffusion_config"]), 1, 3
        sd = weights.random_state_dict(weights.load_json("schema_latent_ddpm.json"), 2)
    b, h = engine.build_ddpm(pc, sd, B, 1000, table, mode, keep_cols=keep, with_noise=False)
    prog = Program(b)
    engine.init_constants(prog, h)
    prog.upload(h["labels"], torch.zeros(B, dtype=torch.int32))
    prog.run_segment("setup")
    prog.upload(h["x"], torch.randn(B * 16, h["C"]))
    first, count = b.segments["forward"]
    prog.set_step(501)
    prog.run(first, count)
    torch.cuda.synchronize()
    L = lib.load()
    L.slide_debug_tc_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    notes = {b.ops[first + j][3]: j for j in range(count)}
    buf = np.zeros((8192, 16), dtype=np.uint64)
    for key in recs:
        i = notes[key]
        for rep in range(2):  # second run = warm L2 (the state inside a step)
            L.slide_debug_tc_timeline(buf.ctypes.data, 8192, 1)
            prog.run(first + i, 1)
            torch.cuda.synchronize()
        L.slide_debug_tc_timeline(buf.ctypes.data, 8192, 1)
        t = buf.astype(np.int64)
        live = t[:, 0] > 0
        t = t[live]
        n = len(t)
        g0 = t[:, 12].min()
        if (t[:, 10] > 0).all() and (t[:, 10] < 1000).all():
            # persistent kernel: per-role cycle accumulators
            life = t[:, 11] - t[:, 0]
            print("== %s [persistent]: %d CTAs, tiles/CTA %.1f, CTA life median %.0f cyc = %.1f us" % (
                key, n, t[:, 10].mean(), np.median(life), np.median(life) / 1.9e3))
            for slot, name in ((9, "copy warp: wait for a free stage"), (8, "transform warp 0: wait for raw tile"),
                               (1, "MMA: wait for operands"), (2, "MMA: wait for a drained accumulator"),
                               (3, "epilogue warp 0: wait for accumulator"), (4, "epilogue warp 0: tcgen05.ld"),
                               (5, "epilogue warp 0: transpose stores"), (6, "epilogue warp 0: rest of chunk")):
                print("   %-42s median %8.0f cyc (%4.1f%% of life)" % (name, np.median(t[:, slot]),
                                                                     100.0 * np.median(t[:, slot] / life)))
            continue
        print("== %s: %d CTAs; kernel span %.1f us; CTA life median %.0f cyc (p90 %.0f)" % (
            key, n, (t[:, 12].max() - g0) / 1e3, np.median(t[:, 11] - t[:, 0]), np.percentile(t[:, 11] - t[:, 0], 90)))
        starts = np.sort(t[:, 12] - g0)
        print("   CTA start times (us): p10 %.1f p50 %.1f p90 %.1f max %.1f" % tuple(
            np.percentile(starts, q) / 1e3 for q in (10, 50, 90, 100)))
        for s in list(range(1, 9)) + [14, 15] + list(range(9, 12)):
            if (t[:, s] == 0).all():
                continue
            d = t[:, s] - t[:, 0]
            name = NAMES[s] if s < 12 else {14: "chunk0-read", 15: "chunk0-done"}[s]
            print("   %-12s at median %7.0f cyc  p90 %7.0f" % (name, np.median(d), np.percentile(d, 90)))


if __name__ == "__main__":
    main()
