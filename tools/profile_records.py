"""Per-record device time of one DDPM step (CUDA events on the launching stream, warm, median of N runs).
Usage: python tools/profile_records.py [pos|lat] [B] [auto|simt]   -> table on stdout + gpurun_out/records_<which>.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from slide_b200 import engine, weights  # noqa: E402
from slide_b200.program import Program, KIND_NAME, V  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "lat"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    backend = sys.argv[3] if len(sys.argv) > 3 else "auto"
    reps = 7
    cfg = weights.load_json("pipeline_airplane.json")
    if which == "refine":
        # the SAP refinement network: B clouds of 4096 mirrored points, one forward + the point split
        rc = weights.load_json("sap_refine.json")
        sd = weights.random_state_dict(weights.load_json("schema_sap_refine.json"), 21)
        b, h = engine.build_refine(rc["pointnet_config"], sd, B, 4096)
        b.segments["forward"] = b.segments["refine"]
        prog = Program(b)
        prog.set_gemm_backend(backend)
        engine.init_constants(prog, h)
        prog.upload(h["labels"], torch.zeros(B, dtype=torch.int32))
        prog.run_segment("setup")
        g = torch.Generator().manual_seed(0)
        x = torch.cat([torch.rand(B * 4096, 3, generator=g) - 0.5,
                       torch.nn.functional.normalize(torch.randn(B * 4096, 3, generator=g), dim=1),
                       torch.ones(B * 4096, 1)], dim=1)
        prog.upload(h["x"], x)
        return profile(b, prog, which, B, backend, reps)
    if which == "pos":
        pc = cfg["position_ddpm"]["pointnet_config"]
        d = cfg["position_ddpm"]["diffusion_config"]
        table, mode, keep = engine.position_table(d["T"], d["beta_0"], d["beta_T"]), 0, 0
        sd = weights.random_state_dict(weights.load_json("schema_position_ddpm.json"), 1)
    else:
        pc = cfg["latent_ddpm"]["pointnet_config"]
        table, mode, keep = engine.latent_table(cfg["latent_ddpm"]["standard_diffusion_config"]), 1, 3
        sd = weights.random_state_dict(weights.load_json("schema_latent_ddpm.json"), 2)
    T = 1000
    b, h = engine.build_ddpm(pc, sd, B, T, table, mode, keep_cols=keep, with_noise=False)
    prog = Program(b)
    prog.set_gemm_backend(backend)
    engine.init_constants(prog, h)
    prog.upload(h["labels"], torch.zeros(B, dtype=torch.int32))
    prog.run_segment("setup")
    prog.upload(h["x"], torch.randn(B * 16, h["C"]))
    return profile(b, prog, which, B, backend, reps)


def profile(b, prog, which, B, backend, reps):
    first, count = b.segments["forward"]
    prog.set_step(501)
    prog.run(first, count)
    torch.cuda.synchronize()
    only = os.environ.get("PROFILE_ONLY")
    if only:
        # ncu mode: `ncu --profile-from-start off ...` captures just these records (one launch each, cold L2)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        notes = {b.ops[first + j][3]: j for j in range(count)}
        for key in only.split(","):
            i = int(key) if key.isdigit() else notes[key]
            flush.zero_()
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            prog.run(first + i, 1)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
            print("profiled record", i, b.ops[first + i][3])
        return
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(first, first + count):
        kind, fields, _f, note = b.ops[i]
        ts = []
        for r in range(reps):
            if KIND_NAME[kind] == "SLIDE_OP_STEP_BEGIN":
                prog.set_step(501)
            flush.zero_()  # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            prog.run(i, 1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(ts))
        row = {"i": i, "kind": KIND_NAME[kind][9:], "note": note, "us": us}
        if KIND_NAME[kind] == "SLIDE_OP_GEMM":
            M, K, N = fields["GEMM_M"], fields["GEMM_K"], fields["GEMM_N"]
            row.update(M=M, K=K, N=N, gflop=2.0 * M * K * N / 1e9, tflops=2.0 * M * K * N / us / 1e6,
                       bytes=4.0 * (M * K + M * N + N * K))
        rows.append(row)
    prog.set_step(501)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        prog.run(first, count)
    e1.record()
    torch.cuda.synchronize()
    total = sum(r["us"] for r in rows)
    print("%s B=%d backend=%s: sum of records %.1f us (cold L2), back-to-back forward %.1f us" %
          (which, B, backend, total, e0.elapsed_time(e1) * 1e3 / 5))
    for r in rows:
        extra = ""
        if "M" in r:
            extra = "M=%6d K=%4d N=%4d %7.1f TFLOP/s %6.0f GB/s(min traffic)" % (r["M"], r["K"], r["N"], r["tflops"],
                                                                                 r["bytes"] / r["us"] / 1e3)
        print("%3d %-14s %-30s %9.1f us  %s" % (r["i"], r["kind"], r["note"], r["us"], extra))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "records_%s_%s_b%d.json" % (which, backend, B)), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
