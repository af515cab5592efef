#!/usr/bin/env python
"""Benchmark of the SLIDE sampling hot path on B200: shapes/sec for 1000-step position DDPM + 1000-step feature
DDPM + decode to 2048-point clouds at batch 256 (BASELINE.json), one process per GPU.

  python bench.py --gpus 1 --steps 2 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the CPU path (oracle port of the reference's modules) on the host cores

A "step" is one full pass of the pipeline over one batch of 256 synthetic shapes (random-init weights with the
reference's state-dict schema, labels = airplane).  `value` times the device path with inputs resident in HBM;
`e2e` adds the pinned host->device copies of every input and the device->host read of the clouds.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_SHAPE = 1151.7  # 1000 x 0.0780 + 1000 x 1.0571 + 16.607 (reference formulation)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the tcgen05 GEMM from the committed `ncu --set full`
    capture (profiles/r01_gemm_tc_ncu_full.txt, one cold launch per record), next to the algorithmic bytes of the same
    launches; None when the summary is missing."""
    path = os.path.join(ROOT, "profiles", "r01_gemm_tc_ncu_full.txt")
    try:
        rows = {}
        for line in open(path):
            if line.startswith("#") or "|" not in line:
                continue
            head, rest = line[:76].strip(), line[76:]
            rows[head.split("  ")[0].strip()] = [c.strip() for c in rest.split("|")]
        names = rows["Kernel Name"]
        rd = [c.split()[-1] for c in rows["dram__bytes_read.sum"]]
        wr = [c.split()[-1] for c in rows["dram__bytes_write.sum"]]
        out = []
        for n, r, w in zip(names, rd, wr):
            if "gemm_tc" in n:
                out.append({"kernel": n.replace("void ", "").split("(")[0], "dram_mbytes": float(r) + float(w)})
        return {"per_launch": out, "source": "profiles/r01_gemm_tc_ncu_full.txt",
                "records": "SA1.att.v, SA1.att.w2+softmax, SA1.mlp.conv1 of the feature-DDPM step (algorithmic MB: 269.5, 277.9, 134.5)"}
    except Exception:
        return None


def cpu_path(cfg, seconds_budget=20.0):
    """The reference's CPU path for this pipeline: its python modules as restated in oracle/ref_model.py (checked
    bit-for-bit against the real modules by tests/golden/make_golden.py) over the C oracle ops, all host cores.
    Bounded sample: a few denoiser forwards at batch 16 + one decode of 2 shapes, scaled to 1000+1000 steps."""
    import torch
    from oracle import ref_model
    from slide_b200 import pipeline
    torch.set_num_threads(os.cpu_count() or 1)
    sds = pipeline.default_state_dicts()
    Bs = 16
    label = torch.zeros(Bs, dtype=torch.long)
    per_shape = 0.0
    detail = {}
    with torch.no_grad():
        for key, name, C in (("position_ddpm", "position", 3), ("latent_ddpm", "latent", 51)):
            pc = cfg[key]["pointnet_config"]
            P = ref_model.Params(sds[name])
            x = torch.randn(Bs, 16, C)
            ref_model.cloud_condition_net(x, P, pc, ts=torch.ones(Bs) * 500, label=label)  # warm-up
            n, t0 = 0, time.time()
            while n < 3 or (time.time() - t0 < seconds_budget / 3 and n < 50):
                ref_model.cloud_condition_net(x, P, pc, ts=torch.ones(Bs) * 500, label=label)
                n += 1
            dt = (time.time() - t0) / n
            detail[name + "_s_per_step_b%d" % Bs] = dt
            per_shape += 1000 * dt / Bs
        P = ref_model.Params(sds["autoencoder"])
        kp, feat = torch.rand(2, 16, 3) - 0.5, torch.randn(2, 16, 48)
        t0 = time.time()
        ref_model.decode(kp, feat, P, cfg["autoencoder"]["decoders"], label[:2])
        detail["decode_s_b2"] = time.time() - t0
        per_shape += detail["decode_s_b2"] / 2
    return 1.0 / per_shape, torch.get_num_threads(), detail


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="slide_b200", choices=["slide_b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="global batch (BASELINE: 256)")
    ap.add_argument("--category", default="airplane")
    ap.add_argument("--ddpm-steps", type=int, default=None, help="DEBUG ONLY: truncate both DDPM loops (invalid as a benchmark)")
    ap.add_argument("--backend", default="auto", choices=["auto", "simt"])
    ap.add_argument("--decode-chunk", type=int, default=128)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from slide_b200 import weights
    cfg = weights.load_json("pipeline_%s.json" % args.category)
    workload = "position DDPM 1000 steps + feature DDPM 1000 steps + decode to 2048 pts, %s config" % args.category

    if args.impl == "reference":
        if rank != 0:
            return
        vals = []
        for i in range(args.warmup + args.steps):
            v, cores, detail = cpu_path(cfg, seconds_budget=12.0)
            if i >= args.warmup:
                vals.append(v)
        v = sum(vals) / len(vals)
        sample = "3+ denoiser forwards per DDPM at batch 16 scaled to 1000 steps, decode of 2 shapes; per step of this arm"
        print(json.dumps({
            "impl": "reference", "metric": "shapes/sec (1000-step DDPM + decode to 2048 pts) @ bs256", "value": v,
            "unit": "shapes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * args.batch / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "global_batch": args.batch, "timing": "host wall clock, extrapolated"},
            "cpu_baseline": {"value": v, "unit": "shapes/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "detail": detail}))
        return

    import torch
    import torch.distributed as dist
    from slide_b200 import pipeline, lib
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    # weak scaling: every GPU works on a full batch of `--batch` shapes; the job's batch is gpus x 256
    Bl = args.batch
    B = Bl * world
    pipe = pipeline.SlidePipeline(cfg, B, rank=rank, world=world, ddpm_steps=args.ddpm_steps, backend=args.backend,
                                  decode_chunk=args.decode_chunk)
    label_id = cfg["label"]
    labels = torch.full((B,), label_id, dtype=torch.long)
    torch.manual_seed(0)
    pipe.draw_host_inputs(labels)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = pipe.sample()
        gathered = pipeline.all_gather_outputs(out, world)
    pipe.stage_inputs()  # resident-input arm: everything the step reads is in HBM before the clock starts
    barrier()
    lib.reset_launch_count()

    clocks = ClockSampler(local_rank) if rank == 0 else None
    wall0 = time.time()
    # ---- arm 1: device path, inputs resident (noise / x_T already in HBM from the warm-up staging) ----------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = pipe.sample_resident()
        gathered = pipeline.all_gather_outputs(out, world)
    ev1.record()
    barrier()
    ms_value = ev0.elapsed_time(ev1) / args.steps
    launches = lib.launch_count() // max(args.steps, 1)
    # ---- arm 2: end to end through the public API, host buffers in, host buffer out -------------------------
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        host = pipe.sample_to_host()
        if world > 1:
            gathered = pipeline.all_gather_outputs(pipe.out, world)
    barrier()
    ms_e2e = 1e3 * (time.time() - t0) / args.steps
    wall1 = time.time()
    clock_info = clocks.stop(wall0, wall1) if clocks else None

    t = torch.tensor([ms_value, ms_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_value, ms_e2e = t.tolist()
    finite = bool(torch.isfinite(out).all().item())
    tc_err = lib.load().slide_tc_error()

    roof = None
    cpu = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        which = "measured bf16 sustained (MEASURED_PEAKS.json)" if peaks else "fallback 1.4 PF sustained"
        # dominant kernel = gemm_tc_kernel (tcgen05): time every dense GEMM record of one feature-DDPM step on the
        # launching stream with CUDA events (warm, back to back with its neighbours' data in L2 as in the real step)
        from slide_b200.program import KIND
        lat = pipe.lat
        first, count = lat.builder.segments["forward"]
        lat.prog.set_step(lat.T)
        lat.prog.run(first, count)
        torch.cuda.synchronize()
        flops = t_us = abytes = 0.0
        n_launch = 0
        for i in range(first, first + count):
            kind, f, _fl, _note = lat.builder.ops[i]
            if kind != KIND["SLIDE_OP_GEMM"] or f.get("GEMM_WP_W", -1) < 0 or f["GEMM_M"] < 128:
                continue
            # algorithmic bytes of the launch: A in, W in, C out (+ the residual / soft-max value rows it reads); fp32
            rows_out = f["GEMM_M"] // f["GEMM_SMK"] if f.get("GEMM_SMK", 0) > 0 else f["GEMM_M"]
            abytes += 4.0 * (f["GEMM_M"] * f["GEMM_K"] + f["GEMM_N"] * f["GEMM_K"] + rows_out * f["GEMM_N"] +
                             (f["GEMM_M"] * f["GEMM_N"] if f.get("GEMM_RES", -1) >= 0 else 0))
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(3):
                lat.prog.run(i, 1)
            a1.record()
            torch.cuda.synchronize()
            t_us += a0.elapsed_time(a1) * 1e3 / 3
            flops += 2.0 * f["GEMM_M"] * f["GEMM_K"] * f["GEMM_N"]
            n_launch += 1
        achieved = flops / t_us / 1e6  # TFLOP/s
        whole = (Bl / (ms_value / 1e3)) * GFLOP_PER_SHAPE / 1e3
        traffic = ncu_traffic()
        traffic_mean = (1e6 * sum(x["dram_mbytes"] for x in traffic["per_launch"]) / len(traffic["per_launch"])
                        if traffic and traffic["per_launch"] else None)
        hbm_peak = peaks.get("hbm_gbs", 6500.0)
        hbm_gbs = abytes / t_us / 1e3
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic_mean, "traffic_unit": "bytes per launch (mean of the ncu-captured GEMM launches)",
                "traffic_detail": traffic, "peak_source": which,
                "kernel": "gemm_tcp_kernel / gemm_tc_kernel (tcgen05.mma kind::tf32, persistent + one-tile variants)",
                # the same launches against the HBM roofline: with fp32 activations in HBM most of these GEMMs
                # (K, N <= 256: < 64 FLOP/B) sit left of the TF32 ridge, so this is the view that bounds them
                "hbm_view": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                             "algorithmic_bytes_per_launch_set": abytes},
                "launches_timed": n_launch, "avg_launch_us": t_us / max(n_launch, 1),
                "executed_gflop_per_launch_set": flops / 1e9,
                "whole_step_achieved": whole, "whole_step_frac": whole / peak,
                "note": "achieved = executed FLOPs of the %d tensor-core GEMM launches of one feature-DDPM step / their "
                        "CUDA-event time; whole_step_* = reference-formulation FLOPs (1151.7 GFLOP/shape) / device time. "
                        "Operands are TF32 (nominal dense peak = half of bf16); peak shown is the measured bf16 figure."
                        % n_launch}
        if args.ddpm_steps is None:
            v, cores, detail = cpu_path(cfg)
            cpu = {"value": v, "unit": "shapes/s", "cores": cores, "kind": "port",
                   "sample": "denoiser forwards at batch 16 scaled to 1000+1000 steps + decode of 2 shapes", "detail": detail}
        valid = args.ddpm_steps is None and finite and tc_err == 0
        print(json.dumps({
            "metric": "shapes/sec (1000-step DDPM + decode to 2048 pts) @ bs256", "value": B / (ms_value / 1e3),
            "unit": "shapes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if args.backend == "auto" else "f32",
            "data": "synthetic",
            "config": {"workload": workload, "global_batch": B, "per_gpu_batch": Bl, "parallelism": "dp%d" % world,
                       "l2": "inputs larger than L2 (noise tensors 49 MB + 836 MB per GPU)", "valid": valid,
                       "ddpm_steps": args.ddpm_steps or 1000, "backend": args.backend},
            "e2e": {"value": B / (ms_e2e / 1e3), "unit": "shapes/s", "h2d_bytes_per_step": pipe.h2d_bytes(),
                    "d2h_bytes_per_step": pipe.d2h_bytes(), "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roof, "cpu_baseline": cpu,
            "finite": finite, "tc_error": tc_err}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
