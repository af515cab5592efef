#!/usr/bin/env python
"""Benchmark of the SLIDE sampling hot path on B200: shapes/sec for 1000-step position DDPM + 1000-step feature
DDPM + decode to 2048-point clouds at batch 256 (BASELINE.json), one process per GPU.

  python bench.py --gpus 1 --steps 2 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the reference's own CPU path: its UNMODIFIED python (baseline/_ref) on the host cores

A "step" is one full pass of the pipeline over one batch of 256 synthetic shapes (random-init weights with the
reference's state-dict schema, labels = airplane).  `value` times the device path with inputs resident in HBM;
`e2e` adds, every step, the host-side RNG draws the reference's loop makes (x_T, 1000 position-noise tensors, FPS start
indices), the pinned host->device copies of every input and the device->host read of the clouds.

Besides the headline (weak scaling, 256 shapes per GPU) the same JSON line carries
  strong      BASELINE config 4 as written: global batch 256, chair, sharded 256/N per GPU (scaling "strong")
  configs     BASELINE configs 1 (FPS + ball query, timed beside the reference's own .cu from oracle/_ref), 2, 3 and 5
  reference_gpu_eager   the reference's eager GPU path (its python + its CUDA extension built for sm_100a) on this box
  parity      epsilon of both denoisers at batch 256, default dispatch vs the fp32 FFMA backend, same process
(--no-extras skips them).
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


def run_json(cmd, timeout):
    """Run a helper process (keeps its monkey-patches / module stubs out of this one) and parse its last JSON line."""
    try:
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout,
                           env=dict(os.environ, PYTHONPATH=ROOT))
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (r.stderr or r.stdout)[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def time_events(fn, iters, warm=3):
    """Mean microseconds per call of fn() on the current stream (CUDA events, warm-up first)."""
    import torch
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def config1_index_ops():
    """BASELINE config 1: FPS (2048 -> 1024) + ball query (r = 0.2, nsample = 32) on one 1 x 2048 x 3 cloud -- the
    reference's own unit -- through the C ABI, next to the reference's own kernels (oracle/_ref, its .cu
    compiled for sm_100a) on the same inputs and the same box; outputs compared bit for bit."""
    import torch
    from slide_b200 import install_dropin
    install_dropin()
    from pointnet2_ops import _ext as ours
    torch.manual_seed(7)
    xyz = torch.rand(1, 2048, 3, device="cuda")
    idx = ours.furthest_point_sampling(xyz, 1024)
    new_xyz = ours.gather_points(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    out = {"workload": "FPS 2048->1024 + ball_query r=0.2 ns=32, 1x2048x3",
           "fps_us": time_events(lambda: ours.furthest_point_sampling(xyz, 1024), 50),
           "ball_query_us": time_events(lambda: ours.ball_query(new_xyz, xyz, 0.2, 32), 200)}
    # algorithmic bytes: FPS 12 N + 4 m, ball query 12 N + 12 m + 4 m ns
    out["fps_algorithmic_bytes"] = 12 * 2048 + 4 * 1024
    out["ball_query_algorithmic_bytes"] = 12 * 2048 + 12 * 1024 + 4 * 1024 * 32
    try:
        from oracle import build_ref
        ref = build_ref.load_module()
    except Exception:  # noqa: BLE001
        ref = None
    if ref is not None:
        out["reference_fps_us"] = time_events(lambda: ref.furthest_point_sampling(xyz, 1024), 50)
        out["reference_ball_query_us"] = time_events(lambda: ref.ball_query(new_xyz, xyz, 0.2, 32), 200)
        def same(a, b):
            a, b = (a if isinstance(a, (tuple, list)) else (a,)), (b if isinstance(b, (tuple, list)) else (b,))
            return len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))
        out["bit_exact_vs_reference"] = bool(same(ref.furthest_point_sampling(xyz, 1024), idx) and
                                             same(ref.ball_query(new_xyz, xyz, 0.2, 32),
                                                  ours.ball_query(new_xyz, xyz, 0.2, 32)))
        out["speedup_fps"] = out["reference_fps_us"] / out["fps_us"]
        out["speedup_ball_query"] = out["reference_ball_query_us"] / out["ball_query_us"]
    else:
        out["reference"] = "oracle/_ref not built"
    return out


def sampler_step_us(sampler, replays=4):
    """Device microseconds per sampling step of a DDPMSampler (graph replays of `graph_steps` steps each)."""
    import torch
    sampler.x_view().normal_()
    sampler.run(sampler.graph_steps)  # captures on first use
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sampler.run(sampler.graph_steps * replays)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (sampler.graph_steps * replays)


def eps_parity(sampler, seed):
    """One denoiser forward at t = T/2 under the default dispatch and under the fp32 FFMA backend on identical inputs
    -> max |d eps| / max |eps|.  The default dispatch is what the timed region ran."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(sampler.x_view().shape, device="cuda", generator=g)
    first, count = sampler.builder.segments["forward"]
    eps = []
    for backend in ("auto", "simt"):
        sampler.prog.set_gemm_backend(backend)
        sampler.x_view().copy_(x)
        sampler.refresh_geometry()
        sampler.prog.set_step(sampler.T // 2)
        sampler.prog.run(first, count)
        torch.cuda.synchronize()
        eps.append(sampler.prog.download(sampler.h["eps"]).float().clone())
    sampler.prog.set_gemm_backend("auto")
    scale = float(eps[1].abs().max())
    return float((eps[0] - eps[1]).abs().max()) / max(scale, 1e-12)


def extra_configs(cfg, args):
    """BASELINE configs 2, 3, 5 (and 1) on this GPU -- single-GPU parity-test cases of BASELINE.json, timed here so that
    the driver's record carries them.  Each runs the full workload once after one warm-up pass."""
    import torch
    from slide_b200 import engine, pipeline
    dev = torch.device("cuda", torch.cuda.current_device())
    out = {"config1": config1_index_ops()}
    sds = pipeline.default_state_dicts()
    # config 2: position DDPM, 16 latent points, airplane, 1000 steps, batch 32
    d = cfg["position_ddpm"]["diffusion_config"]
    pos = pipeline.DDPMSampler(cfg["position_ddpm"]["pointnet_config"], sds["position"], 32,
                               engine.position_table(d["T"], d["beta_0"], d["beta_T"]), 0, 0, d["T"], dev, backend=args.backend)
    pos.set_labels(torch.full((32,), cfg["label"], dtype=torch.long, device=dev))
    pos.noise_view().normal_()

    def run_pos():
        pos.x_view().normal_()
        pos.run()
    us = time_events(run_pos, 1, warm=1)
    out["config2"] = {"workload": "position DDPM 1000 steps, batch 32, airplane", "executor": "one kernel per record",
                      "ms": us / 1e3, "us_per_step": us / d["T"],
                      "shapes_per_s": 32 / (us / 1e6), "launches_per_step": pos.launches_per_step(),
                      "finite": bool(torch.isfinite(pos.x_view()).all().item())}
    del pos
    # the same chain as ONE sample-resident kernel per step (what SlidePipeline picks at this batch size)
    pos = pipeline.DDPMSampler(cfg["position_ddpm"]["pointnet_config"], sds["position"], 32,
                               engine.position_table(d["T"], d["beta_0"], d["beta_T"]), 0, 0, d["T"], dev, backend=args.backend,
                               resident=dict(cluster=4, precise=False))
    if pos.resident:
        pos.set_labels(torch.full((32,), cfg["label"], dtype=torch.long, device=dev))
        pos.noise_view().normal_()
        us = time_events(run_pos, 1, warm=1)
        out["config2"]["resident"] = {"executor": "sample-resident kernel, cluster of 4 CTAs per sample", "ms": us / 1e3,
                                      "us_per_step": us / d["T"], "shapes_per_s": 32 / (us / 1e6),
                                      "launches_per_step": pos.launches_per_step(),
                                      "finite": bool(torch.isfinite(pos.x_view()).all().item())}
    del pos
    # config 3: feature DDPM on fixed keypoints + decode to 2048 points, batch 128
    p3 = pipeline.SlidePipeline(cfg, 128, backend=args.backend, decode_chunk=min(args.decode_chunk, 128))
    labels = torch.full((128,), cfg["label"], dtype=torch.long)
    torch.manual_seed(3)
    p3.draw_host_inputs(labels, skip_position=True)
    p3.stage_inputs()
    kp = (torch.rand(128, 16, 3, device=dev) - 0.5)
    us = time_events(lambda: p3.sample_resident(keypoints=kp), 1, warm=1)
    out["config3"] = {"workload": "feature DDPM 1000 steps on fixed keypoints + decode to 2048 pts, batch 128, airplane",
                      "ms": us / 1e3, "shapes_per_s": 128 / (us / 1e6), "finite": bool(torch.isfinite(p3.out).all().item())}
    del p3
    torch.cuda.empty_cache()
    # config 5: autoencoder encode + decode sweep (tools/bench_autoencoder.py), batch 512, one GPU
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_autoencoder.py"), "--batch", "512", "--reps", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    rows = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    out["config5"] = [{k: row[k] for k in ("points", "batch", "encode_ms", "decode_ms", "shapes_per_s", "hbm_roofline", "finite")}
                      for row in rows] or {"error": (r.stderr or r.stdout)[-300:]}
    # the SAP mesh-reconstruction stage (refine network + DPSR), batch 32 = the shipped eval_batch_size, next
    # to the reference's own code for the same stage on this box (eager GPU: its python + its CUDA extension + cuFFT; CPU:
    # its python on the host cores with the C-oracle native ops, a 2-cloud sample)
    sap = run_json([sys.executable, os.path.join(ROOT, "tools", "bench_sap.py"), "--batch", "32", "--reps", "5"], 300)
    sap["reference_gpu_eager"] = run_json([sys.executable, "-m", "oracle.reference_arms", "sap-gpu", "--batch", "32",
                                           "--steps", "3"], 300)
    sap["reference_cpu"] = run_json([sys.executable, "-m", "oracle.reference_arms", "sap-cpu", "--batch", "2", "--steps", "1"], 300)
    if "clouds_per_s" in sap and "clouds_per_s" in sap["reference_gpu_eager"]:
        sap["speedup_vs_reference_gpu_eager"] = sap["clouds_per_s"] / sap["reference_gpu_eager"]["clouds_per_s"]
    out["sap"] = sap
    return out


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
    ap.add_argument("--no-overlap", action="store_true", help="run the steps strictly one after the other (no cross-batch "
                    "overlap of the next step's position DDPM with this step's feature DDPM)")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip strong / configs / reference_gpu_eager")
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
        # The reference's own CPU implementation of the path: its UNMODIFIED python (baseline/_ref: util.sampling,
        # LatentDiffusion.denoise_and_reconstruct, PointNet2CloudCondition, PointAutoencoder) on all host cores, the native
        # ops it only ships for CUDA supplied by the C oracle.  Each step of this arm is a bounded sample (a few steps of
        # each DDPM at batch 16 + one decode) scaled to the full workload; kind = "port" only if the mirror is absent.
        from oracle import reference_arms
        budget = max(4.0, min(12.0, 150.0 / max(args.warmup + args.steps, 1)))
        vals, wall0 = [], time.time()
        for i in range(args.warmup + args.steps):
            v, cores, kind, sample, detail = reference_arms.cpu_arm(cfg, budget, 16)
            if i >= args.warmup:
                vals.append(v)
        v = sum(vals) / len(vals)
        print(json.dumps({
            "impl": "reference", "metric": "shapes/sec (1000-step DDPM + decode to 2048 pts) @ bs256", "value": v,
            "unit": "shapes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * args.batch / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "global_batch": args.batch,
                       "timing": "host wall clock of a bounded sample, scaled to 1000+1000 steps + decode of 256 shapes",
                       "wall_s": time.time() - wall0},
            "cpu_baseline": {"value": v, "unit": "shapes/s", "cores": cores, "kind": kind, "sample": sample + "; per step of this arm"},
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
    # RNG scope: like the reference's ranks, every rank draws its own 256 shapes from its own generators (seed = rank);
    # the world-size-invariant mode (every rank draws the global batch and keeps its rows) is the generation driver's and
    # is what the strong-scaling arm below uses
    scope = "rank" if world > 1 else "global"
    pipe = pipeline.SlidePipeline(cfg, B, rank=rank, world=world, ddpm_steps=args.ddpm_steps, backend=args.backend,
                                  decode_chunk=args.decode_chunk, rng_scope=scope)
    label_id = cfg["label"]
    labels = torch.full((Bl if scope == "rank" else B,), label_id, dtype=torch.long)
    torch.manual_seed(rank)
    pipe.draw_host_inputs(labels)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = pipe.sample()
        gathered = pipeline.all_gather_outputs(out, world)
    stages = pipe.stage_ms()  # of the last warm-up step: stages back to back, nothing overlapped
    pipe.stage_inputs()  # resident-input arm: everything the step reads is in HBM before the clock starts
    barrier()
    lib.reset_launch_count()

    clocks = ClockSampler(local_rank) if rank == 0 else None
    wall0 = time.time()
    # ---- arm 1: device path, inputs resident (noise / x_T already in HBM from the warm-up staging) ----------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    # Steps are pipelined across batches: while step i runs its feature DDPM + decode, the position DDPM of step i+1 runs
    # on a side stream (SlidePipeline.prefetch_position).  All K position chains, K feature chains and K decodes execute
    # inside the timed region: step 1's position chain runs un-overlapped, the last step prefetches nothing.
    overlap = not args.no_overlap
    for i in range(args.steps):
        out = pipe.sample_resident(prefetch_next=overlap and i + 1 < args.steps)
        gathered = pipeline.all_gather_outputs(out, world)
    ev1.record()
    barrier()
    ms_value = ev0.elapsed_time(ev1) / args.steps
    launches = lib.launch_count() // max(args.steps, 1)
    # ---- arm 2: end to end through the public API, host buffers in, host buffer out -------------------------
    # every step: the host RNG draws of the reference's loop (util.py:131-136,225,253; diffusion.py:373; the decoder's FPS
    # start indices), pinned H2D of all inputs, the three stages, D2H of the clouds
    # (the draws of step i+1 are made by the host while the GPU runs step i -- sample_to_host(next_labels=...) -- so only
    # the first draw is exposed; every step still consumes freshly drawn inputs, drawn inside the timed region)
    barrier()
    t0 = time.time()
    pipe.draw_host_inputs(labels)
    ms_rng = 1e3 * (time.time() - t0)
    for i in range(args.steps):
        host = pipe.sample_to_host(next_labels=labels if i + 1 < args.steps else None, overlap=overlap)
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

    # ---- BASELINE config 4 as written: global batch 256, chair, SHARDED 256 / N per GPU (strong scaling); the reference
    # splits the evaluation batch the same way (pointnet2/mesh_evaluation.py:51 int(eval_batch_size / world_size))
    strong = None
    if args.ddpm_steps is None and not args.no_extras and 256 % world == 0:
        cfg_s = weights.load_json("pipeline_chair.json")
        sp = pipeline.SlidePipeline(cfg_s, 256, rank=rank, world=world, backend=args.backend,
                                    decode_chunk=min(args.decode_chunk, 256 // world))
        torch.manual_seed(1)
        sp.draw_host_inputs(torch.full((256,), cfg_s["label"], dtype=torch.long))
        g_s = pipeline.all_gather_outputs(sp.sample(), world)  # warm-up (captures the graphs)
        sp.stage_inputs()
        ns = max(1, min(args.steps, 6))  # enough steps for the cross-batch pipeline to fill (first step: no overlap)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(ns):
            g_s = pipeline.all_gather_outputs(sp.sample_resident(prefetch_next=overlap and i + 1 < ns), world)
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1) / ns], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_s = float(ts.item())
        strong = {"scaling": "strong", "value": 256 / (ms_s / 1e3), "unit": "shapes/s", "ms_per_step": ms_s, "steps": ns,
                  "warmup": 1, "global_batch": 256, "per_gpu_batch": 256 // world, "category": "chair", "n_gpus": world,
                  "gathered_shape": list(g_s.shape), "finite": bool(torch.isfinite(g_s).all().item()),
                  "launches_per_step": {"position": sp.pos.launches_per_step(), "latent": sp.lat.launches_per_step()}}
        del sp, g_s
        torch.cuda.empty_cache()

    # ---- BASELINE config 5 at N > 1 GPUs: the autoencoder sweep with its batch of 512 clouds sharded 512 / N per GPU
    # (every cloud is independent: no collective); time = max over ranks.  At N = 1 the sweep is in `configs` below.
    ae_sweep = None
    if args.ddpm_steps is None and not args.no_extras and world > 1 and 512 % world == 0:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_autoencoder
        rows = bench_autoencoder.sweep(512 // world, reps=1)
        tt = torch.tensor([[r["encode_ms"], r["decode_ms"]] for r in rows], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ae_sweep = [{"points": r["points"], "global_batch": 512, "per_gpu_batch": 512 // world, "n_gpus": world,
                     "encode_ms": float(tt[i, 0]), "decode_ms": float(tt[i, 1]),
                     "shapes_per_s": 512 / (float(tt[i, 0] + tt[i, 1]) / 1e3), "finite": r["finite"]}
                    for i, r in enumerate(rows)]
        torch.cuda.empty_cache()

    roof = None
    cpu = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)  # whole step: a long run -> the sustained figure
        peak_burst = peaks.get("bf16_tflops", 1600.0)       # GEMM launches event-timed in isolation -> the burst figure
        which = "measured bf16 burst (MEASURED_PEAKS.json)" if peaks else "fallback 1.6 PF burst"
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
        # which roofline binds this launch set: the larger of the two floors -- algorithmic bytes / measured HBM bandwidth
        # against executed FLOPs / TF32 peak (TF32's dense peak is half of the measured bf16 figure)
        hbm_floor_us = abytes / (hbm_peak * 1e3)
        tensor_floor_us = flops / (0.5 * peak_burst * 1e6)
        tensor_view = {"achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved / peak_burst,
                       "peak_source": which, "floor_us": tensor_floor_us,
                       "note": "operands are TF32 (nominal dense peak = half of bf16); peak shown is the measured bf16 figure"}
        hbm_view = {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                    "algorithmic_bytes_per_launch_set": abytes, "floor_us": hbm_floor_us,
                    "peak_source": "measured copy bandwidth (MEASURED_PEAKS.json)" if peaks else "fallback 6.5 TB/s"}
        bound = "hbm" if hbm_floor_us >= tensor_floor_us else "tensor"
        lead = hbm_view if bound == "hbm" else tensor_view
        roof = {"bound": bound, "achieved": lead["achieved"], "peak": lead["peak"], "unit": lead["unit"], "frac": lead["frac"],
                "traffic": traffic_mean, "traffic_unit": "bytes per launch (mean of the ncu-captured GEMM launches)",
                "traffic_detail": traffic, "peak_source": lead["peak_source"],
                "kernel": "gemm_tcp_kernel / gemm_tc_kernel (tcgen05.mma kind::tf32, persistent + one-tile variants)",
                "hbm_view": hbm_view, "tensor_view": tensor_view,
                "launches_timed": n_launch, "avg_launch_us": t_us / max(n_launch, 1), "measured_us_per_launch_set": t_us,
                "executed_gflop_per_launch_set": flops / 1e9,
                "whole_step_achieved": whole, "whole_step_frac": whole / peak,
                "note": "the %d tensor-core GEMM launches of one feature-DDPM step, CUDA-event timed: fp32 activations in HBM put "
                        "them left of the TF32 ridge (HBM floor %.0f us > tensor floor %.0f us), so the HBM roofline is the "
                        "binding one; inside the SM the same launches sit at 50-80 %% of the shared-memory (128 B/clk) and "
                        "L2->SM (6300 B/clk) limits of a TF32 SS-mode MMA (DESIGN.md 4.2). whole_step_* = reference-"
                        "formulation FLOPs (1151.7 GFLOP/shape) / device time." % (n_launch, hbm_floor_us, tensor_floor_us)}
        extras = {}
        parity = None
        if args.ddpm_steps is None:
            # same-process check that the launched instantiations compute the right thing: eps of both denoisers at this
            # batch, default dispatch (what was timed) vs the fp32 FFMA backend; record-level parity vs the oracle at the
            # same batch sizes is tests/test_gpu_baseline_sizes.py
            parity = {"position_eps_rel_err": eps_parity(pipe.pos, 11), "latent_eps_rel_err": eps_parity(pipe.lat, 12),
                      "tolerance": 5e-3, "what": "max |eps_auto - eps_fp32| / max |eps_fp32| at batch %d, t = T/2" % Bl}
            parity["ok"] = bool(parity["position_eps_rel_err"] < 5e-3 and parity["latent_eps_rel_err"] < 5e-3)
            cpu = run_json([sys.executable, "-m", "oracle.reference_arms", "cpu", "--budget", "20", "--batch", "16",
                            "--category", args.category], 240)
            cpu = {k: cpu.get(k) for k in ("value", "unit", "cores", "kind", "sample", "detail", "error") if k in cpu}
        if args.ddpm_steps is None and not args.no_extras and world == 1:
            extras["reference_gpu_eager"] = run_json([sys.executable, "-m", "oracle.reference_arms", "gpu", "--batch",
                                                      str(args.batch), "--steps", "10", "--category", args.category], 300)
            try:
                extras["configs"] = extra_configs(cfg, args)
            except Exception as e:  # noqa: BLE001
                extras["configs"] = {"error": repr(e)[:300]}
        valid = args.ddpm_steps is None and finite and tc_err == 0 and bool(parity and parity["ok"])
        print(json.dumps({
            "metric": "shapes/sec (1000-step DDPM + decode to 2048 pts) @ bs256", "value": B / (ms_value / 1e3),
            "unit": "shapes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if args.backend == "auto" else "f32",
            "data": "synthetic",
            "config": {"workload": workload, "global_batch": B, "per_gpu_batch": Bl, "parallelism": "dp%d" % world,
                       "l2": "inputs larger than L2 (noise tensors 49 MB + 836 MB per GPU)", "valid": valid,
                       "ddpm_steps": args.ddpm_steps or 1000, "backend": args.backend, "rng_scope": scope,
                       "cross_batch_overlap": overlap,
                       "overlap_note": "step i+1's position DDPM runs on a side stream under step i's feature DDPM + decode; "
                                       "every chain of all K steps executes inside the timed region (stages_ms = one "
                                       "un-overlapped warm-up step)"},
            "e2e": {"value": B / (ms_e2e / 1e3), "unit": "shapes/s", "h2d_bytes_per_step": pipe.h2d_bytes(),
                    "d2h_bytes_per_step": pipe.d2h_bytes(), "ms_per_step": ms_e2e, "host_rng_ms_per_draw": ms_rng,
                    "includes": "host RNG draws every step (reference call order; the next step's draws overlap the GPU), "
                                "pinned H2D, 3 stages, D2H"},
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roof, "cpu_baseline": cpu,
            "finite": finite, "tc_error": tc_err, "parity": parity, "stages_ms": stages,
            "noise_path": getattr(pipe, "noise_path", None), "strong": strong,
            **({"config5_sharded": ae_sweep} if ae_sweep else {}), **extras}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
