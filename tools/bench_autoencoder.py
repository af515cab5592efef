"""BASELINE config 5: autoencoder encode + decode throughput sweep, N in {2048, 4096, 8192} points, batch 512, with an
HBM-roofline figure.  Usage: python tools/bench_autoencoder.py [--batch 512] [--chunk 32] [--points 2048 4096 8192]

Algorithmic bytes per shape (fp32, single pass, nothing materialised): the encoder reads the cloud
(24 N B), FPS/kNN scans it (12 N B each level-0 pass) and the decoder writes 2048 x 6 floats; the figure reported is
bytes / time against the measured HBM peak -- a latency/compute-bound path is expected to sit far below 1."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from slide_b200 import pipeline, weights, lib  # noqa: E402


def sweep(batch, chunk=32, points=(2048, 4096, 8192), reps=2, device=None):
    """-> one dict per cloud size: encode / decode device milliseconds for `batch` clouds on `device`."""
    cfg = weights.load_json("pipeline_airplane.json")
    aec = cfg["autoencoder"]
    sd = pipeline.default_state_dicts()["autoencoder"]
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    while batch % chunk:
        chunk -= 1
    dec = pipeline.Decoder(aec["decoders"], sd, chunk, dev)
    B = batch
    labels = torch.zeros(B, dtype=torch.long, device=dev)
    starts = torch.zeros(dec.n_levels, B, dtype=torch.long, device=dev)
    rows = []
    for N in points:
        enc = pipeline.Encoder(aec["encoder"], aec["decoders"][0], sd, chunk, N, dev)
        g = torch.Generator(device="cpu").manual_seed(N)
        pts = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
        nrm = torch.nn.functional.normalize(torch.randn(B, N, 3, generator=g), dim=2).to(dev)
        cloud = torch.cat([pts, nrm], dim=2).contiguous()
        kp = pipeline.sample_keypoints(pts, 16).contiguous()
        latent = torch.empty(B, 16, enc.out_dim, device=dev)
        out = torch.empty(B, dec.out_points, dec.out_dim, device=dev)
        times = {}
        for name, fn in (("encode", lambda: enc.run(cloud, kp, labels, latent)),
                         ("decode", lambda: dec.run(kp, latent, labels, starts, out))):
            fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            times[name] = e0.elapsed_time(e1) / reps
        total_ms = times["encode"] + times["decode"]
        bytes_per_shape = 24 * N + 2 * 12 * N + 4 * 1024 + 16 * 48 * 4 * 2 + 2048 * 6 * 4
        rows.append({"workload": "autoencoder encode+decode", "points": N, "batch": B, "chunk": chunk,
                     "encode_ms": times["encode"], "decode_ms": times["decode"],
                     "shapes_per_s": B / (total_ms / 1e3), "finite": bool(torch.isfinite(out).all().item()),
                     "gflop_per_shape": 6.776 + 16.607,
                     "tflops": B * (6.776 + 16.607) / (total_ms / 1e3) / 1e3,
                     "hbm_roofline": {"algorithmic_bytes_per_shape": bytes_per_shape,
                                      "achieved_gbs": B * bytes_per_shape / (total_ms / 1e3) / 1e9, "peak_gbs": hbm,
                                      "frac": B * bytes_per_shape / (total_ms / 1e3) / 1e9 / hbm},
                     "tc_error": lib.load().slide_tc_error()})
        del enc
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--points", type=int, nargs="+", default=[2048, 4096, 8192])
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    for row in sweep(args.batch, args.chunk, args.points, args.reps, torch.device("cuda", 0)):
        print(json.dumps(row))


if __name__ == "__main__":
    main()
