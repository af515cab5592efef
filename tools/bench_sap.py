"""SAP mesh-reconstruction stage on one B200: clouds/s of cloud -> indicator grid through the C ABI,
per-stage device times (CUDA events) and the DPSR kernels against their HBM roofline.

  unit            one cloud of 2048 oriented points -> mirrored 4096 -> refined 20480 points -> 128^3 indicator grid
  batch           32 (the shipped refine JSON's eval_batch_size)
  algorithmic     DPSR: 213 MB per cloud at 128^3 (csrc/sap.cu header); refine network: 2 x MACs of its 1x1 convs

usage: python tools/bench_sap.py [--batch 32] [--reps 5]      -> one JSON line
"""
import argparse
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def dpsr_bytes_per_cloud(R):
    vol, hvol = R ** 3 * 4, R * R * (R // 2 + 1) * 8
    # memset + splat target, Z (read real, write half spectrum) x 3 ch, Y in place x 3 ch, X + solve (3 ch in, 1 out),
    # inverse Y in place, inverse Z, shift / scale in place
    return 3 * vol + 3 * (vol + hvol) + 3 * 2 * hvol + (3 * hvol + hvol) + 2 * hvol + (hvol + vol) + 2 * vol


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--backend", default="auto")
    args = ap.parse_args()
    import torch
    from slide_b200 import lib, sap
    torch.cuda.set_device(0)
    B = args.batch
    rec = sap.load_default(B, gemm_backend=args.backend)
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(B, 2048, 3, generator=g) - 0.5
    nrm = torch.nn.functional.normalize(torch.randn(B, 2048, 3, generator=g), dim=2)
    cloud_host = torch.cat([pts, nrm], dim=2).pin_memory()
    labels = torch.zeros(B, dtype=torch.int32)
    perm = torch.randperm(4096, generator=g).int()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]

    def stages():
        """The reconstructor's own call sequence with an event between the stages."""
        cloud = cloud_host.cuda(non_blocking=True)
        rec.prog.upload(rec.h["labels"], labels.cuda())
        ev[0].record()
        rec.prog.run_segment("setup")
        X = rec.prog.view(rec.h["x"])
        sap.mirror_concat(cloud, perm, axis=2, out=X.view(B, rec.n_in, X.shape[1]))
        ev[1].record()
        rec.prog.run_segment("refine")
        ev[2].record()
        fine = rec.prog.view(rec.h["fine"]).view(B, rec.n_fine, -1)
        p = sap.unit_cube(fine, True, rec.scale)
        ev[3].record()
        phi = rec.dpsr(p, fine[:, :, 3:6])
        ev[4].record()
        return phi

    for _ in range(2):
        stages()
    torch.cuda.synchronize()
    acc = [0.0] * 4
    lib.reset_launch_count()
    for _ in range(args.reps):
        phi = stages()
        torch.cuda.synchronize()
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
    launches = lib.launch_count() // args.reps
    ms = [a / args.reps for a in acc]
    # end to end through the public call, host buffers in, grids back on the host
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_phi = torch.empty(B, 128, 128, 128, pin_memory=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(args.reps):
        out = rec.reconstruct(cloud_host, labels, perm)
        host_phi.copy_(out["phi"], non_blocking=True)
    t1.record()
    torch.cuda.synchronize()
    e2e_ms = t0.elapsed_time(t1) / args.reps
    # iso-surface meshes of the batch's grids (sap.mc_from_psr: one 8-byte read-back per grid sizes its outputs)
    sap.mc_from_psr(out["phi"][:2])
    torch.cuda.synchronize()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    verts, faces, _ = sap.mc_from_psr(out["phi"])
    m1.record()
    torch.cuda.synchronize()
    mesh_ms = m0.elapsed_time(m1)
    R = rec.dpsr.res[0]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm = float(peaks.get("hbm_gbs", 6500.0))
    dbytes = dpsr_bytes_per_cloud(R) * B
    total = sum(ms)
    print(json.dumps({
        "workload": "SAP stage: %d clouds x 2048 pts -> mirror 4096 -> refine x5 -> DPSR %d^3" % (B, R),
        "clouds_per_s": B / (total / 1e3), "ms": {"setup+mirror": ms[0], "refine_network": ms[1], "unit_cube": ms[2], "dpsr": ms[3]},
        "total_ms": total, "e2e_clouds_per_s": B / (e2e_ms / 1e3), "e2e_ms": e2e_ms,
        "e2e_h2d_bytes": int(cloud_host.numel() * 4), "e2e_d2h_bytes": int(host_phi.numel() * 4),
        "mesh_ms": mesh_ms, "mesh_vertices_per_grid": int(sum(v.shape[0] for v in verts) / B),
        "mesh_faces_per_grid": int(sum(f.shape[0] for f in faces) / B),
        "launches": launches, "finite": bool(torch.isfinite(phi).all().item()), "tc_error": int(lib.load().slide_tc_error()),
        "dpsr_roofline": {"bound": "hbm", "algorithmic_bytes": dbytes, "achieved": dbytes / (ms[3] / 1e3) / 1e9, "peak": hbm,
                          "unit": "GB/s", "frac": dbytes / (ms[3] / 1e3) / 1e9 / hbm}}))


if __name__ == "__main__":
    main()
