"""The reference's OWN sampling code timed on this box -- baselines for bench.py (test / measurement infrastructure,
never imported by slide_b200/).

Two arms, both driving the UNMODIFIED python of SLIDE-3D/SLIDE mirrored under baseline/_ref (baseline/fetch_reference.py):
  util.sampling (pointnet2/util.py:197-259), LatentDiffusion.denoise_and_reconstruct (diffusion_utils/diffusion.py:346-404),
  PointNet2CloudCondition / PointAutoencoder (models/*.py) and the reference's `pointnet2_ops` python package.

  cpu   "the reference's CPU path": the native ops the reference only has for CUDA (`pointnet2_ops._ext`) and pytorch3d
        (not installed anywhere offline) are supplied by the C oracle (oracle/ops.py); everything above them is the
        reference's code on CPU tensors with all host threads.  `Tensor.cuda()` / `.to(cuda)` are no-ops here.
        kind = "reference".  Without the mirror the bit-identical port oracle/ref_model.py is timed: kind = "port".
  gpu   "the reference's eager GPU path on the same B200": `pointnet2_ops._ext` = the reference's own CUDA extension
        compiled for sm_100a (oracle/_ref, oracle/build_ref.py); pytorch3d's ops = slide_b200's drop-in (pytorch3d is
        not installable offline -- stated in the result); torch eager with its defaults (cuDNN TF32).

Both time a BOUNDED sample (K steps of each DDPM at a reduced batch + one decode) through the reference's own loops
(`use_a_precomputed_XT` / `n_steps` are the reference's parameters for partial chains) and scale to 1000 + 1000 steps.
usage: python -m oracle.reference_arms cpu|gpu [--budget SECONDS] [--batch B] [--category airplane]   -> one JSON line
"""
import argparse
import copy
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
MIRROR = os.path.join(ROOT, "baseline", "_ref")


def mirror_available():
    return os.path.isdir(os.path.join(MIRROR, "pointnet2", "models"))


def _build_reference_models(cfg, device):
    """The reference's modules from the shipped hyper-parameters, random-init weights with the checkpoint schema."""
    import torch
    from models.pointnet2_with_pcld_condition import PointNet2CloudCondition
    from models.autoencoder import PointAutoencoder
    from slide_b200 import pipeline
    sds = pipeline.default_state_dicts()
    pos = PointNet2CloudCondition(copy.deepcopy(cfg["position_ddpm"]["pointnet_config"])).eval()
    pos.load_state_dict(sds["position"], strict=True)
    lat = PointNet2CloudCondition(copy.deepcopy(cfg["latent_ddpm"]["pointnet_config"])).eval()
    lat.load_state_dict(sds["latent"], strict=True)
    aec = cfg["autoencoder"]
    ae = PointAutoencoder(copy.deepcopy(aec["encoder"]), copy.deepcopy(aec["decoders"]),
                          apply_kl_regularization=aec["apply_kl_regularization"], kl_weight=aec["kl_weight"]).eval()
    ae.load_state_dict(sds["autoencoder"], strict=True)
    return pos.to(device), lat.to(device), ae.to(device)


class ReferenceRunner(object):
    """The reference's models + loops built once; time(K) runs K steps of util.sampling, K steps + decode of
    LatentDiffusion.denoise_and_reconstruct and returns (shapes/s scaled to 1000 + 1000 steps + decode, detail)."""

    def __init__(self, cfg, B, device, sync):
        import torch
        import util as ref_util
        from diffusion_utils.diffusion import LatentDiffusion
        self.cfg, self.B, self.device, self.sync = cfg, B, device, sync
        self.pos, self.lat, self.ae = _build_reference_models(cfg, device)
        d = cfg["position_ddpm"]["diffusion_config"]
        self.dh = ref_util.calc_diffusion_hyperparams(d["T"], d["beta_0"], d["beta_T"])
        for key in ("Alpha", "Alpha_bar", "Sigma"):
            self.dh[key] = self.dh[key].to(device)
        self.label = torch.full((B,), int(cfg["label"]), dtype=torch.long, device=device)
        self.ld = LatentDiffusion(copy.deepcopy(cfg["latent_ddpm"]["standard_diffusion_config"]), self.ae, device=device)
        self.kp = (torch.rand(B, 16, 3) - 0.5).to(device)
        self.ref_util = ref_util
        self.warm = False

    def run_pos(self, k):
        import torch
        return self.ref_util.sampling(self.pos, (self.B, 16, 3), self.dh, label=self.label, verbose=False,
                                      print_every_n_steps=10 ** 9, use_a_precomputed_XT=True, step=k,
                                      XT=torch.zeros(self.B, 16, 3, device=self.device))

    def run_lat(self, k):
        F = self.cfg["latent_ddpm"]["pointnet_config"]["in_fea_dim"]
        return self.ld.denoise_and_reconstruct(self.B, self.lat, 3, (16, 3 + F), label=self.label, n_steps=k,
                                               keypoint=self.kp)

    def time(self, K):
        import torch
        out, sync = {}, self.sync
        with torch.no_grad():
            if not self.warm:
                self.run_pos(1)
                self.run_lat(1)  # (includes one decode)
                self.warm = True
            sync()
            t0 = time.time()
            self.run_pos(K)
            sync()
            out["pos_s_per_step"] = (time.time() - t0) / K
            # denoise_and_reconstruct always ends in a decode: a timing wrapper around the bound method (instrumentation
            # only, the reference code is untouched) splits the call into its K steps and its decode
            dec = {"s": 0.0}
            ref_decode = self.ld.decode

            def timed_decode(*a, **k):
                sync()
                t = time.time()
                r = ref_decode(*a, **k)
                sync()
                dec["s"] += time.time() - t
                return r

            self.ld.decode = timed_decode
            try:
                t0 = time.time()
                self.run_lat(K)
                sync()
                total = time.time() - t0
            finally:
                self.ld.decode = ref_decode
            out["decode_s"] = dec["s"]
            out["lat_s_per_step"] = max(total - dec["s"], 1e-9) / K
        per_shape = (1000 * out["pos_s_per_step"] + 1000 * out["lat_s_per_step"] + out["decode_s"]) / self.B
        out.update(batch=self.B, steps_timed=K)
        return 1.0 / per_shape, out


_CPU_RUNNERS = {}


def cpu_arm(cfg, budget=20.0, B=16):
    """-> (shapes/s, cores, kind, sample, detail)"""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if not mirror_available():
        from oracle import ref_model
        from slide_b200 import pipeline
        sds = pipeline.default_state_dicts()
        label = torch.zeros(B, dtype=torch.long)
        per_shape, detail = 0.0, {}
        with torch.no_grad():
            for key, name, C in (("position_ddpm", "position", 3), ("latent_ddpm", "latent", 51)):
                pc = cfg[key]["pointnet_config"]
                P = ref_model.Params(sds[name])
                x = torch.randn(B, 16, C)
                ref_model.cloud_condition_net(x, P, pc, ts=torch.ones(B) * 500, label=label)
                n, t0 = 0, time.time()
                while n < 3 or (time.time() - t0 < budget / 3 and n < 50):
                    ref_model.cloud_condition_net(x, P, pc, ts=torch.ones(B) * 500, label=label)
                    n += 1
                detail[name + "_s_per_step"] = (time.time() - t0) / n
                per_shape += 1000 * detail[name + "_s_per_step"] / B
            P = ref_model.Params(sds["autoencoder"])
            t0 = time.time()
            ref_model.decode(torch.rand(2, 16, 3) - 0.5, torch.randn(2, 16, 48), P, cfg["autoencoder"]["decoders"], label[:2])
            detail["decode_s_b2"] = time.time() - t0
            per_shape += detail["decode_s_b2"] / 2
        return (1.0 / per_shape, torch.get_num_threads(), "port",
                "oracle/ref_model.py (bit-identical port): denoiser forwards at batch %d scaled to 1000+1000 steps + decode of 2 shapes" % B,
                detail)
    if B not in _CPU_RUNNERS:
        from oracle import ops
        ops.install_reference_stubs(reference_root=MIRROR)
        # the reference moves tensors with .cuda(): no-ops on the CPU arm (device placement only, arithmetic untouched)
        torch.Tensor.cuda = lambda self, *a, **k: self
        _CPU_RUNNERS[B] = ReferenceRunner(cfg, B, torch.device("cpu"), lambda: None)
    runner = _CPU_RUNNERS[B]
    K = 2
    value, detail = runner.time(K)
    # spend the rest of the budget on more steps if the first pass was quick
    spent = K * (detail["pos_s_per_step"] + detail["lat_s_per_step"]) + detail["decode_s"]
    if spent < budget / 3:
        K2 = int(min(20, max(K, (budget - 2 * spent) / max(detail["pos_s_per_step"] + detail["lat_s_per_step"], 1e-6))))
        if K2 > K:
            value, detail = runner.time(K2)
    sample = ("unmodified util.sampling + LatentDiffusion.denoise_and_reconstruct: %d steps of each DDPM at batch %d + one "
              "decode, scaled to 1000+1000 steps; native ops = C oracle" % (detail["steps_timed"], B))
    return value, torch.get_num_threads(), "reference", sample, detail


def gpu_arm(cfg, B=256, K=10):
    """The reference's eager GPU path on cuda:0 -> (shapes/s, detail)."""
    import torch
    import slide_b200
    from oracle import build_ref
    ext = build_ref.load_module()
    if ext is None or not mirror_available():
        raise RuntimeError("needs oracle/_ref (the reference's CUDA extension) and baseline/_ref (its python)")
    sys.modules["pointnet2_ops._ext"] = ext
    # the reference's python package first, then the drop-in directory (only its `pytorch3d` is picked up from there)
    sys.path.insert(0, slide_b200.DROPIN_DIR)
    sys.path.insert(0, os.path.join(MIRROR, "pointnet2"))
    sys.path.insert(0, os.path.join(MIRROR, "pointnet2_ops_lib"))
    import pointnet2_ops
    assert pointnet2_ops.__file__.startswith(MIRROR), pointnet2_ops.__file__
    pointnet2_ops._ext = ext
    dev = torch.device("cuda", 0)
    value, detail = ReferenceRunner(cfg, B, dev, torch.cuda.synchronize).time(K)
    detail["pointnet2_ops"] = "reference python package + its own CUDA extension built for sm_100a (oracle/_ref)"
    detail["pytorch3d"] = "slide_b200 drop-in (pytorch3d 0.7.0 is not installable offline)"
    return value, detail


# ---------------------------------------------------------------------------------------------------------------------
# SAP mesh-reconstruction stage: the reference's refine network + network_output_to_dpsr_grid + DPSR
# ---------------------------------------------------------------------------------------------------------------------
def _stub_absent_packages():
    """dpsr_utils/utils.py imports mesh / rendering packages at module level that exist nowhere offline and that DPSR
    does not use; empty stand-ins let the UNMODIFIED file import."""
    import types
    for name, attrs in (("trimesh", {}), ("plyfile", {"PlyData": None}), ("skimage", {}), ("skimage.measure", {}),
                        ("pytorch3d.renderer", {"PerspectiveCameras": None, "rasterize_meshes": None}),
                        ("igl", {"adjacency_matrix": None, "connected_components": None})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    import pytorch3d.structures as st
    if not hasattr(st, "Meshes"):
        st.Meshes = None


def _reference_grid_function():
    """network_output_to_dpsr_grid + shapenet_psr_normalize taken out of the unmodified dpsr_evaluation.py (the module
    itself imports visualisation packages that are absent offline)."""
    import ast
    import numpy as np
    import torch
    from models.point_upsample_module import point_upsample
    src = open(os.path.join(MIRROR, "pointnet2", "dpsr_evaluation.py")).read()
    ns = {"torch": torch, "np": np, "point_upsample": point_upsample}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("shapenet_psr_normalize", "network_output_to_dpsr_grid"):
            exec(compile(ast.Module([node], []), "dpsr_evaluation.py", "exec"), ns)
    return ns["network_output_to_dpsr_grid"]


def sap_arm(device_kind, B, reps=3):
    """visualize_per_rank's per-batch work between loading the cloud and marching cubes (dpsr_evaluation.py:214-260):
    mirror_and_concat -> PointNet2CloudCondition(refine JSON) -> network_output_to_dpsr_grid(DPSR 128^3), on `cpu`
    (C-oracle native ops, all host threads) or `gpu` (eager torch + the reference's CUDA extension + cuFFT).
    -> dict(clouds_per_s, ms per stage)."""
    import torch
    from slide_b200 import weights
    if not mirror_available():
        raise RuntimeError("needs baseline/_ref (the reference's python)")
    gpu = device_kind == "gpu"
    if gpu:
        import slide_b200
        from oracle import build_ref
        ext = build_ref.load_module()
        if ext is None:
            raise RuntimeError("needs oracle/_ref (the reference's CUDA extension)")
        sys.modules["pointnet2_ops._ext"] = ext
        sys.path.insert(0, slide_b200.DROPIN_DIR)
        sys.path.insert(0, os.path.join(MIRROR, "pointnet2"))
        sys.path.insert(0, os.path.join(MIRROR, "pointnet2_ops_lib"))
        import pointnet2_ops
        pointnet2_ops._ext = ext
        dev, sync = torch.device("cuda", 0), torch.cuda.synchronize
    else:
        from oracle import ops
        ops.install_reference_stubs(reference_root=MIRROR)
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.set_num_threads(os.cpu_count() or 1)
        dev, sync = torch.device("cpu"), (lambda: None)
    _stub_absent_packages()
    from models.pointnet2_with_pcld_condition import PointNet2CloudCondition
    from data_utils.mirror_partial import mirror_and_concat
    from dpsr_utils.dpsr import DPSR
    to_grid = _reference_grid_function()
    cfg = weights.load_json("sap_refine.json")
    pc, dc = cfg["pointnet_config"], cfg["dpsr_config"]
    net = PointNet2CloudCondition(copy.deepcopy(pc)).eval()
    net.load_state_dict(weights.random_state_dict(weights.load_json("schema_sap_refine.json"), 21), strict=True)
    net = net.to(dev)
    dpsr = DPSR(res=(dc["grid_res"],) * 3, sig=dc["psr_sigma"]).to(dev)
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(B, 2048, 3, generator=g) - 0.5
    nrm = torch.nn.functional.normalize(torch.randn(B, 2048, 3, generator=g), dim=2)
    cloud = torch.cat([pts, nrm], dim=2).to(dev)
    label = torch.zeros(B, dtype=torch.long, device=dev)
    t = {"mirror": 0.0, "network": 0.0, "grid": 0.0}
    with torch.no_grad():
        for it in range(reps + 1):
            sync()
            t0 = time.time()
            X = mirror_and_concat(cloud, axis=2, num_points=[], attach_label=True, permute=True)[0]
            sync()
            t1 = time.time()
            disp = net(X, None, ts=None, label=label)
            sync()
            t2 = time.time()
            phi, rp, rn = to_grid(X, disp, dpsr, cfg.get("scale", 1), pc, last_dim_as_indicator=True,
                                  only_original_points_split=False, explicit_normalize=True)
            sync()
            t3 = time.time()
            if it:  # first pass = warm-up
                t["mirror"] += t1 - t0
                t["network"] += t2 - t1
                t["grid"] += t3 - t2
    ms = {k: v / reps * 1e3 for k, v in t.items()}
    total = sum(ms.values())
    return {"clouds_per_s": B / (total / 1e3), "batch": B, "ms": ms, "total_ms": total,
            "finite": bool(torch.isfinite(phi).all().item()),
            "path": ("eager torch + the reference's CUDA extension (oracle/_ref) + cuFFT; pytorch3d = slide_b200 drop-in"
                     if gpu else "unmodified python on %d host threads, native ops = C oracle" % torch.get_num_threads())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("arm", choices=["cpu", "gpu", "sap-cpu", "sap-gpu"])
    ap.add_argument("--budget", type=float, default=20.0)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--category", default="airplane")
    args = ap.parse_args()
    from slide_b200 import weights
    cfg = weights.load_json("pipeline_%s.json" % args.category)
    if args.arm.startswith("sap-"):
        print(json.dumps(sap_arm(args.arm[4:], args.batch or (32 if args.arm == "sap-gpu" else 2), reps=max(1, min(args.steps, 3)))))
        return
    if args.arm == "cpu":
        v, cores, kind, sample, detail = cpu_arm(cfg, args.budget, args.batch or 16)
        print(json.dumps({"value": v, "unit": "shapes/s", "cores": cores, "kind": kind, "sample": sample, "detail": detail}))
    else:
        v, detail = gpu_arm(cfg, args.batch or 256, args.steps)
        print(json.dumps({"value": v, "unit": "shapes/s", "kind": "reference eager on the same GPU", "detail": detail}))


if __name__ == "__main__":
    main()
