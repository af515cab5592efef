"""Worker of tests/test_gpu_reference_unmodified.py (a separate process: `pointnet2_ops` must resolve to the drop-in).

Builds the REAL reference modules (unmodified files mirrored under baseline/_ref by baseline/fetch_reference.py) on top of
slide_b200's drop-in `pointnet2_ops` / `pytorch3d`, loads the seeded state dicts the golden vectors were made with, runs
them on cuda:0 and compares with tests/golden/golden.npz (written by the same modules on the CPU oracle ops).
usage: python tests/ref_unmodified_worker.py [--import-only]
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import slide_b200  # noqa: E402

slide_b200.install_dropin()
sys.path.insert(1, os.path.join(ROOT, "baseline", "_ref", "pointnet2"))
from models.pointnet2_with_pcld_condition import PointNet2CloudCondition  # noqa: E402  (the reference's file)
from models.autoencoder import PointAutoencoder  # noqa: E402
import pointnet2_ops  # noqa: E402
from slide_b200 import weights  # noqa: E402
from tests import common  # noqa: E402

assert pointnet2_ops.__file__.startswith(slide_b200.DROPIN_DIR), pointnet2_ops.__file__
assert sys.modules["models.autoencoder"].__file__.startswith(os.path.join(ROOT, "baseline", "_ref"))


def main():
    cfg = weights.load_json("pipeline_airplane.json")
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "golden.npz")))
    nets = {}
    for which, key in (("pos", "position_ddpm"), ("lat", "latent_ddpm")):
        net = PointNet2CloudCondition(copy.deepcopy(cfg[key]["pointnet_config"])).eval()
        net.load_state_dict(common.state_dict(which), strict=True)
        nets[which] = net
    aec = cfg["autoencoder"]
    ae = PointAutoencoder(copy.deepcopy(aec["encoder"]), copy.deepcopy(aec["decoders"]),
                          apply_kl_regularization=aec["apply_kl_regularization"], kl_weight=aec["kl_weight"]).eval()
    ae.load_state_dict(common.state_dict("ae"), strict=True)
    print("built: reference modules over the drop-in, state dicts loaded strict")
    if "--import-only" in sys.argv:
        return
    dev = torch.device("cuda", 0)
    label = torch.from_numpy(gold["label"]).long().to(dev)
    worst = 0.0
    with torch.no_grad():
        for which, net in nets.items():
            net.to(dev)
            x = torch.from_numpy(gold[which + "_x"]).to(dev)
            for t in (999, 500, 0):
                eps = net(x, ts=torch.ones(x.shape[0], device=dev) * t, label=label).cpu().numpy()
                want = gold["%s_eps_t%d" % (which, t)]
                err = np.abs(eps - want).max() / max(1.0, np.abs(want).max())
                print("%s t=%d rel err %.3g" % (which, t, err))
                worst = max(worst, err)
        ae.to(dev)
        # encode, posterior mode (deterministic: FPS from index 0)
        cloud = torch.from_numpy(gold["enc_cloud"]).to(dev)
        kp = torch.from_numpy(gold["enc_kp"]).to(dev)
        feat = ae.encode(cloud, kp, label=label, sample_posterior=False)
        if isinstance(feat, (tuple, list)):
            feat = feat[0]
        want = gold["enc_mode"]
        err = np.abs(feat.cpu().numpy() - want).max() / max(1.0, np.abs(want).max())
        print("encode rel err %.3g" % err)
        worst = max(worst, err)
    # torch's cuDNN convs run in TF32 by default (the reference never overrides allow_tf32)
    assert worst < 2e-2, worst
    print("OK worst rel err %.3g" % worst)


if __name__ == "__main__":
    main()
