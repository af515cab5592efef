"""Golden vectors for RANDOM architectures of the reference's denoiser (PointNet2CloudCondition): level counts, point counts
with and without down-sampling, neighbour counts, channel widths that are not multiples of the GroupNorm group count,
MLP depths, input feature widths, timestep / class embedding widths -- what a user training a custom config would have.
Produced by the REAL reference module (build container only: needs /root/reference; the C oracle stands in for `_ext` /
pytorch3d):

    python tests/golden/make_golden_random_archs.py   ->  tests/golden/golden_random_archs.npz
                                                          (hparams + state-dict schemas travel inside, as one JSON string)
Also asserts that oracle/ref_model.py is bit-identical to the real module on every drawn architecture."""
import copy
import json
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops, ref_model  # noqa: E402

ops.install_reference_stubs()
from models.pointnet2_with_pcld_condition import PointNet2CloudCondition  # noqa: E402
from slide_b200 import weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N_ARCHS = 10
N_ARCHS_16 = 6      # 16-point architectures without down-sampling: what the sample-resident compiler accepts
WIDTHS = [16, 24, 32, 40, 64, 72]
WIDTHS_16 = [16, 24, 32, 40, 64, 72, 96, 128]


def draw16(rng, base):
    pc = copy.deepcopy(base)
    a = pc["architecture"]
    n_levels = rng.choice([1, 2, 2, 3])
    a["npoint"] = [16] * n_levels
    a["nsample"] = [rng.choice([8, 16, 16]) for _ in range(n_levels)]
    a["radius"] = [0] * n_levels
    a["feature_dim"] = [rng.choice(WIDTHS_16) for _ in range(n_levels + 1)]
    a["decoder_feature_dim"] = [rng.choice(WIDTHS_16) for _ in range(n_levels)] + [a["feature_dim"][-1]]
    a["mlp_depth"] = rng.choice([2, 3])
    a["decoder_mlp_depth"] = rng.choice([2, 3])
    a["K"] = rng.choice([8, 8, 16])
    pc["t_dim"] = rng.choice([64, 128])
    pc["model_name"] = "random_arch_16"
    return pc, 16


def draw(rng, base):
    pc = copy.deepcopy(base)
    n_levels = rng.choice([1, 2, 2, 3])
    n0 = rng.choice([8, 12, 16, 20, 32])
    npoint = [n0]
    for _ in range(n_levels - 1):
        npoint.append(max(4, npoint[-1] // rng.choice([1, 2])))
    a = pc["architecture"]
    a["npoint"] = npoint
    a["nsample"] = [min(rng.choice([4, 6, 8, 16]), src) for src in [n0] + npoint[:-1]]
    a["radius"] = [0] * n_levels
    a["feature_dim"] = [rng.choice(WIDTHS) for _ in range(n_levels + 1)]
    # models/pointnet2_with_pcld_condition.py:230 asserts decoder_feature_dim[-1] == feature_dim[-1]
    a["decoder_feature_dim"] = [rng.choice(WIDTHS) for _ in range(n_levels)] + [a["feature_dim"][-1]]
    a["mlp_depth"] = rng.choice([2, 3])
    a["decoder_mlp_depth"] = rng.choice([2, 3])
    a["K"] = min(rng.choice([3, 4, 8]), min(npoint))
    pc["in_fea_dim"] = rng.choice([0, 0, 5, 13])
    pc["out_dim"] = 3 + pc["in_fea_dim"]
    pc["t_dim"] = rng.choice([64, 128])
    pc["class_condition_dim"] = rng.choice([32, 128])
    pc["model_name"] = "random_arch"
    return pc, n0


def main():
    base = weights.load_json("pipeline_airplane.json")["position_ddpm"]["pointnet_config"]
    gold = {}
    B = 2
    label = torch.tensor([0, 4])
    for prefix, count, drawer, seed in (("a", N_ARCHS, draw, 2), ("r", N_ARCHS_16, draw16, 5)):
        one_family(gold, prefix, count, drawer, random.Random(seed), base, B, label)
    gold["label"] = label.numpy()
    path = os.path.join(OUT, "golden_random_archs.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")


def one_family(gold, prefix, count, drawer, rng, base, B, label):
    meta = []
    for i in range(count):
        pc, n0 = drawer(rng, base)
        net = PointNet2CloudCondition(copy.deepcopy(pc)).eval()
        schema = [[k, list(v.shape)] for k, v in net.state_dict().items()]
        sd = weights.random_state_dict(schema, 200 + i)
        net.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(300 + i)
        x = torch.randn(B, n0, 3 + pc["in_fea_dim"], generator=g)
        gold["%s%d_x" % (prefix, i)] = x.numpy()
        for t in (999, 0):
            with torch.no_grad():
                y = net(x, ts=torch.ones(B) * t, label=label)
                y2 = ref_model.cloud_condition_net(x, ref_model.Params(sd), pc, ts=torch.ones(B) * t, label=label)
            assert torch.equal(y, y2), "oracle/ref_model.py deviates from the reference on architecture %d" % i
            gold["%s%d_eps_t%d" % (prefix, i, t)] = y.numpy()
        meta.append({"pointnet_config": pc, "schema": schema, "n_points": n0, "seed": 200 + i})
        a = pc["architecture"]
        print(prefix, i, "npoint", a["npoint"], "nsample", a["nsample"], "K", a["K"], "feature_dim", a["feature_dim"], "decoder",
              a["decoder_feature_dim"], "depth", a["mlp_depth"], a["decoder_mlp_depth"], "in_fea", pc["in_fea_dim"],
              "t_dim", pc["t_dim"], "class_dim", pc["class_condition_dim"])
    gold["meta_json" if prefix == "a" else "meta16_json"] = np.array(json.dumps(meta, sort_keys=True, separators=(",", ":")))


if __name__ == "__main__":
    main()
