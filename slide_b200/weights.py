"""Deterministic synthetic parameters with the reference's exact state-dict schema.

No pretrained checkpoints are reachable offline (the reference's README points to a Google-Drive folder), so
benchmarks and parity tests use seeded random weights.  The schema files under slide_b200/configs/ list every
(key, shape) of the reference modules' state_dict (exported from the real reference by
tests/golden/make_golden.py, which also checks that these dicts load with strict=True), so a real checkpoint
with the same keys drops in unchanged.
"""
import json
import os

import numpy as np
import torch

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


def load_json(name):
    with open(os.path.join(CONFIG_DIR, name)) as f:
        return json.load(f)


def random_state_dict(schema, seed):
    """schema: list of [key, shape].  Conv / linear weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's
    default init, GroupNorm affine = 1 + 0.1 N(0,1) / 0.1 N(0,1) so that the affine path is exercised,
    embeddings ~ N(0,1).  Generated on the CPU generator -> identical on every machine."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd = {}
    for key, shape in schema:
        shape = tuple(shape)
        if "group_norm.weight" in key or key.endswith("fc_lyaer.1.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "group_norm.bias" in key or key.endswith("fc_lyaer.1.bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("class_emb.weight"):
            t = torch.randn(shape, generator=g)
        elif key.endswith(".weight"):
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            t = (torch.rand(shape, generator=g) * 2 - 1) / np.sqrt(fan_in)
        elif key.endswith(".bias"):
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        else:
            t = torch.randn(shape, generator=g)
        sd[key] = t.float()
    return sd
