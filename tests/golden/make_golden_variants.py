"""Golden vectors for the module VARIANTS of the reference's `pointnet2_ops` package that no shipped sampling config
selects (tests/golden/variant_cases.py: bn_first, swish, first_conv, identity / no residual, no normalisation, second
condition, plain-conv and un-transformed attention, global attention, ball-query abstraction with pooling, multi-scale
grouping, the propagation modules' grouper), produced by the REAL reference classes
(build container only: needs /root/reference; the C oracle stands in for `_ext` and pytorch3d):

    python tests/golden/make_golden_variants.py      ->  tests/golden/golden_variants.npz

Per case `<name>`: `<name>/sd/<key>` the seeded state dict, `<name>/in/<arg>` the inputs, `<name>/out<i>` the outputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops  # noqa: E402

ops.install_reference_stubs()
from pointnet2_ops import pointnet2_modules as ref_modules  # noqa: E402
from pointnet2_ops import attention as ref_attention  # noqa: E402
from tests.golden import variant_cases  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def seed_state(module, g):
    """Non-trivial parameters everywhere (norm scales around 1, biases / shifts non-zero)."""
    sd = {}
    for k, v in module.state_dict().items():
        if v.dim() > 1:
            w = torch.randn(v.shape, generator=g) * (1.5 / max(1.0, float(v.shape[1])) ** 0.5)
        elif k.endswith("group_norm.weight"):
            w = 1.0 + 0.2 * torch.randn(v.shape, generator=g)
        else:
            w = 0.2 * torch.randn(v.shape, generator=g)
        sd[k] = w.float()
    module.load_state_dict(sd, strict=True)
    return sd


def main():
    gold = {}
    for i, (name, cls, kw, kind) in enumerate(variant_cases.cases()):
        g = torch.Generator().manual_seed(1000 + i)
        ctor = getattr(ref_modules, cls, None) or getattr(ref_attention, cls)
        module = ctor(**kw).eval()
        sd = seed_state(module, g)
        inp = variant_cases.make_inputs(kind, variant_cases.cases()[i][2], torch, g)
        with torch.no_grad():
            outs = variant_cases.call(module, kind, inp)
        for k, v in sd.items():
            gold["%s/sd/%s" % (name, k)] = v.numpy()
        for k, v in inp.items():
            if not isinstance(v, str):
                gold["%s/in/%s" % (name, k)] = v.numpy()
        for j, o in enumerate(outs):
            assert torch.isfinite(o).all(), name
            gold["%s/out%d" % (name, j)] = o.numpy()
        print("%-36s %-22s params %6d  outputs %s" % (name, cls, sum(v.numel() for v in sd.values()),
                                                    [tuple(o.shape) for o in outs]))
    path = os.path.join(OUT, "golden_variants.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
