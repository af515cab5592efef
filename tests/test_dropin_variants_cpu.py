"""The drop-in `pointnet2_ops` python package on the module variants no shipped sampling config selects -- bn_first, swish,
first_conv, identity / no residual, no normalisation, second condition, plain-conv and un-transformed attention, global
attention, ball-query abstraction with pooling, multi-scale grouping, the propagation modules' grouper -- against golden
vectors of the REAL reference classes (tests/golden/make_golden_variants.py -> golden_variants.npz; case table in
tests/golden/variant_cases.py).  SURVEY 8 rows a8-a15: the program lowering refuses these variants (NotImplementedError),
the drop-in modules are what covers them.

Runs on the CPU: the drop-in's python layer is torch; its nine `_ext` entry points and `knn_points` are replaced, for this
test only, by the C oracle (the same ops the golden generator gave the reference), so any difference is the python layer's.
"""
import os

import numpy as np
import pytest
import torch

import slide_b200
from oracle import ops
from tests.golden import variant_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXT = ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
       "three_interpolate_grad", "ball_query", "group_points", "group_points_grad")


@pytest.fixture(scope="module")
def gv():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_variants.npz"))


@pytest.fixture()
def dropin_on_oracle(monkeypatch):
    slide_b200.install_dropin()
    from pointnet2_ops import _ext
    from pytorch3d.ops import knn
    for name in EXT:
        monkeypatch.setattr(_ext, name, getattr(ops, name))
    monkeypatch.setattr(knn, "knn_points", ops.knn_points)
    from pointnet2_ops import attention, pointnet2_modules
    return pointnet2_modules, attention


CASES = variant_cases.cases()


@pytest.mark.parametrize("name,cls,kw,kind", CASES, ids=[c[0] for c in CASES])
def test_variant_matches_reference(name, cls, kw, kind, gv, dropin_on_oracle):
    modules, attention = dropin_on_oracle
    ctor = getattr(modules, cls, None) or getattr(attention, cls)
    module = ctor(**kw).eval()
    pre = name + "/sd/"
    sd = {k[len(pre):]: torch.from_numpy(gv[k]) for k in gv.files if k.startswith(pre)}
    module.load_state_dict(sd, strict=True)          # same keys and shapes as the reference's module
    pre = name + "/in/"
    inp = {k[len(pre):]: torch.from_numpy(gv[k]) for k in gv.files if k.startswith(pre)}
    if kind == "att_all":
        inp["count"] = "all"
    with torch.no_grad():
        outs = variant_cases.call(module, kind, inp)
    n_gold = sum(1 for k in gv.files if k.startswith(name + "/out"))
    assert len(outs) == n_gold
    for j, o in enumerate(outs):
        want = torch.from_numpy(gv["%s/out%d" % (name, j)])
        assert o.shape == want.shape, (name, j)
        if o.dtype != torch.float32 or want.abs().max() == 0:
            assert torch.equal(o, want), (name, j)
            continue
        err = float((o - want).abs().max())
        assert err <= 2e-5 * max(1.0, float(want.abs().max())), (name, j, err)


def test_lowering_refuses_what_it_does_not_implement(pipeline_cfg):
    """The fused program path must fail loudly -- not silently compute something else -- on the variants above."""
    import copy
    from slide_b200 import engine
    from tests import common
    pos = pipeline_cfg["position_ddpm"]
    d = pos["diffusion_config"]
    table = engine.position_table(d["T"], d["beta_0"], d["beta_T"])
    for key, value in (("bn_first", True), ("activation", "swish"), ("res_connect", False)):
        pc = copy.deepcopy(pos["pointnet_config"])
        pc[key] = value
        with pytest.raises(NotImplementedError):
            engine.build_ddpm(pc, common.state_dict("pos"), 2, d["T"], table, 0, keep_cols=0)
