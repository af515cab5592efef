"""SAP mesh-reconstruction stage without a GPU: the oracle against the golden vectors of the REAL reference
(tests/golden/make_golden_sap.py) and the refine-network lowering through the CPU interpreter."""
import os

import numpy as np
import pytest
import torch

from oracle import ir_exec, sap_oracle
from slide_b200 import engine, weights

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_sap.npz")))


def test_dpsr_oracle_matches_reference_golden(gold):
    V, N = torch.from_numpy(gold["edge_V"]), torch.from_numpy(gold["edge_N"])
    for shift, scale in ((True, True), (False, False)):
        phi = sap_oracle.dpsr_forward(V, N, (32, 32, 32), 2, shift=shift, scale=scale)
        assert np.array_equal(phi.numpy(), gold["edge_phi_%d%d" % (shift, scale)])
    phi = sap_oracle.dpsr_forward(V, N, (128, 128, 128), 2)
    assert np.array_equal(phi[:, ::8, ::8, ::8].numpy(), gold["edge_phi128_sub"])


def test_dpsr_oracle_properties():
    g = torch.Generator().manual_seed(3)
    V = torch.rand(1, 300, 3, generator=g) * 0.99
    N1 = torch.randn(1, 300, 3, generator=g)
    N2 = torch.randn(1, 300, 3, generator=g)
    f = lambda n: sap_oracle.dpsr_forward(V, n, (16, 16, 16), 2, shift=False, scale=False)
    # the unshifted, unscaled solve is linear in the normals and has zero mean (the DC mode is removed)
    a, b, c = f(N1), f(N2), f(N1 + 2 * N2)
    assert torch.allclose(c, a + 2 * b, atol=1e-5 * float(c.abs().max()))
    assert abs(float(a.mean())) < 1e-6 * float(a.abs().max())
    # shift + scale pin the value at the grid origin to -+0.5 and the mean value at the points to 0
    phi = sap_oracle.dpsr_forward(V, N1, (16, 16, 16), 2)
    assert abs(abs(float(phi[0, 0, 0, 0])) - 0.5) < 1e-6
    assert abs(float(sap_oracle.grid_interp(phi, V).mean())) < 1e-5


def test_refine_lowering_matches_reference_golden(gold):
    cfg = weights.load_json("sap_refine.json")
    pc = cfg["pointnet_config"]
    sd = weights.random_state_dict(weights.load_json("schema_sap_refine.json"), 21)
    B, N = 2, 2048
    b, h = engine.build_refine(pc, sd, B, 2 * N)
    m = ir_exec.Machine(b)
    engine.init_constants(m, h)
    m.upload(h["labels"], gold["label"].astype(np.int32))
    X = sap_oracle.mirror_concat(torch.from_numpy(gold["cloud"]), gold["perm"])
    m.upload(h["x"], X.reshape(-1, 7).numpy())
    m.run(*b.segments["setup"])
    m.run(*b.segments["refine"])
    disp = np.asarray(m.download(h["disp"])).reshape(B, 2 * N, 30)
    fine = np.asarray(m.download(h["fine"])).reshape(B, 2 * N * 5, 6)
    assert np.abs(disp[:, ::8] - gold["disp_rows8"]).max() < 5e-5 * np.abs(gold["disp_rows8"]).max()
    assert np.abs(fine[:, ::64] - gold["fine_rows64"]).max() < 1e-6
    # and the grid the oracle builds from the interpreter's output is the reference's grid
    phi, _, _ = sap_oracle.refine_to_grid(X[:1], torch.from_numpy(disp[:1]), (32, 32, 32), 2, 5, pc["output_scale_factor"])
    assert np.abs(phi.numpy() - gold["phi_r32"]).max() < 2e-4
