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


def test_mesh_oracle_properties():
    """oracle/mesh_oracle.py (the checker of the GPU iso-surface kernels) on a sphere: closed, consistently oriented, genus 0,
    area and radius right, normals outward.  (Parity with scikit-image is unpinned: not available offline.)"""
    from oracle import mesh_oracle
    R, c, rad = 24, np.array([11.3, 12.1, 11.7], dtype=np.float32), 7.4
    g = np.arange(R, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    phi = (np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - rad).astype(np.float32)
    v, n, f = mesh_oracle.extract(phi, 0.0)
    assert mesh_oracle.mesh_checks(v, f) == dict(closed=True, oriented=True, euler=2)
    r = np.sqrt(((v - c) ** 2).sum(1))
    assert np.abs(r - rad).max() < 3.0 / (8 * rad) + 1e-3
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    cr = np.cross(p1 - p0, p2 - p0)
    assert abs(0.5 * np.linalg.norm(cr, axis=1).sum() / (4 * np.pi * rad ** 2) - 1) < 1e-2
    assert ((cr * ((p0 + p1 + p2) / 3 - c)).sum(1) > 0).all()
    assert ((n * (v - c) / r[:, None]).sum(1)).min() > 0.999
    # a grid without a crossing gives an empty mesh
    v, n, f = mesh_oracle.extract(np.ones((6, 6, 6), dtype=np.float32), 0.0)
    assert len(v) == 0 and len(f) == 0


def test_plain_refine_lowering_matches_reference_golden(gold):
    """The shipped configuration without mirroring / input normals (in_fea_dim 3, 10 children per point)."""
    cfg = weights.load_json("sap_refine_plain.json")
    pc = cfg["pointnet_config"]
    assert not cfg["include_normals"] and not cfg["dpsr_config"].get("mirror_before_upsampling")
    sd = weights.random_state_dict(weights.load_json("schema_sap_refine_plain.json"), 22)
    B, N = 2, 2048
    b, h = engine.build_refine(pc, sd, B, N)
    assert h["factor"] == 10 and h["F"] == 6
    m = ir_exec.Machine(b)
    engine.init_constants(m, h)
    m.upload(h["labels"], gold["label"].astype(np.int32))
    X = np.concatenate([gold["cloud"][:, :, :3], np.zeros_like(gold["cloud"][:, :, :3])], axis=2)
    m.upload(h["x"], X.reshape(-1, 6))
    m.run(*b.segments["setup"])
    m.run(*b.segments["refine"])
    disp = np.asarray(m.download(h["disp"])).reshape(B, N, 60)
    assert np.abs(disp[:, ::8] - gold["plain_disp_rows8"]).max() < 5e-5 * np.abs(gold["plain_disp_rows8"]).max()
    phi, _, _ = sap_oracle.refine_to_grid(torch.from_numpy(X), torch.from_numpy(disp), (16, 16, 16), 2, 10,
                                          pc["output_scale_factor"], indicator=False)
    assert np.abs(phi.numpy() - gold["plain_phi_r16"]).max() < 2e-4


@pytest.mark.parametrize("knob", ["SLIDE_PAIR_LITE", "SLIDE_HOIST_GEOMETRY", "SLIDE_FACTOR_GROUP"])
def test_lowering_variants_agree(gold, knob, monkeypatch):
    """The lowering's A/B knobs change the record list, not the function: folding the neighbour-coordinate term into U
    (PAIR-lite), hoisting the coordinate-only records onto level 0's side branch, and the materialised (GROUP + GEMM) form of
    the grouped convs all give the default lowering's output on the CPU interpreter."""
    cfg = weights.load_json("sap_refine.json")
    pc = cfg["pointnet_config"]
    sd = weights.random_state_dict(weights.load_json("schema_sap_refine.json"), 21)
    B, N = 1, 2048
    X = sap_oracle.mirror_concat(torch.from_numpy(gold["cloud"][:1]), gold["perm"]).reshape(-1, 7).numpy()

    def run():
        b, h = engine.build_refine(pc, sd, B, 2 * N)
        m = ir_exec.Machine(b)
        engine.init_constants(m, h)
        m.upload(h["labels"], gold["label"][:1].astype(np.int32))
        m.upload(h["x"], X)
        m.run(*b.segments["setup"])
        m.run(*b.segments["refine"])
        return np.array(m.download(h["disp"])), [(op[0], op[3]) for op in b.ops]

    base, n_base = run()
    monkeypatch.setenv(knob, "0")
    alt, n_alt = run()
    assert n_alt != n_base, "the knob did not change the lowering (record kinds / order)"
    assert np.abs(alt - base).max() < 2e-5 * np.abs(base).max()
