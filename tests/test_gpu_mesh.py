"""Iso-surface extraction (slide_mc_count / slide_mc_emit through slide_b200.sap.mc_from_psr) against oracle/mesh_oracle.py
-- bit for bit: vertices, normals, faces -- and against what any correct extraction of a level set must satisfy.
Parity with the reference's scikit-image call is UNPINNED (scikit-image is not available offline; oracle/mesh_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import mesh_oracle, sap_oracle
from slide_b200 import sap

pytestmark = pytest.mark.gpu


def _sphere(R, c, rad):
    g = np.arange(R, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return (np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - rad).astype(np.float32)


def _cases():
    out = {"sphere24": _sphere(24, (11.3, 12.1, 11.7), 7.4), "two_spheres32": np.minimum(_sphere(32, (9.2, 10.1, 9.7), 5.3),
                                                                                        _sphere(32, (22.4, 20.6, 21.1), 6.1))}
    g = torch.Generator().manual_seed(0)
    V = torch.rand(1, 800, 3, generator=g) * 0.5 + 0.25
    N = torch.nn.functional.normalize(V - 0.5, dim=2)
    out["dpsr32"] = sap_oracle.dpsr_forward(V, N, (32, 32, 32), 2)[0].numpy()
    lat = _sphere(16, (8.0, 8.0, 8.0), 4.0)   # centre on a node, integer radius: exact zeros on nodes (degenerate triangles)
    out["exact_zeros16"] = lat
    out["empty8"] = np.ones((8, 8, 8), dtype=np.float32)
    return out


@pytest.mark.parametrize("name", ["sphere24", "two_spheres32", "dpsr32", "exact_zeros16", "empty8"])
def test_mesh_matches_oracle_bit_for_bit(name):
    phi = _cases()[name]
    want_v, want_n, want_f = mesh_oracle.extract(phi, 0.0, vertex_scale=1.0 / phi.shape[0])
    v, f, n = sap.mc_from_psr(torch.from_numpy(phi)[None].cuda(), zero_level=0.0)
    v, f, n = v[0].cpu().numpy(), f[0].cpu().numpy(), n[0].cpu().numpy()
    assert v.shape == want_v.shape and f.shape == want_f.shape
    assert np.array_equal(v, want_v)
    assert np.array_equal(f, want_f)
    assert np.array_equal(n, want_n)


def test_mesh_properties_at_the_shipped_size():
    R, c, rad = 128, (63.2, 64.9, 62.7), 41.3
    phi = _sphere(R, c, rad)
    phis = torch.from_numpy(np.stack([phi, -phi])).cuda()          # second grid: inside / outside swapped
    verts, faces, normals = sap.mc_from_psr(phis, zero_level=0.0, real_scale=False)
    for i in range(2):
        v, f, n = verts[i].cpu().numpy().astype(np.float64) * R, faces[i].cpu().numpy(), normals[i].cpu().numpy()
        chk = mesh_oracle.mesh_checks(v, f)
        assert chk == dict(closed=True, oriented=True, euler=2), chk
        r = np.sqrt(((v - np.asarray(c)) ** 2).sum(1))
        # vertices sit on the LINEAR crossing of an edge of length <= sqrt(3): off the sphere by at most 3 / (8 rad) = 0.009
        assert np.abs(r - rad).max() < 1.2e-2
        p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        cr = np.cross(p1 - p0, p2 - p0)
        area = 0.5 * np.linalg.norm(cr, axis=1).sum()
        assert abs(area / (4 * np.pi * rad ** 2) - 1) < 2e-3
        outward = ((cr * ((p0 + p1 + p2) / 3 - np.asarray(c))).sum(1) > 0).mean()
        # faces are oriented from phi < 0 to phi >= 0: outward for the sphere's signed distance, inward for its negative
        assert outward == (1.0 if i == 0 else 0.0)
        radial = (v - np.asarray(c)) / r[:, None]
        assert ((n * radial).sum(1) * (1 if i == 0 else -1)).min() > 0.999   # normals = normalised gradient


def test_mesh_from_the_reconstructor_output():
    """cloud -> ... -> DPSR grid -> mesh: closed surfaces only (the indicator grid of a sphere of oriented points)."""
    g = torch.Generator().manual_seed(1)
    d = torch.nn.functional.normalize(torch.randn(2, 6000, 3, generator=g), dim=2)
    V = (0.5 + 0.3 * d).cuda()
    phi = sap.DPSR((64, 64, 64), sig=2)(V, d.cuda())
    verts, faces, _ = sap.mc_from_psr(phi)
    for v, f in zip(verts, faces):
        assert f.shape[0] > 1000
        chk = mesh_oracle.mesh_checks(v.cpu().numpy(), f.cpu().numpy())
        assert chk["closed"] and chk["oriented"]
        r = (v.cpu() - 0.5).norm(dim=1)
        assert float((r - 0.3).abs().max()) < 0.03
