"""TEST INFRASTRUCTURE ONLY -- numpy restatement of slide_b200/csrc/mesh.cu (iso-surface of a DPSR grid by marching tetrahedra).

PARITY UNPINNED against the reference: the reference calls skimage.measure.marching_cubes (Lewiner) on the CPU
(dpsr_utils/utils.py:246-287) and scikit-image exists nowhere offline, so neither its triangulation nor its vertex order can be
reproduced or checked.  What is checked instead (tests/test_gpu_mesh.py): the CUDA path equals this restatement bit for bit
(vertices, normals, faces), and the mesh has the properties any correct extraction of the level set has -- vertices on the
linear zero crossing of a grid edge, closed and consistently oriented surface for a closed level set, Euler characteristic 2
for a sphere, area converging to the analytic area.

Only tests/ may import this file.
"""
import numpy as np

CORNER = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64)
TETS = np.array([[0, 5, 1, 6], [0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6]], dtype=np.int64)
f32 = np.float32


def _gradient(phi):
    """numpy.gradient semantics in float32: central differences inside, one-sided at the faces."""
    out = []
    for ax in range(3):
        n = phi.shape[ax]
        lo = np.maximum(np.arange(n) - 1, 0)
        hi = np.minimum(np.arange(n) + 1, n - 1)
        a = np.take(phi, lo, axis=ax)
        b = np.take(phi, hi, axis=ax)
        shape = [1, 1, 1]
        shape[ax] = n
        out.append(((b - a).astype(f32) / (hi - lo).astype(f32).reshape(shape)).astype(f32))
    return out


def extract(phi, level=0.0, vertex_scale=1.0):
    """phi (R,R,R) float32 -> verts (V,3) f32, normals (V,3) f32, faces (F,3) i32, in the CUDA kernel's order."""
    phi = np.ascontiguousarray(phi, dtype=f32)
    R = phi.shape[0]
    level = f32(level)
    ins = phi < level
    # --- vertices: node-major, edge type t = dx + 2 dy + 4 dz = 1..7 minor
    slot_id = -np.ones((R, R, R, 7), dtype=np.int64)
    cross = np.zeros((R, R, R, 7), dtype=bool)
    for t in range(1, 8):
        dx, dy, dz = t & 1, (t >> 1) & 1, (t >> 2) & 1
        a = ins[:R - dx, :R - dy, :R - dz]
        b = ins[dx:, dy:, dz:]
        cross[:R - dx, :R - dy, :R - dz, t - 1] = a != b
    flat = cross.reshape(-1)
    ids = np.cumsum(flat) - 1
    slot_id.reshape(-1)[flat] = ids[flat]
    V = int(flat.sum())
    node, typ = np.nonzero(cross.reshape(-1, 7))
    x, y, z = node // (R * R), (node // R) % R, node % R
    t = typ + 1
    d = np.stack([t & 1, (t >> 1) & 1, (t >> 2) & 1], axis=1)
    fa = phi[x, y, z]
    fb = phi[x + d[:, 0], y + d[:, 1], z + d[:, 2]]
    s = ((level - fa).astype(f32) / (fb - fa).astype(f32)).astype(f32)
    base = np.stack([x, y, z], axis=1).astype(f32)
    verts = ((base + (s[:, None] * d.astype(f32)).astype(f32)).astype(f32) * f32(vertex_scale)).astype(f32)
    g = _gradient(phi)
    ga = np.stack([gg[x, y, z] for gg in g], axis=1)
    gb = np.stack([gg[x + d[:, 0], y + d[:, 1], z + d[:, 2]] for gg in g], axis=1)
    gi = (ga + (s[:, None] * (gb - ga).astype(f32)).astype(f32)).astype(f32)
    sq = (gi * gi).astype(f32)
    ln = np.sqrt(((sq[:, 0] + sq[:, 1]).astype(f32) + sq[:, 2]).astype(f32)).astype(f32)
    inv = np.where(ln > 0, (f32(1.0) / np.where(ln > 0, ln, f32(1.0))).astype(f32), f32(0.0)).astype(f32)
    normals = (gi * inv[:, None]).astype(f32)
    assert len(verts) == V

    def edge_vertex(cx, cy, cz, u, v):
        su, sv = CORNER[u].sum(), CORNER[v].sum()
        lo, hi = (u, v) if su < sv else (v, u)
        dd = CORNER[hi] - CORNER[lo]
        assert (dd >= 0).all() and dd.sum() > 0, "tetrahedron edge outside the 7 edge types"
        slot = dd[0] + 2 * dd[1] + 4 * dd[2] - 1
        vid = slot_id[cx + CORNER[lo][0], cy + CORNER[lo][1], cz + CORNER[lo][2], slot]
        assert vid >= 0
        return int(vid)

    # --- faces: cell-major (x, y, z), tetrahedron, triangle
    in8 = np.zeros((R - 1, R - 1, R - 1), dtype=np.int64)
    for c in range(8):
        in8 |= ins[CORNER[c][0]:R - 1 + CORNER[c][0], CORNER[c][1]:R - 1 + CORNER[c][1],
                   CORNER[c][2]:R - 1 + CORNER[c][2]].astype(np.int64) << c
    faces = []
    for cx, cy, cz in zip(*np.nonzero((in8 != 0) & (in8 != 255))):
        bits = int(in8[cx, cy, cz])
        for tet in TETS:
            m4 = 0
            for k in range(4):
                m4 |= ((bits >> int(tet[k])) & 1) << k
            cnt = bin(m4).count("1")
            if cnt in (0, 4):
                continue
            direc = np.zeros(3, dtype=f32)
            for k in range(4):
                isin = (m4 >> k) & 1
                w = f32(f32(-1.0 if isin else 1.0) / f32(cnt if isin else 4 - cnt))
                direc = (direc + (w * CORNER[tet[k]].astype(f32)).astype(f32)).astype(f32)
            tris = []
            if cnt in (1, 3):
                lone_bits = m4 if cnt == 1 else (~m4 & 15)
                L = (lone_bits & -lone_bits).bit_length() - 1
                others = [k for k in range(4) if k != L]
                tris.append([edge_vertex(cx, cy, cz, int(tet[L]), int(tet[o])) for o in others])
            else:
                P = [k for k in range(4) if (m4 >> k) & 1]
                Q = [k for k in range(4) if not (m4 >> k) & 1]
                e0 = edge_vertex(cx, cy, cz, int(tet[P[0]]), int(tet[Q[0]]))
                e1 = edge_vertex(cx, cy, cz, int(tet[P[0]]), int(tet[Q[1]]))
                e2 = edge_vertex(cx, cy, cz, int(tet[P[1]]), int(tet[Q[1]]))
                e3 = edge_vertex(cx, cy, cz, int(tet[P[1]]), int(tet[Q[0]]))
                tris += [[e0, e1, e2], [e0, e2, e3]]
            for tr in tris:
                p0, p1, p2 = verts[tr[0]], verts[tr[1]], verts[tr[2]]
                u = (p1 - p0).astype(f32)
                v = (p2 - p0).astype(f32)
                n = np.array([f32(f32(u[1] * v[2]) - f32(u[2] * v[1])), f32(f32(u[2] * v[0]) - f32(u[0] * v[2])),
                              f32(f32(u[0] * v[1]) - f32(u[1] * v[0]))], dtype=f32)
                dot = f32(f32(f32(n[0] * direc[0]) + f32(n[1] * direc[1])) + f32(n[2] * direc[2]))
                faces.append([tr[0], tr[2], tr[1]] if dot < 0 else tr)
    return verts, normals, np.asarray(faces, dtype=np.int32).reshape(-1, 3)


def mesh_checks(verts, faces):
    """-> dict(closed, oriented, euler): every undirected edge in exactly two faces / every directed edge exactly once."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0).astype(np.int64)
    und = np.sort(e, axis=1)
    _, cnt = np.unique(und, axis=0, return_counts=True)
    _, dcnt = np.unique(e, axis=0, return_counts=True)
    used = np.unique(faces)
    return dict(closed=bool((cnt == 2).all()), oriented=bool((dcnt == 1).all()),
                euler=int(len(used) - len(cnt) + len(faces)))
