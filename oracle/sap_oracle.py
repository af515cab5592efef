"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's SAP mesh-reconstruction path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product path
(slide_b200/sap.py -> libslide_b200.so) never does.

Pinned: tests/golden/make_golden_sap.py imports the REAL reference (dpsr_utils/dpsr.py::DPSR,
dpsr_evaluation.py::network_output_to_dpsr_grid / shapenet_psr_normalize, data_utils/mirror_partial.py::mirror,
models/point_upsample_module.py::point_upsample) in the build container and asserts that every function below is
bit-identical to it on CPU (torch.equal) before writing tests/golden/golden_sap.npz.

What it restates (reference file:line):
  mirror_concat          data_utils/mirror_partial.py:8-58   (mirror + attach_label + permutation)
  psr_normalize          dpsr_evaluation.py:22-32            (bounding-box normalisation to the ShapeNet-PSR scale)
  to_unit_cube           dpsr_evaluation.py:72-76            (scale branch / explicit branch + clamp)
  point_rasterize        dpsr_utils/utils.py:139-200         (periodic trilinear splat)
  grid_interp            dpsr_utils/utils.py:73-115          (periodic trilinear read)
  spectral tables        dpsr_utils/utils.py:24-71           (fftfreqs, spec_gaussian_filter)
  dpsr_forward           dpsr_utils/dpsr.py:30-77            (spectral Poisson solve, shift, scale)
  refine_to_grid         dpsr_evaluation.py:46-86            (split, upsample, normalise, solve)
"""
import numpy as np
import torch

from . import ref_model


def mirror_concat(X, perm=None, axis=2):
    """X (B,N,6) points + normals -> (B,2N,7): originals labelled +1, copies reflected through the plane
    through the centroid normal to `axis` labelled -1 (normal component of that axis negated), then the 2N points re-ordered by
    `perm` (the reference draws torch.randperm(2N) on the CPU generator; None = no re-ordering, the reference's
    only_original_points_split case)."""
    B, N, C = X.shape
    assert C == 6
    centre = torch.mean(X[:, :, 0:3], dim=1, keepdim=True)
    flip = X.clone()
    rel = flip[:, :, 0:3] - centre          # every coordinate takes the (x - c) + c round trip, as in the reference
    rel[:, :, axis] = -rel[:, :, axis]
    flip[:, :, 0:3] = rel + centre
    flip[:, :, axis + 3] = -flip[:, :, axis + 3]
    one = torch.ones(B, N, 1, dtype=X.dtype)
    both = torch.cat([torch.cat([X, one], dim=2), torch.cat([flip, -one], dim=2)], dim=1)
    if perm is not None:
        both = both[:, torch.as_tensor(perm, dtype=torch.long), :]
    return both


def psr_normalize(x):
    lo = x.min(dim=1, keepdim=True)[0]
    hi = x.max(dim=1, keepdim=True)[0]
    centre = (hi + lo) / 2
    extent = (hi - lo).max(dim=2, keepdim=True)[0]
    return (x - centre) / extent * 0.99


def to_unit_cube(points, scale=1.0, explicit_normalize=True):
    p = psr_normalize(points) if explicit_normalize else points / scale / 2
    return torch.clamp(p / 1.2 + 0.5, min=0, max=0.99)


def _corners(pts, res):
    """Shared by splat and read: for each of the 8 cell corners (x-major bit order, z fastest) the wrapped integer
    node and the trilinear weight |p - opposite corner| / cell, multiplied x*y then *z in fp32."""
    size = torch.tensor([float(r) for r in res], dtype=pts.dtype)
    cell = 1.0 / size
    q = pts / cell
    i0 = torch.floor(q).long()
    i1 = torch.fmod(torch.ceil(q), size).long()
    lo = i0.to(cell.dtype) * cell
    hi = (i0.to(cell.dtype) + 1) * cell
    out = []
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                c = (cx, cy, cz)
                node = torch.stack([(i1 if c[d] else i0)[..., d] for d in range(3)], dim=-1)
                opp = torch.stack([(lo if c[d] else hi)[..., d] for d in range(3)], dim=-1)
                w = torch.abs(pts - opp) / cell
                out.append((node, (w[..., 0] * w[..., 1]) * w[..., 2]))
    return out


def point_rasterize(pts, vals, res):
    """pts (B,N,3) in [0,1), vals (B,N,F) -> (B,F,r0,r1,r2); accumulation order = the reference's CPU
    scatter_add_ order (batch, point, corner, feature)."""
    B, N, _ = pts.shape
    Fd = vals.shape[2]
    r0, r1, r2 = res
    corners = _corners(pts, res)
    node = torch.stack([c[0] for c in corners], dim=2)          # (B,N,8,3)
    w = torch.stack([c[1] for c in corners], dim=2)             # (B,N,8)
    flat = (node[..., 0] * r1 + node[..., 1]) * r2 + node[..., 2]
    contrib = w.unsqueeze(-1) * vals.unsqueeze(-2)              # (B,N,8,F)
    bidx = torch.arange(B).view(B, 1, 1, 1)
    fidx = torch.arange(Fd).view(1, 1, 1, Fd)
    target = ((bidx * Fd + fidx) * (r0 * r1 * r2) + flat.unsqueeze(-1)).reshape(-1)
    grid = torch.zeros(B * Fd * r0 * r1 * r2, dtype=vals.dtype)
    grid.scatter_add_(0, target, contrib.reshape(-1))
    return grid.view(B, Fd, r0, r1, r2)


def grid_interp(phi, pts):
    """phi (B,r0,r1,r2), pts (B,N,3) -> (B,N): corners summed in the same 8-corner order."""
    B = phi.shape[0]
    res = tuple(phi.shape[1:])
    corners = _corners(pts, res)
    b = torch.arange(B).view(B, 1)
    lat = torch.stack([phi[b, n[..., 0], n[..., 1], n[..., 2]] for n, _ in corners], dim=2)
    w = torch.stack([c[1] for c in corners], dim=2)
    return torch.sum(lat * w, dim=-1)


def frequencies(res):
    """(kx, ky, kz) integer frequency vectors of the half spectrum (full, full, r2/2+1)."""
    return [np.fft.fftfreq(res[0], d=1 / res[0]), np.fft.fftfreq(res[1], d=1 / res[1]),
            np.fft.rfftfreq(res[2], d=1 / res[2])]


def gaussian_table(res, sig):
    """float32 [r0, r1, r2/2+1]: exp(-0.5 (2 sig |k| / r0)^2) evaluated in float64."""
    kx, ky, kz = [torch.tensor(f, dtype=torch.float64) for f in frequencies(res)]
    k2 = kx[:, None, None] ** 2 + ky[None, :, None] ** 2 + kz[None, None, :] ** 2
    return torch.exp(-0.5 * ((sig * 2 * torch.sqrt(k2) / res[0]) ** 2)).float()


def omega_table(res):
    """float32 [r0, r1, r2/2+1, 3]: 2 pi k (the float32 frequency times float32(2 pi))."""
    kx, ky, kz = [torch.tensor(f, dtype=torch.float32) for f in frequencies(res)]
    w = torch.stack(torch.meshgrid(kx, ky, kz, indexing="ij"), dim=-1)
    w *= 2 * np.pi
    return w


def dpsr_forward(V, N, res, sig, shift=True, scale=True):
    """Indicator grid phi (B,r0,r1,r2) of oriented points V (B,n,3) in [0,1), normals N (B,n,3)."""
    ras = point_rasterize(V, N, res)                              # (B,3,r,r,r)
    spec = torch.fft.rfftn(ras, dim=(2, 3, 4))                    # (B,3,r,r,r/2+1) complex
    spec = spec.permute(0, 2, 3, 4, 1)                            # (B,r,r,r/2+1,3)
    G = gaussian_table(res, sig)[..., None, None]
    sm = torch.view_as_real((spec[..., None] * G)[..., 0])        # (B,...,3,2)
    om = omega_table(res).unsqueeze(-1)                           # (...,3,1)
    # -(i * z) = (im, -re)
    rot = torch.stack([sm[..., 1], -sm[..., 0]], dim=-1)
    div = torch.sum(rot * om, dim=-2)                             # (B,...,2)
    lap = -torch.sum(om ** 2, -2)
    Phi = div / (lap + 1e-6)
    Phi[:, 0, 0, 0, :] = 0
    phi = torch.fft.irfftn(torch.view_as_complex(Phi.contiguous()), s=res, dim=(1, 2, 3))
    if shift or scale:
        fv = grid_interp(phi, V)
        if shift:
            phi = phi - torch.mean(fv, dim=-1).view(-1, 1, 1, 1)
        f0 = phi[:, 0, 0, 0]
        if scale:
            phi = -phi / torch.abs(f0.view(-1, 1, 1, 1)) * 0.5
    return phi


def refine_to_grid(X, displacement, res, sig, factor, out_scale, dataset_scale=1.0, indicator=True,
                   only_original=False, explicit_normalize=True):
    """network_output_to_dpsr_grid: X (B,n,6|7) the (mirrored) network input, displacement (B,n,6*factor) its output
    -> (phi, unit-cube points (B,n*factor,3), normals (B,n*factor,3))."""
    coarse = X[:, :, :-1] if indicator else X
    if indicator and only_original:
        half = X.shape[1] // 2
        coarse, displacement = coarse[:, :half], displacement[:, :half]
    fine = ref_model.point_upsample(coarse, displacement, factor, out_scale)
    pts = to_unit_cube(fine[:, :, 0:3], dataset_scale, explicit_normalize)
    nrm = fine[:, :, 3:]
    return dpsr_forward(pts, nrm, res, sig), pts, nrm
