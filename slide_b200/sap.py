"""Host side of the SAP mesh-reconstruction stage over the C ABI of include/slide_sap.h.

Mirrors the reference's interface for this stage:

  DPSR                      dpsr_utils/dpsr.py::DPSR (same constructor arguments, forward(V, N) -> phi)
  mc_from_psr               dpsr_utils/utils.py::mc_from_psr (same arguments; a different, deterministic triangulation)
  SapReconstructor          what dpsr_evaluation.py::visualize_per_rank does per batch between loading the cloud and
                            marching cubes (:214-260): [mirror_and_concat] -> PointNet2CloudCondition(refine JSON) ->
                            network_output_to_dpsr_grid (point_upsample, shapenet_psr_normalize / scale, clamp, DPSR)

torch is used for device memory and streams only; every computation is a kernel of libslide_b200.so and there is no
fallback: a missing library raises.  mc_from_psr extracts the grids' zero level sets on the GPU (marching tetrahedra; the
reference's skimage.measure.marching_cubes runs on the CPU, dpsr_utils/utils.py:246-287, and cannot be reproduced offline).
"""
import ctypes

import torch

from . import engine, lib
from .program import Program


def _bind(l):
    if getattr(l, "_sap_bound", False):
        return l
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    l.slide_sap_mirror_concat.argtypes = [vp, ci, ci, ci, vp, vp, vp, ci, vp]
    l.slide_sap_unit_cube.argtypes = [vp, ci, ci, ci, ci, cf, vp, vp]
    l.slide_dpsr_workspace_bytes.argtypes = [ci, ci, ctypes.POINTER(ctypes.c_size_t)]
    l.slide_dpsr_forward.argtypes = [vp, ci, vp, ci, ci, ci, ci, cf, ci, ci, vp, vp, ctypes.c_size_t, vp]
    l.slide_mc_workspace_bytes.argtypes = [ci, ctypes.POINTER(ctypes.c_size_t)]
    l.slide_mc_count.argtypes = [vp, ci, cf, vp, ctypes.c_size_t, vp, vp]
    l.slide_mc_emit.argtypes = [vp, ci, cf, vp, cf, vp, vp, vp, vp]
    l._sap_bound = True
    return l


def _f32(t):
    assert t.is_cuda and t.dtype == torch.float32
    return t if t.is_contiguous() else t.contiguous()


def mirror_concat(cloud, perm=None, axis=2, out=None):
    """mirror_and_concat(cloud, axis, num_points=[], attach_label=True, permute=perm is not None)[0]
    (data_utils/mirror_partial.py:37-58).  cloud (B,N,6) cuda f32, perm: int32 (2N,) tensor drawn by the caller
    (the reference's torch.randperm on the CPU generator) or None.  -> (B,2N,7)."""
    l = _bind(lib.load())
    cloud = _f32(cloud)
    B, N, C = cloud.shape
    assert C == 6, "points + normals expected"
    if out is None:
        out = torch.empty(B, 2 * N, 7, device=cloud.device, dtype=torch.float32)
    assert out.shape[0] == B and out.shape[1] == 2 * N and out.stride(2) == 1 and out.stride(0) == 2 * N * out.stride(1)
    if perm is not None:
        perm = torch.as_tensor(perm).to(device=cloud.device, dtype=torch.int32).contiguous()
        assert perm.numel() == 2 * N
    centre = torch.empty(B, 3, device=cloud.device, dtype=torch.float32)
    with torch.cuda.device(cloud.device):
        lib.check(l.slide_sap_mirror_concat(lib.ptr(cloud), B, N, int(axis), lib.ptr(perm), lib.ptr(centre),
                                            lib.ptr(out), int(out.stride(1)), lib.stream_of(cloud)),
                  "slide_sap_mirror_concat")
    return out


def unit_cube(points, explicit_normalize=True, scale=1.0):
    """dpsr_evaluation.py:72-76: shapenet_psr_normalize (or / scale / 2), then clamp(x / 1.2 + 0.5, 0, 0.99).
    points (B,n,>=3) cuda f32 with unit column stride (the xyz columns of a wider tensor are fine) -> (B,n,3)."""
    l = _bind(lib.load())
    assert points.is_cuda and points.dtype == torch.float32 and points.stride(2) == 1
    B, n = points.shape[0], points.shape[1]
    assert points.stride(0) == n * points.stride(1)
    out = torch.empty(B, n, 3, device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        lib.check(l.slide_sap_unit_cube(lib.ptr(points), int(points.stride(1)), B, n, int(bool(explicit_normalize)),
                                        float(scale), lib.ptr(out), lib.stream_of(points)), "slide_sap_unit_cube")
    return out


class DPSR(object):
    """dpsr_utils/dpsr.py::DPSR, forward only.  res: (r, r, r) with r a power of two in 8..256."""

    def __init__(self, res, sig=10, scale=True, shift=True):
        res = tuple(int(r) for r in res)
        if len(res) != 3 or len(set(res)) != 1:
            raise NotImplementedError("cubic 3-D grids only (the shipped dpsr_config: grid_res)")
        self.res, self.sig, self.scale, self.shift = res, float(sig), bool(scale), bool(shift)
        self._ws = None

    def _workspace(self, B, device):
        l = _bind(lib.load())
        need = ctypes.c_size_t()
        lib.check(l.slide_dpsr_workspace_bytes(B, self.res[0], ctypes.byref(need)), "slide_dpsr_workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != device:
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=device)
        return self._ws

    def forward(self, V, N, out=None):
        l = _bind(lib.load())
        assert V.shape == N.shape and V.shape[2] == 3
        assert V.is_cuda and V.dtype == torch.float32 and N.dtype == torch.float32
        assert V.stride(2) == 1 and N.stride(2) == 1
        B, n = V.shape[0], V.shape[1]
        assert V.stride(0) == n * V.stride(1) and N.stride(0) == n * N.stride(1)
        r = self.res[0]
        if out is None:
            out = torch.empty(B, r, r, r, device=V.device, dtype=torch.float32)
        ws = self._workspace(B, V.device)
        with torch.cuda.device(V.device):
            lib.check(l.slide_dpsr_forward(lib.ptr(V), int(V.stride(1)), lib.ptr(N), int(N.stride(1)), B, n, r, self.sig,
                                           int(self.shift), int(self.scale), lib.ptr(out), lib.ptr(ws),
                                           ctypes.c_size_t(ws.numel()), lib.stream_of(V)), "slide_dpsr_forward")
        return out

    __call__ = forward


def mc_from_psr(psr_grid, zero_level=0.0, real_scale=False, with_normals=True):
    """dpsr_utils/utils.py::mc_from_psr: iso-surface meshes of a batch of indicator grids, on the GPU.
    psr_grid (B,r,r,r) cuda f32 -> (verts, faces, normals): lists of B device tensors (V_i,3) f32 in [0,1) grid units
    (index / r, or index / (r-1) with real_scale), (F_i,3) int32, (V_i,3) f32 (normalised gradient of the grid).
    The level set is the reference's; the triangulation is marching tetrahedra, not scikit-image's Lewiner marching cubes
    (include/slide_sap.h: slide_mc_count / slide_mc_emit), so vertex / face counts and order differ from the reference's."""
    l = _bind(lib.load())
    assert psr_grid.is_cuda and psr_grid.dtype == torch.float32 and psr_grid.dim() == 4
    B, r = psr_grid.shape[0], psr_grid.shape[1]
    assert psr_grid.shape[2] == r and psr_grid.shape[3] == r
    grid = psr_grid.contiguous()
    dev = grid.device
    need = ctypes.c_size_t()
    lib.check(l.slide_mc_workspace_bytes(r, ctypes.byref(need)), "slide_mc_workspace_bytes")
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    scale = 1.0 / (r - 1) if real_scale else 1.0 / r
    verts, faces, normals = [], [], []
    with torch.cuda.device(dev):
        for i in range(B):
            g = grid[i]
            lib.check(l.slide_mc_count(lib.ptr(g), r, float(zero_level), lib.ptr(ws), ctypes.c_size_t(ws.numel()),
                                       lib.ptr(counts), lib.stream_of(g)), "slide_mc_count")
            nv, nf = (int(c) for c in counts.tolist())  # device -> host: the output sizes
            v = torch.empty(nv, 3, device=dev, dtype=torch.float32)
            n = torch.empty(nv, 3, device=dev, dtype=torch.float32) if with_normals else None
            f = torch.empty(nf, 3, device=dev, dtype=torch.int32)
            if nv:
                lib.check(l.slide_mc_emit(lib.ptr(g), r, float(zero_level), lib.ptr(ws), float(scale), lib.ptr(v),
                                          lib.ptr(n), lib.ptr(f), lib.stream_of(g)), "slide_mc_emit")
            verts.append(v)
            faces.append(f)
            normals.append(n)
    return verts, faces, normals


class SapReconstructor(object):
    """Refinement network + split + DPSR for batches of B clouds of n_points (xyz + normal).

    cfg: dict(pointnet_config, dpsr_config, scale) as in slide_b200/configs/sap_refine.json (the reference's refine
    JSON); sd: state dict with the reference's key names.
    reconstruct(cloud, labels, perm) -> dict(phi (B,r,r,r), points (B,n_fine,3) in DPSR coordinates,
    normals (B,n_fine,3), refined (B,n_fine,6) in network coordinates)."""

    def __init__(self, cfg, sd, B, n_points=2048, device=None, gemm_backend="auto", explicit_normalize=True):
        self.pc, self.dc = cfg["pointnet_config"], cfg["dpsr_config"]
        self.scale = float(cfg.get("scale", 1))
        self.mirror = bool(self.dc.get("mirror_before_upsampling", False))
        if self.dc.get("only_original_points_split", False):
            raise NotImplementedError("only_original_points_split (no shipped refine JSON sets it)")
        self.include_normals = bool(cfg.get("include_normals", True))
        self.explicit_normalize = explicit_normalize
        self.B, self.n_points = B, n_points
        self.n_in = 2 * n_points if self.mirror else n_points
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        assert self.pc["in_fea_dim"] == (4 if self.mirror else 3)
        self.builder, self.h = engine.build_refine(self.pc, sd, B, self.n_in)
        self.prog = Program(self.builder, self.device)
        self.prog.set_gemm_backend(gemm_backend)
        engine.init_constants(self.prog, self.h)
        r = int(self.dc["grid_res"])
        self.dpsr = DPSR((r, r, r), sig=self.dc["psr_sigma"])
        self.n_fine = self.n_in * self.h["factor"]

    def reconstruct(self, cloud, labels, perm=None, mesh=False):
        """mesh=True: also extract the grids' zero level sets (mc_from_psr) -> out["verts"], out["faces"],
        out["vert_normals"] (lists of B device tensors)."""
        B = self.B
        cloud = torch.as_tensor(cloud, dtype=torch.float32).to(self.device, non_blocking=True)
        assert cloud.shape == (B, self.n_points, 6 if self.include_normals or cloud.shape[2] == 6 else 3)
        if not self.include_normals:
            # the refinement network estimates the normals itself (dpsr_evaluation.py:231-233)
            cloud = torch.cat([cloud[:, :, :3], torch.zeros_like(cloud[:, :, :3])], dim=2)
        self.prog.upload(self.h["labels"], torch.as_tensor(labels, dtype=torch.int32).to(self.device))
        self.prog.run_segment("setup")  # class embedding of this batch's labels -> the modules' condition vectors
        X = self.prog.view(self.h["x"])  # (B*n_in, ld)
        if self.mirror:
            if perm is None:
                raise ValueError("mirror_before_upsampling re-orders the points with torch.randperm(2N): pass perm")
            mirror_concat(cloud, perm, axis=2, out=X.view(B, self.n_in, X.shape[1]))
        else:
            X.view(B, self.n_in, X.shape[1])[:, :, :6].copy_(cloud)
        self.prog.run_segment("refine")
        fine = self.prog.view(self.h["fine"]).view(B, self.n_fine, -1)
        pts = unit_cube(fine, self.explicit_normalize, self.scale)
        nrm = fine[:, :, 3:6]
        phi = self.dpsr(pts, nrm)
        if lib.load().slide_tc_error():
            raise lib.SlideError("tcgen05 pipeline wait timed out (results invalid)")
        out = dict(phi=phi, points=pts, normals=nrm, refined=fine[:, :, :6])
        if mesh:
            out["verts"], out["faces"], out["vert_normals"] = mc_from_psr(phi, zero_level=0.0)
        return out


def load_default(B, n_points=2048, seed=21, config="sap_refine", **kw):
    """SapReconstructor on a shipped refine JSON with seeded random weights of the reference's schema (no checkpoints are
    reachable offline).  config: "sap_refine" (mirrored input, normals given, 5 children per point) or "sap_refine_plain"
    (no mirroring, normals estimated by the network, 10 children per point)."""
    from . import weights
    cfg = weights.load_json(config + ".json")
    sd = weights.random_state_dict(weights.load_json("schema_%s.json" % config), seed)
    return SapReconstructor(cfg, sd, B, n_points, **kw)
