"""ctypes front-end of the CPU oracle (oracle/slide_oracle.c) on torch CPU tensors.

TEST INFRASTRUCTURE ONLY -- see the header of slide_oracle.c.  The function names and argument
order follow the reference's pybind module (pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19)
and the pytorch3d 0.7.0 python API used at pointnet2_ops/pointnet2_utils.py:370,506-507 and
models/point_upsample_decoder.py:178-180.
"""
import collections
import ctypes
import os
import subprocess
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libslide_oracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libslide_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32(t):
    assert t.dtype == torch.float32 and t.device.type == "cpu"
    return t.contiguous()


def _i32(t):
    assert t.dtype == torch.int32 and t.device.type == "cpu"
    return t.contiguous()


def opt_n_threads(n):
    return lib().so_opt_n_threads(int(n))


# ---- the nine functions of pointnet2_ops._ext ------------------------------------------------
def furthest_point_sampling(points, nsamples):
    points = _f32(points)
    B, N, _ = points.shape
    out = torch.zeros(B, nsamples, dtype=torch.int32)
    rc = lib().so_fps(_p(points), B, N, int(nsamples), _p(out))
    assert rc == 0
    return out


def gather_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    m = idx.shape[1]
    out = torch.zeros(B, C, m)
    lib().so_gather(_p(points), _p(idx), B, C, N, m, _p(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, m = grad_out.shape
    out = torch.zeros(B, C, n)
    lib().so_gather_grad(_p(grad_out), _p(idx), B, C, int(n), m, _p(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    B, m, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = torch.zeros(B, m, nsample, dtype=torch.int32)
    cnt = torch.zeros(B, m, dtype=torch.int32)
    lib().so_ball_query(_p(new_xyz), _p(xyz), B, N, m, ctypes.c_float(radius), int(nsample), _p(idx), _p(cnt))
    return idx, cnt


def group_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    _, npoint, ns = idx.shape
    out = torch.zeros(B, C, npoint, ns)
    lib().so_group(_p(points), _p(idx), B, C, N, npoint, ns, _p(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, npoint, ns = grad_out.shape
    out = torch.zeros(B, C, n)
    lib().so_group_grad(_p(grad_out), _p(idx), B, C, int(n), npoint, ns, _p(out))
    return out


def three_nn(unknown, known):
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.zeros(B, n, 3)
    idx = torch.zeros(B, n, 3, dtype=torch.int32)
    lib().so_three_nn(_p(unknown), _p(known), B, n, m, _p(dist2), _p(idx))
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(B, C, n)
    lib().so_three_interpolate(_p(points), _p(idx), _p(weight), B, C, m, n, _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    out = torch.zeros(B, C, m)
    lib().so_three_interpolate_grad(_p(grad_out), _p(idx), _p(weight), B, C, n, int(m), _p(out))
    return out


# ---- pytorch3d 0.7.0 surface -------------------------------------------------------------------
_KNN = collections.namedtuple("KNN", "dists idx knn")


def knn_gather(x, idx, lengths=None):
    """x (N,M,U), idx (N,L,K) -> (N,L,K,U); entries beyond `lengths` are zeroed like pytorch3d."""
    N, M, U = x.shape
    _, L, K = idx.shape
    out = x[:, :, None].expand(-1, -1, K, -1).gather(1, idx[:, :, :, None].expand(-1, -1, -1, U))
    if lengths is not None and (lengths < K).any():
        mask = lengths[:, None] <= torch.arange(K, device=x.device)[None]
        out = out.masked_fill(mask[:, None, :, None].expand(-1, L, -1, U), 0.0)
    return out


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False,
               return_sorted=True):
    assert norm == 2
    p1c, p2c = _f32(p1), _f32(p2)
    B, P1, D = p1c.shape
    P2 = p2c.shape[1]
    l1 = lengths1.to(torch.int64).contiguous() if lengths1 is not None else None
    l2 = lengths2.to(torch.int64).contiguous() if lengths2 is not None else None
    dists = torch.zeros(B, P1, K)
    idx = torch.zeros(B, P1, K, dtype=torch.int64)
    rc = lib().so_knn(_p(p1c), _p(p2c), B, P1, P2, D, _p(l1), _p(l2), int(K), _p(dists), _p(idx))
    assert rc == 0
    nn = knn_gather(p2, idx, lengths2) if return_nn else None
    return _KNN(dists, idx, nn)


def masked_gather(points, idx):
    """points (N,P,D), idx (N,K) with -1 padding -> (N,K,D), padded rows zero."""
    D = points.shape[2]
    mask = idx.eq(-1)
    safe = idx.clone()
    safe[mask] = 0
    out = points.gather(1, safe[:, :, None].expand(-1, -1, D))
    out[mask] = 0.0
    return out


def draw_start_indices(lengths):
    """The per-cloud `torch.randint(high=lengths[n], size=(1,)).item()` draws of pytorch3d 0.7.0 (CPU
    default generator, batch order)."""
    return torch.tensor([int(torch.randint(high=int(l), size=(1,)).item()) for l in lengths], dtype=torch.int64)


def sample_farthest_points(points, lengths=None, K=50, random_start_point=False, start_idx=None):
    pts = _f32(points)
    N, P, D = pts.shape
    if lengths is None:
        lengths = torch.full((N,), P, dtype=torch.int64)
    lengths = lengths.to(torch.int64).contiguous()
    if isinstance(K, int):
        Kt = torch.full((N,), K, dtype=torch.int64)
    elif isinstance(K, list):
        Kt = torch.tensor(K, dtype=torch.int64)
    else:
        Kt = K.to(torch.int64).contiguous()
    maxK = int(Kt.max())
    if start_idx is None:
        start_idx = draw_start_indices(lengths) if random_start_point else torch.zeros(N, dtype=torch.int64)
    start_idx = start_idx.to(torch.int64).contiguous()
    idx = torch.zeros(N, maxK, dtype=torch.int64)
    rc = lib().so_fps_p3d(_p(pts), N, P, D, _p(lengths), _p(Kt), _p(start_idx), maxK, _p(idx))
    assert rc == 0
    return masked_gather(points, idx), idx


# ---- make the reference's python importable on CPU -------------------------------------------------
def install_reference_stubs(reference_root="/root/reference"):
    """Register `pointnet2_ops._ext` and a minimal `pytorch3d` backed by this oracle, then put the
    reference's python trees on sys.path so its modules run unmodified on CPU tensors.
    Only usable where the reference tree exists (this container); never on the GPU box."""
    this = sys.modules[__name__]
    ext = types.ModuleType("pointnet2_ops._ext")
    for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn",
                 "three_interpolate", "three_interpolate_grad", "ball_query", "group_points",
                 "group_points_grad"):
        setattr(ext, name, getattr(this, name))
    sys.modules["pointnet2_ops._ext"] = ext

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
        return m

    knn = mod("pytorch3d.ops.knn", knn_points=knn_points, knn_gather=knn_gather)
    utils = mod("pytorch3d.ops.utils", masked_gather=masked_gather)
    ops = mod("pytorch3d.ops", knn=knn, utils=utils, knn_points=knn_points, knn_gather=knn_gather,
              sample_farthest_points=sample_farthest_points)
    pcl = mod("pytorch3d.structures.pointclouds", Pointclouds=type("Pointclouds", (), {}))
    structures = mod("pytorch3d.structures", pointclouds=pcl, Pointclouds=pcl.Pointclouds)
    loss = mod("pytorch3d.loss", chamfer_distance=None)
    mod("pytorch3d", ops=ops, structures=structures, loss=loss)
    for sub in ("pointnet2_ops_lib", "pointnet2"):
        p = os.path.join(reference_root, sub)
        if p not in sys.path:
            sys.path.insert(0, p)
    import pointnet2_ops  # noqa: F401  (the reference's package; its _ext is the stub above)
    pointnet2_ops._ext = ext
