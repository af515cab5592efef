"""Drop-in for the reference's pybind module `pointnet2_ops._ext`
(pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19): the same nine functions, the same
argument order, the same allocation rules (callee allocates zero-filled outputs on the input's device)
and the same failure mode (RuntimeError for non-contiguous / wrong dtype / CPU tensors,
_ext-src/include/utils.h:5-25, sampling.cpp:34), implemented as ctypes calls into libslide_b200.so.
"""
import ctypes

import torch

from slide_b200 import lib as _l


def _chk(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _contig(t, name):
    _chk(t.is_contiguous(), name + " must be a contiguous tensor")


def _float(t, name):
    _chk(t.dtype == torch.float32, name + " must be a float tensor")


def _int(t, name):
    _chk(t.dtype == torch.int32, name + " must be an int tensor")


def _cuda(t, name="tensor"):
    _chk(t.is_cuda, "CPU not supported" if name == "tensor" else name + " must be a CUDA tensor")


def gather_points(points, idx):
    _contig(points, "points"); _contig(idx, "idx"); _float(points, "points"); _int(idx, "idx")
    _cuda(points); _cuda(idx, "idx")
    B, C, N = points.shape
    m = idx.shape[1]
    out = torch.zeros(B, C, m, device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        _l.check(_l.load().slide_gather_points(_l.ptr(points), _l.ptr(idx), B, C, N, m, _l.ptr(out),
                                               _l.stream_of(points)), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _float(grad_out, "grad_out"); _int(idx, "idx")
    _cuda(grad_out); _cuda(idx, "idx")
    B, C, m = grad_out.shape
    out = torch.zeros(B, C, n, device=grad_out.device, dtype=torch.float32)
    with torch.cuda.device(grad_out.device):
        _l.check(_l.load().slide_gather_points_grad(_l.ptr(grad_out), _l.ptr(idx), B, C, int(n), m, _l.ptr(out),
                                                    _l.stream_of(grad_out)), "gather_points_grad")
    return out


def furthest_point_sampling(points, nsamples):
    _contig(points, "points"); _float(points, "points"); _cuda(points)
    B, N, _ = points.shape
    out = torch.zeros(B, nsamples, device=points.device, dtype=torch.int32)
    lib = _l.load()
    # clouds beyond the register-resident kernel's capacity keep their running distances in a scratch tensor, like the
    # reference's `tmp` (sampling.cpp:74-76)
    tmp = torch.empty(B, N, device=points.device) if N > lib.slide_fps_resident_max_points() else None
    with torch.cuda.device(points.device):
        _l.check(lib.slide_furthest_point_sampling_ws(_l.ptr(points), B, N, int(nsamples), _l.ptr(out), _l.ptr(tmp),
                                                      _l.stream_of(points)), "furthest_point_sampling")
    return out


def three_nn(unknowns, knows):
    _contig(unknowns, "unknowns"); _contig(knows, "knows"); _float(unknowns, "unknowns"); _float(knows, "knows")
    _cuda(unknowns); _cuda(knows, "knows")
    B, n, _ = unknowns.shape
    m = knows.shape[1]
    idx = torch.zeros(B, n, 3, device=unknowns.device, dtype=torch.int32)
    dist2 = torch.zeros(B, n, 3, device=unknowns.device, dtype=torch.float32)
    with torch.cuda.device(unknowns.device):
        _l.check(_l.load().slide_three_nn(_l.ptr(unknowns), _l.ptr(knows), B, n, m, _l.ptr(dist2), _l.ptr(idx),
                                          _l.stream_of(unknowns)), "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    _contig(points, "points"); _contig(idx, "idx"); _contig(weight, "weight")
    _float(points, "points"); _int(idx, "idx"); _float(weight, "weight")
    _cuda(points); _cuda(idx, "idx"); _cuda(weight, "weight")
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(B, C, n, device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        _l.check(_l.load().slide_three_interpolate(_l.ptr(points), _l.ptr(idx), _l.ptr(weight), B, C, m, n,
                                                   _l.ptr(out), _l.stream_of(points)), "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _contig(weight, "weight")
    _float(grad_out, "grad_out"); _int(idx, "idx"); _float(weight, "weight")
    _cuda(grad_out); _cuda(idx, "idx"); _cuda(weight, "weight")
    B, C, n = grad_out.shape
    out = torch.zeros(B, C, m, device=grad_out.device, dtype=torch.float32)
    with torch.cuda.device(grad_out.device):
        _l.check(_l.load().slide_three_interpolate_grad(_l.ptr(grad_out), _l.ptr(idx), _l.ptr(weight), B, C, n,
                                                        int(m), _l.ptr(out), _l.stream_of(grad_out)),
                 "three_interpolate_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    _contig(new_xyz, "new_xyz"); _contig(xyz, "xyz"); _float(new_xyz, "new_xyz"); _float(xyz, "xyz")
    _cuda(new_xyz); _cuda(xyz, "xyz")
    B, m, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = torch.zeros(B, m, nsample, device=new_xyz.device, dtype=torch.int32)
    counts = torch.zeros(B, m, device=new_xyz.device, dtype=torch.int32)
    with torch.cuda.device(new_xyz.device):
        _l.check(_l.load().slide_ball_query(_l.ptr(new_xyz), _l.ptr(xyz), B, N, m, ctypes.c_float(radius),
                                            int(nsample), _l.ptr(idx), _l.ptr(counts), _l.stream_of(new_xyz)),
                 "ball_query")
    return idx, counts


def group_points(points, idx):
    _contig(points, "points"); _contig(idx, "idx"); _float(points, "points"); _int(idx, "idx")
    _cuda(points); _cuda(idx, "idx")
    B, C, N = points.shape
    _, npoint, ns = idx.shape
    out = torch.zeros(B, C, npoint, ns, device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        _l.check(_l.load().slide_group_points(_l.ptr(points), _l.ptr(idx), B, C, N, npoint, ns, _l.ptr(out),
                                              _l.stream_of(points)), "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _float(grad_out, "grad_out"); _int(idx, "idx")
    _cuda(grad_out); _cuda(idx, "idx")
    B, C, npoint, ns = grad_out.shape
    out = torch.zeros(B, C, n, device=grad_out.device, dtype=torch.float32)
    with torch.cuda.device(grad_out.device):
        _l.check(_l.load().slide_group_points_grad(_l.ptr(grad_out), _l.ptr(idx), B, C, int(n), npoint, ns,
                                                   _l.ptr(out), _l.stream_of(grad_out)), "group_points_grad")
    return out
