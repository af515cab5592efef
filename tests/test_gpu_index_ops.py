"""Index ops of the C ABI (drop-in pointnet2_ops._ext / pytorch3d entry points) against the C oracle on the GPU,
and -- when oracle/_ref holds the reference's own CUDA extension -- against the reference kernels themselves."""
import numpy as np
import pytest
import torch

import slide_b200
from oracle import ops as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    slide_b200.install_dropin()
    from pointnet2_ops import _ext
    return _ext


def _clouds(B, N, seed, scale=1.0, dup=False, origin=False):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, N, 3, generator=g) * 2 - 1) * scale
    if dup:  # duplicated points -> exact distance ties (tie rules)
        x[:, N // 2:] = x[:, :N - N // 2]
    if origin:  # points inside the |p|^2 <= 1e-3 ball, including index 0
        x[:, 0] = 0.001
        x[:, 5] = 0.01
        x[:, min(17, N - 1)] = -0.02
    return x.contiguous()


@pytest.mark.parametrize("N,m", [(2048, 1024), (256, 128), (1024, 256), (100, 37), (8192, 1024), (16, 16)])
@pytest.mark.parametrize("variant", ["plain", "dup", "origin", "small"])
def test_fps_bit_exact(ext, N, m, variant):
    x = _clouds(3, N, seed=N + m, scale=0.5 if variant == "small" else 1.0, dup=variant == "dup",
                origin=variant == "origin")
    want = O.furthest_point_sampling(x, m)
    got = ext.furthest_point_sampling(x.cuda(), m).cpu()
    assert torch.equal(got, want)


def test_config1_golden(ext, golden):
    xyz = torch.from_numpy(golden["c1_xyz"]).cuda()
    fps = ext.furthest_point_sampling(xyz, 1024)
    assert np.array_equal(fps.cpu().numpy(), golden["c1_fps"])
    new_xyz = xyz[0][fps[0].long()][None].contiguous()
    idx, cnt = ext.ball_query(new_xyz, xyz, 0.2, 32)
    assert np.array_equal(idx.cpu().numpy(), golden["c1_bq_idx"])
    assert np.array_equal(cnt.cpu().numpy(), golden["c1_bq_cnt"])


@pytest.mark.parametrize("N,m,r,ns", [(2048, 1024, 0.1, 32), (2048, 512, 0.2, 32), (500, 77, 0.05, 16), (64, 64, 10.0, 8)])
def test_ball_query_bit_exact(ext, N, m, r, ns):
    x = _clouds(2, N, seed=N)
    q = _clouds(2, m, seed=m + 1)  # queries that are NOT a subset: rows without any hit stay all-zero
    wi, wc = O.ball_query(q, x, r, ns)
    gi, gc = ext.ball_query(q.cuda(), x.cuda(), r, ns)
    assert torch.equal(gi.cpu(), wi) and torch.equal(gc.cpu(), wc)


def test_gather_group_interpolate(ext):
    g = torch.Generator().manual_seed(9)
    pts = torch.randn(2, 37, 300, generator=g)
    idx = torch.randint(0, 300, (2, 50), generator=g, dtype=torch.int32)
    assert torch.equal(ext.gather_points(pts.cuda(), idx.cuda()).cpu(), O.gather_points(pts, idx))
    gidx = torch.randint(0, 300, (2, 20, 8), generator=g, dtype=torch.int32)
    assert torch.equal(ext.group_points(pts.cuda(), gidx.cuda()).cpu(), O.group_points(pts, gidx))
    unknown, known = torch.rand(2, 90, 3, generator=g), torch.rand(2, 40, 3, generator=g)
    wd, wi = O.three_nn(unknown, known)
    gd, gi = ext.three_nn(unknown.cuda(), known.cuda())
    assert torch.equal(gi.cpu(), wi) and torch.equal(gd.cpu(), wd)
    w = torch.rand(2, 90, 3, generator=g)
    feats = torch.randn(2, 11, 40, generator=g)
    assert torch.equal(ext.three_interpolate(feats.cuda(), wi.cuda(), w.cuda()).cpu(), O.three_interpolate(feats, wi, w))
    go = torch.randn(2, 37, 50, generator=g)
    assert torch.allclose(ext.gather_points_grad(go.cuda(), idx.cuda(), 300).cpu(), O.gather_points_grad(go, idx, 300),
                          atol=1e-5)


def test_cpu_tensors_are_rejected(ext):
    with pytest.raises(RuntimeError):
        ext.furthest_point_sampling(torch.rand(1, 16, 3), 4)


@pytest.mark.parametrize("P1,P2,K", [(16, 16, 16), (16, 16, 8), (1024, 256, 16), (1024, 2048, 32), (256, 16, 4)])
def test_knn_points(P1, P2, K):
    slide_b200.install_dropin()
    from pytorch3d.ops.knn import knn_points, knn_gather
    g = torch.Generator().manual_seed(P1 + P2)
    a, b = torch.rand(2, P1, 3, generator=g), torch.rand(2, P2, 3, generator=g)
    b[:, 3] = b[:, 1]
    want = O.knn_points(a, b, K=K, return_nn=True)
    got = knn_points(a.cuda(), b.cuda(), K=K, return_nn=True)
    assert torch.equal(got.idx.cpu(), want.idx) and torch.equal(got.dists.cpu(), want.dists)
    assert torch.equal(got.knn.cpu(), want.knn)
    assert torch.equal(knn_gather(b.cuda(), got.idx).cpu(), O.knn_gather(b, want.idx))


def test_sample_farthest_points_p3d():
    slide_b200.install_dropin()
    from pytorch3d.ops import sample_farthest_points
    from pytorch3d.ops.utils import masked_gather
    g = torch.Generator().manual_seed(21)
    x = torch.rand(3, 512, 6, generator=g)
    torch.manual_seed(77)
    _, want = O.sample_farthest_points(x[:, :, :3].contiguous(), K=256, random_start_point=True)
    torch.manual_seed(77)
    pts, got = sample_farthest_points(x[:, :, :3].cuda(), K=256, random_start_point=True)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(masked_gather(x.cuda(), got).cpu(), O.masked_gather(x, want))


def test_against_reference_cuda_extension(ext):
    """oracle/_ref = the reference's own .cu files compiled for sm_100a (oracle/build_ref.py).  Pins the oracle AND
    the new kernels to the reference's real kernels, tie rules included."""
    from oracle import build_ref
    ref = build_ref.load_module()
    if ref is None:
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    for N, m, variant in [(2048, 1024, "plain"), (2048, 512, "dup"), (1024, 256, "origin"), (300, 100, "dup")]:
        x = _clouds(2, N, seed=N + 3, dup=variant == "dup", origin=variant == "origin").cuda()
        r = ref.furthest_point_sampling(x, m)
        assert torch.equal(ext.furthest_point_sampling(x, m), r), (N, m, variant)
        assert torch.equal(O.furthest_point_sampling(x.cpu(), m), r.cpu()), (N, m, variant)
        q = x[:, :m].contiguous()
        ri = ref.ball_query(q, x, 0.15, 32)
        gi = ext.ball_query(q, x, 0.15, 32)
        assert torch.equal(gi[0], ri[0]) and torch.equal(gi[1], ri[1])
        rd = ref.three_nn(q, x)
        gd = ext.three_nn(q, x)
        assert torch.equal(gd[0], rd[0]) and torch.equal(gd[1], rd[1])


def test_sample_farthest_points_p3d_golden():
    """slide_sample_farthest_points against the vendored pytorch3d reference implementation's golden vectors
    (tests/golden/make_golden_fps.py): ragged lengths with per-cloud K, duplicated points, clouds shorter than K."""
    import os
    import numpy as np
    slide_b200.install_dropin()
    from pytorch3d.ops import sample_farthest_points
    gold = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_fps.npz")))
    for name in ("full", "ragged", "dup", "short", "decode"):
        p = torch.from_numpy(gold[name + "_points"]).cuda()
        lengths = torch.from_numpy(gold[name + "_lengths"]).cuda()
        _, idx = sample_farthest_points(p, lengths, [int(k) for k in gold[name + "_K"]])
        assert np.array_equal(idx.cpu().numpy(), gold[name + "_idx"]), name
