"""The C oracle's remaining index ops against independent numpy / torch-autograd restatements (CPU only).

test_oracle_golden.py pins FPS / ball query / kNN / pytorch3d FPS; this file does the same for the ops whose strongest
pin (the reference's own .cu, oracle/_ref) only exists on a GPU box: ball query on ragged / empty balls, gather / group
and their gradients, three_nn (strict `<` tie rule, fewer than three known points), three_interpolate and its gradient,
kNN with per-cloud lengths, knn_gather / masked_gather padding.  Reference kernels:
pointnet2_ops_lib/pointnet2_ops/_ext-src/src/{ball_query,group_points,sampling,interpolate}_gpu.cu.
"""
import numpy as np
import pytest
import torch

from oracle import ops


def _rand(shape, seed, lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * (hi - lo) + lo


@pytest.mark.parametrize("radius,nsample", [(0.05, 8), (0.3, 4), (2.0, 16)])
def test_ball_query_first_hits_in_index_order(radius, nsample):
    """ball_query_gpu.cu:9-47: the first `nsample` points with d2 < r2 in index order; the first hit fills the tail;
    a ball without any hit keeps the zero-initialised row (and count 0, the patched wrapper's extra output)."""
    xyz = _rand((3, 97, 3), 5)
    xyz[1, 40:] = xyz[1, 7]          # a pile of duplicates: many equal distances
    q = _rand((3, 13, 3), 6)
    q[2, 0] = 5.0                    # far away: empty ball unless the radius is huge
    q[0, 1] = xyz[0, 3]              # a query that IS a point (d2 == 0)
    idx, cnt = ops.ball_query(q, xyz, radius, nsample)
    r2 = np.float32(radius) * np.float32(radius)
    for b in range(3):
        d = q[b].numpy()[:, None, :] - xyz[b].numpy()[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]).astype(np.float32)
        for j in range(13):
            margin = np.abs(d2[j] - r2) < 1e-6     # fp contraction order only matters on the sphere itself
            assert not margin.any()
            hits = np.nonzero(d2[j] < r2)[0]
            assert cnt[b, j] == min(len(hits), nsample)   # the scan stops at cnt == nsample (ball_query_gpu.cu:30,44)
            got = idx[b, j].numpy()
            if len(hits) == 0:
                assert (got == 0).all()
                continue
            k = min(len(hits), nsample)
            assert np.array_equal(got[:k], hits[:k])
            assert (got[k:] == hits[0]).all()


def test_gather_group_and_their_gradients_match_autograd():
    """group_points_gpu.cu:8-75, sampling_gpu.cu:8-57: plain indexing forward, scatter-add backward."""
    B, C, N, npnt, ns = 2, 5, 23, 7, 4
    pts = _rand((B, C, N), 7, -1, 1)
    g = torch.Generator().manual_seed(8)
    idx1 = torch.randint(0, N, (B, npnt), generator=g)
    idx1[0, :3] = 4                  # repeated indices: the gradient must accumulate
    idx2 = torch.randint(0, N, (B, npnt, ns), generator=g)
    idx2[1, 2] = 9
    p = pts.clone().requires_grad_(True)
    want1 = p.gather(2, idx1[:, None, :].expand(-1, C, -1))
    want2 = p[:, :, None, :].expand(-1, -1, npnt, -1).gather(3, idx2[:, None].expand(-1, C, -1, -1))
    i1, i2 = idx1.int(), idx2.int()   # `_ext` takes int32 indices (pointnet2_utils.py casts nothing: FPS / ball query emit int32)
    assert torch.equal(ops.gather_points(pts, i1), want1.detach())
    assert torch.equal(ops.group_points(pts, i2), want2.detach())
    go1, go2 = _rand((B, C, npnt), 9, -1, 1), _rand((B, C, npnt, ns), 10, -1, 1)
    (gw1,) = torch.autograd.grad(want1, p, go1, retain_graph=True)
    (gw2,) = torch.autograd.grad(want2, p, go2)
    assert torch.allclose(ops.gather_points_grad(go1, i1, N), gw1, atol=1e-6)
    assert torch.allclose(ops.group_points_grad(go2, i2, N), gw2, atol=1e-6)
    # untouched source points get exactly zero gradient
    untouched = torch.ones(B, N, dtype=torch.bool)
    untouched.scatter_(1, idx1, False)
    assert (ops.gather_points_grad(go1, i1, N).permute(0, 2, 1)[untouched] == 0).all()


def test_three_nn_order_ties_and_short_inputs():
    """interpolate_gpu.cu:9-59: ascending squared distances, strict `<` (the earlier index wins a tie), and with fewer
    than three known points the unfilled slots keep index 0 and (float)1e40 = inf."""
    known = _rand((2, 31, 3), 11)
    known[0, 20] = known[0, 4]       # exact duplicate
    unknown = _rand((2, 9, 3), 12)
    unknown[0, 0] = known[0, 4] + 1e-3
    d2, idx = ops.three_nn(unknown, known)
    for b in range(2):
        ref = ((unknown[b][:, None] - known[b][None]) ** 2).sum(-1).numpy()
        order = np.argsort(ref, axis=1, kind="stable")[:, :3]
        assert np.array_equal(idx[b].numpy(), order)
        assert np.allclose(d2[b].numpy(), np.take_along_axis(ref, order, 1), atol=1e-6)
    assert idx[0, 0, :2].tolist() == [4, 20] and d2[0, 0, 0] == d2[0, 0, 1]
    d2s, idxs = ops.three_nn(unknown, known[:, :2].contiguous())
    assert (idxs[:, :, 2] == 0).all() and torch.isinf(d2s[:, :, 2]).all()
    assert torch.isfinite(d2s[:, :, :2]).all()


def test_three_interpolate_and_gradient_match_autograd():
    """interpolate_gpu.cu:72-154: out[c, j] = sum_k points[c, idx[j, k]] * weight[j, k]; gradient scatter-adds."""
    B, C, m, n = 2, 6, 11, 17
    pts = _rand((B, C, m), 13, -1, 1)
    g = torch.Generator().manual_seed(14)
    idx = torch.randint(0, m, (B, n, 3), generator=g)
    idx[0, 0] = 5                    # all three neighbours the same point
    w = _rand((B, n, 3), 15)
    w = w / w.sum(-1, keepdim=True)
    p = pts.clone().requires_grad_(True)
    gathered = p[:, :, None, :].expand(-1, -1, n, -1).gather(3, idx[:, None].expand(-1, C, -1, -1))
    want = (gathered * w[:, None]).sum(-1)
    got = ops.three_interpolate(pts, idx.int(), w)
    assert torch.allclose(got, want.detach(), atol=1e-6)
    go = _rand((B, C, n), 16, -1, 1)
    (gw,) = torch.autograd.grad(want, p, go)
    assert torch.allclose(ops.three_interpolate_grad(go, idx.int(), w, m), gw, atol=1e-6)


def test_knn_with_lengths_and_gather_padding():
    """pytorch3d 0.7.0 knn_points semantics (published API; source absent -- parity unpinned): only the first lengths2[b]
    points of a cloud are candidates, queries beyond lengths1[b] and slots beyond lengths2[b] come back as 0 / 0.0, and
    knn_gather zeroes those slots."""
    p1, p2 = _rand((3, 6, 3), 17), _rand((3, 12, 3), 18)
    l1 = torch.tensor([6, 4, 1])
    l2 = torch.tensor([12, 5, 2])
    K = 4
    res = ops.knn_points(p1, p2, lengths1=l1, lengths2=l2, K=K, return_nn=True)
    for b in range(3):
        ref = ((p1[b][:, None] - p2[b][None, :l2[b]]) ** 2).sum(-1)
        k = min(K, int(l2[b]))
        srt, order = ref.sort(dim=1, stable=True)
        n1 = int(l1[b])
        assert torch.equal(res.idx[b, :n1, :k], order[:n1, :k])
        assert torch.allclose(res.dists[b, :n1, :k], srt[:n1, :k], atol=1e-6)
        assert (res.idx[b, :n1, k:] == 0).all() and (res.dists[b, :n1, k:] == 0).all()
        assert (res.idx[b, n1:] == 0).all() and (res.dists[b, n1:] == 0).all()
        assert (res.knn[b, :, k:] == 0).all()
        assert torch.equal(res.knn[b, :n1, :k], p2[b][order[:n1, :k]])


def test_masked_gather_and_ragged_p3d_fps():
    """pytorch3d masked_gather: -1 selects a zero row; sample_farthest_points pads clouds shorter than K with -1."""
    pts = _rand((2, 9, 3), 19)
    lengths = torch.tensor([9, 3])
    sel, idx = ops.sample_farthest_points(pts, lengths=lengths, K=5)
    assert idx.shape == (2, 5) and (idx[1, 3:] == -1).all() and (idx[1, :3] >= 0).all()
    assert sorted(idx[1, :3].tolist()) == [0, 1, 2] and idx[:, 0].tolist() == [0, 0]
    assert len(set(idx[0].tolist())) == 5
    assert (sel[1, 3:] == 0).all() and torch.equal(sel[0], pts[0][idx[0]])
    assert torch.equal(ops.masked_gather(pts, idx), sel)


def test_dropin_knn_backward_matches_autograd():
    """The drop-in knn_points attaches pytorch3d's gradient of `dists` w.r.t. p1 / p2 to the kernel's (graph-less) outputs
    (group_knn feeds d2 and 1 / (d2 + 1e-8) into the features, pointnet2_utils.py:506-517).  The backward is plain torch,
    so it is checked here on the CPU with the oracle's indices: against autograd through a differentiable restatement,
    with ragged lengths (padded neighbour slots and padded query rows carry no gradient)."""
    from slide_b200.dropin.pytorch3d.ops.knn import _KnnDists
    p1, p2 = _rand((3, 6, 3), 21, -1, 1), _rand((3, 12, 3), 22, -1, 1)
    go = _rand((3, 6, 4), 23, -1, 1)
    for l1, l2 in ((None, None), (torch.tensor([6, 4, 1]), torch.tensor([12, 5, 2]))):
        res = ops.knn_points(p1, p2, lengths1=l1, lengths2=l2, K=4)
        a, b = p1.clone().requires_grad_(True), p2.clone().requires_grad_(True)
        out = _KnnDists.apply(a, b, res.idx, res.dists, l2, l1)
        assert torch.equal(out, res.dists)              # forward values stay the kernel's
        ga, gb = torch.autograd.grad(out, (a, b), go)
        a2, b2 = p1.clone().requires_grad_(True), p2.clone().requires_grad_(True)
        nb = b2[:, :, None].expand(-1, -1, 4, -1).gather(1, res.idx[:, :, :, None].expand(-1, -1, -1, 3))
        d = ((a2[:, :, None, :] - nb) ** 2).sum(-1)
        valid = torch.ones(3, 6, 4)
        if l1 is not None:
            valid = valid * (torch.arange(4)[None, None, :] < l2[:, None, None]) * (torch.arange(6)[None, :, None] < l1[:, None, None])
        wa, wb = torch.autograd.grad(d, (a2, b2), go * valid)
        assert torch.allclose(ga, wa, atol=1e-6) and torch.allclose(gb, wb, atol=1e-6)
        assert torch.allclose(d.detach() * valid, res.dists, atol=1e-6)
