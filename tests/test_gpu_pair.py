"""SLIDE_OP_PAIR (the grouped 1x1 conv factored through the gather) on shapes the networks do not hit: odd column counts
(scalar tail columns), point counts that the points-per-CTA block does not divide, gather sources of 16 / 40 / 64 rows,
with and without the interpolation-weight terms (group_knn, pointnet2_utils.py:497-540) and the transformed residual.
The shared-memory kernel (pair_smem_kernel) must equal the gather kernel (pair_kernel) BIT FOR BIT -- same fused
multiply-add chains -- and both must match the CPU interpreter of the records (oracle/ir_exec.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ir_exec
from slide_b200 import lib
from slide_b200.program import XF, Builder, Program

pytestmark = pytest.mark.gpu

CASES = [  # B, nsrc, np, K, N, d2, res
    (3, 16, 16, 16, 32, False, False),
    (5, 16, 16, 16, 75, True, False),
    (40, 16, 16, 16, 265, False, True),
    (7, 40, 10, 8, 523, True, True),
    (2, 64, 16, 5, 139, True, False),
    (300, 16, 16, 8, 64, False, True),
]


def build(B, nsrc, npnt, K, N, d2, res, seed):
    g = np.random.default_rng(seed)
    b = Builder(B)
    U = b.tensor("U", nsrc, N)
    xyz = b.tensor("xyz", nsrc, 3)
    ctr = b.tensor("ctr", npnt, 3)
    idx = b.tensor("idx", npnt, K, dtype="i32")
    dd = b.tensor("d2", npnt, K) if d2 else None
    out = b.tensor("out", npnt * K, N)
    cg = 1
    for c in (16, 8, 4, 2):
        if N % c == 0:
            cg = c
            break
    nnorm = N // cg * cg
    st = b.stats("st", nnorm, cg, npnt * K, npnt * K * cg)
    resid, xfr = None, XF()
    if res:
        resid = b.tensor("res", npnt * K, N)
        rst = b.stats("rst", nnorm, cg, npnt * K, npnt * K * cg)
        addv = b.tensor("addv", 1, N)
        xfr = XF(stats=rst.tensor, cg=cg, nnorm=nnorm, choff=0, gamma=b.weight(1 + 0.1 * g.standard_normal(N)),
                 beta=b.weight(0.1 * g.standard_normal(N)), R=npnt * K, count=npnt * K * cg, relu=True, addvec=addv, addmode=0)
    wx, wc = b.weight(g.standard_normal((N, 3))), b.weight(g.standard_normal((N, 3)))
    bias = b.weight(g.standard_normal(N))
    wd = b.weight(g.standard_normal(N)) if d2 else -1
    ww = b.weight(g.standard_normal(N)) if d2 else -1
    b.begin_segment("pair")
    b.step_begin()
    b.pair(U, xyz, ctr, idx, K, out, wx, wc, bias=bias, d2=dd, wd=wd, ww=ww, act="relu", resid=resid, xfr=xfr, stats=st,
           note="pair")
    b.end_segment()
    vals = {U: g.standard_normal((B * nsrc, N)), xyz: g.random((B * nsrc, 3)) - 0.5, ctr: g.random((B * npnt, 3)) - 0.5,
            idx: g.integers(0, nsrc, (B * npnt, K)).astype(np.int32)}
    if d2:
        vals[dd] = g.random((B * npnt, K)) * 0.3
    if res:
        vals[resid] = g.standard_normal((B * npnt * K, N))
        vals[addv] = g.standard_normal((B, N))
        r = vals[resid].reshape(B, npnt * K, N)[:, :, :nnorm].reshape(B, npnt * K, nnorm // cg, cg)
        vals[rst.tensor] = np.stack([r.sum(axis=(1, 3)), (r * r).sum(axis=(1, 3))], axis=2).reshape(B, -1)
    return b, vals, out, st


@pytest.mark.parametrize("case", CASES)
def test_pair_kernels_agree_and_match_interpreter(case):
    B, nsrc, npnt, K, N, d2, res = case
    b, vals, out, st = build(*case, seed=sum(case[:5]))
    m = ir_exec.Machine(b)
    for t, v in vals.items():
        m.upload(t, v.astype(np.float64 if t.dtype == "f64" else (np.int32 if t.dtype == "i32" else np.float32)))
    m.set_step(1)
    m.run(*b.segments["pair"])
    want = np.asarray(m.download(out))
    want_st = np.asarray(m.download(st.tensor))
    got = {}
    L = lib.load()
    try:
        for mode in ("1", "0"):
            os.environ["SLIDE_PAIR_SMEM"] = mode
            L.slide_tc_reload_tuning()
            prog = Program(b)
            for t, v in vals.items():
                tv = torch.from_numpy(v.astype(np.float64 if t.dtype == "f64" else (np.int32 if t.dtype == "i32" else np.float32)))
                prog.upload(t, tv)
            prog.set_step(1)
            prog.run(*b.segments["pair"])
            torch.cuda.synchronize()
            got[mode] = (prog.download(out).cpu().numpy(), prog.download(st.tensor).cpu().numpy())
            prog.close()
    finally:
        os.environ.pop("SLIDE_PAIR_SMEM", None)
        L.slide_tc_reload_tuning()
    assert np.array_equal(got["1"][0], got["0"][0]), "shared-memory PAIR kernel deviates from the gather kernel"
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got["1"][0].reshape(want.shape) - want).max() <= 2e-5 * scale
    sscale = max(1.0, float(np.abs(want_st).max()))
    for mode in ("1", "0"):
        assert np.abs(got[mode][1].reshape(want_st.shape) - want_st).max() <= 1e-5 * sscale, mode
