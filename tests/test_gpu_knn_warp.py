"""The warp-per-query kNN kernel of the program executor (knn_warp_kernel: large reference clouds, K > 8) against the C
oracle's pytorch3d-semantics kNN: indices and squared distances bit-exact, including clouds full of ties (duplicated
points, points on a lattice) where only the insertion order -- ascending index, strict `<` -- decides the result."""
import numpy as np
import pytest
import torch

from oracle import ir_exec
from slide_b200.program import Builder, Program

pytestmark = pytest.mark.gpu


def _case(B, P1, P2, K, kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "random":
        ref = torch.rand(B, P2, 3, generator=g) - 0.5
    elif kind == "duplicates":  # every point appears ~4 times
        base = torch.rand(B, max(P2 // 4, 1), 3, generator=g) - 0.5
        ref = base[:, torch.randint(0, base.shape[1], (P2,), generator=g)]
    else:  # lattice: many exactly equal distances
        ref = torch.randint(0, 6, (B, P2, 3), generator=g).float() * 0.125
    q = ref[:, torch.randint(0, P2, (P1,), generator=g)] if kind != "random" else torch.rand(B, P1, 3, generator=g) - 0.5
    return q.contiguous(), ref.contiguous()


@pytest.mark.parametrize("B,P1,P2,K", [(3, 32, 128, 9), (2, 100, 1000, 16), (2, 1024, 4096, 32), (1, 37, 2500, 20),
                                       (2, 256, 1024, 32)])
@pytest.mark.parametrize("kind", ["random", "duplicates", "lattice"])
def test_knn_warp_matches_oracle(B, P1, P2, K, kind):
    b = Builder(B)
    q = b.tensor("q", P1, 3)
    ref = b.tensor("ref", P2, 3)
    idx = b.tensor("idx", P1, K, dtype="i32")
    d2 = b.tensor("d2", P1, K)
    b.begin_segment("knn")
    b.knn(q, ref, K, idx, d2=d2, note="knn")
    b.end_segment()
    qv, rv = _case(B, P1, P2, K, kind, seed=P1 * 7 + K)
    m = ir_exec.Machine(b)
    prog = Program(b)
    for mach in (m, prog):
        mach.upload(q, qv.reshape(-1, 3))
        mach.upload(ref, rv.reshape(-1, 3))
    m.run(*b.segments["knn"])
    prog.run_segment("knn")
    torch.cuda.synchronize()
    want_i, want_d = np.asarray(m.download(idx)), np.asarray(m.download(d2))
    got_i, got_d = prog.download(idx).cpu().numpy(), prog.download(d2).cpu().numpy()
    assert np.array_equal(got_i, want_i), float((got_i != want_i).mean())
    assert np.array_equal(got_d, want_d)
