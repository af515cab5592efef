"""Golden vectors for the pytorch3d-semantics farthest point sampling from the copy of pytorch3d's reference
implementation that the SLIDE tree vendors (pointnet2/data_utils/points_sampling.py::sample_farthest_points_naive,
:13-118; build container only).  pytorch3d itself is absent, so this is the strongest pin available for
`slide_sample_farthest_points` / the oracle's `sample_farthest_points`.

    python tests/golden/make_golden_fps.py      ->  tests/golden/golden_fps.npz

Cases: full clouds, ragged lengths with per-cloud K (slots beyond min(K_b, length_b) stay -1), duplicated points
(arg-max = first maximum), and a cloud shorter than K.  Start index 0 (random_start_point=False: the vendored copy
draws with python's `random`, pytorch3d 0.7.0 with torch.randint -- the start index is an input of the C ABI)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops  # noqa: E402

ops.install_reference_stubs()
from data_utils.points_sampling import sample_farthest_points_naive  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(5)
    cases = {}
    pts = torch.rand(3, 257, 3, generator=g) - 0.5
    cases["full"] = (pts, None, 16)
    cases["ragged"] = (torch.rand(4, 300, 3, generator=g) - 0.5, torch.tensor([300, 120, 17, 5]), [32, 16, 17, 8])
    dup = torch.rand(2, 64, 3, generator=g) - 0.5
    dup[:, 32:] = dup[:, :32]  # every point twice: distance ties
    cases["dup"] = (dup, None, 24)
    cases["short"] = (torch.rand(2, 10, 3, generator=g), torch.tensor([10, 3]), 16)
    cases["decode"] = (torch.rand(2, 2048, 3, generator=g) - 0.5, None, 1024)  # level-2 shape of the decoder
    gold = {}
    for name, (p, lengths, K) in cases.items():
        _, idx = sample_farthest_points_naive(p, lengths, K, random_start_point=False)
        _, mine = ops.sample_farthest_points(p, lengths, K)
        assert torch.equal(idx, mine), "oracle deviates from the vendored pytorch3d reference on case %s" % name
        gold[name + "_points"] = p.numpy()
        gold[name + "_lengths"] = (lengths if lengths is not None else torch.full((p.shape[0],), p.shape[1])).numpy()
        gold[name + "_K"] = np.asarray(K if isinstance(K, list) else [K] * p.shape[0], dtype=np.int64)
        gold[name + "_idx"] = idx.numpy()
    np.savez_compressed(os.path.join(OUT, "golden_fps.npz"), **gold)
    print("wrote golden_fps.npz:", sorted(cases))


if __name__ == "__main__":
    main()
