"""The seeked-Philox slice draw (slide_philox_normal_slice, slide_b200/rng.py) against torch itself: the feature DDPM's
per-step noise must be, bit for bit, the reference's `torch.randn_like` sequence (diffusion_utils/diffusion.py:88) on the
full batch, whatever slice of the batch a rank owns, and must leave the CUDA generator where torch would."""
import pytest
import torch

from slide_b200 import rng

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,lo,Bl,T", [(8, 0, 8, 5), (256, 64, 32, 7), (256, 0, 256, 3), (2048, 1792, 256, 4),
                                       (2048, 0, 2048, 2)])
def test_slice_draw_is_torch_randn(B, lo, Bl, T):
    """B = 2048 (the 8-GPU weak-scaling job) needs two grid-stride iterations per thread in ATen's kernel."""
    dev = torch.device("cuda", 0)
    shape = (B, 16, 51)
    torch.manual_seed(1234 + B)
    torch.randn(7, device=dev)  # a non-zero starting offset
    gen = torch.cuda.default_generators[0]
    state = gen.get_state()
    want = torch.stack([torch.randn(shape, device=dev)[lo:lo + Bl] for _ in range(T)])
    end_offset = gen.get_offset()
    follow = torch.randn(5, device=dev)
    gen.set_state(state)
    got = torch.empty(T, Bl, 16, 51, device=dev)
    path = rng.randn_sequence(got, shape, lo, reverse=True)
    assert path == "philox", "ATen's Philox mapping changed: the torch.randn path was taken"
    assert gen.get_offset() == end_offset
    assert torch.equal(got.flip(0), want)
    assert torch.equal(torch.randn(5, device=dev), follow)  # the stream continues exactly where torch's would
    gen.set_state(state)
    got2 = torch.empty(T, Bl, 16, 51, device=dev)
    rng.randn_sequence(got2, shape, lo, reverse=False)
    assert torch.equal(got2, want)


def test_pipeline_noise_independent_of_world_size(pipeline_cfg):
    """Rank r of W draws exactly rows [r B/W, (r+1) B/W) of the single-GPU draw."""
    dev = torch.device("cuda", 0)
    shape = (64, 16, 51)
    torch.manual_seed(5)
    full = torch.empty(6, 64, 16, 51, device=dev)
    gen = torch.cuda.default_generators[0]
    state = gen.get_state()
    rng.randn_sequence(full, shape, 0)
    for W in (2, 4, 8):
        Bl = 64 // W
        for r in range(W):
            gen.set_state(state)
            part = torch.empty(6, Bl, 16, 51, device=dev)
            rng.randn_sequence(part, shape, r * Bl)
            assert torch.equal(part, full[:, r * Bl:(r + 1) * Bl])
