"""Device-side noise of the feature DDPM: the reference's `torch.randn_like` sequence (diffusion_utils/diffusion.py:88),
drawn as ONE seeked-Philox launch of this rank's slice (csrc/philox.cu) instead of T full-batch draws + T slice copies.

`randn_sequence` produces, bit for bit, what T consecutive `torch.randn(full_shape, device=cuda)` calls on the default CUDA
generator would put into rows [lo, lo + Bl) -- and leaves the generator where those T calls would have left it -- so results
do not depend on the world size and identical seeds give the reference's noise.  The mapping from (call, element) to
(Philox subsequence, offset, component) is ATen's (see philox.cu); it is verified against torch itself once per device and
shape, and if this torch build ever maps differently the draw goes through torch.randn (same values, T launches)."""
import ctypes

import torch

from . import lib as _l

_verified = {}


def _policy(numel, device):
    p = torch.cuda.get_device_properties(device)
    grid = min(p.multi_processor_count * (p.max_threads_per_multi_processor // 256), (numel + 255) // 256)
    inc = ((numel - 1) // (256 * grid * 4) + 1) * 4
    return grid, inc


def _launch(out2d, n_calls, reverse, seed, offset, numel, begin, length, device):
    grid, inc = _policy(numel, device)
    lib = _l.load()
    _l.check(lib.slide_philox_normal_slice(ctypes.c_void_p(out2d.data_ptr()), out2d.stride(0), n_calls, int(reverse),
                                           seed, offset, inc, numel, begin, length, grid, _l.stream_of(out2d)),
             "slide_philox_normal_slice")
    return inc


def _torch_path(out, full_shape, lo, reverse, device):
    T, Bl = out.shape[0], out.shape[1]
    full = torch.empty(full_shape, device=device)
    for s in range(T):
        torch.randn(full_shape, device=device, out=full)
        out[T - 1 - s if reverse else s].copy_(full[lo:lo + Bl])


def _verify(full_shape, device):
    """Two calls of torch.randn against the seeked draw of the same two calls (generator state restored afterwards)."""
    gen = torch.cuda.default_generators[device.index]
    state = gen.get_state()
    try:
        numel = 1
        for d in full_shape:
            numel *= d
        seed, off = gen.initial_seed(), gen.get_offset()
        want = torch.stack([torch.randn(full_shape, device=device) for _ in range(2)]).view(2, numel)
        moved = gen.get_offset() - off
        got = torch.empty(2, numel, device=device)
        inc = _launch(got, 2, False, seed, off, numel, 0, numel, device)
        return bool(torch.equal(want, got)) and moved == 2 * inc
    except Exception:  # noqa: BLE001  (e.g. a generator without offsets: use torch's own draws)
        return False
    finally:
        gen.set_state(state)


def randn_sequence(out, full_shape, lo, reverse=True):
    """out: contiguous fp32 CUDA tensor (T, Bl, *full_shape[1:]); fills out[row(s)] with rows [lo, lo+Bl) of the s-th of T
    torch.randn(full_shape) calls, row(s) = T-1-s if reverse (the loop runs t = T-1 .. 0) else s.  Returns "philox" or
    "torch" (which path produced the values)."""
    device = out.device
    T, Bl = out.shape[0], out.shape[1]
    assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape[2:]) == tuple(full_shape[1:])
    key = (device.index, tuple(full_shape))
    if key not in _verified:
        _verified[key] = _verify(tuple(full_shape), device)
    if not _verified[key]:
        _torch_path(out, tuple(full_shape), lo, reverse, device)
        return "torch"
    numel = 1
    for d in full_shape:
        numel *= d
    per = numel // full_shape[0]
    gen = torch.cuda.default_generators[device.index]
    seed, off = gen.initial_seed(), gen.get_offset()
    inc = _launch(out.view(T, Bl * per), T, reverse, seed, off, numel, lo * per, Bl * per, device)
    gen.set_offset(off + T * inc)
    return "philox"
