"""Golden vectors of the reference's FastDPM samplers (pointnet2/util_fastdpmv2.py: fast_sampling_function_v2 ->
STEP_sampling with get_STEP_step) from the REAL reference (build container only).

    python tests/golden/make_golden_fast.py      ->  tests/golden/golden_fast.npz

The reference calls .cuda() and draws std_normal itself; here Tensor.cuda is patched to the identity and std_normal to
replay pinned noise (x_T first, then one draw per iteration).  The denoiser is the closed-form stand-in of
make_golden_sampler.py.  Also asserts that oracle/ref_model.fast_sampling is bit-identical."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops, ref_model  # noqa: E402

ops.install_reference_stubs()
import util_fastdpmv2 as ref_fast  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DCFG = {"T": 1000, "beta_0": 0.0001, "beta_T": 0.02}
LENGTH = 10


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    T = DCFG["T"]
    model = lambda x, ts=None, label=None: 0.5 * torch.tanh(x) + 0.01 * (ts / T).reshape(-1, 1, 1)
    dh = ref_fast.calc_diffusion_hyperparams(**DCFG)
    g = torch.Generator().manual_seed(77)
    B, N, C = 3, 16, 3
    draws = [torch.randn(B, N, C, generator=g) for _ in range(LENGTH + 1)]
    gold = {"draws": torch.stack(draws).numpy(), "length": np.int64(LENGTH)}
    # the reference's VAR sampler trips its own assertion with the shipped schedule: recorded, not restated
    for schedule in ("linear", "quadratic"):
        it = iter(draws)
        ref_fast.std_normal = lambda size: next(it).clone()
        try:
            ref_fast.fast_sampling_function_v2(model, (B, N, C), dh, DCFG, length=LENGTH, sampling_method="var",
                                               schedule=schedule, kappa=0.5, label=None, verbose=False)
            raise SystemExit("VAR_sampling ran: restate it")
        except AssertionError:
            pass
    for method in ("step",):
        for schedule in ("linear", "quadratic"):
            for kappa in (0.0, 0.5, 1.0):
                it = iter(draws)
                ref_fast.std_normal = lambda size: next(it).clone()
                import io
                import contextlib
                with contextlib.redirect_stdout(io.StringIO()):
                    x = ref_fast.fast_sampling_function_v2(model, (B, N, C), dh, DCFG, length=LENGTH,
                                                           sampling_method=method, schedule=schedule, kappa=kappa,
                                                           label=None, verbose=False)
                mine, taus = ref_model.fast_sampling(lambda xx, tt: model(xx, ts=tt), draws[0],
                                                     draws[1:], DCFG, method, LENGTH, schedule, kappa)
                assert torch.equal(x, mine), (method, schedule, kappa)
                key = "%s_%s_%g" % (method, schedule, kappa)
                gold["out_" + key] = x.numpy()
                gold["taus_" + key] = np.asarray(taus, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "golden_fast.npz"), **gold)
    print("wrote golden_fast.npz with", len(gold), "arrays")


if __name__ == "__main__":
    main()
