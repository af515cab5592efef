"""Golden vectors of the feature-DDPM update under the reference's OTHER diffusion schedules (no shipped config selects
them): get_beta_schedule 'quad' / 'const' / 'jsd' besides 'linear', model_var_type 'fixedlarge' besides 'fixedsmall',
other step counts and beta ranges -- from the REAL Diffusion class and denoising_step
(pointnet2/diffusion_utils/diffusion.py:12-28,58-95,158-208; build container only: needs /root/reference).

    python tests/golden/make_golden_schedules.py      ->  tests/golden/golden_schedules.npz

Stand-in denoiser and pinned noise as in make_golden_sampler.py.  'warmup10' / 'warmup50' call a helper the reference
never defines (NameError in the reference itself): recorded, not restated."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops  # noqa: E402

ops.install_reference_stubs()
from diffusion_utils import diffusion as ref_diffusion  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [
    dict(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000, model_var_type="fixedlarge"),
    dict(beta_schedule="quad", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000, model_var_type="fixedsmall"),
    dict(beta_schedule="quad", beta_start=1e-5, beta_end=0.05, num_diffusion_timesteps=200, model_var_type="fixedlarge"),
    dict(beta_schedule="const", beta_start=1e-4, beta_end=0.01, num_diffusion_timesteps=500, model_var_type="fixedsmall"),
    dict(beta_schedule="jsd", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=100, model_var_type="fixedsmall"),
    dict(beta_schedule="linear", beta_start=2e-4, beta_end=0.03, num_diffusion_timesteps=2000, model_var_type="fixedsmall"),
]


def main():
    for name in ("warmup10", "warmup50"):
        try:
            ref_diffusion.get_beta_schedule(name, beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=10)
            raise SystemExit("%s works in the reference: restate it" % name)
        except NameError:
            pass
    g = torch.Generator().manual_seed(11)
    B, N, C = 2, 16, 51
    x0 = torch.randn(B, N, C, generator=g)
    kp = torch.rand(B, N, 3, generator=g) - 0.5
    gold = {"x": x0.numpy(), "keypoint": kp.numpy()}
    for ci, case in enumerate(CASES):
        cfg = dict(case, data_clamp_range=-1, model_output_scale_factor=1.0, loss_type="mse")
        with contextlib.redirect_stdout(io.StringIO()):
            D = ref_diffusion.Diffusion(cfg, device=torch.device("cpu"))
        T = D.num_timesteps
        model = lambda x, ts=None, label=None: 0.5 * torch.tanh(x) + 0.01 * (ts / T).reshape(-1, 1, 1)
        t_list = [T - 1, T // 2, 1, 0]
        noise = torch.randn(len(t_list), B, N, C, generator=g)
        gold["c%d_noise" % ci] = noise.numpy()
        orig = torch.randn_like
        for j, t in enumerate(t_list):   # one loop body of denoise_and_reconstruct per step, each from the same x
            torch.randn_like = lambda like, _j=j: noise[_j]
            x = torch.cat([kp, x0[:, :, 3:]], dim=2)
            with np.errstate(all="ignore"):
                y, _ = ref_diffusion.denoising_step(
                    x, t=torch.ones(B) * t, model=model, logvar=D.logvar,
                    sqrt_recip_alphas_cumprod=D.sqrt_recip_alphas_cumprod,
                    sqrt_recipm1_alphas_cumprod=D.sqrt_recipm1_alphas_cumprod,
                    posterior_mean_coef1=D.posterior_mean_coef1, posterior_mean_coef2=D.posterior_mean_coef2,
                    return_pred_xstart=True, label=None, data_clamp_range=D.data_clamp_range)
            # the loop re-imposes the keypoint columns after every step (diffusion.py:383-385,394-396)
            gold["c%d_out_t%d" % (ci, t)] = torch.cat([kp, y[:, :, 3:]], dim=2).numpy()
        torch.randn_like = orig
        print(ci, case, "finite outputs:", [bool(np.isfinite(gold["c%d_out_t%d" % (ci, t)]).all()) for t in t_list])
    gold["cases_json"] = np.array(json.dumps(CASES))
    path = os.path.join(OUT, "golden_schedules.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
