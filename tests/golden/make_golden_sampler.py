"""Golden vectors of the reference's samplers from the REAL reference (build container only: needs /root/reference):
the feature-DDPM update, with and without local resampling (pointnet2/diffusion_utils/diffusion.py), and the position
DDPM's full 1000-step chain (pointnet2/util.py::sampling).

    python tests/golden/make_golden_sampler.py      ->  tests/golden/golden_sampler.npz

The denoiser is replaced by a fixed closed-form stand-in (eps = 0.5 * tanh(x) + 0.01 * t / T) so that the vectors pin
exactly the part of the path the networks' golden vectors (make_golden.py) do not: Diffusion's float64 schedule,
`extract`, `denoising_step` (clamp off as in the shipped configs, the local-resampling blend, the t == 0 mask) and the
keypoint overwrite of LatentDiffusion.denoise_and_reconstruct's loop.  torch.randn_like is patched to replay pinned
noise.  Also asserts that oracle/ref_model.py's restatement is bit-identical."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops, ref_model  # noqa: E402

ops.install_reference_stubs()
from diffusion_utils import diffusion as ref_diffusion  # noqa: E402
from slide_b200 import weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def stand_in(T):
    return lambda x, ts=None, label=None: 0.5 * torch.tanh(x) + 0.01 * (ts / T).reshape(-1, 1, 1)


def main():
    cfg = weights.load_json("pipeline_airplane.json")["latent_ddpm"]["standard_diffusion_config"]
    D = ref_diffusion.Diffusion(cfg, device=torch.device("cpu"))
    T = D.num_timesteps
    sch = ref_model.latent_schedule(cfg)
    g = torch.Generator().manual_seed(2024)
    B, N, C = 3, 16, 51
    x_T = torch.randn(B, N, C, generator=g)
    kp = torch.rand(B, N, 3, generator=g) - 0.5
    complete_x0 = torch.cat([kp, torch.randn(B, N, C - 3, generator=g)], dim=2)
    mask = (torch.rand(B, N, generator=g) < 0.5).float()
    gold = {"x_T": x_T.numpy(), "keypoint": kp.numpy(), "complete_x0": complete_x0.numpy(), "mask": mask.numpy()}
    model = stand_in(T)
    for tag, t_list in (("top", [T - 1, T - 2, T - 3]), ("bottom", [2, 1, 0])):
        noises = {t: torch.randn(B, N, C, generator=g) for t in t_list}
        gold["noise_" + tag] = np.stack([noises[t].numpy() for t in t_list])
        for local in (False, True):
            x = x_T.clone()
            orig = torch.randn_like
            for t in t_list:  # the loop body of denoise_and_reconstruct (diffusion.py:381-392)
                torch.randn_like = lambda like, _t=t: noises[_t]
                ts = torch.ones(B) * t
                x = torch.cat([kp, x[:, :, 3:]], dim=2)
                x, _x0 = ref_diffusion.denoising_step(
                    x, t=ts, model=model, logvar=D.logvar, sqrt_recip_alphas_cumprod=D.sqrt_recip_alphas_cumprod,
                    sqrt_recipm1_alphas_cumprod=D.sqrt_recipm1_alphas_cumprod,
                    posterior_mean_coef1=D.posterior_mean_coef1, posterior_mean_coef2=D.posterior_mean_coef2,
                    return_pred_xstart=True, label=None, data_clamp_range=D.data_clamp_range,
                    local_resampling=local, complete_x0=complete_x0 if local else None,
                    keypoint_mask=mask if local else None)
            torch.randn_like = orig
            x = torch.cat([kp, x[:, :, 3:]], dim=2)
            net_fn = lambda xx, tt: model(xx, ts=tt)
            mine = ref_model.latent_denoise(net_fn, x_T, kp, noises, sch, t_start=t_list[0], n_steps=len(t_list),
                                            complete_x0=complete_x0 if local else None,
                                            keypoint_mask=mask if local else None)
            assert torch.equal(x, mine), "oracle/ref_model.latent_denoise deviates from the reference (%s, local=%s)" % (tag, local)
            gold["out_%s_%s" % (tag, "local" if local else "plain")] = x.numpy()
    # ---- position DDPM: the REAL util.sampling over its full 1000-step chain (stand-in denoiser, pinned draws: x_T,
    # then one std_normal after every step t > 0, util.py:225,253)
    import io
    import contextlib
    import util as ref_util
    torch.Tensor.cuda = lambda self, *a, **k: self
    pcfg = weights.load_json("pipeline_airplane.json")["position_ddpm"]["diffusion_config"]
    dh = ref_util.calc_diffusion_hyperparams(**pcfg)
    Tp = pcfg["T"]
    Bp = 2
    pdraws = [torch.randn(Bp, N, 3, generator=g) for _ in range(Tp)]
    it = iter(pdraws)
    ref_util.std_normal = lambda size: next(it).clone()
    pmodel = stand_in(Tp)
    with contextlib.redirect_stdout(io.StringIO()):
        xp = ref_util.sampling(pmodel, (Bp, N, 3), dh, label=None, verbose=False)
    noises = {t: pdraws[Tp - t] for t in range(Tp - 1, 0, -1)}  # draw i+1 follows step t = T-1-i
    mine = ref_model.position_sampling(lambda xx, tt: pmodel(xx, ts=tt), pdraws[0], noises,
                                       ref_model.position_schedule(pcfg["T"], pcfg["beta_0"], pcfg["beta_T"]))
    assert torch.equal(xp, mine), "oracle/ref_model.position_sampling deviates from the reference"
    gold["pos_draws"] = torch.stack(pdraws).numpy()
    gold["pos_out"] = xp.numpy()
    np.savez_compressed(os.path.join(OUT, "golden_sampler.npz"), **gold)
    print("wrote golden_sampler.npz:", sorted(gold))


if __name__ == "__main__":
    main()
