"""Program assembly for the sampling path: DDPM loops (position / latent) and the autoencoder decode.

build_* functions only need numpy (they return a Builder plus named handles) so that tests can interpret the
records on CPU; the *Sampler classes run them on the GPU through slide_b200.program.Program.

Reference behaviour reproduced here:
  position DDPM   pointnet2/util.py:167-259 (calc_diffusion_hyperparams, sampling)
  latent DDPM     pointnet2/diffusion_utils/diffusion.py:12-39,58-95,158-208,346-404
  decode          pointnet2/models/autoencoder.py:42-45
"""
import numpy as np

from . import nets
from .program import Builder


# ---------------------------------------------------------------------------------------------------
# schedules -> per-timestep coefficient tables (fp32 [T, 8])
# ---------------------------------------------------------------------------------------------------
def position_table(T, beta_0, beta_T):
    """util.calc_diffusion_hyperparams in fp32 torch (sequential in-place products, util.py:182-189) and the
    scalars sampling() derives from it (util.py:247-253): [k1 = (1-a)/sqrt(1-abar), sqrt(a), sigma]."""
    import torch
    Beta = torch.linspace(beta_0, beta_T, T)
    Alpha = 1 - Beta
    Alpha_bar = Alpha + 0
    Beta_tilde = Beta + 0
    for t in range(1, T):
        Alpha_bar[t] *= Alpha_bar[t - 1]
        Beta_tilde[t] *= (1 - Alpha_bar[t - 1]) / (1 - Alpha_bar[t])
    Sigma = torch.sqrt(Beta_tilde)
    tab = torch.zeros(T, 8)
    for t in range(T):
        tab[t, 0] = (1 - Alpha[t]) / torch.sqrt(1 - Alpha_bar[t])
        tab[t, 1] = torch.sqrt(Alpha[t])
        tab[t, 2] = Sigma[t]
    return tab.numpy()


def latent_table(dcfg):
    """Diffusion.init_diffusion_parameters (float64 numpy, diffusion.py:158-208) then the fp32 casts of
    extract() (diffusion.py:31-39) and exp(0.5*logvar) evaluated in fp32 like denoising_step (:88-92):
    [sqrt_recip_acp, sqrt_recipm1_acp, post_mean_coef1, post_mean_coef2, exp(0.5*logvar)]."""
    import torch
    T = dcfg["num_diffusion_timesteps"]
    betas = beta_schedule(dcfg["beta_schedule"], dcfg["beta_start"], dcfg["beta_end"], T)
    var_type = dcfg.get("model_var_type", "fixedsmall")
    with np.errstate(divide="ignore", invalid="ignore"):  # 'jsd' ends at beta = 1: the reference's tables hold inf there too
        alphas = 1.0 - betas
        acp = np.cumprod(alphas, axis=0)
        acp_prev = np.append(1.0, acp[:-1])
        post_var = betas * (1.0 - acp_prev) / (1.0 - acp)
        if var_type == "fixedsmall":      # the shipped configs (diffusion.py:204-206)
            logvar = np.log(np.maximum(post_var, 1e-20))
        elif var_type == "fixedlarge":    # diffusion.py:201-203
            logvar = np.log(np.append(post_var[1], betas[1:]))
        else:
            raise NotImplementedError("model_var_type %r (the reference knows fixedsmall / fixedlarge)" % var_type)
        tab = torch.zeros(T, 8)
        tab[:, 0] = torch.tensor(np.sqrt(1.0 / acp)).float()
        tab[:, 1] = torch.tensor(np.sqrt(1.0 / acp - 1)).float()
        tab[:, 2] = torch.tensor(betas * np.sqrt(acp_prev) / (1.0 - acp)).float()
        tab[:, 3] = torch.tensor((1.0 - acp_prev) * np.sqrt(alphas) / (1.0 - acp)).float()
        tab[:, 4] = torch.exp(0.5 * torch.tensor(logvar).float())
    return tab.numpy()


def beta_schedule(name, beta_start, beta_end, T):
    """get_beta_schedule (diffusion_utils/diffusion.py:12-28), float64.  Every shipped config is 'linear'; 'quad', 'const'
    and 'jsd' are the reference's other working branches ('warmup10' / 'warmup50' call a helper the reference never defines)."""
    if name == "linear":
        return np.linspace(beta_start, beta_end, T, dtype=np.float64)
    if name == "quad":
        return np.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=np.float64) ** 2
    if name == "const":
        return beta_end * np.ones(T, dtype=np.float64)
    if name == "jsd":  # 1/T, 1/(T-1), ..., 1
        return 1.0 / np.linspace(T, 1, T, dtype=np.float64)
    raise NotImplementedError(name)


def _alpha_bar_fp32(T, beta_0, beta_T):
    import torch
    Beta = torch.linspace(beta_0, beta_T, T)
    ab = 1 - Beta
    for t in range(1, T):
        ab[t] *= ab[t - 1]
    return Beta, ab


def fast_position_schedule(method, length, schedule, kappa, dcfg):
    """FastDPM STEP sampler of the position DDPM (util_fastdpmv2.py: STEP_sampling :384-452, get_STEP_step :239-258,
    entry fast_sampling_function_v2 :455-478) as data for a `mode 2` program:
    returns (ts, table) with `length` rows in STEP-COUNTER order (row s is used when the program's step counter is s;
    the sampler's i-th iteration is s = length-1-i): ts[s] = the (possibly fractional) diffusion step the denoiser is
    evaluated at, table[s] = [a, c, sigma] of the update x = x*a + (c*eps + sigma*z).  All coefficient arithmetic is
    fp32 torch in the reference's operation order."""
    import torch
    T, b0, bT = dcfg["T"], dcfg["beta_0"], dcfg["beta_T"]
    Beta, ab = _alpha_bar_fp32(T, b0, bT)
    if method == "step":
        if schedule == "linear":
            taus = [int(np.floor(i * ((T - 1.0) / (length - 1.0)))) for i in range(length)]
        elif schedule == "quadratic":
            taus = [int(v) for v in np.linspace(0, np.sqrt(T * 0.8), length) ** 2]
        else:
            raise NotImplementedError(schedule)
        taus = sorted(taus, reverse=True)
        ts = [float(t) for t in taus]
        cur = [ab[t] for t in taus]
    elif method == "var":
        # VAR_sampling (util_fastdpmv2.py:307-381) cannot run as shipped: its last continuous step comes out as 0.497
        # for every length / schedule of the (T=1000, 1e-4..0.02) configs and trips its own `assert abs(tau) < 0.1`
        raise NotImplementedError("the reference's VAR sampler asserts on its own schedule; use method='step'")
    else:
        raise NotImplementedError(method)
    tab = torch.zeros(length, 8)
    for i in range(length):
        if i == length - 1:
            nxt, sigma = torch.tensor(1.0), torch.tensor(0.0)
        else:
            nxt = cur[i + 1]
            sigma = kappa * torch.sqrt((1 - nxt) / (1 - cur[i]) * (1 - cur[i] / nxt))
        a = torch.sqrt(nxt / cur[i])
        c = torch.sqrt(1 - nxt - sigma ** 2) - torch.sqrt(1 - cur[i]) * torch.sqrt(nxt / cur[i])
        s = length - 1 - i
        tab[s, 0], tab[s, 1], tab[s, 2] = a, c, sigma
    ts_by_counter = np.asarray(ts[::-1], dtype=np.float32)  # torch.ones(n) * tau in the reference: fp32
    return ts_by_counter, tab.numpy()


# ---------------------------------------------------------------------------------------------------
# builders
# ---------------------------------------------------------------------------------------------------
def can_freeze_geometry(cfg, n_points=16):
    """True when a denoiser's neighbour searches can be hoisted out of the step for frozen coordinates: no level of the
    network down-samples (with down-sampling the FPS pick belongs to the step, nets._sa_geometry refuses to defer it).
    All shipped denoisers qualify (npoint == the keypoint count on every level)."""
    n = n_points
    for npoint in cfg["architecture"]["npoint"]:
        if n > npoint:
            return False
        n = min(n, npoint)
    return True


def build_ddpm(cfg, sd, B, T, table, mode, n_points=16, keep_cols=0, clamp=-1.0, with_noise=True,
               local_resampling=False, ts_values=None, resident=None, frozen_xyz=False):
    """One program = setup segment + step segment (+ forward-only segment) for a DDPM denoiser.

    mode 0: position sampler (util.sampling), mode 1: latent sampler (denoising_step), mode 2: FastDPM sampler
    (util_fastdpmv2.VAR_sampling / STEP_sampling; T = number of reverse steps, ts_values / table from
    fast_position_schedule).
    keep_cols: leading columns of x the update must not touch (3 for keypoint-conditional sampling).
    local_resampling (latent sampler only, diffusion.py:76-79): adds the handles x0c [B*n, C] (complete x0) and
    mask [B*n, 1] (1 = re-sample this point's features); with an all-ones mask the update equals the plain one.
    Handles: x, eps, labels, noise [T*B*n, C], ts_table, class_emb.
    frozen_xyz (needs keep_cols >= 3): the coordinates of x never change during a chain (keypoint-conditional sampling,
    diffusion.py:76-95 only updates the feature columns), so the neighbour searches of every module are taken out of the
    step into a segment "geometry" that the caller runs once per chain, after x's coordinates are in place.
    resident: None, or dict(cluster=2|4, precise=bool): also compile the "step" and "forward" ranges into
    sample-resident plans (slide_b200/resident.py; one kernel per step instead of one per record) -> h["resident_plans"]
    (a list, empty when the network does not fit the resident kernel); Program.set_resident installs them.
    """
    b = Builder(B)
    C = 3 + cfg["in_fea_dim"]
    assert cfg["out_dim"] == C
    X = b.tensor("x", n_points, C)
    labels = b.tensor("labels", 1, B, B=1, dtype="i32")
    noise = b.tensor("noise", T * B * n_points if with_noise else 1, C, B=1, ld=C)
    table_off = b.weight(np.asarray(table, dtype=np.float32).reshape(-1, 8)[:T])
    x0c = b.tensor("x0c", n_points, C) if local_resampling else None
    mask = b.tensor("mask", n_points, 1, ld=1) if local_resampling else None
    P = nets.Params(sd)
    b.begin_segment("step")
    b.step_begin()
    assert not frozen_xyz or keep_cols >= 3
    net = nets.lower_cloud_net(b, P, cfg, X, n_points, "net", T=T, labels=labels, factor_group=bool(resident),
                               defer_geometry=frozen_xyz)
    fwd_count = len(b.ops) - b._seg_open[1]
    b.ddpm_update(mode, X, net["out"], noise, table_off, col0=keep_cols, clamp=clamp, x0c=x0c, mask=mask,
                  note="ddpm_update")
    b.end_segment()
    first = b.segments["step"][0]
    b.segments["forward"] = (first, fwd_count)
    b.begin_segment("setup")
    net["emit_setup"]()
    b.end_segment()
    if frozen_xyz:
        b.begin_segment("geometry")
        net["emit_geometry"]()
        b.end_segment()
    h = dict(x=X, eps=net["out"], labels=labels, noise=noise, T=T, C=C, n_points=n_points)
    if local_resampling:
        h.update(x0c=x0c, mask=mask)
    if ts_values is not None:
        assert len(ts_values) == T
        h["ts_values"] = np.asarray(ts_values, dtype=np.float32)
    h.update(net["inputs"])
    if resident:
        from . import resident as res
        plans = []
        try:
            for seg in ("step", "forward"):
                plans.append(res.plan_segment(b, seg, cluster=resident.get("cluster", 4),
                                              precise=resident.get("precise", False), np_points=n_points))
        except res.Unsupported as e:
            plans = []
            h["resident_unsupported"] = str(e)
        h["resident_plans"] = plans
    return b, h


def build_decode(decoder_cfgs, sd, B, n_keypoints=16):
    """PointAutoencoder.decode for B shapes of n_keypoints sparse latent points (16 in the flagship configs; the reference
    also ships 8- and 32-keypoint ablation configs, configs/shapenet_psr_configs/autoencoder_configs/{8,32}_keypoints)."""
    b = Builder(B)
    kp = b.tensor("keypoint", n_keypoints, 3)
    fdim = decoder_cfgs[1]["feature_mapper_setting"]  # noqa: F841  (documented: level-1 features feed level 2)
    P = nets.Params(sd)
    Fdim = sd["keypoint_encoder.fc_layer.weight"].shape[1] - 3
    feat = b.tensor("feature", n_keypoints, Fdim)
    labels = b.tensor("labels", 1, B, B=1, dtype="i32")
    b.begin_segment("decode")
    b.step_begin()
    dec = nets.lower_decode(b, P, decoder_cfgs, kp, feat, labels)
    b.end_segment()
    b.begin_segment("setup")
    dec["emit_setup"]()
    b.end_segment()
    h = dict(keypoint=kp, feature=feat, labels=labels, out=dec["out"], starts=dec["starts"],
             class_tables=dec["class_tables"], levels=dec["levels"])
    return b, h


def build_encode(enc_cfg, kp_cfg, sd, B, n_points, sample_posterior=False, n_keypoints=16):
    """PointAutoencoder.encode for B clouds of n_points (xyz + normal) and their n_keypoints keypoints."""
    b = Builder(B)
    cloud = b.tensor("cloud", n_points, 3 + enc_cfg["in_fea_dim"])
    kp = b.tensor("keypoint", n_keypoints, 3)
    labels = b.tensor("labels", 1, B, B=1, dtype="i32")
    C1 = kp_cfg["architecture"]["feature_dim"][-1]
    C2 = kp_cfg["feature_mapper_setting"]["out_dim"]
    noises = None
    if sample_posterior:
        noises = (b.tensor("noise1", n_keypoints, C1), b.tensor("noise2", n_keypoints, C2))
    P = nets.Params(sd)
    b.begin_segment("encode")
    b.step_begin()
    enc = nets.lower_encode(b, P, enc_cfg, kp_cfg, cloud, kp, labels, noises)
    b.end_segment()
    h = dict(cloud=cloud, keypoint=kp, labels=labels, out=enc["out"], class_tables=enc["class_tables"], noises=noises)
    return b, h


def build_refine(cfg, sd, B, n_in):
    """The SAP refinement / upsampling network and the point split that follows it:
    PointNet2CloudCondition(refine JSON).forward(X, None, ts=None, label) (pointnet2/dpsr_evaluation.py:253) and
    point_upsample (models/point_upsample_module.py:4-46, the first_refine_coarse_points=False branch the shipped
    JSONs use).  X [B*n_in, 3 + in_fea_dim]: xyz, normal and -- for the mirrored configs -- the +-1 indicator column,
    which the split ignores (network_output_to_dpsr_grid, dpsr_evaluation.py:57-66).
    Handles: x, labels, disp [B*n_in, 6*factor], fine [B*n_in*factor, 6], class_emb."""
    if cfg.get("first_refine_coarse_points", False) or cfg.get("include_displacement_center_to_final_output", False):
        raise NotImplementedError("first_refine_coarse_points (no shipped refine JSON sets it)")
    b = Builder(B)
    C = 3 + cfg["in_fea_dim"]
    factor = cfg["point_upsample_factor"]
    F = cfg["out_dim"] // factor if cfg["out_dim"] % factor == 0 and cfg["out_dim"] > 6 else cfg["out_dim"]
    X = b.tensor("x", n_in, C)
    labels = b.tensor("labels", 1, B, B=1, dtype="i32")
    P = nets.Params(sd)
    n_out = int(sd["fc_lyaer.3.weight"].shape[0])
    assert n_out == F * factor, (n_out, F, factor)
    b.begin_segment("refine")
    b.step_begin()
    net = nets.lower_cloud_net(b, P, cfg, X, n_in, "net", T=None, labels=labels)
    fine = b.tensor("fine", n_in * factor, F)
    b.upsample(X, F, net["out"], fine, factor, cfg["output_scale_factor"], note="point_upsample")
    b.end_segment()
    b.begin_segment("setup")
    net["emit_setup"]()
    b.end_segment()
    h = dict(x=X, labels=labels, disp=net["out"], fine=fine, factor=factor, F=F)
    h.update(net["inputs"])
    return b, h


def init_constants(machine, h):
    """Upload the constants a freshly created program needs before its setup segment runs: the timestep list
    0..T-1 and the class-embedding table(s).  `machine` is a Program or the CPU interpreter (same interface)."""
    if "ts_table" in h:
        machine.upload(h["ts_table"], h["ts_values"] if "ts_values" in h else np.arange(h["T"], dtype=np.float32))
    if "class_emb" in h:
        t, w = h["class_emb"]
        machine.upload(t, w)
    for t, w in h.get("class_tables", []):
        machine.upload(t, w)
