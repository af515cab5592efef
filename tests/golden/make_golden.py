"""Regenerates everything under tests/golden/ and slide_b200/configs/ from the REAL reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

  1. exports the hyper-parameter sections of the shipped airplane / chair JSON configs that the sampling path
     reads (list-valued strings restored with the reference's own json_reader), and the state-dict schema
     (key, shape) of the reference's PointNet2CloudCondition / PointAutoencoder built from them;
  2. checks that slide_b200.weights.random_state_dict(schema) loads into the reference modules with strict=True
     (checkpoint ABI) and that oracle/ref_model.py reproduces the reference modules bit for bit on CPU;
  3. writes golden input/output vectors produced by the reference's own python modules (with the C oracle
     standing in for pointnet2_ops._ext / pytorch3d, which have no CPU path / are not installed).
"""
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops, ref_model  # noqa: E402

ops.install_reference_stubs()
from data_utils.json_reader import read_json_file, autoencoder_read_config  # noqa: E402
from models.pointnet2_with_pcld_condition import PointNet2CloudCondition  # noqa: E402
from models.autoencoder import PointAutoencoder  # noqa: E402
from slide_b200 import weights  # noqa: E402

REF = "/root/reference/pointnet2"
CFG = "configs/shapenet_psr_configs/"
OUT = os.path.dirname(os.path.abspath(__file__))
CONF_OUT = os.path.join(ROOT, "slide_b200", "configs")


def dump(name, obj):
    with open(os.path.join(CONF_OUT, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)


def schema_of(module):
    return [[k, list(v.shape)] for k, v in module.state_dict().items()]


def main():
    os.chdir(REF)
    torch.manual_seed(0)
    for cat, syn in (("airplane", "airplane_02691156"), ("chair", "chair_03001627")):
        pos = read_json_file(CFG + "ddpm_keypoint_training_configs/config_standard_attention_batchsize_32_s3_ema_model_keypoint_%s.json" % syn)
        lat = read_json_file(CFG + "latent_ddpm_training_configs/config_latent_ddpm_s3_dim_16_32_ae_kp_noise_0.04_keypoint_conditional_%s_ae_trained_on_%s.json" % (cat, cat))
        ae_file = lat["autoencoder_config"]["config_file"]
        ae = read_json_file(ae_file)
        enc, declist = autoencoder_read_config(os.path.dirname(ae_file), ae)
        label = {"airplane": 0, "chair": 4}[cat]  # sorted synset ids (shapenet_psr_dataset.py:59-67)
        dump("pipeline_%s.json" % cat, {
            "category": cat, "label": label,
            "position_ddpm": {"pointnet_config": pos["pointnet_config"], "diffusion_config": pos["diffusion_config"]},
            "latent_ddpm": {"pointnet_config": lat["pointnet_config"],
                            "standard_diffusion_config": lat["standard_diffusion_config"]},
            "autoencoder": {"encoder": enc, "decoders": declist,
                            "apply_kl_regularization": ae["pointnet_config"].get("apply_kl_regularization", False),
                            "kl_weight": ae["pointnet_config"].get("kl_weight", 0)}})
    cfg = weights.load_json("pipeline_airplane.json")
    pos_cfg = cfg["position_ddpm"]["pointnet_config"]
    lat_cfg = cfg["latent_ddpm"]["pointnet_config"]
    aec = cfg["autoencoder"]
    pos_net = PointNet2CloudCondition(copy.deepcopy(pos_cfg)).eval()
    lat_net = PointNet2CloudCondition(copy.deepcopy(lat_cfg)).eval()
    ae_net = PointAutoencoder(copy.deepcopy(aec["encoder"]), copy.deepcopy(aec["decoders"]),
                              apply_kl_regularization=aec["apply_kl_regularization"], kl_weight=aec["kl_weight"]).eval()
    dump("schema_position_ddpm.json", schema_of(pos_net))
    dump("schema_latent_ddpm.json", schema_of(lat_net))
    dump("schema_autoencoder.json", schema_of(ae_net))

    gold = {}
    B = 2
    g = torch.Generator().manual_seed(1234)
    label = torch.tensor([0, 4])
    for name, net, pc, C, seed in (("pos", pos_net, pos_cfg, 3, 11), ("lat", lat_net, lat_cfg, 51, 12)):
        sd = weights.random_state_dict(weights.load_json("schema_%s_ddpm.json" % {"pos": "position", "lat": "latent"}[name]), seed)
        net.load_state_dict(sd, strict=True)
        x = torch.randn(B, 16, C, generator=g)
        gold[name + "_x"] = x.numpy()
        for t in (999, 500, 0):
            with torch.no_grad():
                y = net(x, ts=torch.ones(B) * t, label=label)
                y2 = ref_model.cloud_condition_net(x, ref_model.Params(sd), pc, ts=torch.ones(B) * t, label=label)
            assert torch.equal(y, y2), "oracle/ref_model.py deviates from the reference"
            gold["%s_eps_t%d" % (name, t)] = y.numpy()
    sd = weights.random_state_dict(weights.load_json("schema_autoencoder.json"), 13)
    ae_net.load_state_dict(sd, strict=True)
    kp = torch.rand(B, 16, 3, generator=g) - 0.5
    feat = torch.randn(B, 16, 48, generator=g)
    starts = [torch.tensor([3, 100]), torch.tensor([7, 2000]), torch.tensor([4000, 11])]
    # the reference draws the FPS start indices with torch.randint on the CPU generator (pytorch3d 0.7.0); pin them
    # by monkey-patching the draw so that reference and oracle use the same three (B,) vectors
    it = iter(starts)
    orig = ops.draw_start_indices
    ops.draw_start_indices = lambda lengths: next(it)
    with torch.no_grad():
        out = ae_net.decode(kp, feat, label=label)
    ops.draw_start_indices = orig
    with torch.no_grad():
        out2, levels = ref_model.decode(kp, feat, ref_model.Params(sd), aec["decoders"], label, start_idx_list=starts)
    assert torch.equal(out, out2), "oracle/ref_model.decode deviates from the reference"
    gold.update(dec_kp=kp.numpy(), dec_feat=feat.numpy(), dec_out=out.numpy(), dec_l1=levels[1].numpy(),
                dec_l2=levels[2].numpy(), dec_starts=torch.stack(starts).numpy(), label=label.numpy())
    # encode (BASELINE config 5 input shape at N=2048): posterior mode and a pinned posterior sample.  The reference
    # draws torch.randn(mean.shape) with mean of shape (B, C, N) on the CPU generator (distributions.py:15-17).
    N = 2048
    pts = torch.rand(B, N, 3, generator=g) - 0.5
    nrm = torch.nn.functional.normalize(torch.randn(B, N, 3, generator=g), dim=2)
    cloud = torch.cat([pts, nrm], dim=2)
    ekp = pts[:, :16].contiguous()
    with torch.no_grad():
        e_mode = ae_net.encode(cloud, ekp, label=label, sample_posterior=False)
        torch.manual_seed(99)
        e_samp = ae_net.encode(cloud, ekp, label=label, sample_posterior=True)
    torch.manual_seed(99)
    n1 = torch.randn(B, 16, 16).transpose(1, 2).contiguous()
    n2 = torch.randn(B, 32, 16).transpose(1, 2).contiguous()
    P = ref_model.Params(sd)
    with torch.no_grad():
        assert torch.equal(e_mode, ref_model.encode(cloud, ekp, P, aec["encoder"], aec["decoders"][0], label))
        assert torch.equal(e_samp, ref_model.encode(cloud, ekp, P, aec["encoder"], aec["decoders"][0], label, noises=(n1, n2)))
    gold.update(enc_cloud=cloud.numpy(), enc_kp=ekp.numpy(), enc_mode=e_mode.numpy(), enc_sample=e_samp.numpy(),
                enc_n1=n1.numpy(), enc_n2=n2.numpy())
    # config 1: FPS + ball query on a 1x2048x3 cloud (C oracle; the reference CUDA kernels cannot run without a GPU)
    xyz = torch.rand(1, 2048, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1
    fps = ops.furthest_point_sampling(xyz, 1024)
    new_xyz = xyz[0][fps[0].long()][None].contiguous()
    bq_idx, bq_cnt = ops.ball_query(new_xyz, xyz, 0.2, 32)
    gold.update(c1_xyz=xyz.numpy(), c1_fps=fps.numpy(), c1_bq_idx=bq_idx.numpy(), c1_bq_cnt=bq_cnt.numpy())
    np.savez_compressed(os.path.join(OUT, "golden.npz"), **gold)
    print("wrote", os.path.join(OUT, "golden.npz"), {k: v.shape for k, v in gold.items()})


if __name__ == "__main__":
    main()
