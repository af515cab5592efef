"""Golden vectors for the reference's shipped 8- and 32-keypoint ablation configs (position DDPM, feature DDPM and the
autoencoders with latent dims 4_8 / 8_16 / 16_32 / 32_64), produced by the REAL reference modules (build container only:
needs /root/reference; the C oracle stands in for `_ext` / pytorch3d):

    python tests/golden/make_golden_ablation.py
        -> slide_b200/configs/ablation_{8,32}kps.json   hparams (the reference's own json_reader) + state-dict schemas
        -> tests/golden/golden_ablation.npz             inputs / outputs per family and autoencoder set

Reference configs: pointnet2/configs/shapenet_psr_configs/{ddpm_keypoint_training_configs,latent_ddpm_training_configs,
autoencoder_configs}/{8,32}_keypoints/.  Also asserts that oracle/ref_model.py is bit-identical to the real modules on
these configs (it was written against the 16-keypoint ones)."""
import copy
import glob
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
AE_SETS = {8: ["test_configs_latent_dim_16_32", "test_configs_latent_dim_8_16", "test_configs_latent_dim_32_64"],
           32: ["test_configs_latent_dim_16_32_keypoints_32", "test_configs_latent_dim_4_8_keypoints_32",
                "test_configs_latent_dim_8_16_keypoints_32"]}
SEEDS = {"pos": 31, "lat": 32, "ae": 33}
STARTS = [[3, 5], [7, 100], [400, 11]]


def schema_of(module):
    return [[k, list(v.shape)] for k, v in module.state_dict().items()]


def main():
    os.chdir(REF)
    gold = {}
    B = 2
    label = torch.tensor([0, 4])
    for kps in (8, 32):
        g = torch.Generator().manual_seed(500 + kps)
        pos = read_json_file(glob.glob(CFG + "ddpm_keypoint_training_configs/%d_keypoints/*airplane*" % kps)[0])
        lat = read_json_file(glob.glob(CFG + "latent_ddpm_training_configs/%d_keypoints/*airplane*" % kps)[0])
        fam = {"num_keypoints": kps,
               "position_ddpm": {"pointnet_config": pos["pointnet_config"], "diffusion_config": pos["diffusion_config"]},
               "latent_ddpm": {"pointnet_config": lat["pointnet_config"],
                               "standard_diffusion_config": lat["standard_diffusion_config"]},
               "autoencoders": {}}
        for name, key, C in (("pos", "position_ddpm", 3), ("lat", "latent_ddpm", 3 + lat["pointnet_config"]["in_fea_dim"])):
            pc = fam[key]["pointnet_config"]
            net = PointNet2CloudCondition(copy.deepcopy(pc)).eval()
            fam[key]["schema"] = schema_of(net)
            sd = weights.random_state_dict(fam[key]["schema"], SEEDS[name])
            net.load_state_dict(sd, strict=True)
            x = torch.randn(B, kps, C, generator=g)
            gold["k%d_%s_x" % (kps, name)] = x.numpy()
            for t in (999, 0):
                with torch.no_grad():
                    y = net(x, ts=torch.ones(B) * t, label=label)
                    y2 = ref_model.cloud_condition_net(x, ref_model.Params(sd), pc, ts=torch.ones(B) * t, label=label)
                assert torch.equal(y, y2), "oracle/ref_model.py deviates from the reference (%d keypoints, %s)" % (kps, name)
                gold["k%d_%s_eps_t%d" % (kps, name, t)] = y.numpy()
        pts = torch.rand(B, 2048, 3, generator=g) - 0.5
        nrm = torch.nn.functional.normalize(torch.randn(B, 2048, 3, generator=g), dim=2)
        cloud = torch.cat([pts, nrm], dim=2)
        gold["k%d_enc_cloud" % kps] = cloud.numpy()
        ekp = pts[:, :kps].contiguous()
        ae_file = glob.glob(CFG + "autoencoder_configs/%d_keypoints/*airplane*.json" % kps)[0]
        for sub in AE_SETS[kps]:
            ae = read_json_file(ae_file)
            ae["pointnet_config"]["encoder_config_file"] = sub + "/config_encoder.json"
            ae["pointnet_config"]["decoder_config_file"] = [sub + "/decoder_level_%d.json" % i for i in (1, 2, 3)]
            enc, decs = autoencoder_read_config(os.path.dirname(ae_file), ae)
            net = PointAutoencoder(copy.deepcopy(enc), copy.deepcopy(decs),
                                   apply_kl_regularization=ae["pointnet_config"].get("apply_kl_regularization", False),
                                   kl_weight=ae["pointnet_config"].get("kl_weight", 0)).eval()
            tag = sub.replace("test_configs_", "").replace("_keypoints_32", "")
            fam["autoencoders"][tag] = {"encoder": enc, "decoders": decs, "schema": schema_of(net)}
            sd = weights.random_state_dict(fam["autoencoders"][tag]["schema"], SEEDS["ae"])
            net.load_state_dict(sd, strict=True)
            fdim = sd["keypoint_encoder.fc_layer.weight"].shape[1] - 3
            kp = torch.rand(B, kps, 3, generator=g) - 0.5
            feat = torch.randn(B, kps, fdim, generator=g)
            starts = [torch.tensor(s) for s in STARTS]
            it = iter(starts)
            orig = ops.draw_start_indices
            ops.draw_start_indices = lambda lengths: next(it)   # pin pytorch3d's CPU randint draws
            with torch.no_grad():
                out = net.decode(kp, feat, label=label)
            ops.draw_start_indices = orig
            with torch.no_grad():
                out2, levels = ref_model.decode(kp, feat, ref_model.Params(sd), decs, label, start_idx_list=starts)
                e_mode = net.encode(cloud, ekp, label=label, sample_posterior=False)
                e2 = ref_model.encode(cloud, ekp, ref_model.Params(sd), enc, decs[0], label)
            assert torch.equal(out, out2) and torch.equal(e_mode, e2), "ref_model deviates (%d keypoints, %s)" % (kps, tag)
            pre = "k%d_%s_" % (kps, tag)
            gold.update({pre + "dec_kp": kp.numpy(), pre + "dec_feat": feat.numpy(), pre + "dec_l1": levels[1].numpy(),
                         pre + "dec_out": out.numpy(), pre + "enc_mode": e_mode.numpy()})
            print(kps, tag, "latent", fdim, "levels", [tuple(l.shape[1:]) for l in levels], "encode", tuple(e_mode.shape))
        with open(os.path.join(CONF_OUT, "ablation_%dkps.json" % kps), "w") as f:
            json.dump(fam, f, indent=None, sort_keys=True, separators=(",", ":"))
    gold["label"] = label.numpy()
    gold["dec_starts"] = np.asarray(STARTS)
    path = os.path.join(OUT, "golden_ablation.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
