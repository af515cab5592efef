"""The drop-in `pointnet2_ops` python modules (torch layers over the C-ABI index ops) against the oracle's
restatement of the reference modules, same state dict, same inputs."""
import copy

import pytest
import torch

import slide_b200
from oracle import ref_model
from slide_b200 import weights
from tests import common

pytestmark = pytest.mark.gpu


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def test_sa_and_fp_modules_match_oracle(pipeline_cfg):
    slide_b200.install_dropin()
    from pointnet2_ops.pointnet2_modules import PointnetSAModule, PointnetKnnFPModule
    pc = pipeline_cfg["position_ddpm"]["pointnet_config"]
    sd = common.state_dict("pos")
    att = pc["attention_setting"]
    fd, dd = pc["architecture"]["feature_dim"], pc["architecture"]["decoder_feature_dim"]
    g = torch.Generator().manual_seed(3)
    B = 3
    xyz = torch.rand(B, 16, 3, generator=g) - 0.5
    feats = torch.randn(B, fd[1], 16, generator=g)
    t_emb = torch.randn(B, 4 * pc["t_dim"], generator=g)
    cond = torch.randn(B, 128, generator=g)
    sa = PointnetSAModule(npoint=16, radius=0, nsample=16, mlp=[fd[1], fd[1], fd[1], fd[2]], use_xyz=True,
                          t_dim=4 * pc["t_dim"], include_t=True, include_abs_coordinate=True,
                          include_center_coordinate=True, bn_first=False, first_conv=False, first_conv_in_channel=3,
                          res_connect=True, bias=True, include_condition=True, condition_dim=128, neighbor_def="nn",
                          activation="relu", bn=True, attention_setting=copy.deepcopy(att))
    sa.load_state_dict(_sub(sd, "SA_modules.1."), strict=True)
    sa = sa.cuda().eval()
    with torch.no_grad():
        nx, nf = sa(xyz.cuda(), feats.cuda(), t_emb=t_emb.cuda(), condition_emb=cond.cuda())
        wx, wf = ref_model.sa_module(xyz, feats, ref_model.Params(sd, "SA_modules.1."), 16, 16, pc, t_emb, cond)
    assert torch.equal(nx.cpu(), wx)
    assert (nf.cpu() - wf).abs().max() < 2e-3 * max(1.0, wf.abs().max())  # torch convs may run in TF32 on the GPU
    fp = PointnetKnnFPModule(mlp1=[dd[2], dd[1], dd[1]], mlp2=[dd[1] + fd[1], dd[1], dd[1]], K=8, first_conv=False,
                             bn=True, t_dim=4 * pc["t_dim"], include_t=True, bn_first=False, res_connect=True,
                             bias=True, include_condition=True, condition_dim=128, include_grouper=False, radius=0,
                             nsample=16, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=True,
                             neighbor_def="nn", activation="relu", attention_setting=copy.deepcopy(att))
    fp.load_state_dict(_sub(sd, "FP_modules.1."), strict=True)
    fp = fp.cuda().eval()
    known_f = torch.randn(B, dd[2], 16, generator=g)
    skip_f = torch.randn(B, fd[1], 16, generator=g)
    xyz2 = torch.rand(B, 16, 3, generator=g) - 0.5
    with torch.no_grad():
        got = fp(xyz.cuda(), xyz2.cuda(), skip_f.cuda(), known_f.cuda(), t_emb=t_emb.cuda(), condition_emb=cond.cuda())
        want = ref_model.knn_fp_module(xyz, xyz2, skip_f, known_f, ref_model.Params(sd, "FP_modules.1."), 8, pc, t_emb, cond)
    assert (got.cpu() - want).abs().max() < 2e-3 * max(1.0, want.abs().max())
