"""Golden vectors of the SAP mesh-reconstruction path from the REAL reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_sap.py

  1. exports the refine-and-upsample JSON the shipped mesh_reconstruction.py commands use (pointnet_config +
     dpsr_config) to slide_b200/configs/sap_refine.json and the state-dict schema of the REAL
     PointNet2CloudCondition built from it to schema_sap_refine.json;
  2. asserts that oracle/ref_model.cloud_condition_net and every function of oracle/sap_oracle.py are bit-identical
     (torch.equal) to the REAL modules / functions on CPU: mirror, point_upsample, shapenet_psr_normalize,
     DPSR.forward (point_rasterize, grid_interp, spectral solve) and network_output_to_dpsr_grid;
  3. writes tests/golden/golden_sap.npz: a 2 x 2048 input cloud, the permutation, every 8th row of the network
     output, and indicator grids at small resolutions (so the fixture stays small).
"""
import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ops, ref_model, sap_oracle  # noqa: E402

ops.install_reference_stubs()
from data_utils.json_reader import read_json_file  # noqa: E402
from data_utils.mirror_partial import mirror  # noqa: E402
from models.pointnet2_with_pcld_condition import PointNet2CloudCondition  # noqa: E402
from models.point_upsample_module import point_upsample  # noqa: E402
# dpsr_utils/utils.py imports mesh / rendering packages at module level that are absent here and unused by DPSR
import types  # noqa: E402
for _name, _attrs in (("trimesh", {}), ("plyfile", {"PlyData": None}), ("skimage", {}), ("skimage.measure", {}),
                      ("pytorch3d.renderer", {"PerspectiveCameras": None, "rasterize_meshes": None}),
                      ("igl", {"adjacency_matrix": None, "connected_components": None})):
    if _name not in sys.modules:
        _m = types.ModuleType(_name)
        _m.__dict__.update(_attrs)
        sys.modules[_name] = _m
sys.modules["skimage"].measure = sys.modules["skimage.measure"]
sys.modules["pytorch3d.structures"].Meshes = None
from dpsr_utils.dpsr import DPSR  # noqa: E402
from slide_b200 import weights  # noqa: E402

REF = "/root/reference/pointnet2"
CFG = "configs/shapenet_psr_configs/refine_and_upsample_configs/config_refine_and_upsample_standard_attention_s3_noise_0_symmetry.json"
OUT = os.path.dirname(os.path.abspath(__file__))
CONF_OUT = os.path.join(ROOT, "slide_b200", "configs")


def real_psr_normalize(x):
    # dpsr_evaluation.py imports open3d-style packages at module level; take the two functions out of its source
    import ast
    src = open(os.path.join(REF, "dpsr_evaluation.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "point_upsample": point_upsample}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("shapenet_psr_normalize", "network_output_to_dpsr_grid"):
            exec(compile(ast.Module([node], []), "dpsr_evaluation.py", "exec"), ns)
    return ns["shapenet_psr_normalize"], ns["network_output_to_dpsr_grid"]


def main():
    os.chdir(REF)
    full = read_json_file(CFG)
    pc = full["pointnet_config"]
    with open(os.path.join(CONF_OUT, "sap_refine.json"), "w") as f:
        json.dump({"pointnet_config": pc, "dpsr_config": full["dpsr_config"],
                   "scale": full["shapenet_psr_dataset_config"]["scale"]}, f, indent=1, sort_keys=True)
    net = PointNet2CloudCondition(copy.deepcopy(pc)).eval()
    schema = [[k, list(v.shape)] for k, v in net.state_dict().items()]
    with open(os.path.join(CONF_OUT, "schema_sap_refine.json"), "w") as f:
        json.dump(schema, f, indent=1, sort_keys=True)
    sd = weights.random_state_dict(schema, 21)
    net.load_state_dict(sd, strict=True)

    g = torch.Generator().manual_seed(4321)
    B, N = 2, 2048
    pts = (torch.rand(B, N, 3, generator=g) - 0.5) * torch.tensor([1.0, 0.6, 0.8])
    nrm = torch.nn.functional.normalize(torch.randn(B, N, 3, generator=g), dim=2)
    cloud = torch.cat([pts, nrm], dim=2)
    perm = torch.randperm(2 * N, generator=g)
    label = torch.tensor([0, 4])

    # mirror_and_concat(attach_label=True, permute=True) minus its .cuda(): rebuilt from the REAL mirror()
    mir = mirror(cloud, axis=2)
    one = torch.ones(B, N, 1)
    X = torch.cat([torch.cat([cloud, one], 2), torch.cat([mir, -one], 2)], dim=1)[:, perm, :]
    assert torch.equal(X, sap_oracle.mirror_concat(cloud, perm, axis=2)), "mirror_concat deviates"

    with torch.no_grad():
        disp = net(X, None, ts=None, label=label)
        disp2 = ref_model.cloud_condition_net(X, ref_model.Params(sd), pc, ts=None, label=label)
    assert torch.equal(disp, disp2), "ref_model.cloud_condition_net deviates on the refine config"
    factor = pc["point_upsample_factor"]
    assert disp.shape == (B, 2 * N, 6 * factor)

    fine = point_upsample(X[:, :, :-1], disp, factor, output_scale_factor_value=pc["output_scale_factor"])
    assert torch.equal(fine, ref_model.point_upsample(X[:, :, :-1], disp, factor, pc["output_scale_factor"]))
    real_norm, real_to_grid = real_psr_normalize(None)
    assert torch.equal(real_norm(fine[:, :, :3]), sap_oracle.psr_normalize(fine[:, :, :3]))

    gold = dict(cloud=cloud.numpy(), perm=perm.numpy().astype(np.int32), label=label.numpy(),
                disp_rows8=disp[:, ::8].numpy(), fine_rows64=fine[:, ::64].numpy())
    for res, sig, nb in ((16, 2, 2), (32, 2, 1)):
        dpsr = DPSR(res=(res,) * 3, sig=sig)
        with torch.no_grad():
            phi, rp, rn = real_to_grid(X[:nb], disp[:nb], dpsr, 1, pc, last_dim_as_indicator=True,
                                       only_original_points_split=False, explicit_normalize=True)
            phi2, rp2, rn2 = sap_oracle.refine_to_grid(X[:nb], disp[:nb], (res,) * 3, sig, factor,
                                                       pc["output_scale_factor"])
        assert torch.equal(rp, rp2) and torch.equal(rn, rn2), "unit-cube points deviate"
        assert torch.equal(phi, phi2), ("sap_oracle.dpsr_forward deviates", float((phi - phi2).abs().max()))
        gold["phi_r%d" % res] = phi.numpy()
    # a stand-alone DPSR case with points planted on grid nodes / at the clamp limit (ceil == floor, wrap-around)
    V = torch.rand(1, 600, 3, generator=g) * 0.99
    V[0, :8] = torch.tensor([[0.0, 0.0, 0.0], [0.5, 0.25, 0.75], [0.99, 0.99, 0.99], [0.0, 0.99, 0.5],
                             [0.984375, 0.0, 0.0], [0.96875, 0.96875, 0.96875], [0.9899, 0.5, 0.0], [0.125, 0.0, 0.99]])
    Nn = torch.nn.functional.normalize(torch.randn(1, 600, 3, generator=g), dim=2)
    for shift, scale in ((True, True), (False, False)):
        with torch.no_grad():
            phi = DPSR(res=(32, 32, 32), sig=2, shift=shift, scale=scale)(V, Nn)
            phi2 = sap_oracle.dpsr_forward(V, Nn, (32, 32, 32), 2, shift=shift, scale=scale)
        assert torch.equal(phi, phi2), "sap_oracle.dpsr_forward deviates (edge case)"
        gold["edge_phi_%d%d" % (shift, scale)] = phi.numpy()
    gold.update(edge_V=V.numpy(), edge_N=Nn.numpy())
    # the shipped resolution once, checked here only (8 MB per grid is too large for a fixture): checksum
    with torch.no_grad():
        phi = DPSR(res=(128,) * 3, sig=2)(V, Nn)
        phi2 = sap_oracle.dpsr_forward(V, Nn, (128,) * 3, 2)
    assert torch.equal(phi, phi2)
    gold["edge_phi128_sub"] = phi[:, ::8, ::8, ::8].numpy()
    # the shipped configuration without mirroring and without input normals ("zero_normal": the network estimates them):
    # in_fea_dim 3, 10 children per point, X = [points, zeros] (dpsr_evaluation.py:231-233)
    full2 = read_json_file(CFG.replace("s3_noise_0_symmetry", "s3_zero_normal_noise_0"))
    pc2 = full2["pointnet_config"]
    with open(os.path.join(CONF_OUT, "sap_refine_plain.json"), "w") as f:
        json.dump({"pointnet_config": pc2, "dpsr_config": full2["dpsr_config"], "scale": full2["shapenet_psr_dataset_config"]["scale"],
                   "include_normals": full2["shapenet_psr_dataset_config"].get("include_normals", True)}, f, indent=1, sort_keys=True)
    net2 = PointNet2CloudCondition(copy.deepcopy(pc2)).eval()
    schema2 = [[k, list(v.shape)] for k, v in net2.state_dict().items()]
    with open(os.path.join(CONF_OUT, "schema_sap_refine_plain.json"), "w") as f:
        json.dump(schema2, f, indent=1, sort_keys=True)
    sd2 = weights.random_state_dict(schema2, 22)
    net2.load_state_dict(sd2, strict=True)
    X2 = torch.cat([cloud[:, :, :3], torch.zeros_like(cloud[:, :, :3])], dim=2)
    with torch.no_grad():
        disp_p = net2(X2, None, ts=None, label=label)
        assert torch.equal(disp_p, ref_model.cloud_condition_net(X2, ref_model.Params(sd2), pc2, ts=None, label=label))
        dpsr = DPSR(res=(16, 16, 16), sig=2)
        phi_p, rp, rn = real_to_grid(X2, disp_p, dpsr, 1, pc2, last_dim_as_indicator=False, only_original_points_split=False,
                                     explicit_normalize=True)
        phi_p2, rp2, rn2 = sap_oracle.refine_to_grid(X2, disp_p, (16, 16, 16), 2, pc2["point_upsample_factor"],
                                                     pc2["output_scale_factor"], indicator=False)
    assert torch.equal(phi_p, phi_p2) and torch.equal(rp, rp2) and torch.equal(rn, rn2)
    gold.update(plain_disp_rows8=disp_p[:, ::8].numpy(), plain_phi_r16=phi_p.numpy())
    np.savez_compressed(os.path.join(OUT, "golden_sap.npz"), **gold)
    print("wrote golden_sap.npz", {k: v.shape for k, v in gold.items()})


if __name__ == "__main__":
    main()
